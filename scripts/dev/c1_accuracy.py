"""max |GPU − oracle| on the Gaussian global configs (C1 and the reference's own 2-D / 1-D problems)"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import gskrige, oracle_py as O
from _cases import ref_problem_1d, ref_problem_2d
ctx = gskrige.Context(0)
for name, spec in (("C1", gskrige.synth.config_spec("C1")), ("C1 40x40 n=300", gskrige.synth.config_spec("C1", grid=(40, 40), n=300)),
                   ("ref2d global", ref_problem_2d(gskrige, 0, None)), ("ref1d global", ref_problem_1d(gskrige, 0, None)),
                   ("C4-shaped n=1500", gskrige.synth.config_spec("C4", grid=(128, 128), n=1500))):
    m, v = ctx.krige(spec)
    om, ov = O.krige(spec)
    print(f"{name:22s} max|dmean| {np.abs(m-om).max():.3e}  max|dvar| {np.abs(v-ov).max():.3e}", flush=True)
