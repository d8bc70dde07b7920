#!/bin/bash
python scripts/dev/sgs_profile_case.py > gpurun_out/sgs_case.log 2>&1 || { tail -5 gpurun_out/sgs_case.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_sgs_launches.csv \
  python scripts/dev/sgs_profile_case.py > gpurun_out/sgs_ncu.log 2>&1
echo "launch list rc=$?"
python scripts/ncu_launches.py gpurun_out/r02_sgs_launches.csv | tee gpurun_out/r02_sgs_launch_list_summary.txt
ncu --set full --clock-control none --import-source on -k regex:"search_kernel|sgs_weights|sgs_level" -c 6 -o /tmp/r02_sgs -f \
  python scripts/dev/sgs_profile_case.py > gpurun_out/sgs_ncu_full.log 2>&1
echo "ncu full rc=$?"
python scripts/ncu_summary.py /tmp/r02_sgs.ncu-rep > gpurun_out/r02_sgs_summary.txt 2>&1
tail -5 gpurun_out/sgs_case.log
