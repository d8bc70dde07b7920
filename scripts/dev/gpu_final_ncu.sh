#!/bin/bash
# final captures of round 2; summaries are produced on the box (the .ncu-rep files together exceed gpurun's 64 MiB return limit)
mkdir -p gpurun_out/final
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/final/r02_launchlist_bench.json 2> /dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final/r02_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/final/r02_launches.log 2>&1
echo "launch list rc=$?"
python scripts/ncu_launches.py gpurun_out/final/r02_launches.csv > gpurun_out/final/r02_launch_list_summary.txt
for cfg in C5:2097152 C3a:2097152 C3b:2097152 C2:0; do
  c=${cfg%%:*}; t=${cfg##*:}; targ=""; [ "$t" != "0" ] && targ="--targets $t"
  ncu --set full --clock-control none --import-source on -k regex:"search_kernel|local_solve" -c 3 -o /tmp/r02_final_$c -f \
    python bench.py --config $c $targ --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final/r02_final_$c.log 2>&1
  echo "ncu $c rc=$?"
  python scripts/ncu_summary.py /tmp/r02_final_$c.ncu-rep > gpurun_out/final/r02_final_${c}_summary.txt 2>&1
  ncu -i /tmp/r02_final_$c.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_$c.csv 2>/dev/null
  python scripts/ncu_lines.py /tmp/src_$c.csv 40 > gpurun_out/final/r02_final_${c}_lines.txt 2>&1
  python scripts/ncu_smem.py /tmp/src_$c.csv 14 > gpurun_out/final/r02_final_${c}_smem_lines.txt 2>&1
done
cp /tmp/r02_final_C5.ncu-rep gpurun_out/final/ 2>/dev/null
du -sh gpurun_out/final
