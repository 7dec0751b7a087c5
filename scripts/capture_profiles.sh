#!/bin/bash
# Round-2 evidence under gpurun_out/ (kept small: raw CSV pages, no .ncu-rep): launch list of the bench command, ncu --set full
# of the shipped opening kernels (W = 1, 4 fused; W = 18 erosion + dilation passes) and of the inpainter's kernels.
set -u
B='python bench.py --steps 2 --warmup 3 --c3 0 --cpu-points 0 --f64-points 0 --parity-points 0'
$B > gpurun_out/r2_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_bench_launches_final.csv $B > /dev/null 2>&1
python scripts/profile_open.py 8192 1 > gpurun_out/r2_plain_open.log 2>&1 || exit 1
for spec in "w01 0 1" "w04 3 1" "w18 28 2"; do
  set -- $spec
  ncu --set full --clock-control none -k regex:'open_(march|pass)_kernel' -s $2 -c $3 -o /tmp/r2_open_$1 python scripts/profile_open.py 8192 1 > /dev/null 2>&1
  ncu -i /tmp/r2_open_$1.ncu-rep --page raw --csv > gpurun_out/r2_open_$1.raw.csv 2>/dev/null
done
python scripts/profile_inpaint.py 5001 1 > gpurun_out/r2_plain_inp.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:'(down_kernel|up_kernel|apply_kernel|update_kernel|p_update_kernel|tail_kernel)' -s 0 -c 14 -o /tmp/r2_inpaint python scripts/profile_inpaint.py 5001 1 > /dev/null 2>&1
ncu -i /tmp/r2_inpaint.ncu-rep --page raw --csv > gpurun_out/r2_inpaint_kernels.raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
