#!/bin/bash
# Round-end check on one B200: the whole GPU suite, smoke(), the default bench line, then (only after those exited 0
# without a profiler) the ncu launch list of the bench command and a --set full capture of the compact CG kernels.
set -u
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_parity_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_gpu_parity_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r2_bench_n1_line.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_n1.err
B='python bench.py --steps 2 --warmup 3 --c3 0 --cpu-points 0 --f64-points 0 --parity-points 0'
timeout 120 $B > gpurun_out/r2_plain_bench.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_bench_launches_final.csv $B > /dev/null 2>&1
timeout 100 python scripts/profile_inpaint.py 5001 1 > gpurun_out/r2_plain_inp.log 2>&1 || exit 1
timeout 300 ncu --set full --clock-control none -k regex:'compact_(p|apply|update)_kernel' -s 3 -c 3 -o /tmp/r2_compact python scripts/profile_inpaint.py 5001 1 > /dev/null 2>&1
ncu -i /tmp/r2_compact.ncu-rep --page raw --csv > gpurun_out/r2_inpaint_compact_kernels.raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
