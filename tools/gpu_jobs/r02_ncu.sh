#!/bin/bash
# ncu evidence of round 2 (run on the GPU box from the repo root): full-set captures of the hot kernels at 2^27 rows per launch
# (summarised on the box: the reports themselves exceed what gpurun brings back), the launch list of bench.py.
export ROWS=134217728 REPS=1
SHAPES="M0 filter,M1 dense,M2 q1-shaped fused,M3 shuffle write 200" timeout 900 ncu --set full --clock-control none -k regex:"agg_tile_dense_kernel|agg_lean_dense_kernel|filter_count_lean|filter_apply_lean|shuffle_encode_kernel|shuffle_pids_kernel" -c 8 -f -o /tmp/r02_hot python tools/bench_shapes.py > gpurun_out/r02_ncu_hot.log 2>&1; tail -3 gpurun_out/r02_ncu_hot.log
python tools/ncu_summary.py /tmp/r02_hot.ncu-rep > gpurun_out/r02_ncu_hot_kernels.txt; ls -la /tmp/r02_hot.ncu-rep
SHAPES="wide tile" timeout 600 ncu --set full --clock-control none -k regex:"agg_tile_wide_kernel" -c 5 -f -o /tmp/r02_wide python tools/bench_shapes.py > gpurun_out/r02_ncu_wide.log 2>&1; tail -3 gpurun_out/r02_ncu_wide.log
python tools/ncu_summary.py /tmp/r02_wide.ncu-rep > gpurun_out/r02_ncu_wide_kernels.txt
unset ROWS REPS
B200Q_BENCH_E2E_ROWS=16777216 B200Q_BENCH_E2E_SMALL_ROWS=1000000 B200Q_BENCH_CPU_ROWS=4194304 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches_bench.log 2>&1; tail -c 300 gpurun_out/r02_launches_bench.log; wc -l gpurun_out/r02_launches.csv
du -sh gpurun_out
