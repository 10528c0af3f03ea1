#!/bin/bash
# ncu evidence of the §8(f) kernels at the end of round 2 (run on the GPU box from the repo root; summaries only: reports stay in /tmp)
export ROWS=67108864 REPS=1
SHAPES="M4 hash join" timeout 600 ncu --set full --clock-control none -k regex:"join_probe_count|join_probe_fused|join_probe_pairs|join_gather" -c 6 -f -o /tmp/r02_join python tools/bench_shapes.py > gpurun_out/r02_ncu_join.log 2>&1; tail -2 gpurun_out/r02_ncu_join.log | cut -c1-200
python tools/ncu_summary.py /tmp/r02_join.ncu-rep > gpurun_out/r02_ncu_join_final.txt
SHAPES="M5 sort" timeout 600 ncu --set full --clock-control none -k regex:"sort_scatter_kernel|sort_tile_hist_kernel|sort_normalise_kernel|sort_digit_hist_kernel" -c 6 -f -o /tmp/r02_sort python tools/bench_shapes.py > gpurun_out/r02_ncu_sort.log 2>&1; tail -2 gpurun_out/r02_ncu_sort.log | cut -c1-200
python tools/ncu_summary.py /tmp/r02_sort.ncu-rep > gpurun_out/r02_ncu_sort_final.txt
SHAPES="M6 parquet scan store_sales-like 4 columns, uncompressed" timeout 600 ncu --set full --clock-control none -k regex:"pq_decode_kernel" -c 8 -f -o /tmp/r02_pq python tools/bench_shapes.py > gpurun_out/r02_ncu_pq.log 2>&1; tail -2 gpurun_out/r02_ncu_pq.log | cut -c1-200
python tools/ncu_summary.py /tmp/r02_pq.ncu-rep > gpurun_out/r02_ncu_parquet_decode.txt
unset ROWS REPS
B200Q_BENCH_E2E_ROWS=16777216 B200Q_BENCH_E2E_SMALL_ROWS=1000000 B200Q_BENCH_CPU_ROWS=4194304 B200Q_BENCH_X_ROWS=67108864 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches_bench.log 2>&1; tail -c 200 gpurun_out/r02_launches_bench.log; wc -l gpurun_out/r02_launches_final.csv
wc -l gpurun_out/r02_ncu_*final.txt gpurun_out/r02_ncu_parquet_decode.txt; du -sh gpurun_out
