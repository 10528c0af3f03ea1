#!/bin/bash
# 2-GPU validation (run with gpurun --gpus 2): the N>1 test and the bench line at N=2 with a bounded row count
(timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_exchange_threads.py -x -q 2>&1 | tail -8)
B200Q_BENCH_ROWS=${B200Q_BENCH_ROWS:-500000000} B200Q_BENCH_E2E_ROWS=67108864 B200Q_BENCH_E2E_SMALL_ROWS=4000000 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -3 gpurun_out/r02_bench_n2.err; head -c 1500 gpurun_out/r02_bench_n2.json
