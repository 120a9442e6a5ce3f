#!/bin/bash
mkdir -p gpurun_out
ARGS="--steps 5 --warmup 3 --window-gib 1 --file-gib 1 --cpu-sample-gib 0.25 --threads 2 --cpu-threads 2 --single-thread-sample-gib 0"
echo "== plain"; MALLOC_CHECK_=3 timeout 300 python bench.py $ARGS > gpurun_out/dbg_plain.json 2> gpurun_out/dbg_plain.err; echo "rc=$?"; tail -2 gpurun_out/dbg_plain.err
echo "== ncu, no graphs"
ELB_NO_CUDA_GRAPHS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py $ARGS > gpurun_out/bench_under_ncu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300; wc -l gpurun_out/launches_bench.csv
echo "== ncu, graphs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_graphs.csv python bench.py $ARGS > gpurun_out/bench_under_ncu_graphs.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_under_ncu_graphs.log | cut -c1-300; wc -l gpurun_out/launches_bench_graphs.csv
