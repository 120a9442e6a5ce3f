#!/bin/bash
# first GPU validation: tests, smoke, small bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --file-gib 8 --cpu-sample-gib 4 --threads 4 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench rc=$?"; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
