#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo "ref rc=$?"; cut -c1-900 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.time
( time timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time; echo "bench rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.time
bash scripts/gpu_profile.sh
