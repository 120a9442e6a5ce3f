#!/bin/bash
# full GPU suite, full bench (both arms), ncu evidence after the tiled-kernel change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py > gpurun_out/bench_full4.json 2> gpurun_out/bench_full4.err ) 2> gpurun_out/bench_full4.time
echo "bench rc=$?"; cat gpurun_out/bench_full4.json | cut -c1-1500
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_ref4.json 2> gpurun_out/bench_ref4.err ) 2> gpurun_out/bench_ref4.time
echo "bench ref rc=$?"; cat gpurun_out/bench_ref4.json | cut -c1-600
bash scripts/gpu_profile.sh 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
