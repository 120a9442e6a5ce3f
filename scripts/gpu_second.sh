#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; tail -3 gpurun_out/pytest_k.log
timeout 300 python bench.py --skip-e2e --skip-cpu > gpurun_out/bench_kernel.json 2> gpurun_out/bench_kernel.err; cat gpurun_out/bench_kernel.json | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], json.dumps(d['roofline']['all_kernels']))"
timeout 900 python scripts/explore_e2e.py /dev/shm 8 > gpurun_out/explore_shm.jsonl 2> gpurun_out/explore_shm.err; cat gpurun_out/explore_shm.jsonl; tail -3 gpurun_out/explore_shm.err
