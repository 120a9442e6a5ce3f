#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_worker_gpu.py -m gpu -x -q -k "multi_gpu or rank_offset" 2>&1 | tail -3
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2> gpurun_out/bench_n2.time
echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(json.dumps({k:d.get(k) for k in ('value','n_gpus','ms_per_step','e2e')}, indent=0))"; tail -5 gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.time
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 20 --warmup 3 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err ) 2> gpurun_out/bench_n2_ref.time; cut -c1-400 gpurun_out/bench_n2_ref.json; tail -3 gpurun_out/bench_n2_ref.time
