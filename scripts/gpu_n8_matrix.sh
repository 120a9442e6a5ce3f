#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/scale_matrix.py /dev/shm 16 2 > gpurun_out/scale_matrix.jsonl 2> gpurun_out/scale_matrix.err
echo "matrix rc=$?"; tail -3 gpurun_out/scale_matrix.err
cat gpurun_out/scale_matrix.jsonl
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 \
	--master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 \
	> gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err ) 2> gpurun_out/bench_n4.time
echo "bench n4 rc=$?"; cut -c1-400 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.time
