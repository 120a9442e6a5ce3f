#!/bin/bash
# final validation of the round on 2 GPUs: whole GPU suite, smoke, both bench arms at N=1, N=2 point
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > gpurun_out/bench_full5.json 2> gpurun_out/bench_full5.err ) 2> gpurun_out/bench_full5.time
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_full5.json; tail -3 gpurun_out/bench_full5.time
( time timeout 900 python bench.py --impl reference > gpurun_out/bench_ref5.json 2> gpurun_out/bench_ref5.err ) 2> gpurun_out/bench_ref5.time
echo "bench ref rc=$?"; cut -c1-200 gpurun_out/bench_ref5.json
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
	--master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 \
	> gpurun_out/bench_n2_v2.json 2> gpurun_out/bench_n2_v2.err ) 2> gpurun_out/bench_n2_v2.time
echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/bench_n2_v2.json
