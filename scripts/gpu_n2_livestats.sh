#!/bin/bash
# 2-GPU run: live statistics reduce (gather kernel + grouped ncclReduce) tests, then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/livestats_n2.log 2>&1
timeout 600 python -m pytest tests/test_livestats_gpu.py -x -q >> gpurun_out/livestats_n2.log 2>&1
echo "pytest livestats exit: $?" >> gpurun_out/livestats_n2.log
ELB_FORCE_LIVESTATS=1 timeout 120 elbencho_b200/elbencho-b200 -w -r -t 8 -b 1M -s 8G --verify 1 \
	--gpuids 0,1 --liveint 200 /dev/shm/elb_live_cli.bin > gpurun_out/cli_n2_live.log 2>&1
echo "cli exit: $?" >> gpurun_out/livestats_n2.log
tr '\r' '\n' < gpurun_out/cli_n2_live.log | tail -30 >> gpurun_out/livestats_n2.log
rm -f /dev/shm/elb_live_cli.bin
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_n2.log 2>&1
echo "pytest gpu suite exit: $?" >> gpurun_out/livestats_n2.log
tail -5 gpurun_out/pytest_gpu_n2.log >> gpurun_out/livestats_n2.log
tail -60 gpurun_out/livestats_n2.log
