#!/bin/bash
# 8-GPU run: scaling point of bench.py (torchrun, one process per GPU) and the in-process
# 8-GPU worker pool (--gpuids 0..7) with the NCCL statistics reduce.
mkdir -p gpurun_out
LOG=gpurun_out/n8.log
{
	nvidia-smi -L
	echo "nproc: $(nproc)"
	echo "cgroup memory.max: $(cat /sys/fs/cgroup/memory.max 2>&1)"
	echo "cgroup memory.current: $(cat /sys/fs/cgroup/memory.current 2>&1)"
	df -h /dev/shm
	free -g
} > $LOG 2>&1

( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
	--master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 \
	> gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err ) 2> gpurun_out/bench_n8.time
echo "bench n8 exit: $?" >> $LOG
cat gpurun_out/bench_n8.json >> $LOG
tail -3 gpurun_out/bench_n8.time >> $LOG
rm -rf /dev/shm/elb_bench_* 2>/dev/null

# in-process pool: 64 workers over 8 GPUs, 8 files x 4 GiB, live line + NCCL stats reduce
FILES=""
for i in 0 1 2 3 4 5 6 7; do FILES="$FILES /dev/shm/elb_pool_$i.bin"; done
ELB_FORCE_LIVESTATS=1 timeout 300 elbencho_b200/elbencho-b200 -w -r -t 64 -b 1M -s 4G --verify 1 \
	--gpuids 0,1,2,3,4,5,6,7 --liveint 500 --lat $FILES > gpurun_out/cli_n8_pool.log 2>&1
echo "cli pool exit: $?" >> $LOG
tr '\r' '\n' < gpurun_out/cli_n8_pool.log | grep -v "^.\[2K$" | tail -45 >> $LOG
rm -f /dev/shm/elb_pool_*.bin

# config 5 shape, scaled: dir tree 128 threads x 8 dirs x 128 files x 64 KiB over 8 GPUs
mkdir -p /dev/shm/elb_tree
timeout 300 elbencho_b200/elbencho-b200 -d -w -r -F -D -t 128 -n 8 -N 128 -s 64K -b 64K --verify 1 \
	--gpuids 0,1,2,3,4,5,6,7 --nolive /dev/shm/elb_tree > gpurun_out/cli_n8_tree.log 2>&1
echo "cli tree exit: $?" >> $LOG
cat gpurun_out/cli_n8_tree.log >> $LOG
rm -rf /dev/shm/elb_tree
tail -150 $LOG
