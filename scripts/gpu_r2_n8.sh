# round 2: 8-GPU point. Box facts, the bench under torchrun (incl. the in-process pool leg on rank 0),
# the reference arm on the same box, and the NUMA binding off for comparison.
mkdir -p gpurun_out
{
nvidia-smi topo -m | head -12
cat /sys/fs/cgroup/cpu.max /sys/fs/cgroup/memory.max
for d in /sys/bus/pci/devices/*; do if [ -e $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null && grep -q 0x0302 $d/class 2>/dev/null; then echo "$d $(cat $d/numa_node)"; fi; done
free -g | head -2
} > gpurun_out/r02_n8_box.log 2>&1
run8() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
( time NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 900 bash -c "$(declare -f run8); run8 29611" > gpurun_out/r02_bench_c2_n8.json 2> gpurun_out/r02_bench_c2_n8.err ) 2> gpurun_out/r02_bench_c2_n8.time
echo "n8 rc=$?"; cut -c1-300 gpurun_out/r02_bench_c2_n8.json; tail -3 gpurun_out/r02_bench_c2_n8.time
( time timeout 900 python bench.py --impl reference --gpus 8 > gpurun_out/r02_bench_c2_n8_ref.json 2> gpurun_out/r02_bench_c2_n8_ref.err ) 2> gpurun_out/r02_bench_c2_n8_ref.time
echo "n8 ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_c2_n8_ref.json
timeout 600 bash -c "$(declare -f run8); run8 29612 --no-gpu-numa --skip-pool --skip-kernels --steps 6 --warmup 2" > gpurun_out/r02_bench_c2_n8_unbound.json 2> gpurun_out/r02_bench_c2_n8_unbound.err
echo "n8 unbound rc=$?"; cut -c1-300 gpurun_out/r02_bench_c2_n8_unbound.json
grep -h "NCCL INFO.*nranks\|comm 0x.*rank" gpurun_out/r02_bench_c2_n8.err | head -12
