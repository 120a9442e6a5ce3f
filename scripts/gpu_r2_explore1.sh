# round 2, exploration 1: host topology + host-side path measurements (1 GPU)
set -x
mkdir -p gpurun_out
{
lscpu | egrep -i 'model name|socket|core|thread|numa|l2|l3|mhz'
cat /sys/devices/system/node/node*/cpulist
nvidia-smi topo -m
cat /sys/kernel/mm/transparent_hugepage/shmem_enabled /sys/kernel/mm/transparent_hugepage/enabled
cat /sys/kernel/mm/lru_gen/enabled
cat /proc/sys/kernel/perf_event_paranoid
cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/cpu.max 2>/dev/null
cat /proc/self/status | egrep 'Cpus_allowed_list|Mems_allowed_list'
grep -i -E 'ddio|iommu' /proc/cmdline
} > gpurun_out/r02_topo.log 2>&1
timeout 500 scripts/explore_hostpath.bin pcie,zcopy,tmpfs,stacks,wfiles,chase,chasek,wpipe 8 > gpurun_out/r02_hostpath.jsonl 2> gpurun_out/r02_hostpath.err
tail -5 gpurun_out/r02_hostpath.jsonl
