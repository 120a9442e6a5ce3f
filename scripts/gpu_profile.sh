#!/bin/bash
# ncu evidence for profiles/: launch list of a (small) full bench command + one full-set capture
# of each of the three block kernels. Run under gpurun on ONE GPU.
mkdir -p gpurun_out
# (ELB_NO_CUDA_GRAPHS=1: ncu 2025.x crashes on graphs captured/launched concurrently by several
#  worker threads; the kernels and copies of the GPU stage are the same, enqueued call by call)
ELB_NO_CUDA_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --window-gib 1 --file-gib 1 --cpu-sample-gib 0.25 --threads 2 --cpu-threads 2 \
    > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:elb_blocks -s 2 -c 3 \
    -f -o gpurun_out/prof_blocks_kernels \
    python bench.py --steps 1 --warmup 1 --window-gib 1 --skip-e2e --skip-cpu \
    > gpurun_out/prof_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
