#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
for p in /dev/shm/cufile_probe.bin /tmp/cufile_probe.bin; do
  echo "=== cufile probe on $p"; timeout 120 ./build/cufile_probe $p 2>&1 | tail -25
done > gpurun_out/cufile_probe.log 2>&1
cat gpurun_out/cufile_probe.log
ls /usr/local/cuda/gds/ 2>/dev/null; cat /etc/cufile.json 2>/dev/null | grep -v '^\s*//' | head -60 > gpurun_out/cufile_json.txt
( time timeout 900 python bench.py > gpurun_out/bench_full2.json 2> gpurun_out/bench_full2.err ) 2> gpurun_out/bench_full2.time; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_full2.json')); print(json.dumps({k:d[k] for k in ('value','e2e','cpu_baseline','storage_roofline','single_thread')}, indent=0))"; tail -5 gpurun_out/bench_full2.err; tail -3 gpurun_out/bench_full2.time
