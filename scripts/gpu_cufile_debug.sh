#!/bin/bash
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
sed 's/"level": "ERROR"/"level": "TRACE"/' /etc/cufile.json > /tmp/cufile_debug.json
export CUFILE_ENV_PATH_JSON=/tmp/cufile_debug.json
cd /tmp && rm -f cufile.log
timeout 60 $GRAFT_REPO_ROOT/build/cufile_probe /dev/shm/cufile_probe.bin 2>&1 | tail -8
echo "---- cufile.log (filtered)"; grep -v "Batch Ctx state\|Thread\|thread" /tmp/cufile.log | head -150
cat /proc/mounts | head -40
