#!/bin/bash
# ncu full capture at the bench's own window size (for roofline.traffic) + compute-sanitizer pass
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:elb_blocks -s 2 -c 3 \
    -f -o gpurun_out/prof_blocks_kernels_4g \
    python bench.py --steps 1 --warmup 1 --window-gib 4 --skip-e2e --skip-cpu \
    > gpurun_out/prof_full_4g.log 2>&1
echo "full capture rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "not large" \
    > gpurun_out/sanitizer_memcheck_kernels.log 2>&1
echo "memcheck kernels rc=$?"; tail -5 gpurun_out/sanitizer_memcheck_kernels.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 \
    python -m pytest tests/test_worker_gpu.py -x -q -k "file_mode_offset_variants or dir_mode_full_cycle" \
    > gpurun_out/sanitizer_memcheck_worker.log 2>&1
echo "memcheck worker rc=$?"; tail -5 gpurun_out/sanitizer_memcheck_worker.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "ragged" \
    > gpurun_out/sanitizer_racecheck_kernels.log 2>&1
echo "racecheck kernels rc=$?"; tail -5 gpurun_out/sanitizer_racecheck_kernels.log
