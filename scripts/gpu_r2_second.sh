# round 2, call 2: GPU test suite with the new staging engines, then the pipeline sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_a.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu_a.log
timeout 900 python scripts/explore_r2_pipeline.py /dev/shm 16 > gpurun_out/r02_pipeline_sweep.jsonl 2> gpurun_out/r02_pipeline_sweep.err
echo "sweep rc=$?"; tail -3 gpurun_out/r02_pipeline_sweep.jsonl; tail -5 gpurun_out/r02_pipeline_sweep.err
