# round 2, call 2: GPU test suite with the new staging engines, then the pipeline sweep and bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_a.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu_a.log
timeout 900 python scripts/explore_r2_pipeline.py /dev/shm 16 > gpurun_out/r02_pipeline_sweep.jsonl 2> gpurun_out/r02_pipeline_sweep.err
echo "sweep rc=$?"; tail -3 gpurun_out/r02_pipeline_sweep.jsonl; tail -5 gpurun_out/r02_pipeline_sweep.err
( time timeout 600 python bench.py > gpurun_out/r02_bench_c2_a.json 2> gpurun_out/r02_bench_c2_a.err ) 2> gpurun_out/r02_bench_c2_a.time
echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench_c2_a.json; tail -5 gpurun_out/r02_bench_c2_a.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_c2_ref_a.json 2> gpurun_out/r02_bench_c2_ref_a.err ) 2> gpurun_out/r02_bench_c2_ref_a.time
echo "bench ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_c2_ref_a.json
