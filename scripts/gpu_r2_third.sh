# round 2, call 3: GPU suite (both staging engines), JIT-gated write sweep, bench both arms
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_b.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_gpu_b.log
timeout 900 python scripts/explore_r2_pipeline2.py /dev/shm 16 > gpurun_out/r02_pipeline_sweep2.jsonl 2> gpurun_out/r02_pipeline_sweep2.err
echo "sweep rc=$?"; cat gpurun_out/r02_pipeline_sweep2.jsonl; tail -5 gpurun_out/r02_pipeline_sweep2.err
( time timeout 600 python bench.py > gpurun_out/r02_bench_c2_b.json 2> gpurun_out/r02_bench_c2_b.err ) 2> gpurun_out/r02_bench_c2_b.time
echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_c2_b.json; tail -5 gpurun_out/r02_bench_c2_b.err
( time timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_c2_ref_b.json 2> gpurun_out/r02_bench_c2_ref_b.err ) 2> gpurun_out/r02_bench_c2_ref_b.time
echo "bench ref rc=$?"; cut -c1-200 gpurun_out/r02_bench_c2_ref_b.json
timeout 300 scripts/explore_hostpath.bin dmasrc 8 > gpurun_out/r02_hostpath_dmasrc.jsonl 2> gpurun_out/r02_hostpath_dmasrc.err
cat gpurun_out/r02_hostpath_dmasrc.jsonl
