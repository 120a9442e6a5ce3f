# round 2, last call (1 GPU, <= 13 min): targeted GPU tests of what changed since the last full
# suite run, smoke, final bench of both arms, gate A/B, one staged-kernel ncu capture
mkdir -p gpurun_out
timeout 330 python -m pytest tests/test_worker_variants_gpu.py tests/test_worker_gpu.py tests/test_kernels_gpu.py \
	-q -m gpu --durations=12 -k "stonewall or aio_rate or nofdsharing or file_mode_seq or dir_mode_full_cycle or aio_and_direct or block_variance or plain_write or live_stats or small_block or staged or mid_size_multi_batch_run_bytes_equal_oracle[kernel]" \
	> gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02_pytest_gpu_final.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 200 python bench.py > gpurun_out/r02_bench_c2_n1_final.json 2> gpurun_out/r02_bench_c2_n1_final.err ) 2> gpurun_out/r02_bench_c2_n1_final.time
echo "bench rc=$?"; cut -c1-160 gpurun_out/r02_bench_c2_n1_final.json; tail -2 gpurun_out/r02_bench_c2_n1_final.err
( time timeout 150 python bench.py --impl reference > gpurun_out/r02_bench_c2_n1_final_ref.json 2> gpurun_out/r02_bench_c2_n1_final_ref.err ) 2> gpurun_out/r02_bench_c2_n1_final_ref.time
echo "ref rc=$?"; cut -c1-160 gpurun_out/r02_bench_c2_n1_final_ref.json
ELB_GATE_PREFETCH=0 timeout 120 python bench.py --skip-cpu --skip-kernels > gpurun_out/r02_bench_c2_n1_noprefetch.json 2> gpurun_out/r02_bench_c2_n1_noprefetch.err
echo "noprefetch rc=$?"; cut -c1-130 gpurun_out/r02_bench_c2_n1_noprefetch.json
ELB_GATE_NEAR=2 timeout 120 python bench.py --skip-cpu --skip-kernels > gpurun_out/r02_bench_c2_n1_near2.json 2> gpurun_out/r02_bench_c2_n1_near2.err
echo "near2 rc=$?"; cut -c1-130 gpurun_out/r02_bench_c2_n1_near2.json
timeout 150 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
	-k "regex:elb_blocks_tiled_kernel<[01], 2>" -s 8 -c 120 -f -o gpurun_out/r02_ncu_staged_1m \
	python bench.py --steps 2 --warmup 1 --file-gib 1 --threads 2 --skip-cpu --skip-kernels \
	> gpurun_out/r02_ncu_staged_1m.log 2>&1
echo "staged capture rc=$?"
timeout 60 ncu -i gpurun_out/r02_ncu_staged_1m.ncu-rep --page raw --csv > gpurun_out/r02_ncu_staged_1m_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_staged_1m.ncu-rep
ls -la gpurun_out/r02_ncu_staged_1m_raw.csv
