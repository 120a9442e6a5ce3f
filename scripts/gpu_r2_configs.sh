# round 2: BASELINE configs[2..4] through bench.py at stated size, both arms, N=1
mkdir -p gpurun_out
for cfg in c3 c4 c5; do
	( time timeout 900 python bench.py --config $cfg > gpurun_out/r02_bench_${cfg}_n1.json 2> gpurun_out/r02_bench_${cfg}_n1.err ) 2> gpurun_out/r02_bench_${cfg}_n1.time
	echo "$cfg rc=$?"; cut -c1-260 gpurun_out/r02_bench_${cfg}_n1.json; tail -3 gpurun_out/r02_bench_${cfg}_n1.err
	( time timeout 900 python bench.py --config $cfg --impl reference > gpurun_out/r02_bench_${cfg}_n1_ref.json 2> gpurun_out/r02_bench_${cfg}_n1_ref.err ) 2> gpurun_out/r02_bench_${cfg}_n1_ref.time
	echo "$cfg ref rc=$?"; cut -c1-200 gpurun_out/r02_bench_${cfg}_n1_ref.json; tail -2 gpurun_out/r02_bench_${cfg}_n1_ref.err
done
# cuFile probe of the round (in case a box has nvidia-fs)
timeout 120 python -m pytest tests/test_cufile_gpu.py -q -m gpu -k "real" > gpurun_out/r02_cufile_probe.log 2>&1; tail -3 gpurun_out/r02_cufile_probe.log
