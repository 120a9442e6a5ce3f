# round 2, very last call: full GPU suite on the final tree + the staged-kernel ncu capture
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu --durations=10 > gpurun_out/r02_pytest_gpu_final2.log 2>&1
echo "pytest rc=$?"; tail -18 gpurun_out/r02_pytest_gpu_final2.log
timeout 130 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
	-k "regex:elb_blocks_tiled_kernel<\(int\)[01], \(int\)2>" -s 8 -c 60 -f -o gpurun_out/r02_ncu_staged_1m \
	python bench.py --steps 2 --warmup 1 --file-gib 1 --threads 2 --skip-cpu --skip-kernels \
	> gpurun_out/r02_ncu_staged_1m.log 2>&1
echo "staged capture rc=$?"; tail -2 gpurun_out/r02_ncu_staged_1m.log | cut -c1-200
timeout 60 ncu -i gpurun_out/r02_ncu_staged_1m.ncu-rep --page raw --csv > gpurun_out/r02_ncu_staged_1m_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_staged_1m.ncu-rep
ls -la gpurun_out/r02_ncu_staged_1m_raw.csv
