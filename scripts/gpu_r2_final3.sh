# round 2: one ncu capture of the stage-in + verify kernel inside a small worker run
mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
	-k "regex:elb_blocks_tiled_kernel<\(int\)1, \(int\)2>" -s 4 -c 24 -f -o gpurun_out/r02_ncu_staged_verify \
	python bench.py --steps 1 --warmup 1 --file-gib 0.5 --threads 2 --skip-cpu --skip-kernels \
	> gpurun_out/r02_ncu_staged_verify.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r02_ncu_staged_verify.log | cut -c1-200
timeout 60 ncu -i gpurun_out/r02_ncu_staged_verify.ncu-rep --page raw --csv > gpurun_out/r02_ncu_staged_verify_raw.csv 2>/dev/null
rm -f gpurun_out/r02_ncu_staged_verify.ncu-rep
ls -la gpurun_out/r02_ncu_staged_verify_raw.csv
