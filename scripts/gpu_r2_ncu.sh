# round 2: ncu evidence. Full-set captures of K1/K2/K3 (resident) at 4 KiB, 64 KiB and 1 MiB
# blocks, of the staged forms inside a worker run, and the launch list of a bench command.
mkdir -p gpurun_out
for spec in c3:4:4 c5:64:4 c2:1024:4; do
	cfg=${spec%%:*}; rest=${spec#*:}; kib=${rest%%:*}; win=${rest#*:}
	ncu --set full --clock-control none --import-source on -k regex:elb_blocks -s 9 -c 3 \
		-f -o gpurun_out/r02_ncu_resident_${kib}k \
		python bench.py --only-kernels --config $cfg --window-gib $win \
		> gpurun_out/r02_ncu_resident_${kib}k.log 2>&1
	echo "capture $kib KiB rc=$?"
	ncu -i gpurun_out/r02_ncu_resident_${kib}k.ncu-rep --page raw --csv \
		> gpurun_out/r02_ncu_resident_${kib}k_raw.csv 2>/dev/null
done
# staged kernels inside a small worker run (2 threads; one capture of each staged kernel)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
	-k "regex:elb_blocks_tiled_kernel<[01], 2>" -s 8 -c 200 \
	-f -o gpurun_out/r02_ncu_staged_1m \
	python bench.py --steps 2 --warmup 1 --file-gib 1 --threads 2 --skip-cpu --skip-kernels \
	> gpurun_out/r02_ncu_staged_1m.log 2>&1
echo "staged capture rc=$?"
ncu -i gpurun_out/r02_ncu_staged_1m.ncu-rep --page raw --csv > gpurun_out/r02_ncu_staged_1m_raw.csv 2>/dev/null
# launch list of a (small) full bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
	--log-file gpurun_out/r02_launches_bench.csv \
	python bench.py --steps 3 --warmup 1 --file-gib 2 --window-gib 1 --threads 4 --skip-cpu \
	> gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/ | grep r02_ncu
