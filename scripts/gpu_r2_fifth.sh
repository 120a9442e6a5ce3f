# round 2, call 5 (1 GPU): gate tuning sweep, first-read exploration, staged ncu capture, GPU suite with durations
mkdir -p gpurun_out
timeout 600 python scripts/explore_r2_gate.py /dev/shm 16 > gpurun_out/r02_gate_sweep.jsonl 2> gpurun_out/r02_gate_sweep.err
cat gpurun_out/r02_gate_sweep.jsonl | cut -c1-400; tail -3 gpurun_out/r02_gate_sweep.err
timeout 120 scripts/explore_hostpath.bin firstread 16 > gpurun_out/r02_hostpath_firstread.jsonl 2>&1; cat gpurun_out/r02_hostpath_firstread.jsonl
timeout 1500 python -m pytest tests -q -m gpu --durations=25 > gpurun_out/r02_pytest_gpu_d.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r02_pytest_gpu_d.log
bash scripts/gpu_r2_ncu.sh
