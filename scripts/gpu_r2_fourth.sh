# round 2, call 4 (1 GPU): full GPU suite, configs c3-c5 both arms, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_c.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu_c.log
bash scripts/gpu_r2_configs.sh
bash scripts/gpu_r2_ncu.sh
