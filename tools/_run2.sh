set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "distributed" > gpurun_out/r2_pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dist.log
tail -5 gpurun_out/r2_pytest_dist.log
