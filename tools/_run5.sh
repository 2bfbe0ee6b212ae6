set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu or distributed" > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_2gpu.log
tail -3 gpurun_out/r2_pytest_2gpu.log
