set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gmw or se_family" > gpurun_out/r2_pytest_gmw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gmw.log
tail -30 gpurun_out/r2_pytest_gmw.log
