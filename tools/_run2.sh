set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gmw" > gpurun_out/r2_pytest_gmw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gmw.log
tail -5 gpurun_out/r2_pytest_gmw.log
timeout 600 python tools/gmw_bench.py 1000 2000 4000 > gpurun_out/r2_gmw.log 2>&1
cat gpurun_out/r2_gmw.log
