set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dynamic" > gpurun_out/r2_pytest_dyn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_dyn.log
tail -5 gpurun_out/r2_pytest_dyn.log
timeout 600 python tools/dynamic_bench.py 1000 2000 4000 8000 16384 > gpurun_out/r2_dynamic.log 2>&1
cat gpurun_out/r2_dynamic.log
