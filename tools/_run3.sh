set -x
mkdir -p gpurun_out
timeout 300 python tools/chain_profile.py 8192 > gpurun_out/r2_chain3.log 2>&1
cat gpurun_out/r2_chain3.log
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -25 gpurun_out/r2_pytest.log
timeout 300 python tools/chain_profile.py 16384 >> gpurun_out/r2_chain3.log 2>&1
tail -1 gpurun_out/r2_chain3.log
