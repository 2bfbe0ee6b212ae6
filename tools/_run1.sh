set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_n1.err
timeout 300 python tools/chain_profile.py 16384 > gpurun_out/r2_chain.log 2>&1
timeout 300 python tools/chain_profile.py 8192 >> gpurun_out/r2_chain.log 2>&1
cat gpurun_out/r2_chain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diag_block|k_panel' --launch-skip 40 --launch-count 4 -o gpurun_out/r2_chain_ncu -f python tools/chain_profile.py 8192 --quick > gpurun_out/r2_chain_ncu.log 2>&1; echo "ncu rc=$?"
