set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "known_answer or reference_test_grid_n10_static or families or sizes_around or explosive or batched_many or decompose_solve_multiply or outer_block or distributed_steps" > gpurun_out/r2_pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_quick.log
tail -25 gpurun_out/r2_pytest_quick.log
timeout 300 python tools/chain_profile.py 8192 > gpurun_out/r2_chain8.log 2>&1
timeout 300 python tools/chain_profile.py 16384 >> gpurun_out/r2_chain8.log 2>&1
cat gpurun_out/r2_chain8.log
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 1 -k regex:'k_panel_gemm' --launch-skip 20 --launch-count 1 -o gpurun_out/r2_panel5 -f python tools/chain_profile.py 8192 --quick > gpurun_out/r2_panel5.log 2>&1; echo "ncu rc=$?"
