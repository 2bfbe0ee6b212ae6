set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "upper_triangle or known_answer or reference_test_grid_n10_static or families or sizes_around or batched or decompose_solve_multiply or permutations_and_return" > gpurun_out/r2_pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_quick.log
tail -5 gpurun_out/r2_pytest_quick.log
