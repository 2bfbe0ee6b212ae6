set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "permute_symmetric or known_answer or config2_size or decisions_match_oracle or one_shot_streamed or buffer_cache" > gpurun_out/r2_pytest_perm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_perm.log
tail -25 gpurun_out/r2_pytest_perm.log
timeout 600 python tools/permute_bench.py > gpurun_out/r2_permute.log 2>&1
cat gpurun_out/r2_permute.log
