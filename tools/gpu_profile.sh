set -x
mkdir -p gpurun_out
# 1. launch list of the bench command (one timed step), per-launch durations
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launches_r02.log 2>&1; echo "ncu launches rc=$?"
# 2. full capture of the largest bulk launch of the n = 65536 factorization (first outer block, part b)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tn_tma' --launch-skip 24 --launch-count 12 -o gpurun_out/prof_r02_gemm -f python bench.py --steps 1 --warmup 1 --no-e2e > gpurun_out/ncu_full_r02_gemm.log 2>&1; echo "ncu gemm rc=$?"
# 3. chain kernels and the permutation passes at n = 16384
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diag_block|k_panel_gemm' --launch-skip 60 --launch-count 2 -o gpurun_out/prof_r02_chain -f python tools/chain_profile.py 16384 --quick > gpurun_out/ncu_full_r02_chain.log 2>&1; echo "ncu chain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gather_rows_transpose' --launch-count 2 -o gpurun_out/prof_r02_permute -f python tools/chain_profile.py 16384 --quick > gpurun_out/ncu_full_r02_permute.log 2>&1; echo "ncu permute rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
