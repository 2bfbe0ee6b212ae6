set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_diag_block|k_panel_gemm' --launch-skip 40 --launch-count 2 -o gpurun_out/r2_chain_ncu2 -f python tools/chain_profile.py 8192 --quick > gpurun_out/r2_chain_ncu2.log 2>&1; echo "ncu rc=$?"
