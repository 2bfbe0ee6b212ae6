set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:'k_diag_block' --launch-skip 20 --launch-count 1 -o gpurun_out/r2_diag4 -f python tools/chain_profile.py 8192 --quick > gpurun_out/r2_diag4.log 2>&1; echo "ncu rc=$?"
