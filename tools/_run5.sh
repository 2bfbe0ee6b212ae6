set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu or distributed" > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_2gpu.log
tail -5 gpurun_out/r2_pytest_2gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n2.err
head -c 1500 gpurun_out/r2_bench_n2.json
