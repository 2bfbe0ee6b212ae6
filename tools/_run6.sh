set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu" > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_2gpu.log
tail -3 gpurun_out/r2_pytest_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config'].get('frac_of_fp64_dmma_peak'), d['e2e'], d['parity']['ok'])
PY
