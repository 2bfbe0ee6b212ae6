set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n8.err
MDB200_P2P=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_n8_p2p.json 2> gpurun_out/r2_bench_n8_p2p.err; echo "bench p2p rc=$?"
tail -2 gpurun_out/r2_bench_n8_p2p.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n8.json','gpurun_out/r2_bench_n8_p2p.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['config'].get('frac_of_fp64_dmma_peak'), d['e2e'].get('value'), d['config']['workload'][-60:])
    except Exception as e: print(f, 'ERR', e)
PY
