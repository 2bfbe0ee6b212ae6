mkdir -p gpurun_out
: > gpurun_out/r2_agg_sweep.log
for T in "8192,16384,32768" "4096,8192,16384" "4096,8192,32768" "2048,4096,16384" "6144,12288,24576" "4096,12288,24576"; do
  echo "THRESH $T" >> gpurun_out/r2_agg_sweep.log
  MDB200_AGG_THRESH=$T timeout 300 python tools/chain_profile.py 16384 >> gpurun_out/r2_agg_sweep.log 2>&1
done
python - <<'PY'
import json,re
t=None
for l in open('gpurun_out/r2_agg_sweep.log'):
    if l.startswith('THRESH'): t=l.split()[1]
    elif l.startswith('{'):
        d=json.loads(l); print(t, 'approx %.2f ms  decompose %.2f ms  trailing %.2f'%(d['approximate']['ms'], d['decompose']['ms'], d['decompose']['phases_ms']['trailing']))
PY
