set -x
O=gpurun_out/r2b; mkdir -p $O
S=$(date +%s); MKSNAP_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench8 rc=$? wall=$(( $(date +%s) - S ))s"
grep "mksnap exchange" $O/bench_n8.err | tail -3
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2b/bench_n8.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','n_gpus')}, j['e2e']['value'])
for k,v in j['strong']['runs'].items(): print(k, {a:v[a] for a in ('value','ms_per_step','imbalance_max_over_mean','dedup_ratio_unique_over_total')}, v['rank0_ms'])
print([ (k['name'][:20],round(k['ms'],3)) for k in j['kernels']])
PY
