set -x
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_exchange.py -q -m gpu > $O/gputest_2gpus.txt 2>&1; echo "gputest2 rc=$?"; tail -3 $O/gputest_2gpus.txt
S=$(date +%s); MKSNAP_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc=$? wall=$(( $(date +%s) - S ))s"
grep "mksnap exchange" $O/bench_n2.err | tail -3
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2b/bench_n2.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','n_gpus')}, j['e2e']['value'])
print(json.dumps(j.get('strong'))[:1800])
print([ (k['name'][:20],round(k['ms'],3)) for k in j['kernels']])
PY
