set -x
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/gputest_1gpu_final.txt 2>&1; echo "gputest rc=$?"; tail -3 $O/gputest_1gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
S=$(date +%s); timeout 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s"
S=$(date +%s); timeout 600 python bench.py --impl reference > $O/bench_n1_reference.json 2> $O/bench_n1_reference.err; echo "ref rc=$? wall=$(( $(date +%s) - S ))s"
for t in 8 16 24 40 48; do
  timeout 600 python bench.py --fs-only --fs-threads $t > $O/fs2_t$t.json 2> $O/fs2_t$t.err; echo "fs $t rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/fs2_t*.json')):
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f,{k:round(v,2) for k,v in j.items() if 'GiBps' in k})
j=json.loads(open('gpurun_out/r2b/bench_n1.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','gpu_launches')}, j['e2e'], j['roofline']['frac'])
print(json.dumps(j.get('deliverables'))[:1500])
print(json.dumps(j.get('e2e_fs'))[:800])
print(json.dumps(j.get('tar_digest'))[:600])
PY
