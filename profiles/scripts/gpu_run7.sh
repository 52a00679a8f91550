set -x
O=gpurun_out/r2b; mkdir -p $O
nvidia-smi topo -m 2>/dev/null | head -8 > $O/topo.txt; cat $O/topo.txt
for pin in 1 0; do for t in 16 24 32; do
  MKHOST_NUMA_PIN=$pin MKHOST_TRACE=1 timeout 600 python bench.py --fs-only --fs-threads $t > $O/fs3_p${pin}_t$t.json 2> $O/fs3_p${pin}_t$t.err; echo "pin $pin t $t rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/fs3_*.json')):
    j=json.loads(open(f).read().strip().splitlines()[-1]); print(f,{k:round(v,2) for k,v in j.items() if 'GiBps' in k})
PY
