set -x
mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2b/parity.txt 2>&1; echo "parity rc=$?" 
tail -5 gpurun_out/r2b/parity.txt
for c in 0 1 2 3 4 5 6 7 8 9; do
  MKSNAP_SCAN_CFG=$c timeout 300 python bench.py --files 24000 --steps 5 --warmup 3 --no-e2e --no-cpu --no-deliverables --fs-files 0 --no-strong > gpurun_out/r2b/scancfg_$c.json 2> gpurun_out/r2b/scancfg_$c.err; echo "cfg $c rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/scancfg_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        k={x['name']:x for x in j['roofline_kernels']} if 'roofline_kernels' in j else None
        print(f, j['ms_per_step'], j['roofline']['frac'], j['roofline']['achieved'])
    except Exception as e:
        print(f, 'ERR', e)
PY
