set -x
O=gpurun_out/r2b; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lw makisu_b200/csrc/lone_warp_microbench.cu && /tmp/lw partial > $O/lone_warp_partial.txt 2>&1
cat $O/lone_warp_partial.txt
timeout 900 python -m pytest tests -q -m gpu > $O/gputest_1gpu.txt 2>&1; echo "gputest rc=$?"; tail -3 $O/gputest_1gpu.txt
MKSNAP_CRC_OVERLAP=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -q -m gpu > $O/gputest_overlap.txt 2>&1; echo "overlap parity rc=$?"; tail -3 $O/gputest_overlap.txt
B="--steps 5 --warmup 3 --no-e2e --no-cpu --no-deliverables --fs-files 0 --no-strong"
for c in 0 9 10 11 12 13; do
  MKSNAP_SCAN_CFG=$c timeout 300 python bench.py --files 24000 $B > $O/scancfg2_$c.json 2> $O/scancfg2_$c.err; echo "cfg $c rc=$?"
done
for ov in 0 1; do
  MKSNAP_CRC_OVERLAP=$ov timeout 600 python bench.py $B > $O/overlap_$ov.json 2> $O/overlap_$ov.err; echo "overlap $ov rc=$?"
done
# ncu: launch list, full set of the scan kernel, C3-size traffic of the scan kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --files 8000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-deliverables --fs-files 0 --no-strong > $O/launches_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_roll_scan -s 2 -c 1 -f -o $O/roll_scan_full python bench.py --files 8000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-deliverables --fs-files 0 --no-strong > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_roll_scan|k_crc32_extents' -s 4 -c 2 --csv --log-file $O/c3_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-deliverables --fs-files 0 --no-strong > $O/c3_traffic_bench.log 2>&1; echo "ncu traffic rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b/scancfg2_*.json'))+sorted(glob.glob('gpurun_out/r2b/overlap_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j['ms_per_step'],3), round(j['value'],1), round(j['roofline']['frac'],4), [(k['name'][:14],round(k['ms'],3)) for k in j['kernels']][:6])
    except Exception as e:
        print(f, 'ERR', e)
PY
