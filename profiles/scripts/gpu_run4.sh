set -x
O=gpurun_out/r2b; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lw makisu_b200/csrc/lone_warp_microbench.cu && /tmp/lw split > $O/lone_warp_split.txt 2>&1
cat $O/lone_warp_split.txt
(nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|MHz|L3"; free -g | head -2; df -h /dev/shm | tail -1) > $O/box.txt 2>&1; cat $O/box.txt
for t in 32 64 128; do
  MKHOST_TRACE=1 timeout 600 python bench.py --fs-only --fs-threads $t > $O/fs_t$t.json 2> $O/fs_t$t.err; echo "fs $t rc=$?"
  python -c "
import json;j=json.loads(open('$O/fs_t$t.json').read().strip().splitlines()[-1]);print({k:(round(v,2) if isinstance(v,float) else v) for k,v in j.items() if 'GiBps' in k or k in ('host_threads','create_s')})"
  grep -i "trace\|phase\|walk\|ms" $O/fs_t$t.err | tail -6
done
