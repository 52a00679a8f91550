set -x
O=gpurun_out/r2b; mkdir -p $O
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/k4 makisu_b200/csrc/k4_microbench.cu && timeout 300 /tmp/k4 > $O/k4_microbench.txt 2>&1
cat $O/k4_microbench.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_exchange.py -q -m gpu > $O/gputest_split.txt 2>&1; echo "gputest split rc=$?"; tail -5 $O/gputest_split.txt
MKSNAP_SCAN_CFG=11 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "chunk_table_random" > $O/sanitizer_cfg11.txt 2>&1; echo "sanitizer rc=$?"; grep -v "^$" $O/sanitizer_cfg11.txt | head -40
