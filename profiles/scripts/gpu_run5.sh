set -x
O=gpurun_out/r2b; mkdir -p $O
for r in 0 1; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -DSS_ROLL_ROUNDS=$r -o /tmp/k4_$r makisu_b200/csrc/k4_microbench.cu && timeout 300 /tmp/k4_$r > $O/k4_microbench_roll$r.txt 2>&1
  grep -E "MODE|CTA0|streams      1 |streams     32|streams    256|streams   4736|streams   9472|debug" $O/k4_microbench_roll$r.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_host.py tests/test_gpu_exchange.py -q -m gpu > $O/gputest_split2.txt 2>&1; echo "gputest rc=$?"; tail -3 $O/gputest_split2.txt
