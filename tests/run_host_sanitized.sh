#!/bin/bash
# Build libmkhost with AddressSanitizer + UndefinedBehaviorSanitizer and run the host-side CPU tests under it
# (tar reader on damaged archives, MemFS sequences, Copier, untar).  Restores the normal library afterwards.
#   tests/run_host_sanitized.sh            -> prints every sanitizer report line, then the pytest summary
set -u
cd "$(dirname "$0")/.."
LIB=makisu_b200/lib/libmkhost.so
cp "$LIB" "$LIB.keep"
trap 'mv -f "$LIB.keep" "$LIB"' EXIT
g++ -O1 -g -std=c++17 -fPIC -shared -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -I include \
    -o "$LIB" makisu_b200/host/mkhost.cpp -Lmakisu_b200/lib -lmksnap -Wl,-rpath,'$ORIGIN' || exit 1
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0
python -m pytest tests/test_host_cpu.py tests/test_host_tar_ingest_cpu.py tests/test_host_copier_cpu.py \
    tests/test_host_fuzz_cpu.py tests/test_host_mock_engine_cpu.py -q 2>&1 | grep -E "runtime error|AddressSanitizer|SUMMARY|passed|failed"
