#!/bin/bash
# Build libmkhost with ThreadSanitizer and run the multi-threaded host paths (reader pool, threaded context walk, batch
# committer) under it: the host CPU tests in-process, the mock-engine scenarios in their own processes with their stderr
# shown.  Restores the normal library afterwards.
#   tests/run_host_tsan.sh            -> one line per scenario with its number of ThreadSanitizer reports
set -u
cd "$(dirname "$0")/.."
LIB=makisu_b200/lib/libmkhost.so
cp "$LIB" "$LIB.keep"
trap 'mv -f "$LIB.keep" "$LIB"' EXIT
g++ -O1 -g -std=c++17 -fPIC -shared -pthread -fsanitize=thread -fno-omit-frame-pointer -I include \
    -o "$LIB" makisu_b200/host/mkhost.cpp -Lmakisu_b200/lib -lmksnap -Wl,-rpath,'$ORIGIN' || exit 1
T=$(mktemp -d)
gcc -O2 -fPIC -c oracle/mkoracle.c -o "$T/mkoracle.o" || exit 1
g++ -O1 -g -std=c++17 -fPIC -shared -Wall -I include tests/mock_engine/mock_mksnap.cpp "$T/mkoracle.o" -o "$T/libmock.so" || exit 1
export LD_PRELOAD="$(gcc -print-file-name=libtsan.so)"
export TSAN_OPTIONS="report_signal_unsafe=0 exitcode=0"
python -m pytest tests/test_host_cpu.py -q 2>&1 | grep -E "WARNING: ThreadSanitizer|passed|failed" | sort | uniq -c
for sc in cache_id_and_commit ingest_untar_and_file_digests content_aware_scan materialize_from_the_arena random_trees_small_arenas \
          table_limits_and_arena_leases batch_of_layers_in_one_session ingest_member_larger_than_the_arena incremental_cache_id; do
    rm -rf "$T/t"; mkdir -p "$T/t"
    python -m tests.mock_engine.run "$T/libmock.so" $sc "$T/t" > "$T/$sc.out" 2> "$T/$sc.err"
    echo "$sc: $(tail -1 "$T/$sc.out")  ThreadSanitizer reports: $(grep -c 'WARNING: ThreadSanitizer' "$T/$sc.err")"
done
rm -rf "$T"
