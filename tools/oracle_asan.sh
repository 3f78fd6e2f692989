#!/bin/bash
# The CPU oracle (test infrastructure) under AddressSanitizer + UndefinedBehaviorSanitizer: rebuilds oracle/libfmx_oracle.so with the
# sanitizers, runs the oracle-centred CPU tests with the runtimes preloaded into python, restores the normal build.
cd "$(dirname "$0")/.."
gcc -O1 -g -std=c11 -fPIC -fopenmp -mpopcnt -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer \
    -shared -o oracle/libfmx_oracle.so oracle/fmx_oracle.c || exit 1
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
    python -m pytest tests/test_oracle_golden.py tests/test_host_builder.py tests/test_host_properties.py -x -q
rc=$?
make -C oracle -B libfmx_oracle.so > /dev/null
exit $rc
