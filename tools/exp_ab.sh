#!/bin/bash
# A/B of a compile-time switch on the same box: $1 = nvcc define(s) for variant B
timeout 600 python tools/bench_short.py A1
export PC_NVCC_EXTRA="$1"
timeout 900 python tools/bench_short.py B1
timeout 600 python tools/bench_short.py B2
timeout 300 python tools/chain_bench.py 16 2>&1 | grep "ops=all\|ops=gu" | sed 's/^/B /'
unset PC_NVCC_EXTRA
timeout 900 python tools/bench_short.py A2
timeout 300 python tools/chain_bench.py 16 2>&1 | grep "ops=all\|ops=gu" | sed 's/^/A /'
