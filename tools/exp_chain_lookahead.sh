#!/bin/bash
export PC_NVCC_EXTRA=-DPC_DEBUG
for f in 32 1024 2048 4096 6144 8192; do
PC_CHAIN_FLAGS=$f timeout 300 python tools/chain_bench.py 16 2>&1 | grep "ops=all\|ops=gu"
done
PC_CHAIN_FLAGS=$((4096+16)) timeout 300 python tools/chain_bench.py 16 2>&1 | grep "ops=all\|ops=gu"
