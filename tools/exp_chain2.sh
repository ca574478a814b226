#!/bin/bash
for f in 0 2 1 8 4 16 20; do
PC_CHAIN_FLAGS=$f timeout 600 python tools/bench_short.py "flags=$f"
done
PC_LINEAR_CHAIN=0 timeout 600 python tools/bench_short.py nochain
