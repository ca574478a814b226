#!/bin/bash
O=gpurun_out
mkdir -p $O
ncu --clock-control none --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $O/launches_chain.csv python bench.py --profile --no-extras > /dev/null 2>&1
python tools/launch_summary.py $O/launches_chain.csv
PC_PDL=0 timeout 600 python tools/bench_short.py chain_nopdl
PC_PDL=0 PC_LINEAR_CHAIN=0 timeout 600 python tools/bench_short.py nochain_nopdl
PC_CHAIN_FLAGS=16 timeout 600 python tools/bench_short.py flags16
timeout 300 python tools/chain_bench.py 1
