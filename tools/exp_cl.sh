#!/bin/bash
export PC_NVCC_EXTRA=-DPC_DEBUG
timeout 900 python -m pytest tests/test_linear_gpu.py -m gpu -q -x 2>&1 | tail -2
for s in 3 4; do
PC_LIN_CL_SIZE=$s timeout 600 python tools/bench_short.py "cl_size=$s"
PC_LIN_CL_SIZE=$s timeout 300 python tools/chain_bench.py 16 2>&1 | grep "ops=all\|ops=qkv" | cut -c1-60
done
PC_LIN_CL_SIZE=3 timeout 600 python tools/bench_short.py "cl_size=3 again"
