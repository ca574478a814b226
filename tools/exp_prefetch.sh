#!/bin/bash
# A/B of the in-kernel L2 prefetch of the next linear's weight stream (PC_LINEAR_PREFETCH, PC_LIN_PREFETCH_UNITS)
python -m pytest tests/test_linear_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -3
PC_LINEAR_PREFETCH=1 python tools/bench_short.py prefetch_on
PC_LINEAR_PREFETCH=0 python tools/bench_short.py prefetch_off
PC_LINEAR_PREFETCH=1 python tools/bench_short.py prefetch_on_again
for u in 3 12; do PC_NVCC_EXTRA=-DPC_LIN_PREFETCH_UNITS=$u python tools/bench_short.py units_$u; done
