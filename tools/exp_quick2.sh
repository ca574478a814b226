#!/bin/bash
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_models_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/bench_short.py acqrel
timeout 600 python tools/bench_short.py acqrel_again
timeout 300 python tools/chain_bench.py 16 2>&1 | grep "ops=all\|ops=gu"
