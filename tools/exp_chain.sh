#!/bin/bash
timeout 600 python -m pytest tests/test_linear_gpu.py -m gpu -q -x -k "chain" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_models_gpu.py tests/test_engine_gpu.py tests/test_virtual_splice_gpu.py -m gpu -q -x 2>&1 | tail -5
PC_LINEAR_CHAIN=0 timeout 600 python tools/bench_short.py nochain
timeout 600 python tools/bench_short.py chain
PC_LINEAR_CHAIN=0 timeout 600 python tools/bench_short.py nochain_again
timeout 600 python tools/bench_short.py chain_again
