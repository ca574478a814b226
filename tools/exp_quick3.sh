#!/bin/bash
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_models_gpu.py tests/test_engine_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/bench_short.py norm_prefetch
timeout 600 python tools/bench_short.py norm_prefetch_again
