#!/bin/bash
python -m pytest tests/test_virtual_splice_gpu.py tests/test_engine_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -3
python tools/bench_short.py now
python tools/bench_short.py now_again
