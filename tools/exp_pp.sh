#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_reference_gpu.py tests/test_models_gpu.py tests/test_properties_gpu.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python tools/attn_point.py 32,32,128,1024,8192 32,32,128,1024,2048 32,32,128,2048,0 32,32,128,256,2048 32,32,128,64,2048 71,1,64,256,2048 71,1,64,21,256 2>/dev/null | grep "^{"
