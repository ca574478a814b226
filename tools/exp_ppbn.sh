#!/bin/bash
# A/B of the long-suffix attention kernel's key-tile width (PC_PP_BN = 128 default | 64)
P="32,32,128,1024,8192 32,32,128,256,2048 32,32,128,64,2048 32,32,128,1024,2048 32,32,128,2048,0 71,1,64,256,2048 71,1,64,21,256 40,40,128,512,4096"
echo "== BN=128"; python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py tests/test_properties_gpu.py tests/test_reference_gpu.py -m gpu -q -x -k "not c2_llama7b" 2>&1 | tail -4
python tools/attn_point.py $P 2>&1 | grep "^{"
echo "== BN=64"; PC_NVCC_EXTRA=-DPC_PP_BN=64 python tools/attn_point.py $P 2>&1 | grep "^{"
echo "== BN=128 NO_SOFTMAX"; PC_NVCC_EXTRA=-DPC_EXP_NO_SOFTMAX python tools/attn_point.py 32,32,128,1024,8192 32,32,128,2048,0 71,1,64,256,2048 2>&1 | grep "^{"
echo "== BN=128 NO_MUFU"; PC_NVCC_EXTRA=-DPC_EXP_NO_MUFU python tools/attn_point.py 32,32,128,1024,8192 32,32,128,2048,0 71,1,64,256,2048 2>&1 | grep "^{"
