#!/bin/bash
# key-tile width of the long-suffix kernel: 64 (default) vs 128 (-DPC_PP_BN=128), Falcon MQA (D=64) and Llama (D=128) points
SH="71,1,64,256,2048 71,1,64,21,256 71,1,64,1,2048 71,1,64,16,2048 32,32,128,1024,8192 32,32,128,256,2048"
run() { timeout 300 python tools/attn_point.py $SH 2>/dev/null | grep "^{" | python -c "
import sys, json
print('  '.join('H%d/q%d/N%d %.2f' % (d['H'], d['q'], d['N'], d['us']) for d in map(json.loads, sys.stdin)))"; }
echo "BN=64"; run
export PC_NVCC_EXTRA="-DPC_PP_BN=128"
echo "BN=128"; run
