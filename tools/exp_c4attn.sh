#!/bin/bash
export PC_NVCC_EXTRA=-DPC_DEBUG
SH="71,1,64,16,2048 71,1,64,1,2048 71,1,64,1,600 71,1,64,256,2048 32,32,128,64,2048 32,32,128,256,2048 32,32,128,300,8192"
for s in 1 2 4 6 8 12 16 default; do
echo "splits=$s"
if [ $s = default ]; then unset PC_ATTN_SPLITS; else export PC_ATTN_SPLITS=$s; fi
timeout 300 python tools/attn_point.py $SH 2>/dev/null | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   H%d q=%d N=%d: %.2f us' % (d['H'], d['q'], d['N'], d['us']))"
done
