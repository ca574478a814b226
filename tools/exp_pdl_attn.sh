#!/bin/bash
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_reference_gpu.py tests/test_models_gpu.py tests/test_properties_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -3
SH="32,32,128,1024,8192 32,32,128,256,2048 32,32,128,64,2048 71,1,64,256,2048 71,1,64,21,256 71,1,64,1,2048 71,1,64,16,2048 71,1,64,1,600 32,32,128,300,8192"
timeout 300 python tools/attn_point.py $SH 2>/dev/null | grep "^{" | python -c "
import sys, json
print('  '.join('H%d/q%d/N%d %.2f' % (d['H'], d['q'], d['N'], d['us']) for d in map(json.loads, sys.stdin)))"
python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, '.')
from bench import family_block
dev = torch.device('cuda', 0)
b = family_block("falcon-7b", "C4", dev, 10, 3)
print("falcon C4", {k: round(b[k], 4) for k in ("ttft_ms", "e2e_ttft_ms", "decode_ms_per_token")})
PY
