#!/bin/bash
timeout 900 python -m pytest tests/test_virtual_splice_gpu.py tests/test_models_gpu.py tests/test_engine_gpu.py tests/test_reference_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -5
python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, '.')
from bench import family_block
dev = torch.device('cuda', 0)
b = family_block("mpt-7b", "C2", dev, 10, 3)
print("mpt C2", {k: round(b[k], 4) for k in ("ttft_ms", "e2e_ttft_ms", "decode_ms_per_token")})
PY
