#!/bin/bash
# Falcon on the weight-streaming linears: parity tests, then C4 (21-token suffix) and a 16-token variant, native vs cuBLAS
python -m pytest tests/test_linear_gpu.py tests/test_models_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -5
for mq in 16 32; do
PC_LINEAR_MAX_Q_FALCON=$mq python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, '.')
from bench import family_block
dev = torch.device('cuda', 0)
b = family_block("falcon-7b", "C4", dev, 10, 3)
print("max_q", os.environ["PC_LINEAR_MAX_Q_FALCON"], {k: b[k] for k in ("ttft_ms", "e2e_ttft_ms", "decode_ms_per_token")}, b["roofline_step"]["frac"])
PY
done
