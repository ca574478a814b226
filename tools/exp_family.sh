#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -12
python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
from bench import family_block
dev = torch.device('cuda', 0)
for mk, cfg in (("falcon-7b", "C4"), ("mpt-7b", "C2")):
    b = family_block(mk, cfg, dev, 10, 3)
    print(mk, {k: b[k] for k in ("ttft_ms", "e2e_ttft_ms", "decode_ms_per_token")}, b["roofline_step"]["frac"])
PY
