#!/bin/bash
# after the acq_rel arrival: are the native linears now ahead for Falcon (decode, C4 prefill) and at 32 token columns (C5)?
for mq in 0 16 32; do
PC_LINEAR_MAX_Q_FALCON=$mq python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, '.')
from bench import family_block
dev = torch.device('cuda', 0)
b = family_block("falcon-7b", "C4", dev, 10, 3)
print("falcon max_q", os.environ["PC_LINEAR_MAX_Q_FALCON"], {k: round(b[k], 4) for k in ("ttft_ms", "e2e_ttft_ms", "decode_ms_per_token")})
PY
done
for mq in 16 32; do
PC_LINEAR_MAX_Q=$mq python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, '.')
from bench import build_lm, c5_block
dev = torch.device('cuda', 0)
lm, md = build_lm("llama2-7b", dev)
b = c5_block(lm, dev, 1, 0)
print("c5 linear_max_q", os.environ["PC_LINEAR_MAX_Q"], {k: round(b["virtual_splice"][k], 3) for k in ("batch_ms", "ttft_ms_mean", "prefill_ms_mean")})
PY
done
