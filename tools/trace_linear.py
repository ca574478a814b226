"""Debug: per-CTA %globaltimer stamps of pc_linear_small (PC_LIN_TRACE_PTR) -> where a launch spends its time."""
# NOTE: needs a debug build of the library: PC_NVCC_EXTRA=-DPC_DEBUG python tools/trace_linear.py  (product builds compile the env / trace hooks out)
import os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat
DEV = "cuda"
NAMES = ["start", "setup done", "first stage landed", "last MMA committed", "last accumulator ready", "partial published",
         "epilogue done", "end"]
ws = nat.linear_workspace(DEV)
for name, q, k, n, mode in [("o_proj (partials, regular split-K)", 16, 4096, 4096, "part"), ("qkv-sized plain (one whole tile per CTA)", 16, 4096, 12288, "plain"),
                            ("gate_up silu (hybrid)", 16, 4096, 22016, "silu"), ("down (partials)", 16, 11008, 4096, "part")]:
    silu = mode == "silu"
    epi = {"plain": nat.PC_LIN_PLAIN, "silu": nat.PC_LIN_SILU_MUL, "part": nat.PC_LIN_PARTIALS}[mode]
    W = (torch.randn(4, n, k, device=DEV) * 0.02).half()
    WT = [nat.tile_weights(W[i], epi) for i in range(4)]
    x = torch.randn(q, k, device=DEV).half()
    out = torch.empty(q, n // 2 if silu else n, dtype=torch.half, device=DEV)
    part = torch.empty(nat.linear_partial_splits(n, k), q, n, dtype=torch.float32, device=DEV)
    tr = torch.zeros(8 * 148, dtype=torch.int64, device=DEV)
    os.environ["PC_LIN_TRACE_PTR"] = hex(tr.data_ptr())
    for i in range(4):
        tr.zero_()
        if mode == "part":
            nat.linear_partials(x, W[i], part, ws, w_tiled=WT[i])
        else:
            nat.linear_small(x, W[i], out, ws, silu_mul=silu, w_tiled=WT[i])
        torch.cuda.synchronize()
    t = tr.cpu().numpy().reshape(148, 8).astype(np.float64)
    t0 = t[:, 0][t[:, 0] > 0].min()
    print(f"--- {name} q={q} [{n}x{k}] {n*k*2/1e6:.0f} MB (ideal {n*k*2/6.5e6:.1f} us)")
    for j, nm in enumerate(NAMES):
        col = t[:, j][t[:, j] > 0]
        if len(col):
            col = (col - t0) / 1000
            print(f"  {nm:24s} n={len(col):3d} min {col.min():6.2f}  median {np.median(col):6.2f}  max {col.max():6.2f} us")
    del os.environ["PC_LIN_TRACE_PTR"]
