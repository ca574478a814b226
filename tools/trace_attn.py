"""Debug: %globaltimer stamps of CTA 0 of the short-suffix attention kernel (PC_ATTN_TRACE_PTR)."""
# NOTE: needs a debug build of the library: PC_NVCC_EXTRA=-DPC_DEBUG python tools/trace_attn.py  (product builds compile the env / trace hooks out)
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat
DEV = "cuda"
H = Hkv = 32; D = 128
for q, N in [(16, 2048), (1, 2064), (32, 8192)]:
    max_ctx = ((N + q + 127) // 128) * 128
    K = torch.randn(Hkv, max_ctx, D, device=DEV).half(); V = torch.randn(Hkv, max_ctx, D, device=DEV).half()
    Q = torch.randn(H, q, D, device=DEV).half(); out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
    ws = torch.zeros(nat.attn_workspace_bytes(H, q, D), dtype=torch.uint8, device=DEV)
    tr = torch.zeros(64 + 2 * 1024, dtype=torch.int64, device=DEV)
    os.environ["PC_ATTN_TRACE_PTR"] = hex(tr.data_ptr())
    for _ in range(3):
        tr.zero_()
        nat.attention(Q, K, V, out, N, q, H, Hkv, D, nat.softmax_scale(D), nat.PC_ROUND_LLAMA, workspace=ws, decode=(q == 1))
        torch.cuda.synchronize()
    t = tr.cpu().tolist(); t0 = t[0]
    names = {0: "start", 1: "setup done", 5: "Q landed", 6: "S0 issued", 2: "last PV landed", 3: "partials published", 4: "done"}
    ev = [(v - t0, names.get(i, f"tile{(i-8)//4} " + ["S landed", "max exchanged", "P published"][(i - 8) % 4] if i >= 8 else str(i))) for i, v in enumerate(t[:40]) if v]
    print(f"--- q={q} N={N}")
    import numpy as np
    per = np.array(t[64:64 + 2 * 1024]).reshape(-1, 2)
    per = per[per[:, 0] > 0]
    st, en = per[:, 0] - per[:, 0].min(), per[:, 1] - per[:, 0].min()
    print(f"CTAs {len(per)}: start spread {st.max()/1000:.2f} us; end min/median/max {en.min()/1000:.2f}/{np.median(en)/1000:.2f}/{en.max()/1000:.2f} us; "
          f"duration min/median/max {(en-st).min()/1000:.2f}/{np.median(en-st)/1000:.2f}/{(en-st).max()/1000:.2f} us")
    for dt, nm in sorted(ev):
        print(f"{dt/1000:8.2f} us  {nm}")
