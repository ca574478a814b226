"""One decoder layer's linears + norms (o_proj -> norm -> gate/up -> down -> norm -> QKV), Llama-2-7B shapes, over LAYERS distinct
weight sets (> L2), as six launches per layer vs ONE pc_linear_chain launch per layer; CUDA-graph replayed, CUDA-event timed.
   python tools/chain_bench.py [q_len]"""
import contextlib, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat

dev = "cuda"
q = int(sys.argv[1]) if len(sys.argv) > 1 else 16
hid, inter, H, D, LAYERS = 4096, 11008, 32, 128, 8
g = torch.Generator(device="cpu").manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).half().to(dev)
ws = nat.linear_workspace(dev)
layers = []
for l in range(LAYERS):
    wo, wgu, wd, wqkv = (torch.empty(hid, hid, dtype=torch.half, device=dev).normal_(0, 0.02), torch.empty(2 * inter, hid, dtype=torch.half, device=dev).normal_(0, 0.02),
                         torch.empty(hid, inter, dtype=torch.half, device=dev).normal_(0, 0.02), torch.empty(3 * hid, hid, dtype=torch.half, device=dev).normal_(0, 0.02))
    layers.append(dict(wo=wo, wgu=wgu, wd=wd, wqkv=wqkv, t_o=nat.tile_weights(wo, nat.PC_LIN_PARTIALS), t_gu=nat.tile_weights(wgu, nat.PC_LIN_SILU_MUL),
                       t_d=nat.tile_weights(wd, nat.PC_LIN_PARTIALS), t_qkv=nat.tile_weights(wqkv, nat.PC_LIN_ROPE_APPEND),
                       n1=torch.ones(hid, dtype=torch.half, device=dev), n2=torch.ones(hid, dtype=torch.half, device=dev)))
attn_out, x = rnd(q, hid), rnd(q, hid)
h = torch.zeros(q, hid, dtype=torch.half, device=dev)
act = torch.zeros(q, inter, dtype=torch.half, device=dev)
p_o = torch.zeros(nat.linear_partial_splits(hid, hid), q, hid, dtype=torch.float32, device=dev)
p_d = torch.zeros(nat.linear_partial_splits(hid, inter), q, hid, dtype=torch.float32, device=dev)
qo = torch.zeros(H, q, D, dtype=torch.half, device=dev)
kc, vc = torch.zeros(H, 64, D, dtype=torch.half, device=dev), torch.zeros(H, 64, D, dtype=torch.half, device=dev)
pos = torch.arange(q, device=dev)
from promptcache_b200.model._base import build_rope_tables
cos, sin = build_rope_tables(D, 2048, 10000.0, dev)
wbytes = 2 * (hid * hid + 2 * inter * hid + hid * inter + 3 * hid * hid)


def step(chained, only=None):
    for L in layers:
        with (nat.chain(ws) if chained else contextlib.nullcontext()):
            if only in (None, "o"):
                nat.linear_partials(attn_out, L["wo"], p_o, ws, w_tiled=L["t_o"])
            if only in (None, "norm"):
                nat.add_rmsnorm(x, p_o, L["n1"], h, 1e-5)
            if only in (None, "gu"):
                nat.linear_small(h, L["wgu"], act, ws, silu_mul=True, w_tiled=L["t_gu"])
            if only in (None, "d"):
                nat.linear_partials(act, L["wd"], p_d, ws, w_tiled=L["t_d"])
            if only in (None, "norm"):
                nat.add_rmsnorm(x, p_d, L["n2"], h, 1e-5)
            if only in (None, "qkv"):
                nat.linear_rope_append(h, L["wqkv"], pos, cos, sin, qo, kc, vc, 0, H, H, ws, w_tiled=L["t_qkv"])


def timed(chained, only=None, reps=20):
    step(chained, only)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step(chained, only)
    for _ in range(3):
        gr.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps / LAYERS


for only in (None, "o", "gu", "d", "qkv", "norm"):
    t0, t1 = timed(False, only), timed(True, only)
    extra = f"  {wbytes / t0 / 1e3:.0f} / {wbytes / t1 / 1e3:.0f} GB/s" if only is None else ""
    print(f"q={q} ops={only or 'all'}: six launches {t0:.1f} us/layer, chain {t1:.1f} us/layer{extra}  flags={os.environ.get('PC_CHAIN_FLAGS', '')}")
