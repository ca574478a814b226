"""Kernel micro-benchmarks (CUDA events, inputs larger than L2).  Usage: python tools/microbench.py [splice|attn|linear|all]"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from promptcache_b200 import _native as nat  # noqa: E402

DEV = "cuda"
PEAK = 6538.3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit_graph(fn, reps=16, iters=10):
    """GPU-side time per call: capture `reps` back-to-back calls in a CUDA graph (no host launch overhead)."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / (iters * reps)


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def bench_splice():
    for name, L, Hkv, D, lens in [("C2", 32, 32, 128, [24] + [400] * 5 + [24]), ("C5", 32, 32, 128, [256] * 32),
                                  ("C3", 40, 40, 128, [512] * 8), ("C4-falcon", 32, 1, 64, [40, 60, 70, 50, 36])]:
        segs = [torch.randn(L, 2, Hkv, n, D, device=DEV).half() for n in lens]
        N = sum(lens)
        cache = torch.zeros(L, 2, Hkv, N + 256, D, dtype=torch.half, device=DEV)
        plan, r = [], 0
        for s in segs:
            plan.append((s, r))
            r += s.shape[3]
        ms = timeit(lambda: nat.splice_kv(plan, cache))
        byts = 2 * sum(s.numel() * 2 for s in segs)
        # torch baseline: the reference's copy loop on device-resident segments
        def ref():
            r0 = 0
            for s in segs:
                n = s.shape[3]
                for l in range(L):
                    cache[l, 0, :, r0:r0 + n].copy_(s[l, 0])
                    cache[l, 1, :, r0:r0 + n].copy_(s[l, 1])
                r0 += n
        ms_ref = timeit(ref, iters=5, warmup=1)
        big = torch.empty(byts // 4, dtype=torch.half, device=DEV)
        big2 = torch.empty_like(big)
        ms_cp = timeit(lambda: big2.copy_(big))
        print(f"splice {name}: N={N} {byts/2**30:.2f} GiB r+w  kernel {ms:.3f} ms = {byts/ms/1e6:.0f} GB/s "
              f"({byts/ms/1e6/PEAK:.2f} of measured {PEAK:.0f})  torch-loop {ms_ref:.3f} ms  torch.copy_ same bytes {ms_cp:.3f} ms",
              flush=True)
        del segs, cache, big, big2


def bench_attn(only=None):
    for name, H, Hkv, D, q, N in [c for c in [("C2 q16", 32, 32, 128, 16, 2048), ("C2 q64", 32, 32, 128, 64, 2048),
                                  ("C2 q256", 32, 32, 128, 256, 2048), ("C2 q1024", 32, 32, 128, 1024, 2048),
                                  ("full prefill 2K", 32, 32, 128, 2048, 0), ("decode N=2064", 32, 32, 128, 1, 2064),
                                  ("C5 q32", 32, 32, 128, 32, 8192), ("C5 q1024 N8K", 32, 32, 128, 1024, 8192), ("falcon C4", 71, 1, 64, 21, 256),
                                  ("falcon q256 N2K", 71, 1, 64, 256, 2048)] if only is None or only in c[0]]:
        L = 8  # rotate over several layers' KV so the stream does not sit in L2
        max_ctx = ((N + q + 127) // 128) * 128
        K = (torch.randn(L, Hkv, max_ctx, D, device=DEV)).half()
        V = torch.randn(L, Hkv, max_ctx, D, device=DEV).half()
        Q = torch.randn(H, q, D, device=DEV).half()
        out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
        ws = torch.zeros(max(nat.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=DEV)
        state = {"i": 0}
        for impl, nm in [(nat.PC_ATTN_TCGEN05, "tcgen05"), (nat.PC_ATTN_SIMT, "simt")]:
            if impl == nat.PC_ATTN_SIMT and q > 64:
                continue
            def run():
                l = state["i"] % L
                state["i"] += 1
                nat.attention(Q, K[l], V[l], out, N, q, H, Hkv, D, nat.softmax_scale(D), nat.PC_ROUND_LLAMA, workspace=ws,
                              impl=impl, decode=(q == 1))
            ms = timeit_graph(run, reps=16, iters=10)
            G = H // Hkv
            # causal-aware flop count: key j>=N visible to rows r>=j-N
            vis = q * N + q * (q + 1) // 2
            flops = 4.0 * H * vis * D
            byts = 4.0 * Hkv * (N + q) * D + 4.0 * H * q * D
            print(f"attn {name} [{nm}]: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.0f} GB/s "
                  f"({byts/ms/1e6/PEAK:.2f} of HBM peak)", flush=True)


def bench_linear(only=None):
    """pc_linear_small vs cuBLAS (F.linear) on the Llama-2 step's linears; weights rotate over > L2 worth of copies."""
    import torch.nn.functional as F
    ws = nat.linear_workspace(DEV)
    shapes = [("7B qkv+rope", 16, 4096, 12288, "rope"), ("7B o_proj", 16, 4096, 4096, "plain"),
              ("7B gate_up+silu", 16, 4096, 22016, "silu"), ("7B down", 16, 11008, 4096, "plain"),
              ("7B lm_head", 16, 4096, 32000, "plain"), ("7B qkv+rope q1", 1, 4096, 12288, "rope"),
              ("7B down q32", 32, 11008, 4096, "plain"), ("13B gate_up+silu q32", 32, 5120, 27648, "silu")]
    tot = {"native": 0.0, "cublas": 0.0}
    tiled = os.environ.get("PC_MB_TILED", "1") != "0"  # stream the pre-tiled copy of the weights (1-D bulk loads)
    print(f"pc_linear_small weights: {'pre-tiled (bulk loads)' if tiled else 'row-major (tensor-map loads)'}", flush=True)
    epi_code = {"plain": nat.PC_LIN_PLAIN, "silu": nat.PC_LIN_SILU_MUL, "rope": nat.PC_LIN_ROPE_APPEND}
    for name, q, k, n, epi in [c for c in shapes if only is None or only in c[0]]:
        rot = max(2, int(math.ceil(600e6 / (n * k * 2))))
        W = (torch.randn(rot, n, k, device=DEV) * 0.02).half()
        WT = [nat.tile_weights(W[i], epi_code[epi]) if tiled else None for i in range(rot)]
        x = torch.randn(q, k, device=DEV).half()
        byts = n * k * 2.0
        state = {"i": 0}
        if epi == "rope":
            H = Hkv = n // 384
            pos = torch.arange(2048, 2048 + q, device=DEV)
            inv = 1.0 / (10000.0 ** (torch.arange(0, 128, 2, device=DEV).float() / 128))
            fr = torch.outer(torch.arange(4096, device=DEV).float(), inv)
            emb = torch.cat([fr, fr], -1)
            cos, sin = emb.cos().half().contiguous(), emb.sin().half().contiguous()
            q_rot = torch.empty(H, q, 128, dtype=torch.half, device=DEV)
            kc = torch.zeros(Hkv, 2304, 128, dtype=torch.half, device=DEV)
            vc = torch.zeros(Hkv, 2304, 128, dtype=torch.half, device=DEV)

            def run_native():
                i = state["i"] = (state["i"] + 1) % rot
                nat.linear_rope_append(x, W[i], pos, cos, sin, q_rot, kc, vc, 2048, H, Hkv, ws, w_tiled=WT[i])

            def run_cublas():
                i = state["i"] = (state["i"] + 1) % rot
                qkv = F.linear(x, W[i])
                nat.rope_append(qkv[:, :H * 128], qkv[:, H * 128:(H + Hkv) * 128], qkv[:, (H + Hkv) * 128:], pos, cos, sin, q_rot,
                                kc, vc, 2048, H, Hkv, 128)
        elif epi == "silu":
            out = torch.empty(q, n // 2, dtype=torch.half, device=DEV)

            def run_native():
                i = state["i"] = (state["i"] + 1) % rot
                nat.linear_small(x, W[i], out, ws, silu_mul=True, w_tiled=WT[i])

            def run_cublas():
                i = state["i"] = (state["i"] + 1) % rot
                nat.silu_mul(F.linear(x, W[i]), out, n // 2)
        else:
            out = torch.empty(q, n, dtype=torch.half, device=DEV)

            def run_native():
                i = state["i"] = (state["i"] + 1) % rot
                nat.linear_small(x, W[i], out, ws, w_tiled=WT[i])

            def run_cublas():
                i = state["i"] = (state["i"] + 1) % rot
                F.linear(x, W[i])
        ms_n = timeit_graph(run_native, reps=16, iters=10)
        ms_c = timeit_graph(run_cublas, reps=16, iters=10)
        if q == 16 and name.startswith("7B") and "lm_head" not in name:
            tot["native"] += ms_n
            tot["cublas"] += ms_c
        print(f"linear {name} q={q} [{n}x{k}] {byts/1e6:.0f} MB: native {ms_n*1e3:.1f} us = {byts/ms_n/1e6:.0f} GB/s "
              f"({byts/ms_n/1e6/PEAK:.2f} of HBM peak) | cuBLAS(+epilogue kernel) {ms_c*1e3:.1f} us = {byts/ms_c/1e6:.0f} GB/s "
              f"({byts/ms_c/1e6/PEAK:.2f})", flush=True)
        del W, WT
    if tot["native"] > 0:
        print(f"linear 7B per layer (qkv+o+gate_up+down, q=16): native {tot['native']*1e3:.1f} us vs cuBLAS {tot['cublas']*1e3:.1f} us "
              f"-> x32 layers {tot['native']*32:.3f} ms vs {tot['cublas']*32:.3f} ms", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("device:", torch.cuda.get_device_name(0), flush=True)
    if what in ("splice", "all"):
        bench_splice()
    if what in ("attn", "all"):
        bench_attn(sys.argv[2] if len(sys.argv) > 2 else None)
    if what in ("linear", "all"):
        bench_linear(sys.argv[2] if len(sys.argv) > 2 and what == "linear" else None)
