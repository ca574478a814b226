"""GPU parity of the three kernels (through the C ABI) against the oracle on the same seeded inputs."""
import math

import pytest
import torch

from oracle import pc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


# ------------------------------------------------------------------ splice
def _make_segments(L, Hkv, D, lens, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(L, 2, Hkv, n, D, generator=g).half().to(DEV) for n in lens]


@pytest.mark.parametrize("L,Hkv,D,lens,max_ctx", [
    (2, 4, 64, [3, 1, 4, 2], 16),                 # tiny, ragged (golden-sized)
    (3, 1, 64, [5, 130, 1], 200),                 # Falcon MQA rows of 128 B, > 1 tile
    (4, 8, 128, [24, 400, 400, 7, 129], 1100),    # partial tiles
    (32, 32, 128, [24, 400, 400, 400, 400, 400, 24], 2304),  # C2 full size
])
def test_splice_bit_exact(native, L, Hkv, D, lens, max_ctx):
    segs = _make_segments(L, Hkv, D, lens, seed=2000 + L)
    cache = torch.zeros(L, 2, Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    rows, plan = 0, []
    for s in segs:
        plan.append((s, rows))
        rows += s.shape[3]
    native.splice_kv(plan, cache)
    torch.cuda.synchronize()
    ref, n = O.splice([[(s[l, 0], s[l, 1]) for l in range(L)] for s in segs], [0] * len(segs), max_ctx)
    assert n == rows
    for l in range(L):
        assert torch.equal(cache[l, 0, :, :n], ref[l][0])
        assert torch.equal(cache[l, 1, :, :n], ref[l][1])
    assert int(cache[:, :, :, n:].abs().max()) == 0  # nothing written past the spliced rows


def test_splice_usage_order_and_offsets(native):
    L, Hkv, D = 2, 2, 64
    segs = _make_segments(L, Hkv, D, [4, 2, 3], seed=7)
    usage = [1, 5, 1]
    order = O.splice_order(usage)
    assert order == [1, 0, 2]
    cache = torch.zeros(L, 2, Hkv, 32, D, dtype=torch.half, device=DEV)
    rows, plan = 0, []
    for i in order:
        plan.append((segs[i], rows))
        rows += segs[i].shape[3]
    native.splice_kv(plan, cache)
    ref, n = O.splice([[(s[l, 0], s[l, 1]) for l in range(L)] for s in segs], usage, 32)
    for l in range(L):
        assert torch.equal(cache[l, 0, :, :n], ref[l][0]) and torch.equal(cache[l, 1, :, :n], ref[l][1])


def test_splice_reference_layout(native):
    """pc_splice_kv_tensors: one descriptor per `tgt[:, st:ed, :].copy_(src)` of the reference."""
    L, Hkv, D, max_ctx = 3, 4, 128, 300
    g = torch.Generator(device="cpu").manual_seed(11)
    lens = [17, 130, 64]
    segs = [[(torch.randn(Hkv, n, D, generator=g).half().to(DEV), torch.randn(Hkv, n, D, generator=g).half().to(DEV))
             for _ in range(L)] for n in lens]
    tgt = [(torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV), torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV))
           for _ in range(L)]
    copies, st = [], 0
    for m, n in enumerate(lens):
        for l in range(L):
            copies.append((segs[m][l][0], tgt[l][0][:, st:st + n, :]))
            copies.append((segs[m][l][1], tgt[l][1][:, st:st + n, :]))
        st += n
    native.splice_kv_tensors(copies)
    torch.cuda.synchronize()
    ref, n = O.splice(segs, [0, 0, 0], max_ctx)
    for l in range(L):
        assert torch.equal(tgt[l][0][:, :n], ref[l][0]) and torch.equal(tgt[l][1][:, :n], ref[l][1])


def test_splice_empty_and_errors(native):
    cache = torch.zeros(1, 2, 1, 8, 64, dtype=torch.half, device=DEV)
    native.splice_kv([], cache)  # no-op
    seg = torch.zeros(1, 2, 1, 9, 64, dtype=torch.half, device=DEV)
    with pytest.raises(native.NativeError):
        native.splice_kv([(seg, 0)], cache)  # does not fit


# ------------------------------------------------------------------ rope + append
@pytest.mark.parametrize("H,Hkv,D,q,N,max_ctx", [(4, 4, 64, 5, 10, 32), (32, 32, 128, 16, 2048, 2304),
                                                  (71, 1, 64, 21, 256, 512), (8, 2, 128, 1, 77, 128)])
def test_rope_append_bit_exact(native, H, Hkv, D, q, N, max_ctx):
    g = torch.Generator(device="cpu").manual_seed(100 + H)
    qkv = (torch.randn(q, (H + 2 * Hkv) * D, generator=g) * 1.3).half().to(DEV)
    # discontiguous, non-monotone ids (parameter slots + trailing text)
    pos = torch.randperm(4000, generator=g)[:q].to(DEV)
    cos, sin = O.rope_tables(D, 4096, 10000.0, DEV, torch.float16)
    cos, sin = cos.contiguous(), sin.contiguous()
    kc = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    vc = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    q_out = torch.empty(H, q, D, dtype=torch.half, device=DEV)
    native.rope_append(qkv[:, :H * D], qkv[:, H * D:(H + Hkv) * D], qkv[:, (H + Hkv) * D:], pos, cos, sin, q_out,
                       kc, vc, N, H, Hkv, D)
    torch.cuda.synchronize()
    qs = qkv[:, :H * D].view(1, q, H, D).transpose(1, 2)
    ks = qkv[:, H * D:(H + Hkv) * D].view(1, q, Hkv, D).transpose(1, 2)
    vs = qkv[:, (H + Hkv) * D:].view(1, q, Hkv, D).transpose(1, 2)
    q_ref, k_ref = O.apply_rotary_pos_emb(qs, ks, cos, sin, pos.view(1, q))
    assert torch.equal(q_out, q_ref[0])
    assert torch.equal(kc[:, N:N + q], k_ref[0])
    assert torch.equal(vc[:, N:N + q], vs[0])
    assert int(kc[:, :N].abs().max()) == 0 and int(kc[:, N + q:].abs().max()) == 0


def test_rope_append_no_rotary(native):
    H = Hkv = 4
    D, q, N = 64, 3, 2
    qkv = torch.randn(q, 3 * H * D, device=DEV).half()
    kc = torch.zeros(Hkv, 16, D, dtype=torch.half, device=DEV)
    vc = torch.zeros_like(kc)
    q_out = torch.empty(H, q, D, dtype=torch.half, device=DEV)
    native.rope_append(qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:], None, None, None, q_out, kc, vc, N,
                       H, Hkv, D)
    assert torch.equal(q_out, qkv[:, :H * D].view(q, H, D).transpose(0, 1))
    assert torch.equal(kc[:, N:N + q], qkv[:, H * D:2 * H * D].view(q, H, D).transpose(0, 1))


# ------------------------------------------------------------------ attention
def _attn_case(H, Hkv, D, q, N, seed, std=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    max_ctx = ((N + q + 127) // 128) * 128 + 64
    Q = (torch.randn(H, q, D, generator=g) * std).half().to(DEV)
    K = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    V = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    K[:, :N + q] = (torch.randn(Hkv, N + q, D, generator=g) * std).half().to(DEV)
    V[:, :N + q] = torch.randn(Hkv, N + q, D, generator=g).half().to(DEV)
    return Q, K, V


def _oracle_llama(Q, K, V, N, q):
    """Reference eager attention (llama2.py:368-398) on already-rotated Q / K."""
    H, _, D = Q.shape
    k = K[:, :N + q].unsqueeze(0)
    v = V[:, :N + q].unsqueeze(0)
    n_rep = H // k.shape[1]
    kr, vr = O.repeat_kv(k, n_rep), O.repeat_kv(v, n_rep)
    attn = torch.matmul(Q.unsqueeze(0), kr.transpose(2, 3)) / math.sqrt(D)
    attn = attn + O.llama_causal_mask(q, N, Q.dtype, Q.device)
    attn = torch.softmax(attn, dim=-1, dtype=torch.float32).to(Q.dtype)
    out = torch.matmul(attn, vr)
    return out.transpose(1, 2).reshape(q, H * D)


def _oracle_fp32(Q, K, V, N, q):
    H, _, D = Q.shape
    k = O.repeat_kv(K[:, :N + q].unsqueeze(0).float(), H // K.shape[0])
    v = O.repeat_kv(V[:, :N + q].unsqueeze(0).float(), H // K.shape[0])
    s = torch.matmul(Q.unsqueeze(0).float(), k.transpose(2, 3)) / math.sqrt(D)
    s = s.masked_fill(O.bool_causal_mask(q, N, Q.device), float("-inf"))
    out = torch.matmul(torch.softmax(s, -1), v)
    return out.transpose(1, 2).reshape(q, H * D)


CASES = [
    (4, 4, 64, 5, 10),        # golden-sized
    (4, 4, 128, 16, 300),     # 3 tiles, suffix crosses a tile boundary
    (32, 32, 128, 16, 2048),  # C2
    (8, 8, 128, 1, 1000),     # decode
    (8, 8, 128, 200, 0),      # no-cache prefill, 2 row blocks, causal only
    (71, 1, 64, 21, 256),     # Falcon MQA fold (71*21 rows)
    (8, 2, 128, 40, 500),     # GQA fold
    (8, 2, 128, 4, 700),      # GQA fold on the short-suffix kernel (G*q = 16)
    (4, 4, 128, 32, 1000),    # short-suffix kernel, 32 columns
    (4, 4, 128, 17, 129),     # short-suffix kernel, padded columns, suffix straddles a tile boundary
    (4, 4, 128, 5, 0),        # short-suffix kernel without any cached rows
    (32, 32, 128, 32, 8192),  # C5
    (8, 8, 128, 300, 100),    # ping-pong kernel: 3 row tiles (pair + half pair), suffix-dominated causal part
    (4, 4, 128, 64, 1000),    # ping-pong kernel with a single half-filled tile and split-KV
    (2, 2, 128, 1024, 2048),  # long suffix, 4 tile pairs per head
    (4, 1, 128, 100, 3000),   # MQA at D=128: 400 folded rows
    (2, 2, 64, 130, 70),      # D=64, two tiles, tile B has 2 rows
    (2, 2, 128, 256, 64),     # one tile pair per head, unequal KV extents (regression: o_full phase aliasing)
    (2, 2, 128, 256, 192),
    (40, 40, 128, 32, 4096),  # C3 (Llama-2-13B heads)
]


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("H,Hkv,D,q,N", CASES)
def test_attention_llama_rounding(native, impl, H, Hkv, D, q, N):
    Q, K, V = _attn_case(H, Hkv, D, q, N, seed=H * 1000 + q, std=1.28)
    out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
    ws_bytes = native.attn_workspace_bytes(H, q, D)
    ws = torch.zeros(max(ws_bytes, 1), dtype=torch.uint8, device=DEV)
    native.attention(Q, K, V, out, N, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws,
                     impl=native.PC_ATTN_SIMT if impl == "simt" else native.PC_ATTN_TCGEN05, decode=(q == 1))
    torch.cuda.synchronize()
    ref = _oracle_llama(Q, K, V, N, q)
    # north_star tolerance: 1e-3 fp16-relative against the reference's own (rounding) attention path
    assert _rel(out, ref) <= 1e-3, _rel(out, ref)
    assert torch.allclose(out.float(), ref.float(), rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("H,Hkv,D,q,N", CASES)
def test_attention_exact_mode(native, impl, H, Hkv, D, q, N):
    Q, K, V = _attn_case(H, Hkv, D, q, N, seed=H * 77 + q)
    out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
    ws = torch.zeros(max(native.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=DEV)
    native.attention(Q, K, V, out, N, q, H, Hkv, D, 1.0 / math.sqrt(D), 0, workspace=ws,
                     impl=native.PC_ATTN_SIMT if impl == "simt" else native.PC_ATTN_TCGEN05, decode=(q == 1))
    torch.cuda.synchronize()
    ref = _oracle_fp32(Q, K, V, N, q)
    assert _rel(out, ref) <= 2e-3, _rel(out, ref)  # fp16 P and fp16 output rounding only


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_attention_alibi(native, impl):
    H = Hkv = 8
    D, q, N = 128, 9, 300
    Q, K, V = _attn_case(H, Hkv, D, q, N, seed=5)
    g = torch.Generator(device="cpu").manual_seed(6)
    key_pos = torch.cat([torch.randperm(400, generator=g)[:N], torch.arange(400, 400 + q)]).to(torch.int32).to(DEV)
    slopes = O.mpt_alibi_slopes(H, 8, DEV)
    scale = 1.0 / math.sqrt(D)
    out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
    ws = torch.zeros(max(native.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=DEV)
    native.attention(Q, K, V, out, N, q, H, Hkv, D, scale, native.PC_ROUND_MPT, workspace=ws, alibi_slopes=slopes,
                     key_pos=key_pos, alibi_max_pos=int(key_pos.max()),
                     impl=native.PC_ATTN_SIMT if impl == "simt" else native.PC_ATTN_TCGEN05)
    torch.cuda.synchronize()
    ctx, _, _ = O.mpt_attention_core(Q.unsqueeze(0), K[:, N:N + q].unsqueeze(0), V[:, N:N + q].unsqueeze(0),
                                     K[:, :N].unsqueeze(0), V[:, :N].unsqueeze(0), key_pos.long().view(1, -1), scale)
    assert _rel(out, ctx[0]) <= 1e-3, _rel(out, ctx[0])


def test_attention_workspace_is_left_clean_and_reusable(native):
    """Split-KV arrival counters are self-cleaning: the same workspace serves alternating shapes/kernels."""
    H = Hkv = 8
    D = 128
    ws = torch.zeros(native.attn_workspace_bytes(H, 64, D), dtype=torch.uint8, device=DEV)
    for q, N in [(16, 900), (64, 900), (1, 901), (16, 300), (1, 5000)]:
        Q, K, V = _attn_case(H, Hkv, D, q, N, seed=q + N, std=1.28)
        out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
        for _ in range(2):
            native.attention(Q, K, V, out, N, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws,
                             decode=(q == 1))
            assert _rel(out, _oracle_llama(Q, K, V, N, q)) <= 1e-3
        assert int(ws[:4096].view(torch.int32).abs().max()) == 0


def test_splice_many_segments_chunked_launches(native):
    """More segments than one launch's inline table holds (48): the API chunks the launches; ragged lengths 1..9."""
    L, Hkv, D = 2, 2, 64
    g = torch.Generator(device="cpu").manual_seed(3)
    lens = [1 + (i * 7) % 9 for i in range(131)]
    segs = [torch.randn(L, 2, Hkv, n, D, generator=g).half().to(DEV) for n in lens]
    max_ctx = sum(lens) + 5
    cache = torch.zeros(L, 2, Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    plan, row = [], 0
    for s in segs:
        plan.append((s, row))
        row += s.shape[3]
    native.splice_kv(plan, cache)
    ref, n = O.splice([[(s[l, 0], s[l, 1]) for l in range(L)] for s in segs], [0] * len(segs), max_ctx)
    for l in range(L):
        assert torch.equal(cache[l, 0, :, :n], ref[l][0]) and torch.equal(cache[l, 1, :, :n], ref[l][1])
    assert int(cache[:, :, :, n:].abs().max()) == 0


def test_rope_append_runtime_row(native):
    """pc_rope_append_dyn: destination row read from the device equals the host-row variant."""
    H = Hkv = 4
    D, q, N, max_ctx = 128, 3, 41, 64
    g = torch.Generator(device="cpu").manual_seed(8)
    qkv = torch.randn(q, 3 * H * D, generator=g).half().to(DEV)
    pos = torch.tensor([5, 900, 17], device=DEV)
    cos, sin = O.rope_tables(D, 1024, 10000.0, DEV, torch.float16)
    cos, sin = cos.contiguous(), sin.contiguous()
    outs = []
    for row in (N, torch.tensor([N], dtype=torch.int32, device=DEV)):
        kc = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
        vc = torch.zeros_like(kc)
        qo = torch.empty(H, q, D, dtype=torch.half, device=DEV)
        native.rope_append(qkv[:, :H * D], qkv[:, H * D:2 * H * D], qkv[:, 2 * H * D:], pos, cos, sin, qo, kc, vc, row, H, Hkv, D)
        outs.append((qo, kc, vc))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_attention_runtime_length_equals_host_length(native):
    """n_cached_dev (graph-replayable decode): same bits as the host-length launch, including surplus splits."""
    H = Hkv = 8
    D, max_ctx = 128, 2048
    ws = torch.zeros(native.attn_workspace_bytes(H, 16, D), dtype=torch.uint8, device=DEV)
    for q, N in [(1, 1900), (1, 130), (1, 5), (16, 700), (9, 1)]:
        Q, K, V = _attn_case(H, Hkv, D, q, N, seed=N)
        K2 = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
        V2 = torch.zeros_like(K2)
        K2[:, :N + q], V2[:, :N + q] = K[:, :N + q], V[:, :N + q]
        a = torch.empty(q, H * D, dtype=torch.half, device=DEV)
        b = torch.empty_like(a)
        native.attention(Q, K2, V2, a, N, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws)
        n_dev = torch.tensor([N], dtype=torch.int32, device=DEV)
        native.attention(Q, K2, V2, b, max_ctx - q, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws,
                         n_cached_dev=n_dev, impl=native.PC_ATTN_TCGEN05)
        assert _rel(b, _oracle_llama(Q, K, V, N, q)) <= 1e-3
        assert float((a.float() - b.float()).abs().max()) <= 2e-3  # split partition may differ (sized for the upper bound)
    with pytest.raises(native.NativeError):  # the CUDA-core cross-check kernel has no run-time-length path
        Q, K, V = _attn_case(H, Hkv, D, 200, 100, seed=1)
        native.attention(Q, K, V, torch.empty(200, H * D, dtype=torch.half, device=DEV), 100, 200, H, Hkv, D, 0.1, 3, workspace=ws,
                         n_cached_dev=torch.tensor([100], dtype=torch.int32, device=DEV), impl=native.PC_ATTN_SIMT)


@pytest.mark.parametrize("H,Hkv,q,N", [
    (32, 32, 16, 2048),   # C2: 4 splits per KV head -> clusters of 4
    (32, 32, 1, 2064),    # decode
    (32, 32, 32, 8192),   # C5, 32 columns
    (40, 40, 32, 4096),   # C3: 3 splits
    (20, 20, 7, 1500),    # 7 splits
    (64, 32, 4, 900),     # GQA fold, 4 splits
    (24, 24, 9, 300),     # fewer tiles (3) than SMs / heads (6): 3 splits
])
def test_attention_cluster_merge_equals_global_merge(native, monkeypatch, H, Hkv, q, N):
    """Split-KV merge through the cluster leader's shared memory (DSMEM) vs through global partials + arrival counter:
    same partials, same summation order -> the same bits; both within the north-star tolerance of the reference."""
    D = 128
    Q, K, V = _attn_case(H, Hkv, D, q, N, seed=H + q + N, std=1.28)
    ws = torch.zeros(max(native.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=DEV)
    outs = []
    for flag in (0, native.PC_ATTN_GLOBAL_MERGE, 0):
        out = torch.full((q, H * D), float("nan"), dtype=torch.half, device=DEV)
        native.attention(Q, K, V, out, N, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws,
                         impl=native.PC_ATTN_TCGEN05 | flag, decode=(q == 1))
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert _rel(outs[0], _oracle_llama(Q, K, V, N, q)) <= 1e-3
    assert int(ws[:4096].view(torch.int32).abs().max()) == 0


def test_attention_cluster_merge_with_runtime_length_and_surplus_ranks(native, monkeypatch):
    """Run-time N (decode graph): the cluster is sized for the upper bound, ranks without tiles contribute an empty
    partial instead of leaving early."""
    H = Hkv = 32
    D, max_ctx = 128, 1024  # upper bound: 8 tiles -> 4 splits per head
    ws = torch.zeros(native.attn_workspace_bytes(H, 16, D), dtype=torch.uint8, device=DEV)
    for q, N in [(1, 1000), (1, 130), (1, 5), (16, 300), (9, 1)]:
        Q, K, V = _attn_case(H, Hkv, D, q, N, seed=N + 3)
        K2 = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
        V2 = torch.zeros_like(K2)
        K2[:, :N + q], V2[:, :N + q] = K[:, :N + q], V[:, :N + q]
        n_dev = torch.tensor([N], dtype=torch.int32, device=DEV)
        res = {}
        for flag in ("1", "0"):
            out = torch.full((q, H * D), float("nan"), dtype=torch.half, device=DEV)
            native.attention(Q, K2, V2, out, max_ctx - q, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA,
                             workspace=ws, n_cached_dev=n_dev,
                             impl=native.PC_ATTN_TCGEN05 | (native.PC_ATTN_GLOBAL_MERGE if flag == "0" else 0))
            torch.cuda.synchronize()
            res[flag] = out
        assert torch.equal(res["1"], res["0"])
        assert _rel(res["1"], _oracle_llama(Q, K, V, N, q)) <= 1e-3


# ------------------------------------------------------------------ fused residual + LayerNorm (Falcon / MPT)
@pytest.mark.parametrize("rows,hidden,two,bias", [(21, 4544, True, True), (16, 4096, False, False), (1, 4544, True, True),
                                                  (300, 512, False, True), (7, 4096, True, False)])
def test_add_layernorm_matches_eager_reference_arithmetic(native, rows, hidden, two, bias):
    """x + (mlp += attn) then F.layer_norm, as falcon.py:781,796-810 / mpt.py:246-268 evaluate them in fp16."""
    g = torch.Generator(device="cpu").manual_seed(rows + hidden)
    x = (torch.randn(rows, hidden, generator=g) * 2 + 0.3).half().to(DEV)
    d1 = torch.randn(rows, hidden, generator=g).half().to(DEV)
    d2 = torch.randn(rows, hidden, generator=g).half().to(DEV) if two else None
    w = (1 + 0.1 * torch.randn(hidden, generator=g)).half().to(DEV)
    b = (0.1 * torch.randn(hidden, generator=g)).half().to(DEV) if bias else None
    t = d1.clone()
    if two:
        t += d2                      # mlp_output += attention_output
    want_x = x + t                   # residual + out
    want_h = torch.nn.functional.layer_norm(want_x, (hidden,), w, b, 1e-5)
    got_x, got_h = x.clone(), torch.empty_like(x)
    native.add_layernorm(got_x, d1, d2, w, b, got_h, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(got_x, want_x)
    assert _rel(got_h, want_h) <= 1e-3
    # no deltas: a plain LayerNorm
    h2 = torch.empty_like(x)
    native.add_layernorm(got_x, None, None, w, b, h2, 1e-5)
    assert torch.equal(h2, got_h)
