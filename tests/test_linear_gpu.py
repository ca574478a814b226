"""GPU parity of pc_linear_small (stream-K tcgen05 skinny linear + fused epilogues) through the C ABI.

Reference semantics (promptcache/model/llama2.py): an fp16 nn.Linear = fp32-accumulated dot products rounded once to
fp16 (:345-351, :405, :241, :1050); SiLU(gate)*up in fp16 (:241); RoPE at explicit ids + KV concat (:195-210,357-364).
Integer-valued inputs make every fp32 partial sum exact, so those cases must match BIT FOR BIT whatever the
summation order; random inputs are checked against an fp64 product within one fp16 rounding.
"""
import pytest
import torch

from oracle import pc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ints(shape, lo, hi, scale, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randint(lo, hi + 1, shape, generator=g).float() * scale).half().to(DEV)


def _exact_linear(x, w):
    return (x.double() @ w.double().t()).half()


SHAPES = [  # (q, k, n)
    (16, 4096, 4096),     # o_proj 7B
    (1, 4096, 12288),     # decode, qkv-sized
    (32, 11008, 4096),    # down_proj, NPAD 32
    (7, 256, 512),        # tiny: fewer units than SMs
    (16, 4544, 4672),     # Falcon hidden (71 k-blocks), rows not a multiple of 128
    (16, 4096, 32000),    # lm_head
    (3, 64, 128),         # a single unit
    (17, 5120, 5120),     # 13B, NPAD 32 with ragged tokens
    (4, 256, 37888),      # 296 row tiles = 2 whole tiles per SM, empty stream-K remainder
    (9, 384, 19200),      # 150 row tiles: one whole tile per SM + a 2-tile remainder shared by all SMs
    (8, 256, 1000),       # ragged last row tile (CodeLlama-style vocabularies are not multiples of 128)
    (16, 4096, 32016),    # CodeLlama lm_head: 251 row tiles, the last one 16 rows
]


def _tiled(native, w, epi, tiled):
    """Pre-tiled copy of the weight (1-D bulk-load path) or None (tensor-map path)."""
    if not tiled:
        return None
    wt = native.tile_weights(w, epi)
    if wt is None:
        pytest.skip("k is not a multiple of 64: no pre-tiled layout")
    return wt


@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("q,k,n", SHAPES)
def test_linear_plain_exact_on_integer_inputs(native, q, k, n, tiled):
    x = _ints((q, k), -4, 4, 0.25, seed=q + k)
    w = _ints((n, k), -2, 2, 0.125, seed=n + 1)
    ws = native.linear_workspace(DEV)
    wt = _tiled(native, w, native.PC_LIN_PLAIN, tiled)
    out = torch.full((q, n), float("nan"), dtype=torch.half, device=DEV)
    native.linear_small(x, w, out, ws, w_tiled=wt)
    torch.cuda.synchronize()
    assert torch.equal(out, _exact_linear(x, w))
    assert int(ws[:16384].view(torch.int32).abs().max()) == 0  # arrival counters are left zeroed
    # replays on the same workspace give the same bits
    out2 = torch.empty_like(out)
    for _ in range(3):
        native.linear_small(x, w, out2, ws, w_tiled=wt)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


@pytest.mark.parametrize("q,k,n", [(16, 4096, 4096), (32, 11008, 4096), (5, 4096, 22016)])
def test_linear_plain_random_within_one_rounding(native, q, k, n):
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(q, k, generator=g).half().to(DEV)
    w = (torch.randn(n, k, generator=g) * 0.02).half().to(DEV)
    ws = native.linear_workspace(DEV)
    out = torch.empty(q, n, dtype=torch.half, device=DEV)
    native.linear_small(x, w, out, ws)
    ref = x.double() @ w.double().t()
    err = (out.double() - ref).abs()
    # fp32 accumulation error is far below the final fp16 rounding (2^-11 relative) at these sizes
    assert bool((err <= ref.abs() * 2.0 ** -10 + 1e-4).all()), float(err.max())
    cub = torch.nn.functional.linear(x, w)
    assert float((out.float() - cub.float()).abs().max()) <= 2e-3 * float(ref.abs().max())


@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("q,k,inter", [(16, 4096, 11008), (32, 512, 192), (1, 5120, 13824)])
def test_linear_silu_mul(native, q, k, inter, tiled):
    x = _ints((q, k), -4, 4, 0.25, seed=3)
    w = _ints((2 * inter, k), -2, 2, 1.0 / 64, seed=4)
    ws = native.linear_workspace(DEV)
    out = torch.empty(q, inter, dtype=torch.half, device=DEV)
    native.linear_small(x, w, out, ws, silu_mul=True, w_tiled=_tiled(native, w, native.PC_LIN_SILU_MUL, tiled))
    gu = _exact_linear(x, w)  # exact fp16 gate / up projections
    want = torch.empty_like(out)
    native.silu_mul(gu, want, inter)  # the unfused kernel on the same fp16 inputs
    ref = (torch.nn.functional.silu(gu[:, :inter].float()).half() * gu[:, inter:])
    torch.cuda.synchronize()
    assert torch.allclose(out.float(), want.float(), rtol=1e-3, atol=1e-6)
    assert torch.allclose(out.float(), ref.float(), rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("q,H,Hkv,k,row0,rotary,D", [
    (16, 32, 32, 4096, 2048, True, 128),   # C2
    (1, 32, 32, 4096, 777, True, 128),     # decode
    (32, 8, 2, 512, 5, True, 128),         # GQA, NPAD 32
    (9, 4, 4, 256, 0, False, 128),         # MPT: no rotation, plain append
    (12, 32, 32, 4096, 100, False, 128),   # MPT-7B: 96 head tiles -> owner + helper clusters (pre-tiled), no rotation
    (5, 36, 36, 512, 9, True, 128),        # 108 head tiles -> 36 clusters of 4 = 144 CTAs
    (16, 71, 1, 4544, 256, True, 64),      # Falcon-7B: head_dim 64 (two heads per tile, last tile half padding), 71 k-blocks
    (1, 71, 1, 4544, 1999, True, 64),      # Falcon decode
    (21, 6, 2, 320, 7, True, 64),          # D=64 GQA, 5 k-blocks, NPAD 32
    (5, 3, 1, 192, 0, False, 64),          # D=64 without rotation, 5 heads = 2.5 tiles
])
@pytest.mark.parametrize("tiled", [False, True])
def test_linear_rope_append_bit_exact_vs_unfused(native, q, H, Hkv, k, row0, rotary, D, tiled):
    max_ctx = row0 + q + 3
    x = _ints((q, k), -4, 4, 0.25, seed=5)
    w = _ints(((H + 2 * Hkv) * D, k), -2, 2, 1.0 / 32, seed=6)
    g = torch.Generator(device="cpu").manual_seed(7)
    pos = torch.randint(0, 4096, (q,), generator=g).to(DEV)  # discontiguous, non-monotone ids
    cos, sin = O.rope_tables(D, 4096, dtype=torch.float16)
    cos, sin = cos.contiguous().to(DEV), sin.contiguous().to(DEV)
    ws = native.linear_workspace(DEV)
    wt = _tiled(native, w, native.PC_LIN_ROPE_APPEND, tiled)

    def fresh():
        return (torch.zeros(H, q, D, dtype=torch.half, device=DEV), torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV),
                torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV))

    q1, k1, v1 = fresh()
    native.linear_rope_append(x, w, pos if rotary else None, cos if rotary else None, sin if rotary else None, q1, k1, v1,
                              row0, H, Hkv, ws, w_tiled=wt, head_dim=D)
    qkv = _exact_linear(x, w)
    q2, k2, v2 = fresh()
    native.rope_append(qkv[:, :H * D], qkv[:, H * D:(H + Hkv) * D], qkv[:, (H + Hkv) * D:], pos if rotary else None,
                       cos if rotary else None, sin if rotary else None, q2, k2, v2, row0, H, Hkv, D)
    torch.cuda.synchronize()
    assert torch.equal(q1, q2) and torch.equal(k1, k2) and torch.equal(v1, v2)
    # run-time destination row (graph-replayable decode)
    q3, k3, v3 = fresh()
    row_dev = torch.tensor([row0], dtype=torch.int32, device=DEV)
    native.linear_rope_append(x, w, pos if rotary else None, cos if rotary else None, sin if rotary else None, q3, k3, v3,
                              row_dev, H, Hkv, ws, w_tiled=wt, head_dim=D)
    torch.cuda.synchronize()
    assert torch.equal(q3, q2) and torch.equal(k3, k2) and torch.equal(v3, v2)


@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("q,k,n", [(16, 4096, 4096), (16, 11008, 4096), (1, 4096, 4096), (32, 13824, 5120), (5, 512, 256)])
def test_linear_partials_and_consumer_side_reduction(native, q, k, n, tiled):
    """PC_LIN_PARTIALS + pc_add_rmsnorm_partials == PC_LIN_PLAIN + pc_add_rmsnorm, bit for bit on integer-valued inputs:
    the split-K partials are summed by the norm kernel instead of by the last CTA of the GEMM."""
    x = _ints((q, k), -4, 4, 0.25, seed=q + 2 * k)
    w = _ints((n, k), -2, 2, 0.125, seed=n + 7)
    ws = native.linear_workspace(DEV)
    S = native.linear_partial_splits(n, k)
    assert S >= 2
    part = torch.full((S, q, n), float("nan"), dtype=torch.float32, device=DEV)
    assert native.linear_partials(x, w, part, ws, w_tiled=_tiled(native, w, native.PC_LIN_PARTIALS, tiled)) == S
    torch.cuda.synchronize()
    assert torch.equal(part.sum(0).half(), _exact_linear(x, w))
    g = torch.Generator(device="cpu").manual_seed(5)
    resid = torch.randn(q, n, generator=g).half().to(DEV)
    wn = torch.randn(n, generator=g).half().to(DEV)
    xa, xb = resid.clone(), resid.clone()
    ha, hb = torch.empty_like(xa), torch.empty_like(xb)
    native.add_rmsnorm(xa, part, wn, ha, 1e-5)
    native.add_rmsnorm(xb, _exact_linear(x, w), wn, hb, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(xa, xb) and torch.equal(ha, hb)


def test_linear_rejects_bad_arguments(native):
    ws = native.linear_workspace(DEV)
    x = torch.zeros(33, 128, dtype=torch.half, device=DEV)
    w = torch.zeros(128, 128, dtype=torch.half, device=DEV)
    with pytest.raises(native.NativeError):
        native.linear_small(x, w, torch.zeros(33, 128, dtype=torch.half, device=DEV), ws)  # q > 32
    with pytest.raises(native.NativeError):
        native.linear_small(x[:4, :96].contiguous(), w[:, :96].contiguous(), torch.zeros(4, 128, dtype=torch.half, device=DEV), ws)
    with pytest.raises(native.NativeError):
        native.linear_small(x[:4], w, torch.zeros(4, 128, dtype=torch.half, device=DEV), ws[:1024])  # workspace too small


def test_model_native_linears_match_cublas_linears(native):
    """The same 2-layer Llama-2-7B-shaped step with pc_linear_small and with cuBLAS linears: logits agree to fp16
    round-off, the appended K/V rows agree, and a replayed CUDA graph of the native step equals its eager run."""
    from promptcache_b200 import CacheEngine, Prompt
    from promptcache_b200.model import Llama2, LlamaConfigLite, LlamaForCausalLM
    from promptcache_b200.synthetic import IdTokenizer, make_case
    sd = O.random_llama_state_dict(4000, 4096, 11008, 2, 32, 32, seed=0, dtype=torch.float16, device=DEV)
    model = LlamaForCausalLM(LlamaConfigLite(vocab_size=4000, num_hidden_layers=2), device=DEV)
    model.load_hf_state_dict(sd)
    lm = Llama2("llama7b-2l", model=model, tokenizer=IdTokenizer(4000))
    case = make_case("C2", vocab=4000)
    eng = CacheEngine(2304, lm)
    eng.add_schema(case.schema_xml)
    res = {}
    for impl in ("native", "cublas", "native", "native"):
        model.linear_impl = impl
        before = native.launch_count()
        ids, pos, _, cache = eng.process(Prompt(case.prompt_xml))
        out = lm(input_ids=torch.tensor([ids], device=DEV), position_ids=torch.tensor([pos], device=DEV),
                 past_key_values=[(k.unsqueeze(0), v.unsqueeze(0)) for k, v in cache], use_cache=True)
        torch.cuda.synchronize()
        rows = eng.prompt_cache.buffer[:, :, :, 2048:2064].clone()
        res.setdefault(impl, []).append((out.logits.clone(), rows, native.launch_count() - before))
    a, b = res["native"][0], res["cublas"][0]
    scale = float(b[0].abs().max())
    assert float((a[0] - b[0]).abs().max()) < 3e-3 * scale
    assert float((a[1].float() - b[1].float()).abs().max()) < 3e-3 * float(b[1].abs().max())
    assert a[2] >= 2 * 7 and b[2] >= 2 * 5  # native: 7 launches of this library per layer; cuBLAS path: 5 (+4 library GEMMs)
    for other in res["native"][1:]:  # eager, then captured and replayed
        assert torch.equal(other[0], a[0]) and torch.equal(other[1], a[1])
    # the same step with everything between two attention launches chained into one persistent launch (experiment switch)
    model.linear_impl, model.linear_chain = "native", True
    before = native.launch_count()
    ids, pos, _, cache = eng.process(Prompt(case.prompt_xml))
    out = lm(input_ids=torch.tensor([ids], device=DEV), position_ids=torch.tensor([pos], device=DEV),
             past_key_values=[(k.unsqueeze(0), v.unsqueeze(0)) for k, v in cache], use_cache=True)
    torch.cuda.synchronize()
    assert native.launch_count() - before <= 2 * 2 + 2  # per layer: one chain + one attention
    assert float((out.logits - a[0]).abs().max()) < 2e-3 * scale


@pytest.mark.parametrize("q,k,n", [(16, 4096, 16384), (1, 512, 2048), (9, 4096, 640)])
def test_gelu_epilogue_equals_gelu_of_the_fp16_projection(native, q, k, n):
    """PLAIN + PC_LIN_ACT_GELU: nn.GELU() of the fp16-rounded projection (mpt.py:200-208), on integer-valued inputs so the
    projection itself is exact."""
    g = torch.Generator(device="cpu").manual_seed(q + k + n)
    x = torch.randint(-2, 3, (q, k), generator=g).half().to(DEV)
    w = (torch.randint(-2, 3, (n, k), generator=g).float() / 64).half().to(DEV)
    ws = native.linear_workspace(DEV)
    want = torch.nn.functional.gelu((x.double() @ w.double().t()).half())
    for tiled in (False, True):
        wt = native.tile_weights(w, native.PC_LIN_PLAIN) if tiled else None
        out = torch.empty(q, n, dtype=torch.half, device=DEV)
        native.linear_small(x, w, out, ws, gelu=True, w_tiled=wt)
        torch.cuda.synchronize()
        assert float((out.float() - want.float()).abs().max()) <= 2e-3 * float(want.float().abs().max())
        assert float((out != want).float().mean()) < 1e-3  # erf implementations agree to the last bit almost everywhere


def test_add_layernorm_partials_equal_linear_then_add_layernorm(native):
    g = torch.Generator(device="cpu").manual_seed(3)
    q, k, n = 12, 4096, 4096
    x_in = torch.randint(-2, 3, (q, k), generator=g).half().to(DEV)
    w = (torch.randint(-2, 3, (n, k), generator=g).float() / 64).half().to(DEV)
    res = torch.randn(q, n, generator=g).half().to(DEV)
    lw = (1 + 0.1 * torch.randn(n, generator=g)).half().to(DEV)
    ws = native.linear_workspace(DEV)
    part = torch.empty(native.linear_partial_splits(n, k), q, n, dtype=torch.float32, device=DEV)
    native.linear_partials(x_in, w, part, ws, w_tiled=native.tile_weights(w, native.PC_LIN_PARTIALS))
    a_x, a_h = res.clone(), torch.empty_like(res)
    native.add_layernorm(a_x, part, None, lw, None, a_h, 1e-5)
    out = torch.empty(q, n, dtype=torch.half, device=DEV)
    native.linear_small(x_in, w, out, ws)
    b_x, b_h = res.clone(), torch.empty_like(res)
    native.add_layernorm(b_x, out, None, lw, None, b_h, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(a_x, b_x) and torch.equal(a_h, b_h)


def test_add_layernorm_two_partial_stacks_equal_two_fp16_deltas(native):
    """Falcon's parallel block: x + (mlp += attn) with both deltas as split-K partial stacks == the same add with the
    fp16 outputs of the two linears (falcon.py:796-810), bit for bit on integer-valued inputs; K = 4544 (71 k-blocks)."""
    g = torch.Generator(device="cpu").manual_seed(9)
    q, hid = 13, 4544
    ctx = torch.randint(-2, 3, (q, hid), generator=g).half().to(DEV)
    act = torch.randint(-2, 3, (q, 4 * hid), generator=g).half().to(DEV)
    wd = (torch.randint(-2, 3, (hid, hid), generator=g).float() / 64).half().to(DEV)
    w4 = (torch.randint(-2, 3, (hid, 4 * hid), generator=g).float() / 64).half().to(DEV)
    res = torch.randn(q, hid, generator=g).half().to(DEV)
    lw = (1 + 0.1 * torch.randn(hid, generator=g)).half().to(DEV)
    lb = (0.1 * torch.randn(hid, generator=g)).half().to(DEV)
    ws = native.linear_workspace(DEV)
    pa = torch.empty(native.linear_partial_splits(hid, hid), q, hid, dtype=torch.float32, device=DEV)
    pm = torch.empty(native.linear_partial_splits(hid, 4 * hid), q, hid, dtype=torch.float32, device=DEV)
    native.linear_partials(ctx, wd, pa, ws, w_tiled=native.tile_weights(wd, native.PC_LIN_PARTIALS))
    native.linear_partials(act, w4, pm, ws, w_tiled=native.tile_weights(w4, native.PC_LIN_PARTIALS))
    a_x, a_h = res.clone(), torch.empty_like(res)
    native.add_layernorm(a_x, pm, pa, lw, lb, a_h, 1e-5)
    b_x, b_h = res.clone(), torch.empty_like(res)
    native.add_layernorm(b_x, _exact_linear(act, w4), _exact_linear(ctx, wd), lw, lb, b_h, 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(a_x, b_x) and torch.equal(a_h, b_h)


@pytest.mark.parametrize("q,hid,inter,H", [(9, 512, 1024, 4), (16, 4096, 11008, 32), (1, 4096, 11008, 32), (23, 1024, 2816, 8)])
def test_linear_chain_equals_the_launches_it_replaces(native, q, hid, inter, H):
    """pc_linear_chain (ONE persistent launch: o_proj partials -> add+RMSNorm -> gate/up+SiLU -> down partials -> add+RMSNorm ->
    QKV+RoPE+append) against the same six ops launched one by one: the linears' outputs are bit-identical given identical
    inputs, the norms differ only by the summation order of the row statistics (128 vs 512 threads)."""
    D, row0, max_ctx = 128, 40, 96
    g = torch.Generator(device="cpu").manual_seed(q + hid)
    rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).half().to(DEV)
    attn_out, x0 = rnd(q, hid), rnd(q, hid)
    wo, wgu, wd, wqkv = rnd(hid, hid, scale=0.03), rnd(2 * inter, hid, scale=0.03), rnd(hid, inter, scale=0.03), rnd(3 * hid, hid, scale=0.03)
    n1, n2 = (1 + 0.1 * torch.randn(hid, generator=g)).half().to(DEV), (1 + 0.1 * torch.randn(hid, generator=g)).half().to(DEV)
    pos = torch.randint(0, 2048, (q,), generator=g).to(DEV)
    cos, sin = O.rope_tables(D, 2048, dtype=torch.float16)
    cos, sin = cos.contiguous().to(DEV), sin.contiguous().to(DEV)
    ws = native.linear_workspace(DEV)
    t_o, t_gu = native.tile_weights(wo, native.PC_LIN_PARTIALS), native.tile_weights(wgu, native.PC_LIN_SILU_MUL)
    t_d, t_qkv = native.tile_weights(wd, native.PC_LIN_PARTIALS), native.tile_weights(wqkv, native.PC_LIN_ROPE_APPEND)

    def run(chained):
        x, h = x0.clone(), torch.zeros(q, hid, dtype=torch.half, device=DEV)
        act = torch.zeros(q, inter, dtype=torch.half, device=DEV)
        p_o = torch.zeros(native.linear_partial_splits(hid, hid), q, hid, dtype=torch.float32, device=DEV)
        p_d = torch.zeros(native.linear_partial_splits(hid, inter), q, hid, dtype=torch.float32, device=DEV)
        qo = torch.zeros(H, q, D, dtype=torch.half, device=DEV)
        kc, vc = torch.zeros(H, max_ctx, D, dtype=torch.half, device=DEV), torch.zeros(H, max_ctx, D, dtype=torch.half, device=DEV)
        import contextlib
        before = native.launch_count()
        with (native.chain(ws) if chained else contextlib.nullcontext()):
            native.linear_partials(attn_out, wo, p_o, ws, w_tiled=t_o)
            native.add_rmsnorm(x, p_o, n1, h, 1e-5)
            native.linear_small(h, wgu, act, ws, silu_mul=True, w_tiled=t_gu)
            native.linear_partials(act, wd, p_d, ws, w_tiled=t_d)
            native.add_rmsnorm(x, p_d, n2, h, 1e-5)
            native.linear_rope_append(h, wqkv, pos, cos, sin, qo, kc, vc, row0, H, H, ws, w_tiled=t_qkv)
        torch.cuda.synchronize()
        return dict(x=x, h=h, act=act, p_o=p_o, p_d=p_d, q=qo, k=kc, v=vc), native.launch_count() - before

    a, n_a = run(False)
    for rep in range(3):  # the chain leaves its barrier counters zeroed: replays give the same bits
        b, n_b = run(True)
        assert (n_a, n_b) == (6, 1)
        assert torch.equal(a["p_o"], b["p_o"])  # same inputs, same partition: same bits
        rel = lambda u, v: float((u.float() - v.float()).abs().max() / v.float().abs().max().clamp_min(1e-9))
        for name in ("x", "h", "act", "q", "k", "v"):
            assert rel(b[name], a[name]) <= 2e-3, (name, rel(b[name], a[name]))
        assert rel(b["p_d"].sum(0), a["p_d"].sum(0)) <= 2e-3
        if rep:
            assert all(torch.equal(b[k], prev[k]) for k in b)
        prev = b
    assert int(ws.view(torch.int32).abs().max()) == 0 or int(ws[:16384].view(torch.int32).abs().max()) == 0
    bars = ws[-256:].view(torch.int32)
    assert int(bars.abs().max()) == 0  # grid-barrier counters are left zeroed
