"""Size-independent properties at BASELINE.json's full sizes (where the oracle would take too long or too much
memory): checksums, idempotence, linearity, row-stochasticity."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _checksum(t: torch.Tensor) -> int:
    """Order-independent checksum of the raw fp16 bit patterns."""
    return int(t.contiguous().view(torch.int16).to(torch.int64).sum())


def test_splice_c5_full_size_checksum_and_idempotence(native):
    """C5: 32 segments x 256 tokens at Llama-2-7B dims = 4 GiB moved; a checksum of checksums instead of a 4 GiB oracle."""
    L, Hkv, D, n_seg, seg_len = 32, 32, 128, 32, 256
    g = torch.Generator(device=DEV).manual_seed(5)
    pool = [torch.randn(L, 2, Hkv, seg_len, D, device=DEV, generator=g).half() for _ in range(4)]  # 4 distinct segments reused
    max_ctx = n_seg * seg_len
    cache = torch.zeros(L, 2, Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    plan = [(pool[i % 4], i * seg_len) for i in range(n_seg)]
    native.splice_kv(plan, cache)
    torch.cuda.synchronize()
    want = sum(_checksum(p) for p, _ in plan)
    assert _checksum(cache) == want
    # every destination block equals its source (spot-check all 32 blocks on 3 layers)
    for i, (src, row) in enumerate(plan):
        for l in (0, 13, 31):
            assert torch.equal(cache[l, :, :, row:row + seg_len], src[l])
    before = cache.clone()
    native.splice_kv(plan, cache)  # idempotent
    assert torch.equal(cache, before)
    del before
    # ragged re-splice of a prefix leaves the other rows untouched
    native.splice_kv([(pool[3], 0)], cache)
    assert torch.equal(cache[:, :, :, :seg_len], pool[3]) and torch.equal(cache[5, :, :, seg_len:2 * seg_len], pool[1][5])


@pytest.mark.parametrize("H,Hkv,D,q,N", [(32, 32, 128, 32, 8192), (32, 32, 128, 1, 8223), (40, 40, 128, 32, 4096),
                                         (32, 32, 128, 1024, 8192), (71, 1, 64, 21, 256)])
def test_attention_rows_are_stochastic_and_linear_in_v(native, H, Hkv, D, q, N):
    g = torch.Generator(device=DEV).manual_seed(q + N)
    max_ctx = ((N + q + 127) // 128) * 128
    Q = (torch.randn(H, q, D, device=DEV, generator=g) * 1.2).half()
    K = (torch.randn(Hkv, max_ctx, D, device=DEV, generator=g) * 1.2).half()
    V1 = torch.randn(Hkv, max_ctx, D, device=DEV, generator=g).half()
    V2 = torch.randn(Hkv, max_ctx, D, device=DEV, generator=g).half()
    ws = torch.zeros(max(native.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=DEV)
    rounding = native.PC_ROUND_LLAMA if D == 128 else native.PC_ROUND_FALCON

    def run(V):
        out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
        native.attention(Q, K, V, out, N, q, H, Hkv, D, native.softmax_scale(D), rounding, workspace=ws, decode=(q == 1))
        return out.float()

    ones = run(torch.ones_like(V1))
    assert float((ones - 1).abs().max()) < 2e-3          # softmax rows sum to one (P rounded to fp16, l in fp32)
    o1, o2, o12 = run(V1), run(V2), run((V1.float() + V2.float()).half())
    assert float((o12 - (o1 + o2)).abs().max()) < 6e-3   # linear in V up to fp16 rounding of V1+V2 and of the outputs
    # keys beyond N+q never contribute: poison the padding rows
    Kp, Vp = K.clone(), V1.clone()
    Kp[:, N + q:] = 30000.0
    Vp[:, N + q:] = 30000.0
    out = torch.empty(q, H * D, dtype=torch.half, device=DEV)
    native.attention(Q, Kp, Vp, out, N, q, H, Hkv, D, native.softmax_scale(D), rounding, workspace=ws, decode=(q == 1))
    assert torch.equal(out.float(), o1)


def test_causal_suffix_rows_ignore_later_suffix_keys(native):
    """Row r of the suffix must not change when later suffix keys change (sequence-order causality, llama2.py:62-76)."""
    H = Hkv = 8
    D, q, N = 128, 300, 500
    g = torch.Generator(device=DEV).manual_seed(1)
    max_ctx = 1024
    Q = torch.randn(H, q, D, device=DEV, generator=g).half()
    K = torch.randn(Hkv, max_ctx, D, device=DEV, generator=g).half()
    V = torch.randn(Hkv, max_ctx, D, device=DEV, generator=g).half()
    ws = torch.zeros(max(native.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=DEV)
    a = torch.empty(q, H * D, dtype=torch.half, device=DEV)
    b = torch.empty_like(a)
    native.attention(Q, K, V, a, N, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws)
    K2, V2 = K.clone(), V.clone()
    K2[:, N + 150:N + q] = torch.randn(Hkv, 150, D, device=DEV, generator=g).half()
    V2[:, N + 150:N + q] = 7.0
    native.attention(Q, K2, V2, b, N, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws)
    assert torch.equal(a[:150], b[:150]) and not torch.equal(a[150:], b[150:])
