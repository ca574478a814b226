"""KV-buffer plumbing shared by the model families.

The reference re-materialises the whole KV with torch.cat on every forward (llama2.py:361-364,
falcon.py:446-452, mpt.py:148-151).  Here `past_key_values` are row-contiguous views
[1,Hkv,N,D] of a buffer whose head stride (max_ctx*D) leaves room, and the new rows are written in place
behind them by pc_rope_append; the tuple handed back is the same buffer viewed N+q rows long.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch


@dataclass
class CausalLMOutput:
    """The two fields of HF's CausalLMOutputWithPast the reference reads (generation_engine.py:120-121)."""
    logits: torch.Tensor
    past_key_values: Optional[Tuple[Tuple[torch.Tensor, torch.Tensor], ...]] = None

    def __getitem__(self, i):
        return (self.logits, self.past_key_values)[i]


def _appendable(t: torch.Tensor, extra: int, head_dim: int) -> bool:
    """Can `extra` rows be written behind view t = [1,Hkv,N,D] without leaving its storage?"""
    if t.dim() != 4 or t.shape[0] != 1 or t.dtype != torch.float16 or not t.is_cuda:
        return False
    _, hkv, n, d = t.shape
    if d != head_dim or t.stride(3) != 1 or t.stride(2) != d:
        return False
    hs = t.stride(1)  # sliced views keep the buffer's head stride even when Hkv == 1 (Falcon MQA)
    if hs % d != 0 or hs < (n + extra) * d:
        return False
    if t.data_ptr() % 16 != 0:
        return False
    need = t.storage_offset() + (hkv - 1) * hs + (n + extra) * d
    return t.untyped_storage().nbytes() // t.element_size() >= need


class LayerKV:
    """Per-forward view of the per-layer K/V buffers: base pointers as [Hkv, max_ctx, D] strided tensors."""

    def __init__(self, k: List[torch.Tensor], v: List[torch.Tensor], n_cached: int, max_ctx: int):
        self.k, self.v, self.n_cached, self.max_ctx = k, v, n_cached, max_ctx

    def views(self, rows: int) -> "PastKV":
        return PastKV(self, rows)


class PastKV:
    """`past_key_values` as the models hand it back: behaves like the reference's tuple of L (K, V) pairs, each
    [1, Hkv, rows, D] (llama2.py:366), but the views are only materialised when somebody indexes them.  Passing it
    back into the next forward is O(1): the buffers are recognised instead of re-validated tensor by tensor."""

    def __init__(self, kv: LayerKV, rows: int):
        self.kv, self.rows = kv, rows
        self._views = {}

    def __len__(self) -> int:
        return len(self.kv.k)

    def __getitem__(self, l):
        if isinstance(l, slice):
            return tuple(self[i] for i in range(*l.indices(len(self))))
        if l < 0:
            l += len(self)
        if l not in self._views:
            k, v = self.kv.k[l], self.kv.v[l]
            self._views[l] = (k[:, :self.rows, :].unsqueeze(0), v[:, :self.rows, :].unsqueeze(0))
        return self._views[l]

    def __iter__(self):
        return (self[l] for l in range(len(self)))


def _as_strided_head_major(t: torch.Tensor, max_ctx: int) -> torch.Tensor:
    """[1,Hkv,N,D] view -> [Hkv,max_ctx,D] view over the same storage."""
    _, hkv, _, d = t.shape
    return t.as_strided((hkv, max_ctx, d), (max_ctx * d, d, 1), t.storage_offset())


def resolve_kv(past: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]], q_len: int, n_layers: int,
               n_kv_heads: int, head_dim: int, device, headroom: int = 256) -> LayerKV:
    """Find (or build) buffers that already hold the past rows and have room for q_len more."""
    if isinstance(past, PastKV) and past.rows + q_len <= past.kv.max_ctx:
        return LayerKV(past.kv.k, past.kv.v, past.rows, past.kv.max_ctx)  # our own buffers coming back: O(1)
    if past is None or len(past) == 0 or past[0] is None:
        n_cached = 0
    else:
        n_cached = int(past[0][0].shape[-2])
    if n_cached > 0:
        ok = all(_appendable(k, q_len, head_dim) and _appendable(v, q_len, head_dim) for k, v in past)
        if ok:
            ok = len({t.stride(1) for kv in past for t in kv}) == 1
        if ok:
            max_ctx = past[0][0].stride(1) // head_dim
            ks = [_as_strided_head_major(k, max_ctx) for k, _ in past]
            vs = [_as_strided_head_major(v, max_ctx) for _, v in past]
            return LayerKV(ks, vs, n_cached, max_ctx)
    # own buffer (no past, or past without room: copy it once)
    cap = ((n_cached + q_len + headroom + 127) // 128) * 128
    buf = torch.zeros(n_layers, 2, n_kv_heads, cap, head_dim, dtype=torch.float16, device=device)
    if n_cached > 0:
        for l, (k, v) in enumerate(past):
            buf[l, 0, :, :n_cached].copy_(k.reshape(n_kv_heads, n_cached, head_dim))
            buf[l, 1, :, :n_cached].copy_(v.reshape(n_kv_heads, n_cached, head_dim))
    return LayerKV([buf[l, 0] for l in range(n_layers)], [buf[l, 1] for l in range(n_layers)], n_cached, cap)
