"""KV-buffer plumbing shared by the model families.

The reference re-materialises the whole KV with torch.cat on every forward (llama2.py:361-364,
falcon.py:446-452, mpt.py:148-151).  Here the KV of a sequence lives in a `KVBuffer` -- per-layer `[Hkv, max_ctx, D]`
bases with room behind the valid rows -- and the new rows are written in place behind them by pc_rope_append (or the QKV
projection's epilogue); the `past_key_values` handed back is the same buffer viewed N+q rows long.

In-place appends must not change what a caller can observe compared with the reference's functional forward, so every
buffer keeps two counters:
  * `high`   rows written so far.  A past that ends BELOW it (the same past used twice, e.g. to score two continuations,
             or a second generation started from the same spliced cache while the first is still running) is forked:
             its rows are copied to a fresh buffer instead of clobbering what the first call appended;
  * `epoch`  bumped whenever the owner re-writes the cached rows (PromptCache.update).  A past from an older epoch
             refers to rows that no longer exist; using it raises `StaleKVError` instead of silently attending over
             another request's segments (the reference's generator would still hold its private torch.cat copy).
Only buffers the models / the engine created themselves are appended to in place; foreign tensors are copied once.
"""
from __future__ import annotations

import itertools
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch


class StaleKVError(RuntimeError):
    """`past_key_values` / a `process()` cache from before the engine re-spliced its buffer."""


@dataclass
class CausalLMOutput:
    """The two fields of HF's CausalLMOutputWithPast the reference reads (generation_engine.py:120-121)."""
    logits: torch.Tensor
    past_key_values: Optional[Tuple[Tuple[torch.Tensor, torch.Tensor], ...]] = None

    def __getitem__(self, i):
        return (self.logits, self.past_key_values)[i]


_serial = itertools.count(1)
_registry: "dict[int, weakref.ref]" = {}  # storage data_ptr -> KVBuffer (to recognise plain views of our own buffers)


class KVBuffer:
    """Per-layer K/V bases `[Hkv, max_ctx, D]` (strided views of one allocation) + the append bookkeeping."""

    def __init__(self, k: List[torch.Tensor], v: List[torch.Tensor], max_ctx: int, storage: Optional[torch.Tensor] = None):
        self.k, self.v, self.max_ctx = k, v, max_ctx
        self.serial = next(_serial)   # never reused: identifies the buffer in CUDA-graph keys
        self.epoch = 0
        self.high = 0
        self._storage = storage
        if storage is not None:
            _registry[storage.untyped_storage().data_ptr()] = weakref.ref(self)

    def __del__(self):
        try:
            if self._storage is not None:
                _registry.pop(self._storage.untyped_storage().data_ptr(), None)
        except Exception:
            pass

    def rewritten(self, rows: int) -> None:
        """The owner re-wrote rows [.., rows): everything handed out before is stale."""
        self.epoch += 1
        self.high = rows

    @staticmethod
    def allocate(n_layers: int, n_kv_heads: int, cap: int, head_dim: int, device) -> "KVBuffer":
        buf = torch.zeros(n_layers, 2, n_kv_heads, cap, head_dim, dtype=torch.float16, device=device)
        return KVBuffer([buf[l, 0] for l in range(n_layers)], [buf[l, 1] for l in range(n_layers)], cap, storage=buf)


def buffer_of(t: torch.Tensor) -> Optional[KVBuffer]:
    ref = _registry.get(t.untyped_storage().data_ptr())
    return ref() if ref is not None else None


class VirtualKV:
    """The cached rows [0, rows) of a KVBuffer that have NOT been copied into it ("virtual splice"):
    they are read in place from the schema's segment arena [L,2,Hkv,R,D] through `tiles` (int32 [T,2] on the device:
    first arena row, valid rows).  `materialise()` makes the copy after all, for consumers that need the contiguous rows
    (long-suffix / ALiBi attention kernels, callers indexing the reference-shaped views)."""

    __slots__ = ("arena", "tiles", "n_tiles", "rows", "_materialise", "done", "slot_rows")

    def __init__(self, arena: torch.Tensor, tiles: torch.Tensor, n_tiles: int, rows: int, materialise,
                 slot_rows: Optional[torch.Tensor] = None):
        self.arena, self.tiles, self.n_tiles, self.rows = arena, tiles, n_tiles, rows
        self._materialise, self.done = materialise, False
        # int64 [n_tiles * 128] on the device: the buffer row each tile slot WOULD occupy after a copy (0 for slots past a
        # tile's valid rows) -- what a model with per-key position ids (MPT's ALiBi) needs to put its ids in tile order
        self.slot_rows = slot_rows

    def materialise(self) -> None:
        if not self.done:
            self._materialise()
            self.done = True

    def key(self):
        """What a captured CUDA graph bakes in (tensor maps over the arena, tile-list address, tile and row counts)."""
        return (self.arena.data_ptr(), int(self.arena.shape[3]), self.tiles.data_ptr(), self.n_tiles, self.rows,
                self.slot_rows.data_ptr() if self.slot_rows is not None else 0)

    def planes(self, layer: int):
        return (self.arena[layer, 0], self.arena[layer, 1], self.tiles, self.n_tiles, self.rows)


class LayerKV:
    """What one forward works on: a buffer and the number of valid rows in front of the rows it will write."""

    def __init__(self, buf: KVBuffer, n_cached: int, virtual: Optional[VirtualKV] = None):
        self.buf, self.n_cached, self.virtual = buf, n_cached, virtual

    k = property(lambda self: self.buf.k)
    v = property(lambda self: self.buf.v)
    max_ctx = property(lambda self: self.buf.max_ctx)

    def views(self, rows: int) -> "PastKV":
        return PastKV(self.buf, rows, virtual=self.virtual)

    def wrote(self, rows: int) -> None:
        self.buf.high = max(self.buf.high, rows)


class PastKV:
    """`past_key_values` as the models hand it back: behaves like the reference's tuple of L (K, V) pairs, each
    [1, Hkv, rows, D] (llama2.py:366), but the views are only materialised when somebody indexes them.  Passing it
    back into the next forward is O(1): the buffer is recognised instead of re-validated tensor by tensor."""

    def __init__(self, buf: KVBuffer, rows: int, epoch: Optional[int] = None, virtual: Optional[VirtualKV] = None):
        self.buf, self.rows = buf, rows
        self.epoch = buf.epoch if epoch is None else epoch
        self.virtual = virtual  # leading rows still live in the segment arena only (see VirtualKV)
        self._views = {}

    def __len__(self) -> int:
        return len(self.buf.k)

    def __getitem__(self, l):
        if isinstance(l, slice):
            return tuple(self[i] for i in range(*l.indices(len(self))))
        if l < 0:
            l += len(self)
        if l not in self._views:
            if self.virtual is not None:  # somebody wants the contiguous rows after all
                if self.epoch != self.buf.epoch:
                    raise StaleKVError("this past_key_values was handed out before the engine re-spliced its buffer")
                self.virtual.materialise()
            k, v = self.buf.k[l], self.buf.v[l]
            self._views[l] = (k[:, :self.rows, :].unsqueeze(0), v[:, :self.rows, :].unsqueeze(0))
        return self._views[l]

    def __iter__(self):
        return (self[l] for l in range(len(self)))


def _own_view_rows(past, head_dim: int) -> Optional[Tuple[KVBuffer, int]]:
    """Is `past` a list of plain `[1,Hkv,N,D]` views that start at row 0 of every layer of one of OUR buffers (what
    eval_sys.py:340 builds from `process()`'s cache)?  Returns (buffer, N)."""
    k0 = past[0][0]
    buf = buffer_of(k0)
    if buf is None or len(past) != len(buf.k) or k0.dim() != 4 or k0.shape[0] != 1:
        return None
    n = int(k0.shape[2])
    for l, (k, v) in enumerate(past):
        for t, base in ((k, buf.k[l]), (v, buf.v[l])):
            if (t.dim() != 4 or t.shape[0] != 1 or t.data_ptr() != base.data_ptr() or t.shape[2] != n or t.stride(3) != 1
                    or t.stride(2) != head_dim or (t.shape[1] > 1 and t.stride(1) != base.stride(0))):
                return None
    return buf, n


def _fork(src_k, src_v, n_cached: int, q_len: int, n_layers: int, n_kv_heads: int, head_dim: int, device,
          headroom: int) -> LayerKV:
    cap = ((n_cached + q_len + headroom + 127) // 128) * 128
    buf = KVBuffer.allocate(n_layers, n_kv_heads, cap, head_dim, device)
    if n_cached > 0:
        for l in range(n_layers):
            buf.k[l][:, :n_cached].copy_(src_k(l).reshape(n_kv_heads, n_cached, head_dim))
            buf.v[l][:, :n_cached].copy_(src_v(l).reshape(n_kv_heads, n_cached, head_dim))
    buf.high = n_cached
    return LayerKV(buf, n_cached)


def resolve_kv(past: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]], q_len: int, n_layers: int,
               n_kv_heads: int, head_dim: int, device, headroom: int = 256) -> LayerKV:
    """Find (or build) a buffer that already holds the past rows and has room for q_len more."""
    buf = rows = None
    if isinstance(past, PastKV):
        if past.epoch != past.buf.epoch:
            raise StaleKVError("this past_key_values / process() cache was handed out before the engine re-spliced its "
                               "buffer (another CacheEngine.process ran); its rows no longer exist")
        buf, rows = past.buf, past.rows
    elif past is not None and len(past) > 0 and past[0] is not None and int(past[0][0].shape[-2]) > 0:
        own = _own_view_rows(past, head_dim)
        if own is not None:
            buf, rows = own
    if buf is not None:
        virtual = past.virtual if isinstance(past, PastKV) else None
        if rows == buf.high and rows + q_len <= buf.max_ctx:
            return LayerKV(buf, rows, virtual)             # the tip of the sequence: append in place, O(1)
        # somebody already appended behind these rows (or there is no room): continue on a private copy
        if virtual is not None:
            virtual.materialise()
        return _fork(lambda l: buf.k[l][:, :rows], lambda l: buf.v[l][:, :rows], rows, q_len, n_layers, n_kv_heads,
                     head_dim, device, headroom)
    if past is None or len(past) == 0 or past[0] is None:
        n_cached = 0
    else:
        n_cached = int(past[0][0].shape[-2])
    # no past, or foreign tensors: the forward stays functional (the caller's storage is never written)
    return _fork(lambda l: past[l][0], lambda l: past[l][1], n_cached, q_len, n_layers, n_kv_heads, head_dim, device, headroom)
