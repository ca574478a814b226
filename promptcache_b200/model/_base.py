"""Shared plumbing of the B200 model families: HF-style forward signature, in-place KV buffers, RoPE tables,
attention workspace and CUDA-graph replay of the single-sequence forward."""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from ._kv import CausalLMOutput, PastKV, resolve_kv


def build_rope_tables(head_dim: int, rows: int, theta: float, device, dtype=torch.float16):
    """cos/sin tables as LlamaRotaryEmbedding._set_cos_sin_cache (llama2.py:128-137) and
    FalconRotaryEmbedding._set_cos_sin_cache (falcon.py:104-118) build them: fp32 angles, cast at the end."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device).float() / head_dim))
    t = torch.arange(rows, device=device, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype).contiguous(), emb.sin().to(dtype).contiguous()


class B200CausalLM(nn.Module):
    """Subclasses set n_layers/n_heads/n_kv_heads/head_dim/vocab_size, `rotary` and implement `_layers`."""

    accepts_max_position = True   # forward() takes the host-known max position id (avoids a device sync)
    full_position_ids = False     # MPT: position_ids cover cached + new tokens (mpt.py:172)
    rotary = True

    def _init_runtime(self):
        self._cos = self._sin = None
        self._ws = None
        self._lin_ws = None
        self.attn_impl = nat.PC_ATTN_AUTO
        # "native": skinny linears (q <= 32) run in pc_linear_small; "cublas": torch.nn.functional.linear everywhere
        self.linear_impl = os.environ.get("PC_LINEAR", "native")
        self.linear_native_ops = set(os.environ.get("PC_LINEAR_OPS", "qkv,o,gu,down,head").split(","))
        self.linear_max_q = min(32, int(os.environ.get("PC_LINEAR_MAX_Q", "16")))
        self.linear_partials = os.environ.get("PC_LINEAR_PARTIALS", "1") != "0"
        self.linear_tiled = os.environ.get("PC_LINEAR_TILED", "1") != "0"  # native linears stream pre-tiled weight copies
        # experiment, off: each native linear asks L2 for the head of the next one's weight stream while its own tail drains --
        # measured SLOWER (3.43 vs 3.31 ms per C2 step, worse with more stages): the tail is bandwidth-contended, not idle
        self.linear_prefetch = os.environ.get("PC_LINEAR_PREFETCH", "0") != "0"
        # experiment, off: the linears and norms between two attention launches as ONE persistent launch (pc_linear_chain, grid
        # barriers between the ops).  Measured SLOWER: 106 vs 87 us per Llama-2-7B layer (4.22 vs 3.30 ms per C2 step) -- each
        # of the 5 grid barriers + 2 in-kernel norms of a layer is a chain of L2 round trips under a saturated memory system,
        # ~33 us of consumer stalls per layer against the ring's 4.3 us of run-ahead
        self.linear_chain = os.environ.get("PC_LINEAR_CHAIN", "0") != "0"
        # experiment, off: the short-suffix attention kernel asks L2 for the output projection's weight stream once its own K/V
        # loads are out (that projection's CTAs cannot start before the attention CTAs leave their SMs).  Measured same-box:
        # C2 step 3.250 -> 3.268 ms, decode 3.18 -> 3.17 ms/token -- like every other L2 prefetch tried here, no gain
        self.attn_prefetch_o = os.environ.get("PC_ATTN_PREFETCH", "0") != "0"
        # CUDA graphs of the whole single-sequence forward, keyed by (q_len, n_cached, KV buffer): the per-layer
        # launch sequence is fixed once those are, so a repeated request shape is replayed with one launch
        self.ablate = frozenset()  # measurement only: kernel families left out of the step (see LlamaForCausalLM._layers_ablated)
        self.use_cuda_graphs = True
        self._graphs = {}
        self._graph_seen = {}
        self._decode_pos_len = {}
        self._tiled = {}  # pre-tiled copies of the weights for pc_linear_small, built on first native use
        self.max_graphs = 8

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    # ---- RoPE tables / workspace -------------------------------------------------------------------------
    def ensure_rope(self, rows: int) -> None:
        """Grow-on-demand like llama2.py:141-142, but driven by host-known ids (no device sync)."""
        if not self.rotary:
            return
        rows = max(int(rows), int(getattr(self.config, "max_position_embeddings", 2048)))
        if self._cos is None or self._cos.shape[0] < rows or self._cos.device != self.device:
            rows = ((rows + 1023) // 1024) * 1024
            self._graphs.clear()  # captured graphs hold the old table pointers
            self._cos, self._sin = build_rope_tables(self.head_dim, rows, self.theta, self.device)

    def _workspace(self, q_len: int) -> Optional[torch.Tensor]:
        need = nat.attn_workspace_bytes(self.n_heads, q_len, self.head_dim)
        if self._ws is None or self._ws.numel() < need or self._ws.device != self.device:
            # zero-initialised once: the first bytes hold split-KV arrival counters the kernels leave zeroed
            self._graphs.clear()  # captured graphs hold the old workspace pointer
            self._ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _wt(self, key, w: torch.Tensor, epilogue: int):
        """Pre-tiled copy of a weight (pc_linear_tile_weights): doubles the memory of the matrices it is used for, and
        lets one SM stream ~1.5x more bytes per second than 128-byte-row tensor-map boxes do."""
        t = self._tiled.get(key)
        if t is None:
            t = nat.tile_weights(w, epilogue) if self.linear_tiled else None
            self._tiled[key] = t if t is not None else False
        return t if t is not False else None

    def _linear_workspace(self) -> torch.Tensor:
        if self._lin_ws is None or self._lin_ws.device != self.device:
            self._graphs.clear()  # captured graphs hold the old workspace pointer
            self._lin_ws = nat.linear_workspace(self.device)  # zeroed once; the kernel leaves its counters zeroed
        return self._lin_ws

    # ---- forward ---------------------------------------------------------------------------------------
    def _layers(self, ids: torch.Tensor, pos: torch.Tensor, kv, extra) -> torch.Tensor:
        raise NotImplementedError

    def _graph_extra(self, pos_host_max: int):
        """Host scalars baked into kernel parameters besides (q_len, n_cached); part of the graph key."""
        return ()

    supports_dynamic_n = False    # _layers(..., n_dev=) reads the cached length on the device (decode graph)

    def virtual_kv_max_q(self) -> int:
        """Longest suffix whose attention can read cached segments in place from the schema arena (virtual splice);
        0 = this model always needs the contiguous rows."""
        return 0

    virtual_kv_full_ids = False   # the model takes position ids of ALL keys and can put them in tile order (MPT)

    def _virtual_ok(self, virt, q_len: int) -> bool:
        """Can this forward read the cached rows in place (else they are copied first)?"""
        return 0 < q_len <= self.virtual_kv_max_q()

    def _launch_shape_key(self):
        """Every switch that changes the launch sequence of `_layers`; part of each graph key so that toggling one
        after a capture can never replay a stale sequence."""
        return (self.linear_impl, tuple(sorted(self.linear_native_ops)), self.linear_max_q, self.linear_partials,
                self.linear_tiled, self.linear_prefetch, self.linear_chain, self.attn_prefetch_o, self.attn_impl,
                tuple(sorted(self.ablate)))

    def _note_seen(self, key) -> None:
        if len(self._graph_seen) >= 256:  # bounded: forked / short-lived buffers would otherwise grow it forever
            self._graph_seen.clear()
        self._graph_seen[key] = self._graph_seen.get(key, 0) + 1

    def _weights_changed(self) -> None:
        """Captured graphs and pre-tiled weight copies refer to the old parameter storage / values."""
        self._graphs.clear()
        for cache in (getattr(self, "_tiled", None), getattr(self, "_hints", None)):
            if cache is not None:
                cache.clear()

    def _apply(self, fn, *args, **kwargs):  # .to() / .half() / .cuda()
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_graphs"):
            self._weights_changed()
            self._cos = self._sin = self._ws = self._lin_ws = None
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._weights_changed()
        return out

    def _decode_replay(self, ids, pos, kv, static_ok=False):
        """One-token decode through ONE captured graph per KV buffer: the cached length is a device scalar the
        kernels read at run time (pc_rope_append_dyn, pc_attn_params.n_cached_dev), so the same graph serves every
        step of every request on that buffer (reference loop: generation_engine.py:123-147)."""
        key = ("decode", kv.buf.serial, kv.max_ctx, kv.virtual.key() if kv.virtual is not None else None) + self._launch_shape_key()
        entry = self._graphs.get(key)
        if entry is None:
            if self._graph_seen.get(key, 0) < 1:  # first decode step on this buffer runs eagerly
                self._note_seen(key)
                return None
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            s_ids = ids.clone()
            if self.full_position_ids:  # ids of every key: a max_ctx-long static buffer, filled up to N+1 per step
                s_pos = torch.zeros(kv.max_ctx, dtype=pos.dtype, device=self.device)
                s_pos[:pos.shape[0]].copy_(pos)
            else:
                s_pos = pos.clone()
            s_n = torch.full((1,), kv.n_cached, dtype=torch.int32, device=self.device)
            self._layers(s_ids, s_pos, kv, (), n_dev=s_n)  # warm-up; rewrites row n_cached with the same values
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            before = nat.launch_count()
            with torch.cuda.graph(graph):
                s_logits = self._layers(s_ids, s_pos, kv, (), n_dev=s_n)
            entry = (graph, s_ids, s_pos, s_n, s_logits, nat.launch_count() - before)
            self._graphs[key] = entry
        graph, s_ids, s_pos, s_n, s_logits, n_native = entry
        s_ids.copy_(ids, non_blocking=True)
        if self.full_position_ids:
            n = pos.shape[0]
            if n < self._decode_pos_len.get(key, 0):  # a shorter sequence than the last one on this buffer: clear stale ids
                s_pos[n:].zero_()
            self._decode_pos_len[key] = n
            s_pos[:n].copy_(pos, non_blocking=True)
        else:
            s_pos.copy_(pos, non_blocking=True)
        s_n.fill_(kv.n_cached)
        graph.replay()
        nat.note_replayed_launches(n_native)
        return s_logits if static_ok else s_logits.clone()

    def _forward_one(self, ids: torch.Tensor, pos: torch.Tensor, past, use_cache: bool, max_position: int,
                     static_ok: bool = False):
        """`static_ok`: the caller consumes the logits before the next forward, so a graph replay may hand back its
        static output buffer instead of a copy (GenerationEngine's device-side greedy loop)."""
        q_len = ids.shape[0]
        kv = resolve_kv(past, q_len, self.n_layers, self.n_kv_heads, self.head_dim, self.device)
        N = kv.n_cached
        if kv.virtual is not None and not self._virtual_ok(kv.virtual, q_len):
            kv.virtual.materialise()  # this forward needs the contiguous rows: make the copy the engine skipped
            kv.virtual = None
        if (q_len == 1 and N > 0 and self.use_cuda_graphs and self.supports_dynamic_n
                and N + 1 < kv.max_ctx and not torch.cuda.is_current_stream_capturing()):
            logits = self._decode_replay(ids, pos, kv, static_ok)
            if logits is not None:
                kv.wrote(N + 1)
                return logits, (kv.views(N + 1) if use_cache else None)
        extra = self._graph_extra(max_position)
        key = (q_len, N, kv.buf.serial, kv.max_ctx, pos.shape[0], kv.virtual.key() if kv.virtual is not None else None) \
            + self._launch_shape_key() + tuple(extra)
        if self.use_cuda_graphs and N > 0 and q_len > 1 and not torch.cuda.is_current_stream_capturing():
            entry = self._graphs.get(key)
            if entry is None and self._graph_seen.get(key, 0) >= 1:
                entry = self._capture(key, ids, pos, kv, extra)
            self._note_seen(key)
            if entry is not None:
                graph, s_ids, s_pos, s_logits, n_native = entry
                s_ids.copy_(ids, non_blocking=True)
                s_pos.copy_(pos, non_blocking=True)
                graph.replay()
                nat.note_replayed_launches(n_native)
                kv.wrote(N + q_len)
                return (s_logits if static_ok else s_logits.clone()), (kv.views(N + q_len) if use_cache else None)
        logits = self._layers(ids, pos, kv, extra)
        kv.wrote(N + q_len)
        return logits, (kv.views(N + q_len) if use_cache else None)

    def _capture(self, key, ids, pos, kv, extra):
        """Record the forward for this (q_len, n_cached, KV buffer) once; later calls replay it."""
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        s_ids, s_pos = ids.clone(), pos.clone()
        self._layers(s_ids, s_pos, kv, extra)  # warm-up on the capture inputs (function attributes, workspaces, cuBLAS)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = nat.launch_count()
        with torch.cuda.graph(graph):
            s_logits = self._layers(s_ids, s_pos, kv, extra)
        self._graphs[key] = (graph, s_ids, s_pos, s_logits, nat.launch_count() - before)
        return self._graphs[key]

    @torch.inference_mode()
    def forward(self, input_ids: torch.Tensor = None, position_ids: Optional[torch.Tensor] = None,
                past_key_values=None, attention_mask: Optional[torch.Tensor] = None, inputs_embeds=None, labels=None,
                use_cache: Optional[bool] = None, output_attentions=None, output_hidden_states=None,
                return_dict=None, max_position: Optional[int] = None, static_logits_ok: bool = False, **_):
        if input_ids is None:
            raise ValueError("promptcache_b200 models need input_ids")
        if not input_ids.is_cuda:
            raise nat.NativeError("promptcache_b200 models run on CUDA only (no CPU fallback)")
        use_cache = True if use_cache is None else use_cache
        bsz, q_len = input_ids.shape
        if isinstance(past_key_values, PastKV):
            past_len = past_key_values.rows
        else:
            past_len = 0 if not past_key_values else int(past_key_values[0][0].shape[-2])
        if position_ids is None:
            first = 0 if self.full_position_ids else past_len
            position_ids = torch.arange(first, past_len + q_len, device=input_ids.device).unsqueeze(0).expand(bsz, -1)
        position_ids = position_ids.view(bsz, -1).long()
        want = past_len + q_len if self.full_position_ids else q_len
        if position_ids.shape[1] != want:
            raise ValueError(f"position_ids has {position_ids.shape[1]} entries, expected {want}")
        if max_position is None:
            max_position = int(position_ids.max())  # device sync; GenerationEngine passes the host value instead
        self.ensure_rope(max_position + 1)
        if bsz == 1:
            logits, pkv = self._forward_one(input_ids[0], position_ids[0].contiguous(), past_key_values, use_cache,
                                            max_position, static_logits_ok)
            return CausalLMOutput(logits.unsqueeze(0), pkv)
        # batch > 1 (schema encoding with right padding, cache_engine.py:240-248): rows are independent
        if past_key_values:
            raise ValueError("batched forward with past_key_values is not supported")
        lens = [q_len] * bsz if attention_mask is None else [int(v) for v in attention_mask.to(torch.int64).sum(-1).tolist()]
        logits = torch.zeros(bsz, q_len, self.vocab_size, dtype=torch.float32, device=input_ids.device)
        pk = [[None] * bsz for _ in range(self.n_layers)]
        pv = [[None] * bsz for _ in range(self.n_layers)]
        for b in range(bsz):
            n = max(lens[b], 1)
            lg, pkv = self._forward_one(input_ids[b, :n], position_ids[b, :n].contiguous(), None, True,
                                        int(position_ids[b, :n].max()))
            logits[b, :n] = lg
            for l, (k, v) in enumerate(pkv):
                pk[l][b] = F.pad(k[0], (0, 0, 0, q_len - n))
                pv[l][b] = F.pad(v[0], (0, 0, 0, q_len - n))
            del pkv
        pkv = tuple((torch.stack(pk[l]), torch.stack(pv[l])) for l in range(self.n_layers)) if use_cache else None
        return CausalLMOutput(logits, pkv)
