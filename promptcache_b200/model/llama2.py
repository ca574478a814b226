"""Llama-2 / CodeLlama on the B200 hot path.

Same forward signature and outputs as reference promptcache/model/llama2.py (LlamaForCausalLM.forward
:986-1076 -> LlamaModel.forward :822-951 -> LlamaDecoderLayer :600-654 -> LlamaAttention :315-410).
The per-layer attention core (llama2.py:353-398: RoPE at explicit ids, KV concat, repeat_kv, mask, QK^T,
softmax, PV, transpose) is TWO kernel launches here -- pc_rope_append and pc_prefill_attn/pc_decode_attn --
called through the C ABI; projections, norms and the MLP stay torch/cuBLAS.
"""
from __future__ import annotations

import contextlib

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from ._base import B200CausalLM, build_rope_tables  # noqa: F401  (re-exported)


class LlamaConfigLite:
    """The config fields the hot path needs (a transformers LlamaConfig works too)."""

    def __init__(self, vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=None, max_position_embeddings=4096, rms_norm_eps=1e-5,
                 rope_theta=10000.0, initializer_range=0.02, **_):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.num_key_value_heads = num_key_value_heads or num_attention_heads
        self.max_position_embeddings = max_position_embeddings
        self.rms_norm_eps = rms_norm_eps
        self.rope_theta = rope_theta
        self.initializer_range = initializer_range


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm.forward (llama2.py:104-109): fp32 statistics, round to the input dtype, then scale."""
    h = x.to(torch.float32)
    h = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps)
    return weight * h.to(x.dtype)


class LlamaForCausalLM(B200CausalLM):
    """fp16 Llama with fused QKV / gate-up weights; attention runs in the sm_100a kernels."""

    @property
    def supports_dynamic_n(self) -> bool:  # the short-suffix kernel reads N on the device (D=128, G <= 32)
        return self.head_dim == 128 and self.n_heads // self.n_kv_heads <= 32

    def virtual_kv_max_q(self) -> int:  # the short-suffix kernel reads segments in place: D=128, G*q <= 32
        return 32 // (self.n_heads // self.n_kv_heads) if self.head_dim == 128 else 0

    def __init__(self, config, device="cuda", dtype=torch.float16):
        super().__init__()
        self.config = config
        c = config
        self.vocab_size = c.vocab_size
        self.n_layers = c.num_hidden_layers
        self.n_heads = c.num_attention_heads
        self.n_kv_heads = getattr(c, "num_key_value_heads", None) or c.num_attention_heads
        self.head_dim = c.hidden_size // c.num_attention_heads
        self.eps = c.rms_norm_eps
        self.theta = float(getattr(c, "rope_theta", 10000.0))
        hid, inter = c.hidden_size, c.intermediate_size
        qkv_out = (self.n_heads + 2 * self.n_kv_heads) * self.head_dim
        std = getattr(c, "initializer_range", 0.02)

        def w(*shape):
            return nn.Parameter(torch.empty(*shape, device=device, dtype=dtype).normal_(0.0, std), requires_grad=False)

        self.embed_tokens = w(c.vocab_size, hid)
        self.wqkv = nn.ParameterList([w(qkv_out, hid) for _ in range(self.n_layers)])
        self.wo = nn.ParameterList([w(hid, self.n_heads * self.head_dim) for _ in range(self.n_layers)])
        self.w_gate_up = nn.ParameterList([w(2 * inter, hid) for _ in range(self.n_layers)])
        self.w_down = nn.ParameterList([w(hid, inter) for _ in range(self.n_layers)])
        ones = lambda: nn.Parameter(torch.ones(hid, device=device, dtype=dtype), requires_grad=False)
        self.ln1 = nn.ParameterList([ones() for _ in range(self.n_layers)])
        self.ln2 = nn.ParameterList([ones() for _ in range(self.n_layers)])
        self.norm = ones()
        self.lm_head = w(c.vocab_size, hid)
        self._init_runtime()
        self._hints = {}  # prefetch_next descriptions of the native linears
        self.rounding = nat.PC_ROUND_LLAMA
        self.scale = nat.softmax_scale(self.head_dim)

    # ---- weights ------------------------------------------------------------------------------------
    @torch.no_grad()
    def load_hf_state_dict(self, sd) -> None:
        """Load a HuggingFace Llama state dict (model.layers.N.self_attn.q_proj.weight ...)."""
        g = lambda k: sd[k].to(self.embed_tokens.device, self.embed_tokens.dtype)
        self._tiled.clear()
        self._hints.clear()
        self._graphs.clear()
        self.embed_tokens.copy_(g("model.embed_tokens.weight"))
        self.norm.copy_(g("model.norm.weight"))
        self.lm_head.copy_(g("lm_head.weight"))
        for l in range(self.n_layers):
            p = f"model.layers.{l}."
            self.wqkv[l].copy_(torch.cat([g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"),
                                          g(p + "self_attn.v_proj.weight")], 0))
            self.wo[l].copy_(g(p + "self_attn.o_proj.weight"))
            self.w_gate_up[l].copy_(torch.cat([g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")], 0))
            self.w_down[l].copy_(g(p + "mlp.down_proj.weight"))
            self.ln1[l].copy_(g(p + "input_layernorm.weight"))
            self.ln2[l].copy_(g(p + "post_attention_layernorm.weight"))

    @classmethod
    def from_pretrained(cls, name: str, device="cuda", **kwargs):
        """Load through transformers, then repack (needs local weights; there is no network in CI)."""
        from transformers import AutoConfig, AutoModelForCausalLM
        kwargs.pop("load_in_8bit", None)  # bitsandbytes int8 linears are out of scope; fp16 linears are used
        kwargs.pop("device_map", None)
        cfg = AutoConfig.from_pretrained(name)
        hf = AutoModelForCausalLM.from_pretrained(name, torch_dtype=torch.float16, **kwargs)
        model = cls(cfg, device=device)
        model.load_hf_state_dict(hf.state_dict())
        return model

    def _layers_ablated(self, x, h, pos, kv, q_rot, attn_out, act, ws, lws, pbuf_o, pbuf_d, n_dev):
        """The native q <= linear_max_q step with the kernel families named in `self.ablate` left out ("linear", "attn",
        "norm").  Results are meaningless; bench.py times it to attribute the step: cost(X) = T(full) - T(without X),
        which includes the overlap X has with its neighbours (programmatic dependent launch) -- the in-step number."""
        ab = self.ablate
        q_len, H, Hkv, D, N = x.shape[0], self.n_heads, self.n_kv_heads, self.head_dim, kv.n_cached
        if lws is None or pbuf_o is None or pbuf_d is None:
            raise nat.NativeError("ablation runs need the native skinny-linear step (q_len <= linear_max_q)")
        for l in range(self.n_layers):
            if "norm" not in ab:
                nat.add_rmsnorm(x, pbuf_d if l else None, self.ln1[l], h, self.eps)
            if "linear" not in ab:
                nat.linear_rope_append(h, self.wqkv[l], pos, self._cos, self._sin, q_rot, kv.k[l], kv.v[l], N, H, Hkv, lws,
                                       w_tiled=self._wt(("qkv", l), self.wqkv[l], nat.PC_LIN_ROPE_APPEND))
            if "attn" not in ab:
                virt = kv.virtual
                nat.attention(q_rot, kv.k[l], kv.v[l], attn_out, N, q_len, H, Hkv, D, self.scale, self.rounding, workspace=ws,
                              decode=(q_len == 1), impl=self.attn_impl if virt is None else nat.PC_ATTN_TCGEN05,
                              virtual=virt.planes(l) if virt is not None else None)
            if "linear" not in ab:
                nat.linear_partials(attn_out, self.wo[l], pbuf_o, lws, w_tiled=self._wt(("o", l), self.wo[l], nat.PC_LIN_PARTIALS))
            if "norm" not in ab:
                nat.add_rmsnorm(x, pbuf_o, self.ln2[l], h, self.eps)
            if "linear" not in ab:
                nat.linear_small(h, self.w_gate_up[l], act, lws, silu_mul=True,
                                 w_tiled=self._wt(("gu", l), self.w_gate_up[l], nat.PC_LIN_SILU_MUL))
                nat.linear_partials(act, self.w_down[l], pbuf_d, lws, w_tiled=self._wt(("down", l), self.w_down[l], nat.PC_LIN_PARTIALS))
        if "norm" not in ab:
            nat.add_rmsnorm(x, pbuf_d, self.norm, h, self.eps)
        logits = torch.zeros(q_len, self.vocab_size, dtype=torch.float16, device=self.device)
        if "linear" not in ab:
            nat.linear_small(h, self.lm_head, logits, lws, w_tiled=self._wt(("lm_head", 0), self.lm_head, nat.PC_LIN_PLAIN))
        return logits.float()

    @torch.inference_mode()
    def forward_many(self, items, max_position: Optional[int] = None):
        """Cached prefill of SEVERAL independent sequences in one pass (schema encoding; the reference
        pads whole scaffolds into a batch instead, cache_engine.py:228-248).  `items` = [(ids [n_b], position ids [n_b],
        past_key_values or None)].  The tokens of all sequences run through the norms / projections / MLP as ONE
        concatenated [sum n_b, hidden] matrix (no padding, one weight stream for all of them); RoPE + append and
        attention run per sequence over that sequence's own KV buffer.  Returns the per-sequence past_key_values
        (no logits: the encoder only stores KV)."""
        from ._kv import resolve_kv
        H, Hkv, D = self.n_heads, self.n_kv_heads, self.head_dim
        lens = [int(ids.shape[0]) for ids, _, _ in items]
        kvs = [resolve_kv(past, n, self.n_layers, Hkv, D, self.device) for (_, _, past), n in zip(items, lens)]
        for kv in kvs:
            if kv.virtual is not None:
                kv.virtual.materialise()
                kv.virtual = None
        ids = torch.cat([i for i, _, _ in items])
        pos = torch.cat([p for _, p, _ in items]).long().contiguous()
        if max_position is None:
            max_position = int(pos.max())
        self.ensure_rope(max_position + 1)
        T, inter = ids.shape[0], self.config.intermediate_size
        x = F.embedding(ids, self.embed_tokens)
        h = torch.empty_like(x)
        attn_out = torch.empty(T, H * D, dtype=torch.float16, device=self.device)
        act = torch.empty(T, inter, dtype=torch.float16, device=self.device)
        q_rots = [torch.empty(H, n, D, dtype=torch.float16, device=self.device) for n in lens]
        ws = self._workspace(max(lens))
        delta = None
        for l in range(self.n_layers):
            nat.add_rmsnorm(x, delta, self.ln1[l], h, self.eps)
            qkv = F.linear(h, self.wqkv[l])
            off = 0
            for kv, n, q_rot in zip(kvs, lens, q_rots):
                rows = qkv[off:off + n]
                nat.rope_append(rows[:, :H * D], rows[:, H * D:(H + Hkv) * D], rows[:, (H + Hkv) * D:], pos[off:off + n], self._cos,
                                self._sin, q_rot, kv.k[l], kv.v[l], kv.n_cached, H, Hkv, D)
                nat.attention(q_rot, kv.k[l], kv.v[l], attn_out[off:off + n], kv.n_cached, n, H, Hkv, D, self.scale, self.rounding,
                              workspace=ws, impl=self.attn_impl, decode=False)
                off += n
            if l + 1 == self.n_layers:
                break  # the last layer's K/V are stored: its output projection and MLP feed nothing the encoder keeps
            delta = F.linear(attn_out, self.wo[l])
            nat.add_rmsnorm(x, delta, self.ln2[l], h, self.eps)
            nat.silu_mul(F.linear(h, self.w_gate_up[l]), act, inter)
            delta = F.linear(act, self.w_down[l])
        out = []
        for kv, n in zip(kvs, lens):
            kv.wrote(kv.n_cached + n)
            out.append(kv.views(kv.n_cached + n))
        return out

    def _hint(self, name: str, l: int, partials: bool):
        """`prefetch_next` description of the linear that follows (its CTAs' first weight stages are asked from L2 by
        the tail of the current one); None when that linear is not a pre-tiled native launch."""
        if not self.linear_prefetch or name not in self.linear_native_ops:
            return None
        key = (name, l, partials)
        h = self._hints.get(key)
        if h is None:
            H, Hkv = self.n_heads, self.n_kv_heads
            if name == "qkv":
                h = nat.linear_hint(self.wqkv[l], nat.PC_LIN_ROPE_APPEND, self._wt(("qkv", l), self.wqkv[l], nat.PC_LIN_ROPE_APPEND), H, Hkv)
            elif name == "gu":
                h = nat.linear_hint(self.w_gate_up[l], nat.PC_LIN_SILU_MUL, self._wt(("gu", l), self.w_gate_up[l], nat.PC_LIN_SILU_MUL))
            elif name == "head":
                h = nat.linear_hint(self.lm_head, nat.PC_LIN_PLAIN, self._wt(("lm_head", 0), self.lm_head, nat.PC_LIN_PLAIN))
            else:
                w = self.wo[l] if name == "o" else self.w_down[l]
                epi = nat.PC_LIN_PARTIALS if partials else nat.PC_LIN_PLAIN
                h = nat.linear_hint(w, epi, self._wt((name, l), w, epi))
            self._hints[key] = h if h is not None else False
        return h if h is not False else None

    # ---- forward (one sequence; batching, KV buffers and graph replay live in B200CausalLM) ---------------
    def _layers(self, ids: torch.Tensor, pos: torch.Tensor, kv, extra=(), n_dev=None) -> torch.Tensor:
        q_len = ids.shape[0]
        H, Hkv, D = self.n_heads, self.n_kv_heads, self.head_dim
        N = kv.n_cached
        x = F.embedding(ids, self.embed_tokens)  # [q, hidden]: the residual stream, updated in place
        h = torch.empty_like(x)
        q_rot = torch.empty(H, q_len, D, dtype=torch.float16, device=self.device)
        attn_out = torch.empty(q_len, H * D, dtype=torch.float16, device=self.device)
        inter = self.config.intermediate_size
        act = torch.empty(q_len, inter, dtype=torch.float16, device=self.device)
        ws = self._workspace(q_len)
        # few tokens: the linears are weight streams -> pc_linear_small (tcgen05, fused RoPE+append / SiLU*up); the kernel
        # takes up to 32 tokens but only beats the library up to 16 (its 32-column variant has one pipeline stage less),
        # longer suffixes are compute-bound GEMMs: both stay in cuBLAS
        lin = self.linear_impl == "native" and q_len <= self.linear_max_q and D == 128 and self.config.hidden_size % 64 == 0 \
            and inter % 64 == 0
        lws = self._linear_workspace() if lin else None
        dbuf = torch.empty_like(x) if lin else None
        hid = self.config.hidden_size
        # o_proj / down_proj leave their split-K partials to the next pc_add_rmsnorm (no reduction tail in the GEMM)
        pbuf_o = pbuf_d = None
        if lin and self.linear_partials:
            s_o, s_d = nat.linear_partial_splits(hid, hid), nat.linear_partial_splits(hid, inter)
            if s_o >= 2:
                pbuf_o = torch.empty(s_o, q_len, hid, dtype=torch.float32, device=self.device)
            if s_d >= 2:
                pbuf_d = torch.empty(s_d, q_len, hid, dtype=torch.float32, device=self.device)
        which = self.linear_native_ops  # which linears run natively (A/B switch, PC_LINEAR_OPS)
        lin_qkv, lin_o, lin_gu, lin_down, lin_head = (lin and k in which for k in ("qkv", "o", "gu", "down", "head"))
        delta = None
        virt = kv.virtual  # cached rows still in the schema arena: attention reads them through the tile list
        if self.ablate:  # measurement only (bench.py: in-step cost of a kernel family = full step - step without it)
            return self._layers_ablated(x, h, pos, kv, q_rot, attn_out, act, ws, lws, pbuf_o, pbuf_d, n_dev)
        # chain mode: everything between two attention launches (o_proj, norm, gate/up, down, norm, next QKV) is recorded and
        # goes out as ONE persistent launch when the next non-chainable call (attention) or the end of the block flushes it
        chained = lin and self.linear_chain and self.linear_tiled and pbuf_o is not None and pbuf_d is not None \
            and lin_qkv and lin_o and lin_gu and lin_down and lin_head and not self.linear_prefetch
        with (nat.chain(lws) if chained else contextlib.nullcontext()):
            for l in range(self.n_layers):
                nat.add_rmsnorm(x, delta, self.ln1[l], h, self.eps)          # x += mlp(l-1); h = norm(x)
                # n_dev: run-time cached length (decode graph); N is then only the sizing upper bound
                if lin_qkv:
                    nat.linear_rope_append(h, self.wqkv[l], pos, self._cos, self._sin, q_rot, kv.k[l], kv.v[l],
                                           N if n_dev is None else n_dev, H, Hkv, lws,
                                           w_tiled=self._wt(("qkv", l), self.wqkv[l], nat.PC_LIN_ROPE_APPEND),
                                           nxt=self._hint("o", l, pbuf_o is not None) if lin_o else None)
                else:
                    qkv = F.linear(h, self.wqkv[l])                          # [q, (H+2Hkv)*D]
                    nat.rope_append(qkv[:, :H * D], qkv[:, H * D:(H + Hkv) * D], qkv[:, (H + Hkv) * D:], pos, self._cos,
                                    self._sin, q_rot, kv.k[l], kv.v[l], N if n_dev is None else n_dev, H, Hkv, D)
                nat.attention(q_rot, kv.k[l], kv.v[l], attn_out, N if n_dev is None else kv.max_ctx - q_len, q_len, H, Hkv,
                              D, self.scale, self.rounding, workspace=ws, decode=(q_len == 1), n_cached_dev=n_dev,
                              impl=self.attn_impl if (n_dev is None and virt is None) else nat.PC_ATTN_TCGEN05,
                              virtual=virt.planes(l) if virt is not None else None,
                              prefetch=self._wt(("o", l), self.wo[l], nat.PC_LIN_PARTIALS if pbuf_o is not None else nat.PC_LIN_PLAIN)
                              if (lin_o and self.attn_prefetch_o) else None)
                if lin_o and pbuf_o is not None:
                    nat.linear_partials(attn_out, self.wo[l], pbuf_o, lws, w_tiled=self._wt(("o", l), self.wo[l], nat.PC_LIN_PARTIALS),
                                        nxt=self._hint("gu", l, False) if lin_gu else None)
                    delta = pbuf_o
                elif lin_o:
                    nat.linear_small(attn_out, self.wo[l], dbuf, lws, w_tiled=self._wt(("o", l), self.wo[l], nat.PC_LIN_PLAIN),
                                     nxt=self._hint("gu", l, False) if lin_gu else None)
                    delta = dbuf
                else:
                    delta = F.linear(attn_out, self.wo[l])
                nat.add_rmsnorm(x, delta, self.ln2[l], h, self.eps)          # x += attn; h = norm(x)
                if lin_gu:
                    nat.linear_small(h, self.w_gate_up[l], act, lws, silu_mul=True,
                                     w_tiled=self._wt(("gu", l), self.w_gate_up[l], nat.PC_LIN_SILU_MUL),
                                     nxt=self._hint("down", l, pbuf_d is not None) if lin_down else None)
                else:
                    gu = F.linear(h, self.w_gate_up[l])
                    nat.silu_mul(gu, act, inter)
                if lin_down and pbuf_d is not None:
                    nxt_d = (self._hint("qkv", l + 1, False) if lin_qkv else None) if l + 1 < self.n_layers else \
                        (self._hint("head", 0, False) if lin_head else None)
                    nat.linear_partials(act, self.w_down[l], pbuf_d, lws, w_tiled=self._wt(("down", l), self.w_down[l], nat.PC_LIN_PARTIALS),
                                        nxt=nxt_d)
                    delta = pbuf_d
                elif lin_down:
                    nxt_d = (self._hint("qkv", l + 1, False) if lin_qkv else None) if l + 1 < self.n_layers else \
                        (self._hint("head", 0, False) if lin_head else None)
                    nat.linear_small(act, self.w_down[l], dbuf, lws, w_tiled=self._wt(("down", l), self.w_down[l], nat.PC_LIN_PLAIN), nxt=nxt_d)
                    delta = dbuf
                else:
                    delta = F.linear(act, self.w_down[l])
            nat.add_rmsnorm(x, delta, self.norm, h, self.eps)
            if lin_head:
                logits = torch.empty(q_len, self.vocab_size, dtype=torch.float16, device=self.device)
                nat.linear_small(h, self.lm_head, logits, lws, w_tiled=self._wt(("lm_head", 0), self.lm_head, nat.PC_LIN_PLAIN))
        if lin_head:
            return logits.float()
        return F.linear(h, self.lm_head).float()
