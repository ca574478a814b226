"""Llama-2 / CodeLlama on the B200 hot path.

Same forward signature and outputs as reference promptcache/model/llama2.py (LlamaForCausalLM.forward
:986-1076 -> LlamaModel.forward :822-951 -> LlamaDecoderLayer :600-654 -> LlamaAttention :315-410).
The per-layer attention core (llama2.py:353-398: RoPE at explicit ids, KV concat, repeat_kv, mask, QK^T,
softmax, PV, transpose) is TWO kernel launches here -- pc_rope_append and pc_prefill_attn/pc_decode_attn --
called through the C ABI; projections, norms and the MLP stay torch/cuBLAS.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from ._kv import CausalLMOutput, resolve_kv


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


def build_rope_tables(head_dim: int, rows: int, theta: float, device, dtype=torch.float16):
    """cos/sin tables as LlamaRotaryEmbedding._set_cos_sin_cache builds them (llama2.py:128-137)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, device=device).float() / head_dim))
    t = torch.arange(rows, device=device, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype).contiguous(), emb.sin().to(dtype).contiguous()


class LlamaForCausalLM(nn.Module):
    """fp16 Llama with fused QKV / gate-up weights; attention runs in the sm_100a kernels."""

    accepts_max_position = True  # forward() takes the host-known max position id (avoids a device sync)

    def __init__(self, config, device="cuda", dtype=torch.float16):
        super().__init__()
        self.config = config
        c = config
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
        self._cos = self._sin = None
        self._ws = None
        # CUDA graphs of the whole single-sequence forward, keyed by (q_len, n_cached, KV buffer): the per-layer
        # launch sequence is fixed once those are, so a repeated request shape is replayed with one launch
        self.use_cuda_graphs = True
        self._graphs = {}
        self._graph_seen = {}
        self.max_graphs = 8
        self.attn_impl = nat.PC_ATTN_AUTO
        self.rounding = nat.PC_ROUND_LLAMA
        self.scale = nat.softmax_scale(self.head_dim)

    # ---- weights ------------------------------------------------------------------------------------
    @torch.no_grad()
    def load_hf_state_dict(self, sd) -> None:
        """Load a HuggingFace Llama state dict (model.layers.N.self_attn.q_proj.weight ...)."""
        g = lambda k: sd[k].to(self.embed_tokens.device, self.embed_tokens.dtype)
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

    @property
    def device(self) -> torch.device:
        return self.embed_tokens.device

    @property
    def dtype(self) -> torch.dtype:
        return self.embed_tokens.dtype

    # ---- RoPE tables ----------------------------------------------------------------------------------
    def ensure_rope(self, rows: int) -> None:
        """Grow-on-demand like llama2.py:141-142, but driven by host-known ids (no device sync)."""
        rows = max(int(rows), int(getattr(self.config, "max_position_embeddings", 2048)))
        if self._cos is None or self._cos.shape[0] < rows or self._cos.device != self.device:
            rows = ((rows + 1023) // 1024) * 1024
            self._graphs.clear()  # captured graphs hold the old table pointers
            self._cos, self._sin = build_rope_tables(self.head_dim, rows, self.theta, self.device)

    def _workspace(self, q_len: int) -> Optional[torch.Tensor]:
        need = nat.attn_workspace_bytes(self.n_heads, q_len, self.head_dim)
        if need == 0:
            return None
        if self._ws is None or self._ws.numel() < need or self._ws.device != self.device:
            # zero-initialised once: the first bytes hold split-KV arrival counters the kernels leave zeroed
            self._graphs.clear()  # captured graphs hold the old workspace pointer
            self._ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- forward --------------------------------------------------------------------------------------
    def _forward_one(self, ids: torch.Tensor, pos: torch.Tensor, past, use_cache: bool):
        q_len = ids.shape[0]
        kv = resolve_kv(past, q_len, self.n_layers, self.n_kv_heads, self.head_dim, self.device)
        N = kv.n_cached
        key = (q_len, N, kv.k[0].data_ptr(), kv.max_ctx)
        if self.use_cuda_graphs and N > 0 and q_len > 1 and not torch.cuda.is_current_stream_capturing():
            entry = self._graphs.get(key)
            if entry is None and self._graph_seen.get(key, 0) >= 1:
                entry = self._capture(key, ids, pos, kv)
            self._graph_seen[key] = self._graph_seen.get(key, 0) + 1
            if entry is not None:
                graph, s_ids, s_pos, s_logits, n_native = entry
                s_ids.copy_(ids, non_blocking=True)
                s_pos.copy_(pos, non_blocking=True)
                graph.replay()
                nat.note_replayed_launches(n_native)
                return s_logits.clone(), (kv.views(N + q_len) if use_cache else None)
        logits = self._layers(ids, pos, kv)
        return logits, (kv.views(N + q_len) if use_cache else None)

    def _capture(self, key, ids, pos, kv):
        """Record the forward for this (q_len, n_cached, KV buffer) once; later calls replay it."""
        if len(self._graphs) >= self.max_graphs:
            self._graphs.pop(next(iter(self._graphs)))
        s_ids, s_pos = ids.clone(), pos.clone()
        self._layers(s_ids, s_pos, kv)  # warm-up on the capture inputs (function attributes, workspaces, cuBLAS)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before = nat.launch_count()
        with torch.cuda.graph(graph):
            s_logits = self._layers(s_ids, s_pos, kv)
        self._graphs[key] = (graph, s_ids, s_pos, s_logits, nat.launch_count() - before)
        return self._graphs[key]

    def _layers(self, ids: torch.Tensor, pos: torch.Tensor, kv) -> torch.Tensor:
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
        delta = None
        for l in range(self.n_layers):
            nat.add_rmsnorm(x, delta, self.ln1[l], h, self.eps)          # x += mlp(l-1); h = norm(x)
            qkv = F.linear(h, self.wqkv[l])                              # [q, (H+2Hkv)*D]
            nat.rope_append(qkv[:, :H * D], qkv[:, H * D:(H + Hkv) * D], qkv[:, (H + Hkv) * D:], pos, self._cos,
                            self._sin, q_rot, kv.k[l], kv.v[l], N, H, Hkv, D)
            nat.attention(q_rot, kv.k[l], kv.v[l], attn_out, N, q_len, H, Hkv, D, self.scale, self.rounding,
                          workspace=ws, impl=self.attn_impl, decode=(q_len == 1))
            delta = F.linear(attn_out, self.wo[l])
            nat.add_rmsnorm(x, delta, self.ln2[l], h, self.eps)          # x += attn; h = norm(x)
            gu = F.linear(h, self.w_gate_up[l])
            nat.silu_mul(gu, act, inter)
            delta = F.linear(act, self.w_down[l])
        nat.add_rmsnorm(x, delta, self.norm, h, self.eps)
        return F.linear(h, self.lm_head).float()

    @torch.inference_mode()
    def forward(self, input_ids: torch.Tensor = None, attention_mask: Optional[torch.Tensor] = None,
                position_ids: Optional[torch.Tensor] = None, past_key_values=None, inputs_embeds=None, labels=None,
                use_cache: Optional[bool] = None, output_attentions=None, output_hidden_states=None,
                return_dict=None, max_position: Optional[int] = None, **_):
        if input_ids is None:
            raise ValueError("promptcache_b200 Llama needs input_ids")
        if not input_ids.is_cuda:
            raise nat.NativeError("promptcache_b200 models run on CUDA only (no CPU fallback)")
        use_cache = True if use_cache is None else use_cache
        bsz, q_len = input_ids.shape
        past_len = 0 if not past_key_values else int(past_key_values[0][0].shape[-2])
        if position_ids is None:
            position_ids = torch.arange(past_len, past_len + q_len, device=input_ids.device).unsqueeze(0).expand(bsz, -1)
        position_ids = position_ids.view(-1, q_len).long()
        if max_position is None:
            max_position = int(position_ids.max())  # device sync; GenerationEngine passes the host value instead
        self.ensure_rope(max_position + 1)
        if bsz == 1:
            logits, pkv = self._forward_one(input_ids[0], position_ids[0].contiguous(), past_key_values, use_cache)
            return CausalLMOutput(logits.unsqueeze(0), pkv)
        # batch > 1 (schema encoding with right padding, cache_engine.py:240-248): rows are independent
        if past_key_values:
            raise ValueError("batched forward with past_key_values is not supported")
        lens = [q_len] * bsz if attention_mask is None else [int(v) for v in attention_mask.to(torch.int64).sum(-1).tolist()]
        logits = torch.zeros(bsz, q_len, self.config.vocab_size, dtype=torch.float32, device=input_ids.device)
        pk = [[None] * bsz for _ in range(self.n_layers)]
        pv = [[None] * bsz for _ in range(self.n_layers)]
        for b in range(bsz):
            n = max(lens[b], 1)
            lg, pkv = self._forward_one(input_ids[b, :n], position_ids[b, :n].contiguous(), None, True)
            logits[b, :n] = lg
            for l, (k, v) in enumerate(pkv):
                pk[l][b] = F.pad(k[0], (0, 0, 0, q_len - n))
                pv[l][b] = F.pad(v[0], (0, 0, 0, q_len - n))
        pkv = tuple((torch.stack(pk[l]), torch.stack(pv[l])) for l in range(self.n_layers)) if use_cache else None
        return CausalLMOutput(logits, pkv)
