"""Falcon-7B style models (multi-query attention, parallel attention+MLP block, rotary) on the B200 hot path.

Same forward contract as reference promptcache/model/falcon.py (FalconForCausalLM.forward :1276-1341 ->
FalconModel.forward :1040-1178 -> FalconDecoderLayer :767-817 -> FalconAttention :416-498).  The attention core
(falcon.py:443-491: rotary gathered at position_ids, KV concat, additive float-min mask, SDPA with one KV head
broadcast to 71 query heads) is pc_rope_append + pc_prefill_attn/pc_decode_attn: the 71 heads are FOLDED into
the MMA's M dimension (71*q rows share each K/V tile), scores stay fp32 (PC_ROUND_FALCON).
Only the 7B layout is supported: multi_query, parallel_attn, no alibi, old decoder architecture.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from ._base import B200CausalLM


class FalconConfigLite:
    def __init__(self, vocab_size=65024, hidden_size=4544, num_hidden_layers=32, num_attention_heads=71,
                 layer_norm_epsilon=1e-5, initializer_range=0.02, multi_query=True, parallel_attn=True, bias=False,
                 alibi=False, new_decoder_architecture=False, rope_theta=10000.0, max_position_embeddings=2048, **_):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.layer_norm_epsilon = layer_norm_epsilon
        self.initializer_range = initializer_range
        self.multi_query = multi_query
        self.parallel_attn = parallel_attn
        self.bias = bias
        self.alibi = alibi
        self.new_decoder_architecture = new_decoder_architecture
        self.rope_theta = rope_theta
        self.max_position_embeddings = max_position_embeddings


class FalconForCausalLM(B200CausalLM):
    def __init__(self, config, device="cuda", dtype=torch.float16):
        super().__init__()
        c = config
        if getattr(c, "new_decoder_architecture", False) or getattr(c, "alibi", False) or not getattr(c, "multi_query", True) \
                or not getattr(c, "parallel_attn", True):
            raise NotImplementedError("promptcache_b200 Falcon supports the 7B layout: multi_query, parallel_attn, rotary")
        self.config = c
        self.vocab_size = c.vocab_size
        self.n_layers = c.num_hidden_layers
        self.n_heads = c.num_attention_heads
        self.n_kv_heads = 1
        self.head_dim = c.hidden_size // c.num_attention_heads
        self.eps = c.layer_norm_epsilon
        self.theta = float(getattr(c, "rope_theta", 10000.0))
        self.has_bias = bool(getattr(c, "bias", False))
        hid = c.hidden_size
        std = getattr(c, "initializer_range", 0.02)

        def w(*shape):
            return nn.Parameter(torch.empty(*shape, device=device, dtype=dtype).normal_(0.0, std), requires_grad=False)

        def vec(val):
            return nn.Parameter(torch.full((hid,), val, device=device, dtype=dtype), requires_grad=False)

        self.word_embeddings = w(c.vocab_size, hid)  # lm_head is tied to it (falcon.py:1262)
        self.ln_w = nn.ParameterList([vec(1.0) for _ in range(self.n_layers)])
        self.ln_b = nn.ParameterList([vec(0.0) for _ in range(self.n_layers)])
        self.wqkv = nn.ParameterList([w((self.n_heads + 2) * self.head_dim, hid) for _ in range(self.n_layers)])
        self.wdense = nn.ParameterList([w(hid, hid) for _ in range(self.n_layers)])
        self.w_h4h = nn.ParameterList([w(4 * hid, hid) for _ in range(self.n_layers)])
        self.w_4hh = nn.ParameterList([w(hid, 4 * hid) for _ in range(self.n_layers)])
        self.lnf_w, self.lnf_b = vec(1.0), vec(0.0)
        self._init_runtime()
        # native skinny linears up to this many tokens (the kernel takes up to 32 token columns).  Measured on C4 (B200):
        # decode 3.34 ms/token native vs 3.19 cuBLAS, 21-token prefill 3.65 vs 3.29 ms -- Falcon's shapes (37 / 36 row tiles,
        # K = 71 k-blocks) partition worse than Llama's, so the default stays cuBLAS (0); tests switch it on.
        self.linear_max_q_falcon = min(32, int(os.environ.get("PC_LINEAR_MAX_Q_FALCON", "0")))
        self.rounding = nat.PC_ROUND_FALCON
        self.scale = nat.softmax_scale(self.head_dim)

    @torch.no_grad()
    def load_hf_state_dict(self, sd) -> None:
        self._weights_changed()
        g = lambda k: sd[k].to(self.word_embeddings.device, self.word_embeddings.dtype)
        self.word_embeddings.copy_(g("transformer.word_embeddings.weight"))
        self.lnf_w.copy_(g("transformer.ln_f.weight"))
        self.lnf_b.copy_(g("transformer.ln_f.bias"))
        for l in range(self.n_layers):
            p = f"transformer.h.{l}."
            self.ln_w[l].copy_(g(p + "input_layernorm.weight"))
            self.ln_b[l].copy_(g(p + "input_layernorm.bias"))
            self.wqkv[l].copy_(g(p + "self_attention.query_key_value.weight"))
            self.wdense[l].copy_(g(p + "self_attention.dense.weight"))
            self.w_h4h[l].copy_(g(p + "mlp.dense_h_to_4h.weight"))
            self.w_4hh[l].copy_(g(p + "mlp.dense_4h_to_h.weight"))

    @classmethod
    def from_pretrained(cls, name: str, device="cuda", **kwargs):
        from transformers import AutoConfig, AutoModelForCausalLM
        kwargs.pop("load_in_8bit", None)
        kwargs.pop("device_map", None)
        hf = AutoModelForCausalLM.from_pretrained(name, torch_dtype=torch.float16, **kwargs)
        model = cls(AutoConfig.from_pretrained(name), device=device)
        model.load_hf_state_dict(hf.state_dict())
        return model

    supports_dynamic_n = True  # decode graph: pc_rope_append_dyn + attention with a run-time cached length

    def _launch_shape_key(self):
        return super()._launch_shape_key() + (self.linear_max_q_falcon,)

    def _layers(self, ids: torch.Tensor, pos: torch.Tensor, kv, extra=(), n_dev=None) -> torch.Tensor:
        q_len = ids.shape[0]
        H, D = self.n_heads, self.head_dim
        N = kv.n_cached
        hid = self.config.hidden_size
        x = F.embedding(ids, self.word_embeddings)
        ln = torch.empty_like(x)
        q_rot = torch.empty(H, q_len, D, dtype=torch.float16, device=self.device)
        ctx = torch.empty(q_len, H * D, dtype=torch.float16, device=self.device)
        ws = self._workspace(q_len)
        nd = N if n_dev is None else n_dev  # run-time cached length (decode graph); N is then only the sizing upper bound
        # few tokens: the block's four matrices are weight streams -> pc_linear_small (RoPE at head_dim 64 + MQA append in the
        # fused QKV epilogue, GELU in h_to_4h's, dense / 4h_to_h as split-K partials that the next add+LayerNorm reduces);
        # the pre-tiled copies pad K = 4544 (71 blocks of 64) with one block of zeros.  Longer suffixes: cuBLAS.
        lin = self.linear_impl == "native" and q_len <= self.linear_max_q_falcon and D in (64, 128) and hid % 64 == 0 and not self.has_bias
        if lin:
            lws = self._linear_workspace()
            act = torch.empty(q_len, 4 * hid, dtype=torch.float16, device=self.device)
            pbuf_a = torch.empty(nat.linear_partial_splits(hid, hid), q_len, hid, dtype=torch.float32, device=self.device)
            pbuf_m = torch.empty(nat.linear_partial_splits(hid, 4 * hid), q_len, hid, dtype=torch.float32, device=self.device)
        attn = mlp = None
        for l in range(self.n_layers):
            # x = x + (mlp += attn) of the previous block (falcon.py:807-810) and this block's input_layernorm (:781): one launch
            nat.add_layernorm(x, mlp, attn, self.ln_w[l], self.ln_b[l], ln, self.eps)
            if lin:
                nat.linear_rope_append(ln, self.wqkv[l], pos, self._cos, self._sin, q_rot, kv.k[l], kv.v[l], nd, H, 1, lws,
                                       w_tiled=self._wt(("qkv", l), self.wqkv[l], nat.PC_LIN_ROPE_APPEND), head_dim=D)
            else:
                qkv = F.linear(ln, self.wqkv[l])                                      # [q, (H+2)*D]: q heads | k | v (:385-388)
                nat.rope_append(qkv[:, :H * D], qkv[:, H * D:(H + 1) * D], qkv[:, (H + 1) * D:], pos, self._cos, self._sin,
                                q_rot, kv.k[l], kv.v[l], nd, H, 1, D)
            nat.attention(q_rot, kv.k[l], kv.v[l], ctx, N if n_dev is None else kv.max_ctx - q_len, q_len, H, 1, D, self.scale,
                          self.rounding, workspace=ws, impl=self.attn_impl if n_dev is None else nat.PC_ATTN_TCGEN05,
                          decode=(q_len == 1), n_cached_dev=n_dev)
            if lin:
                nat.linear_partials(ctx, self.wdense[l], pbuf_a, lws, w_tiled=self._wt(("dense", l), self.wdense[l], nat.PC_LIN_PARTIALS))
                nat.linear_small(ln, self.w_h4h[l], act, lws, gelu=True, w_tiled=self._wt(("h4h", l), self.w_h4h[l], nat.PC_LIN_PLAIN))
                nat.linear_partials(act, self.w_4hh[l], pbuf_m, lws, w_tiled=self._wt(("4hh", l), self.w_4hh[l], nat.PC_LIN_PARTIALS))
                attn, mlp = pbuf_a, pbuf_m
            else:
                attn = F.linear(ctx, self.wdense[l])                                  # :493
                mlp = F.linear(F.gelu(F.linear(ln, self.w_h4h[l])), self.w_4hh[l])    # parallel MLP on the same norm (:796-805)
        nat.add_layernorm(x, mlp, attn, self.lnf_w, self.lnf_b, ln, self.eps)          # last block's residual adds + ln_f
        if lin:
            logits = torch.empty(q_len, self.vocab_size, dtype=torch.float16, device=self.device)
            nat.linear_small(ln, self.word_embeddings, logits, lws,
                             w_tiled=self._wt(("lm_head", 0), self.word_embeddings, nat.PC_LIN_PLAIN))
            return logits.float()
        return F.linear(ln, self.word_embeddings).float()
