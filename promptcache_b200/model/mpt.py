"""MPT-7B style models (ALiBi, no rotary) on the B200 hot path.

Same forward contract as reference promptcache/model/mpt.py (MptForCausalLM.forward :654-720 -> MptModel.forward
:464-584 -> MptBlock :232-270 -> MptAttention :132-188).  ALiBi is the reference's variant: the bias of key j is
slope_h * (position_ids[j] - max(position_ids)) gathered at the KEY position ids, cached rows included
(mpt.py:97-110,172-175,522-523) -- so forward() takes the FULL id list (cached rows in buffer order, then the new
tokens), which is what CacheEngine.process(return_full_position_ids=True) returns.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from ._base import B200CausalLM


class MptConfigLite:
    def __init__(self, d_model=4096, n_heads=32, n_layers=32, vocab_size=50432, layer_norm_epsilon=1e-5,
                 alibi_bias_max=8, softmax_scale=None, initializer_range=0.02, max_seq_len=8192, expansion_ratio=4, **_):
        self.d_model = self.hidden_size = d_model
        self.n_heads = n_heads
        self.n_layers = n_layers
        self.vocab_size = vocab_size
        self.layer_norm_epsilon = layer_norm_epsilon
        self.alibi_bias_max = alibi_bias_max
        self.softmax_scale = softmax_scale
        self.initializer_range = initializer_range
        self.max_seq_len = max_seq_len
        self.expansion_ratio = expansion_ratio


def alibi_slopes(n_heads: int, alibi_bias_max: int = 8, device=None) -> torch.Tensor:
    """Per-head slopes of build_mpt_alibi_tensor (mpt.py:97-107): 2^-(i*bias_max/P), P = next power of two."""
    p2 = 2 ** math.ceil(math.log2(n_heads))
    base = torch.arange(1, p2 + 1, dtype=torch.float32, device=device) * (alibi_bias_max / p2)
    slopes = 1.0 / torch.pow(2, base)
    if p2 != n_heads:
        slopes = torch.cat([slopes[1::2], slopes[::2]])[:n_heads]
    return slopes.contiguous()


class MptForCausalLM(B200CausalLM):
    full_position_ids = True
    rotary = False

    def __init__(self, config, device="cuda", dtype=torch.float16):
        super().__init__()
        c = config
        self.config = c
        hid = getattr(c, "d_model", None) or c.hidden_size
        self.hidden = hid
        self.vocab_size = c.vocab_size
        self.n_layers = c.n_layers
        self.n_heads = self.n_kv_heads = c.n_heads
        self.head_dim = hid // c.n_heads
        self.eps = getattr(c, "layer_norm_epsilon", 1e-5)
        attn_cfg = getattr(c, "attn_config", None)
        bias_max = getattr(attn_cfg, "alibi_bias_max", None) if attn_cfg is not None else None
        self.alibi_bias_max = bias_max if bias_max is not None else getattr(c, "alibi_bias_max", 8)
        scale = getattr(attn_cfg, "softmax_scale", None) if attn_cfg is not None else getattr(c, "softmax_scale", None)
        self.scale = float(scale) if scale is not None else 1.0 / math.sqrt(self.head_dim)  # mpt.py:127-129
        self.inter = 4 * hid  # the reference hard-codes 4x (mpt.py:195-197), whatever expansion_ratio says
        std = getattr(c, "initializer_range", 0.02)

        def w(*shape):
            return nn.Parameter(torch.empty(*shape, device=device, dtype=dtype).normal_(0.0, std), requires_grad=False)

        ones = lambda: nn.Parameter(torch.ones(hid, device=device, dtype=dtype), requires_grad=False)
        self.wte = w(c.vocab_size, hid)  # lm_head is tied to it (mpt.py:599)
        self.norm1 = nn.ParameterList([ones() for _ in range(self.n_layers)])
        self.norm2 = nn.ParameterList([ones() for _ in range(self.n_layers)])
        self.wqkv = nn.ParameterList([w(3 * hid, hid) for _ in range(self.n_layers)])
        self.wo = nn.ParameterList([w(hid, hid) for _ in range(self.n_layers)])
        self.w_up = nn.ParameterList([w(self.inter, hid) for _ in range(self.n_layers)])
        self.w_down = nn.ParameterList([w(hid, self.inter) for _ in range(self.n_layers)])
        self.norm_f = ones()
        self.slopes = nn.Parameter(alibi_slopes(self.n_heads, self.alibi_bias_max, device), requires_grad=False)
        self._init_runtime()
        self.rounding = nat.PC_ROUND_MPT

    @torch.no_grad()
    def load_hf_state_dict(self, sd) -> None:
        self._weights_changed()
        g = lambda k: sd[k].to(self.wte.device, self.wte.dtype)
        self.wte.copy_(g("transformer.wte.weight"))
        self.norm_f.copy_(g("transformer.norm_f.weight"))
        for l in range(self.n_layers):
            p = f"transformer.blocks.{l}."
            self.norm1[l].copy_(g(p + "norm_1.weight"))
            self.norm2[l].copy_(g(p + "norm_2.weight"))
            self.wqkv[l].copy_(g(p + "attn.Wqkv.weight"))
            self.wo[l].copy_(g(p + "attn.out_proj.weight"))
            self.w_up[l].copy_(g(p + "ffn.up_proj.weight"))
            self.w_down[l].copy_(g(p + "ffn.down_proj.weight"))

    @classmethod
    def from_pretrained(cls, name: str, device="cuda", **kwargs):
        from transformers import AutoConfig, AutoModelForCausalLM
        kwargs.pop("load_in_8bit", None)
        kwargs.pop("device_map", None)
        kwargs.pop("max_seq_len", None)
        hf = AutoModelForCausalLM.from_pretrained(name, torch_dtype=torch.float16, **kwargs)
        model = cls(AutoConfig.from_pretrained(name), device=device)
        model.load_hf_state_dict(hf.state_dict())
        return model

    supports_dynamic_n = True  # decode graph: the key ids live in a max_ctx-long static buffer, N is read on the device
    virtual_kv_full_ids = True  # virtual splice with ALiBi: the key ids are gathered into tile order (VirtualKV.slot_rows)

    def virtual_kv_max_q(self) -> int:  # the short-suffix kernel reads segments in place: D=128, one query head per KV head
        return 32 if self.head_dim == 128 else 0

    def _virtual_ok(self, virt, q_len: int) -> bool:
        return virt.slot_rows is not None and 0 < q_len <= self.virtual_kv_max_q()

    def _layers(self, ids: torch.Tensor, pos: torch.Tensor, kv, extra=(), n_dev=None) -> torch.Tensor:
        q_len = ids.shape[0]
        H, D, hid = self.n_heads, self.head_dim, self.hidden
        N = kv.n_cached
        # ids of all N+q keys (cached rows in buffer order, then the new tokens), relative to max(position_ids)
        # (mpt.py:522: the table is built for max(position_ids)+1 positions), computed on the device so that neither a
        # host value nor a kernel parameter changes from step to step; with n_dev `pos` is the max_ctx-long static buffer
        # (zeros behind the live ids, which never exceed the maximum)
        key_pos = (pos - pos.max()).to(torch.int32)
        virt = kv.virtual  # cached rows still in the schema arena: attention walks them in TILE order, so do the ids --
        if virt is not None:  # [128 slots per virtual tile][rows appended behind the cached ones]
            key_pos = torch.cat([key_pos[virt.slot_rows], key_pos[virt.rows:]])
        x = F.embedding(ids, self.wte)
        ln = torch.empty_like(x)
        q_out = torch.empty(H, q_len, D, dtype=torch.float16, device=self.device)
        ctx = torch.empty(q_len, H * D, dtype=torch.float16, device=self.device)
        ws = self._workspace(q_len)
        nd = N if n_dev is None else n_dev  # run-time cached length (decode graph); N is then only the sizing upper bound
        # few tokens: the projections / MLP are weight streams -> pc_linear_small (no-rotary append in the QKV epilogue, GELU in the
        # up projection's, out_proj / down_proj as split-K partials the next add+LayerNorm reduces); otherwise cuBLAS
        lin = self.linear_impl == "native" and q_len <= self.linear_max_q and D == 128 and hid % 128 == 0
        if lin:
            lws = self._linear_workspace()
            act = torch.empty(q_len, self.inter, dtype=torch.float16, device=self.device)
            pbuf_o = torch.empty(nat.linear_partial_splits(hid, hid), q_len, hid, dtype=torch.float32, device=self.device)
            pbuf_d = torch.empty(nat.linear_partial_splits(hid, self.inter), q_len, hid, dtype=torch.float32, device=self.device)
        ffn = None
        for l in range(self.n_layers):
            # x += ffn of the previous block (:208) and norm_1, bias dropped (mpt.py:220-222): one launch
            nat.add_layernorm(x, ffn, None, self.norm1[l], None, ln, self.eps)
            if lin:
                nat.linear_rope_append(ln, self.wqkv[l], None, None, None, q_out, kv.k[l], kv.v[l], nd, H, H, lws,
                                       w_tiled=self._wt(("qkv", l), self.wqkv[l], nat.PC_LIN_ROPE_APPEND))  # q | k | v heads, no rotary
            else:
                qkv = F.linear(ln, self.wqkv[l])                                            # chunk(3): q | k | v (:142-143)
                nat.rope_append(qkv[:, :hid], qkv[:, hid:2 * hid], qkv[:, 2 * hid:], None, None, None, q_out, kv.k[l], kv.v[l],
                                nd, H, H, D)                                                # no rotary: transpose + append (:148-151)
            nat.attention(q_out, kv.k[l], kv.v[l], ctx, N if n_dev is None else kv.max_ctx - q_len, q_len, H, H, D, self.scale,
                          self.rounding, workspace=ws, alibi_slopes=self.slopes, key_pos=key_pos, alibi_max_pos=0,
                          impl=self.attn_impl if (n_dev is None and virt is None) else nat.PC_ATTN_TCGEN05, decode=(q_len == 1),
                          n_cached_dev=n_dev, virtual=virt.planes(l) if virt is not None else None)   # :156-185
            if lin:
                nat.linear_partials(ctx, self.wo[l], pbuf_o, lws, w_tiled=self._wt(("o", l), self.wo[l], nat.PC_LIN_PARTIALS))
                nat.add_layernorm(x, pbuf_o, None, self.norm2[l], None, ln, self.eps)       # x += attn (:251); norm_2
                nat.linear_small(ln, self.w_up[l], act, lws, gelu=True, w_tiled=self._wt(("up", l), self.w_up[l], nat.PC_LIN_PLAIN))
                nat.linear_partials(act, self.w_down[l], pbuf_d, lws, w_tiled=self._wt(("down", l), self.w_down[l], nat.PC_LIN_PARTIALS))
                ffn = pbuf_d                                                                # MptMLP (:200-208)
            else:
                nat.add_layernorm(x, F.linear(ctx, self.wo[l]), None, self.norm2[l], None, ln, self.eps)
                ffn = F.linear(F.gelu(F.linear(ln, self.w_up[l])), self.w_down[l])
        nat.add_layernorm(x, ffn, None, self.norm_f, None, ln, self.eps)
        if lin:
            logits = torch.empty(q_len, self.vocab_size, dtype=torch.float16, device=self.device)
            nat.linear_small(ln, self.wte, logits, lws, w_tiled=self._wt(("lm_head", 0), self.wte, nat.PC_LIN_PLAIN))
            return logits.float()
        return F.linear(ln, self.wte).float()
