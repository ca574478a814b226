"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (yale-sys/prompt-cache).

Imports ``promptcache`` from ``oracle/_ref`` (the byte-for-byte copy staged by ``oracle/make_ref.py``; git-ignored,
travels to the GPU box) or from ``/root/reference`` (build container only) through the shim set of SURVEY.md
section 8c, so that ``oracle/gen_golden.py`` can freeze the reference's outputs under ``tests/golden/``, the
``-m gpu`` tests can run the reference's own modules in fp16 on the B200 next to the CUDA path, and
``bench.py --impl reference`` can time the reference's own classes on the host cores.

Nothing under ``promptcache_b200/`` imports this file.  Only ``oracle/gen_golden.py`` and the
"oracle vs live reference" tests (skipped when ``/root/reference`` is absent) do.

Shims (the reference itself is never edited):
  * ``transformers.utils.is_flash_attn_available``      (promptcache/model/falcon.py:39)
  * stub modules ``lxml.etree`` / ``fire`` / ``termcolor`` (oracle/_shims)
  * ``torch.cuda.Event/synchronize/empty_cache`` on a host without a CUDA driver
    (promptcache/cache_engine.py:391-394,508; promptcache/generation_engine.py:104-113,209)
  * config attributes transformers 5.x dropped (promptcache/model/llama2.py:225,271,286)
"""
from __future__ import annotations

import os
import sys
import time
from typing import List, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "_shims")


def _find_reference() -> str:
    """PC_REFERENCE_ROOT, else the staged byte-for-byte copy oracle/_ref (made by oracle/make_ref.py; the only one
    that exists on the GPU box), else the read-only mount of the build container."""
    for cand in (os.environ.get("PC_REFERENCE_ROOT"), os.path.join(_HERE, "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "promptcache")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "promptcache"))


class _CpuEvent:
    """perf_counter stand-in for torch.cuda.Event on a CUDA-less host."""

    def __init__(self, enable_timing: bool = False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other) -> float:
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


_loaded = None


def load_reference():
    """Return the imported reference package ``promptcache`` (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import transformers.utils as tu

    if not hasattr(tu, "is_flash_attn_available"):
        tu.is_flash_attn_available = lambda: False
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.cuda.Event = _CpuEvent
        torch.cuda.synchronize = lambda *a, **k: None
        torch.cuda.empty_cache = lambda *a, **k: None
    import promptcache  # noqa: E402
    from promptcache.model import falcon as _falcon, mpt as _mpt

    _falcon.FalconForCausalLM._tied_weights_keys = {"lm_head.weight": "transformer.word_embeddings.weight"}
    _mpt.MptForCausalLM._tied_weights_keys = {"lm_head.weight": "transformer.wte.weight"}
    _falcon.FalconModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    _loaded = promptcache
    return promptcache


# ----------------------------------------------------------------------------------------------
# deterministic toy tokenizer shared by the golden generator, the oracle and the CUDA-path tests
# ----------------------------------------------------------------------------------------------
class WordTokenizer:
    """Whitespace tokenizer with a hash vocabulary (ids 4..vocab-1; 0=unk, 1=bos, 2=eos, 3=pad).

    Exposes what ``LanguageModel`` touches (promptcache/model/__init__.py:131-153).
    """

    unk_token = "<unk>"
    unk_token_id = 0
    eos_token = "</s>"
    eos_token_id = 2

    def __init__(self, vocab_size: int = 128):
        self.vocab_size = vocab_size
        self._rev = {}

    def _tid(self, w: str) -> int:
        h = 2166136261
        for ch in w.encode("utf-8"):
            h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
        t = 4 + h % (self.vocab_size - 4)
        self._rev.setdefault(t, w)
        return t

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:
        return [self._tid(w) for w in text.split()]

    def decode(self, ids, **kw) -> str:
        return " ".join(self._rev.get(int(i), f"<{int(i)}>") for i in ids)


def tiny_llama_config(layers=2, heads=4, head_dim=64, vocab=128, inter=None, max_pos=256):
    from transformers import LlamaConfig

    hidden = heads * head_dim
    cfg = LlamaConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter or 2 * hidden,
                      num_hidden_layers=layers, num_attention_heads=heads, num_key_value_heads=heads,
                      max_position_embeddings=max_pos, rms_norm_eps=1e-5)
    cfg.rope_theta = 10000.0
    cfg.rope_scaling = None
    cfg.pretraining_tp = 1
    return cfg


def tiny_falcon_config(layers=2, heads=4, head_dim=64, vocab=128, max_pos=256):
    from transformers import FalconConfig

    cfg = FalconConfig(vocab_size=vocab, hidden_size=heads * head_dim, num_hidden_layers=layers,
                       num_attention_heads=heads, multi_query=True, parallel_attn=True, bias=False,
                       alibi=False, new_decoder_architecture=False)
    cfg.rope_theta = 10000.0
    cfg.rope_scaling = None
    cfg.max_position_embeddings = max_pos
    return cfg


def tiny_mpt_config(layers=2, heads=4, head_dim=64, vocab=128, max_seq_len=256):
    from transformers import MptConfig

    cfg = MptConfig(d_model=heads * head_dim, n_heads=heads, n_layers=layers, vocab_size=vocab,
                    max_seq_len=max_seq_len, expansion_ratio=2)
    return cfg


def make_reference_lm(arch: str, cfg, seed: int = 0, dtype=torch.float32, tokenizer=None):
    """Reference ``LanguageModel`` around the reference's own patched HF model, seeded random init."""
    pc = load_reference()
    from promptcache.model import LanguageModel, llama2, falcon, mpt

    cls = {"llama": llama2.LlamaForCausalLM, "falcon": falcon.FalconForCausalLM, "mpt": mpt.MptForCausalLM}[arch]
    torch.manual_seed(seed)
    model = cls(cfg).to(dtype).eval()
    tok = tokenizer or WordTokenizer(cfg.vocab_size)

    class _LM(LanguageModel):
        def __init__(self):
            super().__init__(f"tiny-{arch}", model, tok, [tok.eos_token_id], [])
            self.use_full_position_ids = arch == "mpt"

        def get_formatter(self):
            return lambda s: s

        def get_cache_shape(self):
            c = model.config
            if arch == "falcon":
                return c.num_hidden_layers, 1, c.hidden_size // c.num_attention_heads
            if arch == "mpt":
                return c.n_layers, c.n_heads, c.d_model // c.n_heads
            return c.num_hidden_layers, c.num_attention_heads, c.hidden_size // c.num_attention_heads

    return _LM()
