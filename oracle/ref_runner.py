"""TEST INFRASTRUCTURE ONLY -- drive the UNMODIFIED reference (oracle/_ref or /root/reference) at real model dims.

Used by `bench.py --impl reference` / the `cpu_baseline` leg (reference's own CacheEngine.process ->
GenerationEngine.generate on the host cores, the timing eval_sys.py:327-359 takes) and by the `-m gpu` parity tests
that run the reference's models in fp16 on the B200 beside the CUDA path.  Nothing under `promptcache_b200/` imports it.
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch

from . import ref_shim


def hf_config(arch: str, dims: Dict):
    """transformers config object for the reference's vendored model files, with the attributes transformers 5 dropped
    restored (SURVEY.md section 8c shim 4/5)."""
    if arch == "llama":
        from transformers import LlamaConfig
        cfg = LlamaConfig(vocab_size=dims["vocab_size"], hidden_size=dims["hidden_size"],
                          intermediate_size=dims["intermediate_size"], num_hidden_layers=dims["num_hidden_layers"],
                          num_attention_heads=dims["num_attention_heads"],
                          num_key_value_heads=dims.get("num_key_value_heads") or dims["num_attention_heads"],
                          max_position_embeddings=dims.get("max_position_embeddings", 4096),
                          rms_norm_eps=dims.get("rms_norm_eps", 1e-5))
        cfg.rope_theta = 10000.0
        cfg.rope_scaling = None
        cfg.pretraining_tp = 1
        return cfg
    if arch == "falcon":
        from transformers import FalconConfig
        cfg = FalconConfig(vocab_size=dims["vocab_size"], hidden_size=dims["hidden_size"],
                           num_hidden_layers=dims["num_hidden_layers"], num_attention_heads=dims["num_attention_heads"],
                           multi_query=True, parallel_attn=True, bias=False, alibi=False, new_decoder_architecture=False)
        cfg.rope_theta = 10000.0
        cfg.rope_scaling = None
        cfg.max_position_embeddings = dims.get("max_position_embeddings", 2048)
        return cfg
    if arch == "mpt":
        from transformers import MptConfig
        return MptConfig(d_model=dims["hidden_size"], n_heads=dims["num_attention_heads"], n_layers=dims["num_hidden_layers"],
                         vocab_size=dims["vocab_size"], max_seq_len=dims.get("max_seq_len", 8192), expansion_ratio=4)
    raise ValueError(arch)


def build_reference_lm(arch: str, dims: Dict, tokenizer, dtype=torch.float32, device="cpu", seed: int = 0,
                       state_dict: Optional[dict] = None, init_std: float = 0.02):
    """The reference's own patched HF model wrapped in its own `LanguageModel`, random weights at the given dims.

    Weights: `state_dict` when given (so that the CUDA path and the reference hold the SAME weights), else seeded
    N(0, init_std) drawn directly in `dtype` on `device` -- HF's own init pass over 6.7 G fp32 parameters on the CPU takes
    minutes and adds nothing to a timing run."""
    ref_shim.load_reference()
    from promptcache.model import LanguageModel, falcon, llama2, mpt
    cls = {"llama": llama2.LlamaForCausalLM, "falcon": falcon.FalconForCausalLM, "mpt": mpt.MptForCausalLM}[arch]
    cfg = hf_config(arch, dims)
    try:
        from transformers.initialization import no_init_weights
    except ImportError:  # transformers 4.x
        from transformers.modeling_utils import no_init_weights
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with no_init_weights(), torch.device(device):
            model = cls(cfg)
    finally:
        torch.set_default_dtype(prev)
    model = model.eval()
    if getattr(cfg, "tie_word_embeddings", False):
        model.tie_weights()  # no_init_weights skipped post_init's tying: Falcon / MPT share lm_head with the embedding
    with torch.no_grad():
        if state_dict is not None:
            missing = model.load_state_dict({k: v.to(device=device, dtype=dtype) for k, v in state_dict.items()}, strict=False)
            bad = [k for k in missing.missing_keys if "rotary_emb" not in k and "lm_head" not in k]
            if bad:
                raise RuntimeError(f"reference model is missing weights: {bad[:4]}")
        else:
            g = torch.Generator(device=device).manual_seed(seed)
            # one seeded pool of N(0, init_std) values, tiled into the big matrices at rolling offsets: drawing 6.7 G
            # normals one by one on the CPU takes over a minute and the values do not matter to a dense GEMM's time
            pool = torch.empty(1 << 24, dtype=dtype, device=device).normal_(0.0, init_std, generator=g)
            off = 0
            for name, p in model.named_parameters():
                if p.dim() >= 2:
                    flat = p.view(-1)
                    done = 0
                    while done < flat.numel():
                        n = min(flat.numel() - done, pool.numel() - off)
                        flat[done:done + n].copy_(pool[off:off + n])
                        done += n
                        off = (off + n) % pool.numel() if n < pool.numel() - off else 0
                        off = (off + 7919) % (pool.numel() // 2)  # de-phase successive rows/matrices
                elif "bias" in name:
                    p.zero_()
                else:
                    p.fill_(1.0)
    tok = tokenizer

    class _LM(LanguageModel):
        def __init__(self):
            super().__init__(f"synthetic-{arch}", model, tok, [tok.eos_token_id], [])
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


class ReferenceTTFT:
    """The reference's own request path, timed as eval_sys.py:327-359 / demo.py:60-84 time it:
    TTFT = `cache_time` of CacheEngine.process (plan + PromptCache.update) + the first Output's prefill time of
    GenerationEngine.generate (one `lm()` call over the spliced cache), both from the reference's own CUDA-event /
    perf_counter stamps, plus wall clock around the two calls."""

    def __init__(self, lm, schema_xml: str, prompt_xml: str, max_ctx: int, device="cpu", synthetic_segments: bool = False):
        """`synthetic_segments`: skip the schema encode (one full-model forward over the whole schema; minutes on host
        cores and outside the TTFT) and fill the reference's own `TokenSequenceCache`s with seeded N(0,1) K/V of the shape
        and dtype its encoder would have stored.  Only for union-free schemas; the timed path is unchanged."""
        ref_shim.load_reference()
        from promptcache import CacheEngine, GenerationEngine, GenerationParameters, Prompt
        self.lm = lm
        self.engine = CacheEngine(max_ctx, lm, target_device=device)
        t0 = time.perf_counter()
        if synthetic_segments:
            from promptcache.cache_engine import TokenSequenceCache
            from promptcache.schema import Path
            self.engine.add_schema(schema_xml, no_cache=True)
            (sc,) = self.engine.schemas.values()
            if sc.schema.contains_union():
                raise ValueError("synthetic_segments needs a union-free schema")
            L, H, D = lm.get_cache_shape()
            g = torch.Generator().manual_seed(2002)
            dt = lm.hf_model.dtype
            for seq in sc.schema.get_scaffold(Path()).select(Path()).all_token_sequences():
                kv = [(torch.randn(H, len(seq), D, generator=g, dtype=dt), torch.randn(H, len(seq), D, generator=g, dtype=dt))
                      for _ in range(L)]
                sc.cache_l1[id(seq)] = TokenSequenceCache(seq, kv)
        else:
            self.engine.add_schema(schema_xml)
        self.encode_s = time.perf_counter() - t0
        self.prompt = Prompt(prompt_xml)
        self.gen = GenerationEngine(lm)
        self.params = GenerationParameters(temperature=0.0, max_new_tokens=1, stop_token_ids=[], stop_str=[], echo=False)
        self.full_ids = bool(getattr(lm, "use_full_position_ids", False))

    def step(self, invalidate: bool = True):
        if invalidate:  # every step splices every segment, as the CUDA arm does (retained-prefix shortcut off)
            self.engine.prompt_cache.staged = []
        t0 = time.perf_counter()
        ids, pos, cache_ms, cache = self.engine.process(self.prompt, return_full_position_ids=self.full_ids)
        t1 = time.perf_counter()
        first = next(iter(self.gen.generate(ids, pos, self.params, cache, stream_interval=1,
                                            use_full_position_ids=self.full_ids)))
        t2 = time.perf_counter()
        return dict(ttft_s=t2 - t0, process_s=t1 - t0, cache_ms=float(cache_ms), prefill_ms=float(first.elapsed_time),
                    n_cached=int(cache[0][0].shape[1]), q_len=len(ids), text=first.new_text)
