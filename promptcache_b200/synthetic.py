"""Synthetic schemas and a numeric tokenizer for benchmarks and tests.

There are no tokenizers, weights or datasets on disk (and no network), so the BASELINE configs are
realised as PML schemas whose text is made of pseudo-words `t<id>`; `IdTokenizer` maps `t123` <-> 123.
Shapes follow SURVEY.md section 8d (C1..C5).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple


class IdTokenizer:
    """`t<id>` <-> id.  Exposes what LanguageModel touches (reference model/__init__.py:131-153)."""

    unk_token = "t0"
    unk_token_id = 0
    eos_token = "t2"
    eos_token_id = 2

    def __init__(self, vocab_size: int = 32000):
        self.vocab_size = vocab_size

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:
        return [int(w[1:]) % self.vocab_size for w in text.split()]

    def decode(self, ids, **kw) -> str:
        return " ".join(f"t{int(i)}" for i in ids)


def words(rng: random.Random, n: int, vocab: int) -> str:
    return " ".join(f"t{rng.randrange(3, vocab)}" for _ in range(n))


@dataclass
class SyntheticCase:
    name: str
    schema_xml: str
    prompt_xml: str
    n_cached: int
    q_len: int
    n_segments: int


def flat_schema(name: str, seg_lens: Sequence[int], suffix: int, vocab: int, seed: int) -> SyntheticCase:
    """Union-free schema: one module per segment, all selected, `suffix` trailing tokens (C2, C3)."""
    rng = random.Random(seed)
    mods = "".join(f'<module name="m{i}">{words(rng, n, vocab)}</module>' for i, n in enumerate(seg_lens))
    schema = f'<schema name="{name}">{mods}</schema>'
    refs = "".join(f"<m{i}/>" for i in range(len(seg_lens)))
    prompt = f'<prompt schema="{name}">{refs}{words(rng, suffix, vocab)}</prompt>'
    return SyntheticCase(name, schema, prompt, sum(seg_lens), suffix, len(seg_lens))


def union_schema(name: str, head: int, slots: int, slot_width: int, member_len: int, tail: int, suffix: int,
                 vocab: int, seed: int) -> SyntheticCase:
    """C1 'persona': head + `slots` unions of width `slot_width` (two members each; the selected one is
    `member_len` long, leaving position gaps) + tail."""
    rng = random.Random(seed)
    parts = [f'<module name="head">{words(rng, head, vocab)}</module>']
    for s in range(slots):
        parts.append(f'<union><module name="s{s}a">{words(rng, slot_width, vocab)}</module>'
                     f'<module name="s{s}b">{words(rng, member_len, vocab)}</module></union>')
    parts.append(f'<module name="tail">{words(rng, tail, vocab)}</module>')
    schema = f'<schema name="{name}">{"".join(parts)}</schema>'
    refs = "<head/>" + "".join(f"<s{s}b/>" for s in range(slots)) + "<tail/>"
    prompt = f'<prompt schema="{name}">{refs}{words(rng, suffix, vocab)}</prompt>'
    return SyntheticCase(name, schema, prompt, head + slots * member_len + tail, suffix, slots + 2)


def parameter_schema(name: str, vocab: int, seed: int) -> SyntheticCase:
    """C4 'travel-like': 5 segments, N=256, parameters `duration` (len 5) and `city` (len 10) filled with 2 and
    3 tokens -> 5 interior-position suffix tokens + 16 trailing (examples/parameterized_prompts.xml:14,21).  The prompt
    names `city` before `duration`, so the suffix ids arrive in argument order, not position order: non-monotone."""
    rng = random.Random(seed)
    w = lambda n: words(rng, n, vocab)
    schema = (f'<schema name="{name}"><module name="sys">{w(40)}</module>'
              f'<module name="plan">{w(60)}<parameter name="duration" length="5"/>{w(70)}'
              f'<parameter name="city" length="10"/>{w(50)}</module><module name="end">{w(36)}</module></schema>')
    prompt = (f'<prompt schema="{name}"><sys/><plan city="{w(3)}" duration="{w(2)}"/><end/>{w(16)}</prompt>')
    return SyntheticCase(name, schema, prompt, 256, 21, 5)


def pool_schema(name: str, pool: int, seg_len: int, vocab: int, seed: int) -> Tuple[str, List[str]]:
    """C5: a pool of `pool` modules of `seg_len` tokens; prompts pick subsets."""
    rng = random.Random(seed)
    mods = "".join(f'<module name="m{i}">{words(rng, seg_len, vocab)}</module>' for i in range(pool))
    return f'<schema name="{name}">{mods}</schema>', [f"m{i}" for i in range(pool)]


def pool_prompt(name: str, picks: Sequence[str], suffix: int, vocab: int, seed: int) -> str:
    rng = random.Random(seed)
    return f'<prompt schema="{name}">{"".join(f"<{p}/>" for p in picks)}{words(rng, suffix, vocab)}</prompt>'


CONFIGS: Dict[str, dict] = {
    # BASELINE.json configs -> concrete shapes
    "C1": dict(model="llama2-7b", desc="persona: 8 segments [24,92x6,24], N=600, q=16, union gaps"),
    "C2": dict(model="llama2-7b", desc="code-generation-game: 7 segments [24,400x5,24], N=2048, q=16"),
    "C3": dict(model="llama2-13b", desc="LongBench multi-doc: 8 segments x512, N=4096, q=32"),
    "C4": dict(model="falcon-7b", desc="parameterized prompt: 5 segments, N=256, q=21 (5 interior ids)"),
    "C5": dict(model="llama2-7b", desc="32 segments x256 of a 64-pool, N=8192, q=32"),
}

MODEL_DIMS = {
    "llama2-7b": dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=32, max_position_embeddings=4096, rms_norm_eps=1e-5),
    "llama2-13b": dict(vocab_size=32000, hidden_size=5120, intermediate_size=13824, num_hidden_layers=40,
                       num_attention_heads=40, num_key_value_heads=40, max_position_embeddings=4096, rms_norm_eps=1e-5),
    # tiiuae/falcon-7b: multi_query, parallel_attn, rotary, no bias (== transformers FalconConfig() defaults)
    "falcon-7b": dict(vocab_size=65024, hidden_size=4544, num_hidden_layers=32, num_attention_heads=71,
                      layer_norm_epsilon=1e-5, max_position_embeddings=2048),
    # mosaicml/mpt-7b-chat with the reference's max_seq_len override (model/__init__.py:266)
    "mpt-7b": dict(vocab_size=50432, hidden_size=4096, num_hidden_layers=32, num_attention_heads=32, max_seq_len=8192),
}


def make_case(cfg: str, vocab: int = 32000) -> SyntheticCase:
    if cfg == "C1":
        return union_schema("persona", 24, 6, 128, 92, 24, 16, vocab, seed=1001)
    if cfg == "C2":
        return flat_schema("codegen", [24] + [400] * 5 + [24], 16, vocab, seed=1002)
    if cfg == "C3":
        return flat_schema("longbench", [512] * 8, 32, vocab, seed=1003)
    if cfg == "C4":
        return parameter_schema("travel", vocab, seed=1004)
    raise ValueError(f"unknown config {cfg}")
