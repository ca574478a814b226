"""Shared test helpers: the deterministic toy tokenizer (same algorithm as oracle/ref_shim.WordTokenizer, restated
here so the GPU box needs nothing from the reference) and golden loading."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class WordTokenizer:
    unk_token = "<unk>"
    unk_token_id = 0
    eos_token = "</s>"
    eos_token_id = 2

    def __init__(self, vocab_size=128):
        self.vocab_size = vocab_size
        self._rev = {}

    def _tid(self, w):
        h = 2166136261
        for ch in w.encode("utf-8"):
            h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
        t = 4 + h % (self.vocab_size - 4)
        self._rev.setdefault(t, w)
        return t

    def encode(self, text, add_special_tokens=False):
        return [self._tid(w) for w in text.split()]

    def decode(self, ids, **kw):
        return " ".join(self._rev.get(int(i), f"<{int(i)}>") for i in ids)


class HostLM:
    """Tokenizer-only stand-in for LanguageModel: enough for Schema/Prompt/plan tests on CPU."""

    def __init__(self, vocab=128):
        self.tok = WordTokenizer(vocab)
        self.unk_token_id = 0
        self.eos_token_id = 2

    def encode(self, text):
        return self.tok.encode(text)

    def decode(self, ids):
        return self.tok.decode(ids)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)
