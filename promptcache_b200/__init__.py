"""promptcache_b200 -- B200-native prompt-cache prefill engine behind the yale-sys/prompt-cache API.

    from promptcache_b200 import Prompt, CompactSpaces, read_file, CacheEngine, GenerationEngine, \
        GenerationParameters, llama2_template

mirrors `from promptcache import ...` (reference promptcache/__init__.py).  The hot path (splice, RoPE at
remapped ids + append, prefill/decode attention) runs in hand-written sm_100a kernels reached through the
C ABI in include/promptcache_b200.h; there is no CPU or eager fallback.
"""
from .prompt import Prompt, CompactSpaces, read_file, apply_preproc
from .schema import Schema
from .cache_engine import CacheEngine, PromptCache, SchemaCache, TokenSequenceCache
from .generation_engine import GenerationEngine, GenerationParameters, Output
from .model._kv import StaleKVError

__all__ = ["Prompt", "CompactSpaces", "read_file", "apply_preproc", "Schema", "CacheEngine", "PromptCache",
           "SchemaCache", "TokenSequenceCache", "GenerationEngine", "GenerationParameters", "Output", "StaleKVError"]
__version__ = "0.1.0"
