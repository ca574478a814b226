"""Model adapters: the `LanguageModel` interface of reference promptcache/model/__init__.py:90-295.

Same attributes and hooks (`encode/decode`, `get_cache_shape`, `store_*_hook/read_*_hook`,
`get_formatter`, `use_full_position_ids`, `stop_token_ids/stop_str`); `hf_model` is one of this
package's B200 models whose attention core runs in the sm_100a kernels.
"""
from __future__ import annotations

import abc
import re
from typing import Callable, List, Optional, Tuple

import torch

from ..prompt import Preprocessor, PreprocessorList, escape_xml
from .llama2 import LlamaConfigLite, LlamaForCausalLM


class FormatConversation(Preprocessor):
    """Replace <system>/<user>/<assistant> tags with a model's chat template (model/__init__.py:48-77)."""

    def __init__(self, system: Tuple[str, str, str], user: Tuple[str, str], assistant: Tuple[str, str]):
        self.system = tuple(escape_xml(s) for s in system)
        self.user = tuple(escape_xml(s) for s in user)
        self.assistant = tuple(escape_xml(s) for s in assistant)

    def __call__(self, prompt: str) -> str:
        prompt = re.sub(r" +<system>", "<system>", prompt)
        table = [("<system>", self.system[0]), ("</system>", self.system[1]), ("<system/>", self.system[2]),
                 ("<user>", self.user[0]), ("</user>", self.user[1]),
                 ("<assistant>", self.assistant[0]), ("</assistant>", self.assistant[1])]
        for old, new in table:
            prompt = prompt.replace(old, new)
        return prompt


class FormatLlama2Conversation(FormatConversation):
    def __init__(self):
        super().__init__(system=("<s>[INST] <<SYS>>\n", "<</SYS>>\n\n", "<s>[INST]"), user=(" ", "[/INST]"),
                         assistant=(" ", "</s><s>[INST]"))


class LanguageModel(abc.ABC):
    name: str
    use_full_position_ids: bool = False

    def __init__(self, name: str, model, tokenizer, stop_token_ids: Optional[List[int]] = None,
                 stop_str: Optional[List[str]] = None):
        self.name = name
        self.hf_tokenizer = tokenizer
        self.hf_model = model
        self.stop_token_ids = stop_token_ids if stop_token_ids is not None else [self.eos_token_id]
        self.stop_str = stop_str if stop_str is not None else []

    @abc.abstractmethod
    def get_formatter(self) -> Callable[[str], str]:
        ...

    def get_cache_shape(self) -> Tuple[int, int, int]:
        c = self.config
        return c.num_hidden_layers, c.num_attention_heads, c.hidden_size // c.num_attention_heads

    # identity hooks, as in the reference (model/__init__.py:116-126)
    def store_k_hook(self, k_cache: torch.Tensor) -> torch.Tensor:
        return k_cache

    def store_v_hook(self, v_cache: torch.Tensor) -> torch.Tensor:
        return v_cache

    def read_k_hook(self, k_cache: torch.Tensor) -> torch.Tensor:
        return k_cache

    def read_v_hook(self, v_cache: torch.Tensor) -> torch.Tensor:
        return v_cache

    def __call__(self, **kwargs):
        return self.hf_model(**kwargs)

    def encode(self, text: str) -> List[int]:
        return self.hf_tokenizer.encode(text, add_special_tokens=False)

    def decode(self, token_ids: List[int]) -> str:
        return self.hf_tokenizer.decode(token_ids, skip_special_tokens=False, spaces_between_special_tokens=False)

    @property
    def unk_token(self) -> str:
        return self.hf_tokenizer.unk_token

    @property
    def unk_token_id(self) -> int:
        return self.hf_tokenizer.unk_token_id

    @property
    def eos_token(self) -> str:
        return self.hf_tokenizer.eos_token

    @property
    def eos_token_id(self) -> int:
        return self.hf_tokenizer.eos_token_id

    @property
    def device(self) -> torch.device:
        return self.hf_model.device

    @property
    def config(self):
        return self.hf_model.config


_LLAMA_CHAT = dict(system=("<s> [INST] <<SYS>>\n", "<</SYS>>\n\n", "<s> [INST] "), user=("", "[/INST]"),
                   assistant=("", "</s><s> [INST] "))


class Llama2(LanguageModel):
    """model/__init__.py:185-203.  Pass `model=`/`tokenizer=` to wrap an already-built model (tests, bench)."""

    def __init__(self, name: str = "meta-llama/Llama-2-7b-chat-hf", model=None, tokenizer=None, **kwargs):
        if tokenizer is None:
            from transformers import LlamaTokenizer
            tokenizer = LlamaTokenizer.from_pretrained(name)
        if model is None:
            model = LlamaForCausalLM.from_pretrained(name, **kwargs)
        self.formatter = FormatConversation(**_LLAMA_CHAT)
        super().__init__(name, model, tokenizer, [tokenizer.eos_token_id], ["</s>"])

    def get_formatter(self) -> Callable[[str], str]:
        return self.formatter

    def get_cache_shape(self) -> Tuple[int, int, int]:
        """(L, KV heads, head_dim).  The reference returns num_attention_heads (model/__init__.py:110-114), which equals
        the KV-head count for every Llama-2 7B/13B checkpoint it was run on; for grouped-query checkpoints (Llama-2-70B,
        CodeLlama-34B) the cached K/V have num_key_value_heads heads, which is what the prompt cache must be sized for."""
        c = self.config
        hkv = getattr(c, "num_key_value_heads", None) or c.num_attention_heads
        return c.num_hidden_layers, hkv, c.hidden_size // c.num_attention_heads


class CodeLlama(Llama2):
    """model/__init__.py:164-182."""

    def __init__(self, name: str = "codellama/CodeLlama-13b-Instruct-hf", model=None, tokenizer=None, **kwargs):
        if tokenizer is None:
            from transformers import CodeLlamaTokenizer
            tokenizer = CodeLlamaTokenizer.from_pretrained(name)
        super().__init__(name, model=model, tokenizer=tokenizer, **kwargs)


_FALCON_STOP_IDS = list(range(12))


class Falcon(LanguageModel):
    """model/__init__.py:206-258 (MQA: cache shape (L, 1, head_dim))."""

    def __init__(self, name: str = "tiiuae/falcon-7b-instruct", model=None, tokenizer=None, **kwargs):
        from .falcon import FalconForCausalLM
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(name)
        if model is None:
            model = FalconForCausalLM.from_pretrained(name, **kwargs)
        rep = lambda prompt: prompt.replace("\r\n", "\n").replace("\n\n", "\n")
        conv = FormatConversation(system=("", "\n\n", ""), user=("User: ", "\n\nAssistant:"), assistant=(" ", "\n\n"))
        self.formatter = PreprocessorList([rep, conv])
        super().__init__(name, model, tokenizer, list(_FALCON_STOP_IDS), ["<|endoftext|>", "\nUser"])

    def get_formatter(self) -> Callable[[str], str]:
        return self.formatter

    def get_cache_shape(self) -> Tuple[int, int, int]:
        c = self.hf_model.config
        return c.num_hidden_layers, 1, c.hidden_size // c.num_attention_heads


class Mpt(LanguageModel):
    """model/__init__.py:261-288 (ALiBi: needs the cached + new position ids)."""

    def __init__(self, name: str = "mosaicml/mpt-7b-chat", model=None, tokenizer=None, **kwargs):
        from .mpt import MptForCausalLM
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(name, trust_remote_code=True)
        if model is None:
            model = MptForCausalLM.from_pretrained(name, max_seq_len=8192, **kwargs)
        self.formatter = FormatConversation(system=("<|im_start|>system\n", "<|im_end|>\n", ""),
                                            user=("<|im_start|>user\n", "<|im_end|>\n<|im_start|>assistant\n"),
                                            assistant=("", "<|im_end|>\n"))
        super().__init__(name, model, tokenizer, [50278, 0], [])
        self.use_full_position_ids = True

    def get_formatter(self) -> Callable[[str], str]:
        return self.formatter

    def get_cache_shape(self) -> Tuple[int, int, int]:
        c = self.hf_model.config
        return c.n_layers, c.n_heads, c.d_model // c.n_heads


__all__ = ["LanguageModel", "Llama2", "CodeLlama", "Falcon", "Mpt", "FormatConversation",
           "FormatLlama2Conversation", "LlamaForCausalLM", "LlamaConfigLite"]
