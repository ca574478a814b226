"""Live cross-check of the oracle and of the product front-end against the unmodified reference.  Runs only where
/root/reference is mounted (the build container); the GPU box relies on the goldens generated from it."""
import pytest
import torch

from oracle import pc_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted")

SCHEMA = ('<schema name="s2"><module name="sys">you are a helpful assistant</module>'
          '<module name="doc">alpha beta<parameter name="who" length="4" scaffold="dear user"/>gamma delta epsilon'
          '<union scaffold="x1"><module name="x1">one two three</module><module name="x2">four</module></union>zeta</module>'
          '<union><module name="u1">eta theta</module><module name="u2">iota kappa lambda mu</module></union>'
          '<module name="end">nu xi omicron</module></schema>')
PROMPTS = ['<prompt schema="s2"><sys/><doc who="bob"><x2/></doc><u2/><end/>what now</prompt>',
           '<prompt schema="s2"><end/><doc/>hello there friend</prompt>',
           '<prompt schema="s2"><sys/>just text</prompt>']


@pytest.fixture(scope="module")
def ref_engine():
    ref_shim.load_reference()
    from promptcache import CacheEngine
    lm = ref_shim.make_reference_lm("llama", ref_shim.tiny_llama_config(), seed=3)
    eng = CacheEngine(128, lm, target_device="cpu")
    eng.add_schema(SCHEMA)
    return eng, lm


@pytest.mark.parametrize("xml", PROMPTS)
@pytest.mark.parametrize("mode", ["cached", "full", "no_cache"])
def test_plan_equals_reference(ref_engine, xml, mode):
    from promptcache import Prompt as RefPrompt
    from promptcache_b200.cache_engine import CacheEngine, SchemaCache
    from promptcache_b200.prompt import Prompt
    from promptcache_b200.schema import Schema
    from tests.helpers import HostLM
    eng, lm = ref_engine
    for t in eng.schemas["s2"].cache_l1.values():
        t.usage_counter = 0
    eng.prompt_cache.staged = []
    kw = {"full": {"return_full_position_ids": True}, "no_cache": {"no_cache": True}}.get(mode, {})
    ids, pos, _, cache = eng.process(RefPrompt(xml), **kw)
    host = HostLM()
    ours = CacheEngine.__new__(CacheEngine)
    ours.lm = host
    ours.schemas = {"s2": SchemaCache(Schema(SCHEMA, host), host, no_cache=True, target_device="cpu")}
    _, used, oi, op, my_ids, my_pos = ours.plan(Prompt(xml), collect_original=True)
    if mode == "cached":
        assert (my_ids, my_pos) == (list(ids), list(pos))
        assert sum(len(s) for s in used) == cache[0][0].shape[1]
    elif mode == "full":
        assert [p for s in op for p in s] + my_pos == list(pos)
    else:
        pairs = sorted(zip([p for s in op for p in s] + my_pos, [t for s in oi for t in s] + my_ids))
        assert [t for _, t in pairs] == list(ids) and list(range(len(pairs))) == list(pos)


def test_oracle_forward_equals_reference_model(ref_engine):
    eng, lm = ref_engine
    sd = lm.hf_model.state_dict()
    cfg = dict(n_layers=2, n_heads=4, n_kv_heads=4, eps=1e-5, theta=10000.0, max_pos=256)
    ids = torch.tensor([[5, 9, 33, 77, 12, 8]])
    pos = torch.tensor([[0, 1, 2, 10, 11, 40]])
    with torch.no_grad():
        ref = lm(input_ids=ids, position_ids=pos, use_cache=True)
    logits, past = O.llama_forward(sd, cfg, ids, pos)
    assert torch.equal(logits, ref.logits)
    for l in range(2):
        assert torch.equal(past[l][0], ref.past_key_values[l][0]) and torch.equal(past[l][1], ref.past_key_values[l][1])


# ---- the reference's own example schemas (examples/*.xml; demo.py:54, eval_sys.py) --------------------------------
EXAMPLES = ["code_generation_bookstore.xml", "parameterized_prompts.xml", "persona_generation.xml", "personalization-education.xml"]


class _FmtLM:
    """Tokenizer-only LanguageModel stand-in with the Llama-2 conversation formatter of either code base."""

    unk_token_id, eos_token_id = 0, 2

    def __init__(self, fmt):
        self.tok, self.fmt = ref_shim.WordTokenizer(32000), fmt

    def encode(self, text):
        return self.tok.encode(text.decode() if isinstance(text, bytes) else text)

    def decode(self, ids):
        return self.tok.decode(ids)

    def get_formatter(self):
        return self.fmt


def _walk(node, out):
    """(kind, name, offset, length[, token ids]) of every node, depth first -- same traversal for both code bases."""
    for c in node.children:
        kind = type(c).__name__
        if kind == "TokenSequence":
            out.append(("T", c.offset, len(c), tuple(c.token_ids())))
        elif kind == "Parameter":
            out.append(("P", c.name, c.offset, c.length))
        elif kind == "UnionModule":
            out.append(("U", c.offset, len(c), c.scaffold_name))
            for m in c.modules:
                out.append(("M", m.name, m.offset, len(m)))
                _walk(m, out)
        else:
            out.append(("M", c.name, c.offset, len(c)))
            _walk(c, out)
    return out


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("max_tokens", [None, 800])  # demo.py:54 truncates modules to 800 tokens
def test_example_schema_layout_equals_reference(name, max_tokens):
    """Module tree, offsets, lengths, token ids, parameter slots and the default scaffold of the reference's example
    schemas: identical in both front-ends (the position ids a schema assigns are what the splice and RoPE consume)."""
    import os
    ref_shim.load_reference()
    from promptcache.model import FormatLlama2Conversation as RefFmt
    from promptcache.schema import Schema as RefSchema
    from promptcache_b200.model import FormatLlama2Conversation
    from promptcache_b200.schema import Schema
    text = open(os.path.join(ref_shim.REFERENCE_ROOT, "examples", name)).read()
    # demo.py:46-54: the model's conversation formatter is applied to the file, then the schema is parsed
    ref_text, our_text = RefFmt()(text), FormatLlama2Conversation()(text)
    assert ref_text == our_text
    ref = RefSchema(ref_text, _FmtLM(RefFmt()), max_tokens=max_tokens)
    ours = Schema(our_text, _FmtLM(FormatLlama2Conversation()), max_tokens=max_tokens)
    assert ours.name == ref.name and len(ours) == len(ref)
    a, b = _walk(ref, []), _walk(ours, [])
    assert sum(1 for n in a if n[0] == "M") >= 2 and a == b
    # the un-preprocessed file (conversation tags kept as literal text, schema.py:362-363) must agree as well
    assert _walk(RefSchema(text, _FmtLM(None)), []) == _walk(Schema(text, _FmtLM(None)), [])
    rs, os_ = ref.get_scaffold(), ours.get_scaffold()
    assert list(rs.token_ids()) == list(os_.token_ids()) and list(rs.position_ids()) == list(os_.position_ids())


def test_malformed_example_is_parsed_tolerantly():
    """examples/code_generation_game.xml (the schema demo.py loads) holds raw `<` in code and needs lxml's recover
    mode (schema.py:285); the stdlib parser behind the reference shim rejects it, so only OUR tolerant front-end can be
    exercised here: it must parse, keep every module and assign contiguous offsets."""
    import os
    from promptcache_b200.model import FormatLlama2Conversation
    from promptcache_b200.schema import Schema
    text = open(os.path.join(ref_shim.REFERENCE_ROOT, "examples", "code_generation_game.xml")).read()
    s = Schema(FormatLlama2Conversation()(text), _FmtLM(None), max_tokens=800)
    nodes = _walk(s, [])
    assert s.name == "code-generation-game" and len(s) > 300
    assert sum(1 for n in nodes if n[0] == "M") >= 5
    pos = s.get_scaffold().position_ids()
    assert list(pos) == sorted(pos) and len(set(pos)) == len(pos)
