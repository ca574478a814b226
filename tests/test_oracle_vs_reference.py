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
