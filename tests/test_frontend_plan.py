"""Host logic on CPU: PML front-end and the request plan against goldens produced by the reference itself
(oracle/gen_golden.py ran promptcache.CacheEngine.process on the same XML)."""
import pytest

from promptcache_b200.cache_engine import CacheEngine, SchemaCache, pad_batch
from promptcache_b200.prompt import CompactSpaces, Prompt, compact_surrounding_spaces, loads_xml
from promptcache_b200.schema import Path, Schema
from tests.helpers import HostLM, load_golden


@pytest.fixture(scope="module")
def gold():
    return load_golden("llama_tiny.pt")


def _engine(gold):
    lm = HostLM()
    eng = CacheEngine.__new__(CacheEngine)  # no PromptCache (that one lives in HBM); plan() is pure host logic
    eng.lm = lm
    eng.schemas = {"demo": SchemaCache(Schema(gold["schema"], lm), lm, no_cache=True, target_device="cpu")}
    return eng


def test_schema_offsets_match_reference(gold):
    schema = Schema(gold["schema"], HostLM())
    assert len(schema) == gold["schema_len"]
    ours = sorted((s.offset, s.token_ids()) for p in SchemaCache(schema, HostLM(), no_cache=True, target_device="cpu").scaffold_paths()
                  for s in schema.get_scaffold(p).select(p).all_token_sequences())
    ref = sorted((s["offset"], s["token_ids"]) for s in gold["segments"])
    assert ours == ref


@pytest.mark.parametrize("tag", ["p1", "p2"])
def test_plan_matches_reference(gold, tag):
    eng = _engine(gold)
    prompt = Prompt(gold["prompt" if tag == "p1" else "prompt2"])
    cached, used, orig_ids, orig_pos, ids, pos = eng.plan(prompt, collect_original=True)
    assert ids == gold[f"{tag}_cached"]["ids"] and pos == gold[f"{tag}_cached"]["pos"]
    # visit order of cached sequences == the reference's row order on a first request
    assert [(s.offset, len(s)) for s in used] == [tuple(o) for o in gold[f"{tag}_cached"]["order"]]
    full = [p for seq in orig_pos for p in seq] + pos
    assert full == gold[f"{tag}_full"]["pos"]
    pairs = sorted(zip([p for s in orig_pos for p in s] + pos, [t for s in orig_ids for t in s] + ids))
    assert [t for _, t in pairs] == gold[f"{tag}_no_cache"]["ids"]
    assert list(range(len(pairs))) == gold[f"{tag}_no_cache"]["pos"]


def test_plan_errors(gold):
    eng = _engine(gold)
    with pytest.raises(ValueError, match="no such layout"):
        eng.plan(Prompt('<prompt schema="nope"><a/>x</prompt>'), False)
    with pytest.raises(ValueError, match="no such module"):
        eng.plan(Prompt('<prompt schema="demo"><zzz/>x</prompt>'), False)
    with pytest.raises(ValueError, match="no such parameter"):
        eng.plan(Prompt('<prompt schema="demo"><a q="1"/>x</prompt>'), False)
    with pytest.raises(ValueError, match="too long"):
        eng.plan(Prompt('<prompt schema="demo"><a p="1 2 3 4"/>x</prompt>'), False)


def test_union_positions_are_shared_and_discontiguous():
    lm = HostLM()
    s = Schema('<schema name="u"><module name="h">a b</module><union scaffold="m1"><module name="m1">c d e f g</module>'
               '<module name="m2">h i</module></union><module name="t">z</module></schema>', lm)
    assert len(s) == 2 + 5 + 1
    assert s.select("m1").offset == s.select("m2").offset == 2
    sc = SchemaCache(s, lm, no_cache=True, target_device="cpu")
    assert [str(p) for p in sc.scaffold_paths()] == ["", "m2"]
    assert s.get_scaffold(Path("m2")).position_ids() == [0, 1, 2, 3, 7]  # gap 4..6


def test_nested_union_paths_follow_reference_order():
    lm = HostLM()
    xml = ('<schema name="n"><module name="o">x<union scaffold="i1"><module name="i1">p q</module><module name="i2">r</module>'
           '</union></module><union><module name="w1">s</module><module name="w2">t u</module></union></schema>')
    sc = SchemaCache(Schema(xml, lm), lm, no_cache=True, target_device="cpu")
    assert [str(p) for p in sc.scaffold_paths()] == ["", "w1", "w2", "o/i2"]


def test_parameter_and_scaffold_tokens():
    lm = HostLM()
    s = Schema('<schema name="p"><module name="m">hi<parameter name="x" length="4" scaffold="a b"/>yo</module></schema>', lm)
    par = s.select("m").parameters()[0]
    assert par.position_ids() == [1, 2, 3, 4] and par.token_ids()[2:] == [0, 0]
    with pytest.raises(ValueError):
        Schema('<schema name="p"><parameter name="x" length="2"/></schema>', lm)
    with pytest.raises(ValueError):
        Schema('<schema name="p"><module name="m"/><module name="m"/></schema>', lm)


def test_max_tokens_head_tail():
    lm = HostLM()
    s = Schema('<schema name="t"><module name="m">a b c d e f g h</module></schema>', lm, max_tokens=4)
    full = lm.encode("a b c d e f g h")
    assert s.select("m").token_ids() == full[:2] + full[-2:]


def test_prompt_parsing():
    p = Prompt('<prompt schema="s"><a x="1"/><b><c/></b> tail text </prompt>', [CompactSpaces()])
    assert p.schema == "s" and [m.name for m in p.modules] == ["a", "b"] and p.modules[1].modules[0].name == "c"
    assert p.text == "tail text" and p.modules[0].args[0].name == "x"
    with pytest.raises(ValueError):
        Prompt('<prompt schema="s">lead<a/></prompt>')
    with pytest.raises(ValueError):
        Prompt('<prompt schema="s"><a>text</a></prompt>')
    assert compact_surrounding_spaces("\n  a  b \n") == " a  b "
    # tolerant parse of raw '<' and '&' in text (lxml recover=True in the reference, schema.py:285)
    e = loads_xml('<schema name="x"><module name="m">if a < b && c: pass</module></schema>')
    assert "a < b && c" in e[0].text
    assert pad_batch([[1, 2, 3], [4]], 0) == ([[1, 2, 3], [4, 0, 0]], [[1, 1, 1], [1, 0, 0]])
