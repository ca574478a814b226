"""Schema encoding (SchemaCache._process, reference cache_engine.py:186-307): union members are encoded as BATCHED cached
prefills over the default scaffold's prefix and must store the same KV as the reference's way of doing
it -- one whole-scaffold forward per member."""
import pytest
import torch

from oracle import pc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


def _lm(heads=4, kv_heads=4, hidden=512, vocab=512, layers=3):
    from promptcache_b200.model import Llama2, LlamaConfigLite, LlamaForCausalLM
    from promptcache_b200.synthetic import IdTokenizer
    sd = O.random_llama_state_dict(vocab, hidden, 2 * hidden, layers, heads, kv_heads, seed=4, dtype=torch.float16, device=DEV)
    cfg = LlamaConfigLite(vocab_size=vocab, hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=layers,
                          num_attention_heads=heads, num_key_value_heads=kv_heads, max_position_embeddings=2048)
    model = LlamaForCausalLM(cfg, device=DEV)
    model.load_hf_state_dict(sd)
    return Llama2("enc", model=model, tokenizer=IdTokenizer(vocab))


def _persona(vocab=512):
    """C1-like: head + 5 unions (3 members of different lengths each, one nested in a module) + tail."""
    import random
    from promptcache_b200.synthetic import words
    rng = random.Random(11)
    w = lambda n: words(rng, n, vocab)
    parts = [f'<module name="head">{w(24)}</module>']
    for s, lens in enumerate([(40, 92, 17), (64, 33, 128), (5, 70, 70), (92, 92, 10), (130, 3, 60)]):
        members = "".join(f'<module name="s{s}m{i}">{w(n)}</module>' for i, n in enumerate(lens))
        parts.append(f"<union>{members}</union>")
    parts.append(f'<module name="tail">{w(24)}<parameter name="p" length="4"/>{w(9)}</module>')
    return f'<schema name="persona">{"".join(parts)}</schema>'


@pytest.mark.parametrize("batch_size", [1, 3])
def test_batched_member_encode_equals_whole_scaffold_encode(native, batch_size):
    from promptcache_b200 import Schema
    from promptcache_b200.cache_engine import SchemaCache
    lm = _lm()
    xml = _persona()
    launches0 = native.launch_count()
    batched = SchemaCache(Schema(xml, lm), lm, batch_size=batch_size)
    launches_batched = native.launch_count() - launches0
    SchemaCache.reuse_prefix = False
    try:
        launches0 = native.launch_count()
        whole = SchemaCache(Schema(xml, lm), lm)  # the reference's way: the whole scaffold again for every member
        launches_whole = native.launch_count() - launches0
    finally:
        SchemaCache.reuse_prefix = True
    a = sorted(batched.cache_l1.values(), key=lambda e: (e.token_sequence.offset, len(e), e.row0))
    b = sorted(whole.cache_l1.values(), key=lambda e: (e.token_sequence.offset, len(e), e.row0))
    assert len(a) == len(b) == 1 + 15 + 2  # head, 5 x 3 members, tail's two text runs
    assert [(e.token_sequence.offset, len(e)) for e in a] == [(e.token_sequence.offset, len(e)) for e in b]
    for ea, eb in zip(a, b):
        assert ea.arena is batched.arena and ea.packed.shape == eb.packed.shape
        assert _rel(ea.packed, eb.packed) <= 2e-3, (ea.token_sequence.offset, _rel(ea.packed, eb.packed))
    assert launches_batched < (launches_whole / 2 if batch_size == 1 else launches_whole), (launches_batched, launches_whole)


def test_forward_many_equals_separate_forwards(native):
    """forward_many([...]) == one forward per sequence: same KV rows (bit-exact for K/V of layer 0, whose inputs do not
    depend on how the GEMM was batched; tolerance deeper down where cuBLAS may pick another kernel for another M)."""
    lm = _lm(heads=8, kv_heads=2, hidden=1024)
    m = lm.hf_model
    g = torch.Generator().manual_seed(3)
    pre_ids = torch.randint(3, 512, (1, 200), generator=g).to(DEV)
    pre = m(input_ids=pre_ids, position_ids=torch.arange(200, device=DEV).unsqueeze(0), use_cache=True).past_key_values
    items, want = [], []
    for n, first, pos0 in [(37, 50, 300), (128, 200, 200), (5, 10, 10), (260, 120, 900)]:
        ids = torch.randint(3, 512, (n,), generator=g).to(DEV)
        pos = torch.arange(pos0, pos0 + n, device=DEV)
        past = [(k[:, :, :first], v[:, :, :first]) for k, v in pre]
        items.append((ids, pos, past))
        want.append(m(input_ids=ids.unsqueeze(0), position_ids=pos.unsqueeze(0), past_key_values=[(k.clone(), v.clone()) for k, v in past],
                      use_cache=True).past_key_values)
    got = m.forward_many(items)
    for kv_g, kv_w, (ids, _, past) in zip(got, want, items):
        first, n = past[0][0].shape[2], ids.shape[0]
        for l in range(3):
            assert kv_g[l][0].shape == kv_w[l][0].shape == (1, 2, first + n, 128)
            assert torch.equal(kv_g[l][0][:, :, :first], kv_w[l][0][:, :, :first])
            assert _rel(kv_g[l][0], kv_w[l][0]) <= 2e-3 and _rel(kv_g[l][1], kv_w[l][1]) <= 2e-3
