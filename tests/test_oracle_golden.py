"""Pin the oracle: its restatement must reproduce what the reference's own classes produced (tests/golden/,
written by oracle/gen_golden.py from /root/reference).  CPU only."""
import pytest
import torch
import torch.nn.functional as F

from oracle import pc_oracle as O
from tests.helpers import load_golden


@pytest.fixture(scope="module")
def gold():
    return load_golden("llama_tiny.pt")


@pytest.fixture(scope="module")
def cores():
    return load_golden("attn_cores.pt")


def _cfg(g):
    c = g["config"]
    return dict(n_layers=c["layers"], n_heads=c["heads"], n_kv_heads=c["heads"], eps=c["eps"], theta=10000.0,
                max_pos=c["max_pos"])


@pytest.mark.parametrize("tag", ["p1", "p2"])
def test_splice_reproduces_promptcache(gold, tag):
    """oracle.splice on the reference's segments == the reference's PromptCache contents, bit for bit."""
    rec = gold[f"{tag}_cached"]
    by_key = {(s["offset"], len(s["token_ids"])): s["kv"] for s in gold["segments"]}
    segs = [by_key[tuple(o)] for o in rec["order"]]
    out, n = O.splice(segs, [1] * len(segs), 64)
    assert n == rec["cache"][0][0].shape[1]
    for l, (k, v) in enumerate(rec["cache"]):
        assert torch.equal(out[l][0], k) and torch.equal(out[l][1], v)
        assert out[l][0].dtype == torch.float16  # the reference hard-codes an fp16 cache (cache_engine.py:105-106)


def test_splice_order_is_stable_descending():
    assert O.splice_order([1, 5, 1, 5, 0]) == [1, 3, 0, 2, 4]


def test_llama_forward_reproduces_encode_and_prefill(gold):
    """Scaffold encode -> per-segment KV, then cached prefill logits and greedy ids, all from the oracle."""
    sd, cfg = gold["state_dict"], _cfg(gold)
    # default scaffold of the demo schema: a(0..2) p(3..5 unk) delta(6) b1? no default -> union skipped, c(11..14)
    from promptcache_b200.cache_engine import SchemaCache
    from promptcache_b200.schema import Schema
    from tests.helpers import HostLM
    lm = HostLM()
    schema = Schema(gold["schema"], lm)
    sc = SchemaCache(schema, lm, no_cache=True, target_device="cpu")
    seg_by_key = {(s["offset"], len(s["token_ids"])): s for s in gold["segments"]}
    for path in sc.scaffold_paths():
        scaf = schema.get_scaffold(path)
        ids, pos = scaf.token_ids(), scaf.position_ids()
        _, past = O.llama_forward(sd, cfg, torch.tensor([ids]), torch.tensor([pos]))
        for seq in scaf.select(path).all_token_sequences():
            st = pos.index(seq.offset)
            ref = seg_by_key[(seq.offset, len(seq))]["kv"]
            for l in range(cfg["n_layers"]):
                assert torch.equal(past[l][0][0, :, st:st + len(seq)], ref[l][0])
                assert torch.equal(past[l][1][0, :, st:st + len(seq)], ref[l][1])
    rec = gold["p1_cached"]
    past = [(k.unsqueeze(0), v.unsqueeze(0)) for k, v in rec["cache"]]
    logits, new_past = O.llama_forward(sd, cfg, torch.tensor([rec["ids"]]), torch.tensor([rec["pos"]]), past)
    assert torch.equal(logits, rec["prefill_logits"])
    nxt, got = int(logits[0, -1].argmax()), []
    for i in range(1, 4):
        got.append(nxt)
        lg, new_past = O.llama_forward(sd, cfg, torch.tensor([[nxt]]), torch.tensor([[max(rec["pos"]) + 1 + i]]), new_past)
        nxt = int(lg[0, -1].argmax())
    got.append(nxt)
    assert got == rec["greedy_ids"]


@pytest.mark.parametrize("tag", ["f32", "f16"])
def test_llama_attention_core_matches_reference_module(cores, tag):
    g = cores[f"llama_{tag}"]
    w, hid = g["weights"], g["hidden"]
    H, D = 2, 128
    q = F.linear(hid, w["q_proj.weight"]).view(1, -1, H, D).transpose(1, 2)
    k = F.linear(hid, w["k_proj.weight"]).view(1, -1, H, D).transpose(1, 2)
    v = F.linear(hid, w["v_proj.weight"]).view(1, -1, H, D).transpose(1, 2)
    cos, sin = g["cos"].to(hid.dtype), g["sin"].to(hid.dtype)
    ctx, kk, vv = O.llama_attention_core(q, k, v, g["past_k"], g["past_v"], cos, sin, g["pos"])
    out = F.linear(ctx, w["o_proj.weight"])
    assert torch.equal(kk, g["k"]) and torch.equal(vv, g["v"])
    assert torch.equal(out, g["out"])
    # and the table builder restates LlamaRotaryEmbedding
    c2, s2 = O.rope_tables(D, g["cos"].shape[0], 10000.0, "cpu", g["cos"].dtype)
    assert torch.equal(c2, g["cos"]) and torch.equal(s2, g["sin"])


def test_mpt_alibi_matches_reference(cores):
    assert torch.equal(O.mpt_alibi(16, 20), cores["mpt_alibi_16_20"])
    assert torch.equal(O.mpt_alibi(32, 8), cores["mpt_alibi_32_8"])
    sl = O.mpt_alibi_slopes(32)
    assert torch.allclose(sl[:, None, None] * torch.arange(-7, 1).view(1, 1, -1), cores["mpt_alibi_32_8"])


def test_rope_three_roundings_is_what_the_reference_does():
    """half(half(x*c) + half(rot(x)*s)) -- the identity pc_rope_append relies on for bit-exactness."""
    torch.manual_seed(0)
    q = torch.randn(1, 2, 7, 64).half()
    cos, sin = O.rope_tables(64, 100, dtype=torch.float16)
    pos = torch.tensor([[3, 9, 50, 51, 98, 0, 7]])
    ref, _ = O.apply_rotary_pos_emb(q, q, cos, sin, pos)
    c, s = cos[pos].unsqueeze(1).float(), sin[pos].unsqueeze(1).float()
    a = (q.float() * c).half().float()
    b = (O.rotate_half(q).float() * s).half().float()
    assert torch.equal((a + b).half(), ref)


def _fm_cfg(g):
    c = g["config"]
    return dict(n_layers=c["layers"], n_heads=c["heads"], eps=c["eps"], theta=10000.0)


def test_falcon_forward_reproduces_reference():
    """Oracle Falcon (MQA, parallel block, rotary at discontiguous ids, SDPA) vs the reference model, fp32 CPU."""
    g = load_golden("falcon_mpt_tiny.pt")["falcon"]
    sd, cfg = g["state_dict"], _fm_cfg(g)
    lg1, past = O.falcon_forward(sd, cfg, g["prefix_ids"], g["prefix_pos"])
    assert torch.equal(lg1, g["prefix_logits"])
    for l in range(2):
        assert torch.equal(past[l][0], g["past"][l][0]) and torch.equal(past[l][1], g["past"][l][1])
    lg2, present = O.falcon_forward(sd, cfg, g["suffix_ids"], g["suffix_pos"], g["past"])
    assert torch.equal(lg2, g["suffix_logits"])
    for l in range(2):
        assert torch.equal(present[l][0], g["present"][l][0]) and torch.equal(present[l][1], g["present"][l][1])


def test_mpt_forward_reproduces_reference():
    g = load_golden("falcon_mpt_tiny.pt")["mpt"]
    sd, cfg = g["state_dict"], _fm_cfg(g)
    lg1, past = O.mpt_forward(sd, cfg, g["prefix_ids"], g["prefix_pos"])
    assert torch.equal(lg1, g["prefix_logits"])
    lg2, present = O.mpt_forward(sd, cfg, g["suffix_ids"], g["full_pos"], g["past"])
    assert torch.equal(lg2, g["suffix_logits"])
    for l in range(2):
        assert torch.equal(present[l][0], g["present"][l][0]) and torch.equal(present[l][1], g["present"][l][1])
