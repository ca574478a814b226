"""TEST INFRASTRUCTURE ONLY -- freeze outputs of the UNMODIFIED reference into tests/golden/.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py
Imports the reference through oracle/ref_shim.py, runs its own classes on CPU with seeded tiny models and
writes
  tests/golden/llama_tiny.pt     state dict, schema/prompt XML, plans (cached / no_cache / full ids), the spliced
                                 PromptCache contents, per-segment KV, prefill logits and greedy decode ids
  tests/golden/attn_cores.pt     inputs/outputs of the reference attention modules (Llama fp32+fp16, Falcon, MPT)
The GPU box has no /root/reference; tests read only these files.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# no whitespace between elements: a whitespace-only text run would become a zero-token TokenSequence, which the
# reference cannot encode (cache_engine.py:278 raises) -- see SURVEY.md section 7
SCHEMA = ('<schema name="demo"><module name="a">alpha beta gamma<parameter name="p" length="3"/>delta</module>'
          '<union><module name="b1">one two three four</module><module name="b2">five six</module></union>'
          '<module name="c">red green blue cyan</module></schema>')
PROMPT = """<prompt schema="demo"><a p="x y"/><b2/><c/>question words</prompt>"""
PROMPT2 = """<prompt schema="demo"><c/><b1/>other tail text here</prompt>"""


def llama_tiny():
    pc = ref_shim.load_reference()
    from promptcache import CacheEngine, GenerationEngine, GenerationParameters, Prompt
    cfg = ref_shim.tiny_llama_config()
    lm = ref_shim.make_reference_lm("llama", cfg, seed=0)
    eng = CacheEngine(64, lm, target_device="cpu")
    eng.add_schema(SCHEMA, max_tokens=None)
    sc = eng.schemas["demo"]
    segs = []
    for key, tsc in sc.cache_l1.items():
        segs.append({"offset": tsc.token_sequence.offset, "token_ids": list(tsc.token_sequence.token_ids()),
                     "kv": [(k.clone(), v.clone()) for k, v in tsc.host_cache]})
    gold = {"schema": SCHEMA, "prompt": PROMPT, "prompt2": PROMPT2, "state_dict": {k: v.clone() for k, v in lm.hf_model.state_dict().items()},
            "config": dict(layers=2, heads=4, head_dim=64, vocab=128, inter=512, max_pos=256, eps=1e-5),
            "segments": segs, "schema_len": len(sc.schema)}
    for tag, xml in (("p1", PROMPT), ("p2", PROMPT2)):
        # fresh engines per mode so usage counters start from zero, as on a first request
        for mode, kw in (("cached", {}), ("full", {"return_full_position_ids": True}), ("no_cache", {"no_cache": True})):
            e2 = CacheEngine(64, lm, target_device="cpu")
            e2.schemas = eng.schemas
            for t in sc.cache_l1.values():
                t.usage_counter = 0
            ids, pos, _, cache = e2.process(Prompt(xml), **kw)
            rec = {"ids": list(ids), "pos": list(pos)}
            if cache is not None:
                rec["cache"] = [(k.clone(), v.clone()) for k, v in cache]
                rec["order"] = [(m.token_sequence.offset, len(m)) for m in
                                sorted(e2.prompt_cache.staged, key=lambda e: e.usage_counter, reverse=True)]
            if mode == "cached":
                ge = GenerationEngine(lm)
                params = GenerationParameters(temperature=0.0, max_new_tokens=4, stop_token_ids=[], stop_str=[])
                new_ids = None
                # replicate generate() but record logits of the prefill
                out = lm(input_ids=torch.tensor([ids]), position_ids=torch.tensor([pos]),
                         past_key_values=[(k.unsqueeze(0), v.unsqueeze(0)) for k, v in cache], use_cache=True)
                rec["prefill_logits"] = out.logits.clone()
                texts = [o.new_text for o in ge.generate(ids, pos, params, cache, stream_interval=1)]
                rec["greedy_text"] = texts[-1]
                # the same decode steps with the ids kept: first new token is fed at max(pos)+1+1 (:82,132)
                past, nxt, gids = out.past_key_values, int(out.logits[0, -1].argmax()), []
                for i in range(1, 4):
                    gids.append(nxt)
                    o2 = lm(input_ids=torch.tensor([[nxt]]), position_ids=torch.tensor([[max(pos) + 1 + i]]),
                            past_key_values=past, use_cache=True)
                    past, nxt = o2.past_key_values, int(o2.logits[0, -1].argmax())
                gids.append(nxt)
                rec["greedy_ids"] = gids
                assert lm.decode(gids) == texts[-1], (lm.decode(gids), texts[-1])
            gold[f"{tag}_{mode}"] = rec
    torch.save(gold, os.path.join(OUT, "llama_tiny.pt"))
    print("llama_tiny.pt:", {k: (v["ids"], v["pos"]) for k, v in gold.items() if isinstance(v, dict) and "ids" in v})


def attn_cores():
    ref_shim.load_reference()
    from promptcache.model import llama2, falcon, mpt
    gold = {}
    torch.manual_seed(1)
    # ---- Llama: module-level run, fp32 and fp16 (CPU), discontiguous ids, past present
    cfg = ref_shim.tiny_llama_config(layers=1, heads=2, head_dim=128, vocab=32, max_pos=64)
    for dt, tag in ((torch.float32, "f32"), (torch.float16, "f16")):
        torch.manual_seed(2)
        attn = llama2.LlamaAttention(cfg).to(dt).eval()
        q_len, N = 5, 37
        hid = torch.randn(1, q_len, cfg.hidden_size).to(dt)
        pk = torch.randn(1, 2, N, 128).to(dt)
        pv = torch.randn(1, 2, N, 128).to(dt)
        pos = torch.tensor([[3, 4, 50, 51, 52]])
        mask = llama2._make_causal_mask((1, q_len), dt, hid.device, past_key_values_length=N)
        with torch.no_grad():
            out, _, (k, v) = attn(hid, attention_mask=mask, position_ids=pos, past_key_value=(pk, pv), use_cache=True)
        gold[f"llama_{tag}"] = dict(hidden=hid, past_k=pk, past_v=pv, pos=pos, out=out, k=k, v=v,
                                    weights={n: p.detach().clone() for n, p in attn.named_parameters()},
                                    cos=attn.rotary_emb.cos_cached[0, 0].clone(), sin=attn.rotary_emb.sin_cached[0, 0].clone())
    # ---- MPT alibi table (the one thing reference tests/test.py prints)
    gold["mpt_alibi_16_20"] = mpt.build_mpt_alibi_tensor(16, 20)
    gold["mpt_alibi_32_8"] = mpt.build_mpt_alibi_tensor(32, 8)
    torch.save(gold, os.path.join(OUT, "attn_cores.pt"))
    print("attn_cores.pt:", list(gold))


def falcon_mpt_tiny():
    """Reference Falcon (7B layout: MQA, parallel attn, rotary) and MPT (ALiBi) tiny models, fp32 on CPU: a prefix
    forward builds the past, then a suffix forward with discontiguous position ids runs over it."""
    ref_shim.load_reference()
    gold = {}
    prefix_ids = torch.tensor([[5, 9, 33, 77, 12, 8, 101, 64, 3, 90]])
    prefix_pos = torch.tensor([[0, 1, 2, 3, 10, 11, 12, 13, 14, 15]])
    suffix_ids = torch.tensor([[44, 45, 46, 47, 48]])
    suffix_pos = torch.tensor([[4, 5, 30, 31, 32]])
    # ---- Falcon
    cfg = ref_shim.tiny_falcon_config(layers=2, heads=2, head_dim=64, vocab=128, max_pos=256)
    lm = ref_shim.make_reference_lm("falcon", cfg, seed=5)
    with torch.no_grad():
        o1 = lm(input_ids=prefix_ids, position_ids=prefix_pos, use_cache=True)
        o2 = lm(input_ids=suffix_ids, position_ids=suffix_pos, past_key_values=o1.past_key_values, use_cache=True)
    gold["falcon"] = dict(state_dict={k: v.clone() for k, v in lm.hf_model.state_dict().items() if k != "lm_head.weight"},
                          config=dict(layers=2, heads=2, head_dim=64, vocab=128, eps=cfg.layer_norm_epsilon),
                          prefix_ids=prefix_ids, prefix_pos=prefix_pos, suffix_ids=suffix_ids, suffix_pos=suffix_pos,
                          prefix_logits=o1.logits.clone(), suffix_logits=o2.logits.clone(),
                          past=[(k.clone(), v.clone()) for k, v in o1.past_key_values],
                          present=[(k.clone(), v.clone()) for k, v in o2.past_key_values])
    # ---- MPT (full position ids: cached rows then new tokens)
    cfg = ref_shim.tiny_mpt_config(layers=2, heads=2, head_dim=64, vocab=128, max_seq_len=256)
    lm = ref_shim.make_reference_lm("mpt", cfg, seed=6)
    full_pos = torch.cat([prefix_pos, suffix_pos], dim=1)
    with torch.no_grad():
        o1 = lm.hf_model(prefix_ids, prefix_pos, use_cache=True)
        o2 = lm.hf_model(suffix_ids, full_pos, past_key_values=o1.past_key_values, use_cache=True)
    gold["mpt"] = dict(state_dict={k: v.clone() for k, v in lm.hf_model.state_dict().items() if k != "lm_head.weight"},
                       config=dict(layers=2, heads=2, head_dim=64, vocab=128, eps=cfg.layer_norm_epsilon),
                       prefix_ids=prefix_ids, prefix_pos=prefix_pos, suffix_ids=suffix_ids, full_pos=full_pos,
                       prefix_logits=o1.logits.clone(), suffix_logits=o2.logits.clone(),
                       past=[(k.clone(), v.clone()) for k, v in o1.past_key_values],
                       present=[(k.clone(), v.clone()) for k, v in o2.past_key_values])
    torch.save(gold, os.path.join(OUT, "falcon_mpt_tiny.pt"))
    print("falcon_mpt_tiny.pt:", {k: tuple(v["suffix_logits"].shape) for k, v in gold.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:] or ["llama", "cores", "falcon_mpt"]
    if "llama" in only:
        llama_tiny()
    if "cores" in only:
        attn_cores()
    if "falcon_mpt" in only:
        falcon_mpt_tiny()
