"""Falcon (MQA, rotary, SDPA rounding) and MPT (ALiBi at key ids, full position ids) adapters on the GPU against the
oracle in fp16 on the same device and against goldens produced by the reference's own models (fp32, CPU)."""
import pytest
import torch

from oracle import pc_oracle as O
from tests.helpers import WordTokenizer, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))


@pytest.fixture(scope="module")
def gold():
    return load_golden("falcon_mpt_tiny.pt")


def _falcon(g):
    from promptcache_b200.model.falcon import FalconConfigLite, FalconForCausalLM
    c = g["config"]
    m = FalconForCausalLM(FalconConfigLite(vocab_size=c["vocab"], hidden_size=c["heads"] * c["head_dim"], num_hidden_layers=c["layers"],
                                           num_attention_heads=c["heads"], layer_norm_epsilon=c["eps"]), device=DEV)
    m.load_hf_state_dict(g["state_dict"])
    return m


def _mpt(g):
    from promptcache_b200.model.mpt import MptConfigLite, MptForCausalLM
    c = g["config"]
    m = MptForCausalLM(MptConfigLite(d_model=c["heads"] * c["head_dim"], n_heads=c["heads"], n_layers=c["layers"], vocab_size=c["vocab"],
                                     layer_norm_epsilon=c["eps"]), device=DEV)
    m.load_hf_state_dict(g["state_dict"])
    return m


def test_falcon_prefix_then_suffix(native, gold):
    g = gold["falcon"]
    m = _falcon(g)
    sd16 = {k: v.to(DEV).half() for k, v in g["state_dict"].items()}
    cfg = dict(n_layers=2, n_heads=g["config"]["heads"], eps=g["config"]["eps"], theta=10000.0)
    pre_ids, pre_pos = g["prefix_ids"].to(DEV), g["prefix_pos"].to(DEV)
    suf_ids, suf_pos = g["suffix_ids"].to(DEV), g["suffix_pos"].to(DEV)
    o1 = m(input_ids=pre_ids, position_ids=pre_pos, use_cache=True)
    assert o1.past_key_values[0][0].shape == (1, 1, 10, 64)  # cache shape (L, 1, head_dim), model/__init__.py:256-258
    r1, rp = O.falcon_forward(sd16, cfg, pre_ids, pre_pos, cos=m._cos, sin=m._sin)
    assert _rel(o1.logits, r1) < 2e-3
    assert _rel(o1.logits.cpu(), g["prefix_logits"]) < 2e-2
    assert torch.equal(o1.past_key_values[0][0], rp[0][0]) or _rel(o1.past_key_values[0][0], rp[0][0]) < 1e-3
    o2 = m(input_ids=suf_ids, position_ids=suf_pos, past_key_values=o1.past_key_values, use_cache=True)
    r2, _ = O.falcon_forward(sd16, cfg, suf_ids, suf_pos, rp, cos=m._cos, sin=m._sin)
    assert _rel(o2.logits, r2) < 2e-3
    assert _rel(o2.logits.cpu(), g["suffix_logits"]) < 2e-2
    assert o2.past_key_values[1][1].shape == (1, 1, 15, 64)
    assert _rel(o2.past_key_values[1][0].cpu(), g["present"][1][0]) < 1e-2


def test_mpt_prefix_then_suffix_full_position_ids(native, gold):
    g = gold["mpt"]
    m = _mpt(g)
    sd16 = {k: v.to(DEV).half() for k, v in g["state_dict"].items()}
    cfg = dict(n_layers=2, n_heads=g["config"]["heads"], eps=g["config"]["eps"])
    pre_ids, pre_pos = g["prefix_ids"].to(DEV), g["prefix_pos"].to(DEV)
    suf_ids, full_pos = g["suffix_ids"].to(DEV), g["full_pos"].to(DEV)
    o1 = m(pre_ids, pre_pos, use_cache=True)  # MPT takes position_ids positionally 2nd (mpt.py:654-666)
    r1, rp = O.mpt_forward(sd16, cfg, pre_ids, pre_pos)
    assert _rel(o1.logits, r1) < 2e-3 and _rel(o1.logits.cpu(), g["prefix_logits"]) < 2e-2
    o2 = m(suf_ids, full_pos, past_key_values=o1.past_key_values, use_cache=True)
    r2, _ = O.mpt_forward(sd16, cfg, suf_ids, full_pos, rp)
    assert _rel(o2.logits, r2) < 2e-3 and _rel(o2.logits.cpu(), g["suffix_logits"]) < 2e-2
    with pytest.raises(ValueError):
        m(suf_ids, g["suffix_ids"].to(DEV), past_key_values=o2.past_key_values)  # needs the full id list


def test_falcon_engine_parameterized_prompt(native, gold):
    """C4 shape in miniature: parameter slots filled at interior positions, non-monotone suffix ids, MQA cache."""
    from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
    from promptcache_b200.model import Falcon
    g = gold["falcon"]
    lm = Falcon("tiny-falcon", model=_falcon(g), tokenizer=WordTokenizer(128))
    assert lm.get_cache_shape() == (2, 1, 64)
    eng = CacheEngine(128, lm)
    eng.add_schema('<schema name="trip"><module name="sys">you plan trips well</module><module name="plan">a trip of'
                   '<parameter name="duration" length="4"/>days to<parameter name="city" length="5"/>please</module></schema>')
    ids, pos, _, cache = eng.process(Prompt('<prompt schema="trip"><sys/><plan duration="three" city="new york"/>now go</prompt>'))
    assert pos == [7, 13, 14, 19, 20] and cache[0][0].shape == (1, 10, 64)  # parameter slots are not cached rows
    sd16 = {k: v.to(DEV).half() for k, v in g["state_dict"].items()}
    cfg = dict(n_layers=2, n_heads=g["config"]["heads"], eps=g["config"]["eps"], theta=10000.0)
    past = [(k.clone().unsqueeze(0), v.clone().unsqueeze(0)) for k, v in cache]
    ref, _ = O.falcon_forward(sd16, cfg, torch.tensor([ids], device=DEV), torch.tensor([pos], device=DEV), past,
                              cos=lm.hf_model._cos, sin=lm.hf_model._sin)
    out = lm(input_ids=torch.tensor([ids], device=DEV), position_ids=torch.tensor([pos], device=DEV),
             past_key_values=[(k.unsqueeze(0), v.unsqueeze(0)) for k, v in cache], use_cache=True)
    assert _rel(out.logits, ref) < 2e-3
    params = GenerationParameters(temperature=0.0, max_new_tokens=3, stop_token_ids=[], stop_str=[])
    outs = list(GenerationEngine(lm).generate(ids, pos, params, eng.process(Prompt(
        '<prompt schema="trip"><sys/><plan duration="three" city="new york"/>now go</prompt>'))[3], stream_interval=1))
    assert len(outs) == 3


def test_mpt_engine_full_ids_generate(native, gold):
    from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
    from promptcache_b200.model import Mpt
    g = gold["mpt"]
    lm = Mpt("tiny-mpt", model=_mpt(g), tokenizer=WordTokenizer(128))
    assert lm.use_full_position_ids and lm.get_cache_shape() == (2, 2, 64)
    eng = CacheEngine(128, lm)
    eng.add_schema('<schema name="u"><module name="h">a b c</module><union><module name="m1">d e f g h</module>'
                   '<module name="m2">i j</module></union><module name="t">k l</module></schema>')
    xml = '<prompt schema="u"><h/><m2/><t/>what is next</prompt>'
    ids, pos, _, cache = eng.process(Prompt(xml), return_full_position_ids=True)
    n = cache[0][0].shape[1]
    assert n == 7 and len(pos) == n + len(ids) and pos[-3:] == [10, 11, 12]
    sd16 = {k: v.to(DEV).half() for k, v in g["state_dict"].items()}
    cfg = dict(n_layers=2, n_heads=g["config"]["heads"], eps=g["config"]["eps"])
    past = [(k.clone().unsqueeze(0), v.clone().unsqueeze(0)) for k, v in cache]
    ref, _ = O.mpt_forward(sd16, cfg, torch.tensor([ids], device=DEV), torch.tensor([pos], device=DEV), past)
    out = lm(input_ids=torch.tensor([ids], device=DEV), position_ids=torch.tensor([pos], device=DEV),
             past_key_values=[(k.unsqueeze(0), v.unsqueeze(0)) for k, v in cache], use_cache=True)
    assert _rel(out.logits, ref) < 2e-3
    params = GenerationParameters(temperature=0.0, max_new_tokens=3, stop_token_ids=[], stop_str=[])
    ids, pos, _, cache = eng.process(Prompt(xml), return_full_position_ids=True)
    outs = list(GenerationEngine(lm).generate(ids, pos, params, cache, stream_interval=1, use_full_position_ids=True))
    assert len(outs) == 3


def test_falcon_decode_graph_with_runtime_length_equals_eager_decode(native, gold):
    """Falcon decode replays ONE captured graph per KV buffer (run-time cached length in pc_rope_append_dyn and in the
    long-suffix attention kernel, MQA fold of 71 heads -> here 2): same greedy tokens as the eager step-by-step loop."""
    from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
    from promptcache_b200.model import Falcon
    g = gold["falcon"]
    lm = Falcon("tiny-falcon", model=_falcon(g), tokenizer=WordTokenizer(128))
    assert lm.hf_model.supports_dynamic_n
    eng = CacheEngine(256, lm)
    eng.add_schema('<schema name="s"><module name="a">' + " ".join(f"w{i}" for i in range(70)) + '</module><module name="b">'
                   + " ".join(f"v{i}" for i in range(45)) + "</module></schema>")
    xml = '<prompt schema="s"><a/><b/>and then what</prompt>'
    params = GenerationParameters(temperature=0.0, max_new_tokens=12, stop_token_ids=[], stop_str=[])

    def run(graphs, device_sampling):
        lm.hf_model.use_cuda_graphs = graphs
        gen = GenerationEngine(lm)
        gen.device_sampling = device_sampling
        texts = []
        for _ in range(2):  # the second request replays the decode graph captured during the first
            ids, pos, _, cache = eng.process(Prompt(xml))
            texts.append([o.new_text for o in gen.generate(ids, pos, params, cache, stream_interval=1)][-1])
        return texts

    eager = run(False, False)
    replay = run(True, True)
    assert eager[0] == eager[1] == replay[0] == replay[1] and len(eager[0].split()) == 12
    assert any(k[0] == "decode" for k in lm.hf_model._graphs), "the decode step was not captured"


@pytest.mark.parametrize("H,Hkv,D,q,N", [(71, 1, 64, 1, 300), (71, 1, 64, 1, 2100), (8, 2, 128, 40, 700), (71, 1, 64, 21, 256)])
def test_long_suffix_kernel_runtime_length(native, H, Hkv, D, q, N):
    """pc_prefill_attn with n_cached_dev on the ping-pong kernel: grid / maps sized for the upper bound, N read on device."""
    g = torch.Generator(device="cpu").manual_seed(N + q)
    max_ctx = 4096
    Q = torch.randn(H, q, D, generator=g).half().to(DEV)
    K = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    V = torch.zeros_like(K)
    K[:, :N + q] = torch.randn(Hkv, N + q, D, generator=g).half().to(DEV)
    V[:, :N + q] = torch.randn(Hkv, N + q, D, generator=g).half().to(DEV)
    ws = torch.zeros(native.attn_workspace_bytes(H, q, D), dtype=torch.uint8, device=DEV)
    rounding = native.PC_ROUND_FALCON if D == 64 else native.PC_ROUND_LLAMA
    want = torch.empty(q, H * D, dtype=torch.half, device=DEV)
    native.attention(Q, K, V, want, N, q, H, Hkv, D, native.softmax_scale(D), rounding, workspace=ws, impl=native.PC_ATTN_TCGEN05)
    got = torch.full_like(want, float("nan"))
    n_dev = torch.tensor([N], dtype=torch.int32, device=DEV)
    native.attention(Q, K, V, got, max_ctx - q, q, H, Hkv, D, native.softmax_scale(D), rounding, workspace=ws,
                     impl=native.PC_ATTN_TCGEN05, n_cached_dev=n_dev)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    assert _rel(got, want) <= 1e-3  # different split boundaries: fp32 summation order only


@pytest.mark.parametrize("head_dim", [64, 128])  # 64: long-suffix kernel with ALiBi; 128: short-suffix kernel with ALiBi
def test_mpt_decode_graph_with_full_position_ids_equals_eager_decode(native, head_dim):
    """MPT decode replays one captured graph per KV buffer although every step carries the ids of ALL keys: they live in a
    max_ctx-long static buffer, max(position_ids) is taken on the device, the cached length is a device scalar."""
    from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
    from promptcache_b200.model import Mpt
    from promptcache_b200.model.mpt import MptConfigLite, MptForCausalLM
    torch.manual_seed(7)
    model = MptForCausalLM(MptConfigLite(d_model=2 * head_dim, n_heads=2, n_layers=2, vocab_size=128), device=DEV)
    lm = Mpt("tiny-mpt", model=model, tokenizer=WordTokenizer(128))
    eng = CacheEngine(512, lm)
    eng.add_schema('<schema name="s"><module name="a">' + " ".join(f"w{i}" for i in range(150)) + '</module><union><module name="b1">'
                   + " ".join(f"v{i}" for i in range(45)) + '</module><module name="b2">x y z</module></union></schema>')
    params = GenerationParameters(temperature=0.0, max_new_tokens=10, stop_token_ids=[], stop_str=[])

    def run(graphs, xml):
        model.use_cuda_graphs = graphs
        gen = GenerationEngine(lm)
        ids, pos, _, cache = eng.process(Prompt(xml), return_full_position_ids=True)
        return [o.new_text for o in gen.generate(ids, pos, params, cache, stream_interval=1, use_full_position_ids=True)][-1]

    long_xml = '<prompt schema="s"><a/><b1/>and then what</prompt>'
    short_xml = '<prompt schema="s"><b2/>tell me</prompt>'  # a much shorter sequence on the same buffer: stale ids must not leak
    eager = [run(False, long_xml), run(False, short_xml), run(False, long_xml)]
    replay = [run(True, long_xml), run(True, short_xml), run(True, long_xml)]
    assert eager == replay and len(eager[0].split()) == 10
    assert any(k[0] == "decode" for k in model._graphs)


def test_mpt_native_linears_match_cublas_step_and_oracle(native):
    """MPT block on the weight-streaming kernels (no-rotary append in the QKV epilogue, GELU in the up projection's,
    split-K partials reduced by the fused add+LayerNorm) vs the cuBLAS step vs the oracle's eager forward."""
    from promptcache_b200.model.mpt import MptConfigLite, MptForCausalLM
    torch.manual_seed(11)
    model = MptForCausalLM(MptConfigLite(d_model=512, n_heads=4, n_layers=3, vocab_size=640), device=DEV)
    g = torch.Generator().manual_seed(5)
    pre_ids = torch.randint(3, 640, (1, 300), generator=g).to(DEV)
    pre_pos = torch.arange(300, device=DEV).unsqueeze(0)
    suf_ids = torch.randint(3, 640, (1, 9), generator=g).to(DEV)
    full_pos = torch.cat([pre_pos, torch.arange(340, 349, device=DEV).unsqueeze(0)], dim=1)
    outs = {}
    for impl in ("native", "cublas"):
        model.linear_impl = impl
        past = model(pre_ids, pre_pos, use_cache=True).past_key_values
        launches0 = native.launch_count()
        outs[impl] = model(suf_ids, full_pos, past_key_values=past, use_cache=True).logits
        outs[impl + "_launches"] = native.launch_count() - launches0
    assert _rel(outs["native"], outs["cublas"]) <= 2e-3
    assert outs["native_launches"] > outs["cublas_launches"]  # the linears went through the C ABI
    sd = {"transformer.wte.weight": model.wte, "transformer.norm_f.weight": model.norm_f}
    for l in range(3):
        p = f"transformer.blocks.{l}."
        sd.update({p + "norm_1.weight": model.norm1[l], p + "norm_2.weight": model.norm2[l], p + "attn.Wqkv.weight": model.wqkv[l],
                   p + "attn.out_proj.weight": model.wo[l], p + "ffn.up_proj.weight": model.w_up[l], p + "ffn.down_proj.weight": model.w_down[l]})
    sd = {k: v.detach() for k, v in sd.items()}
    cfg = dict(n_layers=3, n_heads=4, eps=1e-5)
    r1, rp = O.mpt_forward(sd, cfg, pre_ids, pre_pos)
    r2, _ = O.mpt_forward(sd, cfg, suf_ids, full_pos, rp)
    assert _rel(outs["native"], r2) <= 3e-3


def test_falcon_native_linears_match_cublas_step_and_oracle(native):
    """Falcon-7B dims (H=71, D=64, hidden 4544 = 71 k-blocks, padded to 72 in the pre-tiled copies), 2 layers: the block on
    the weight-streaming kernels (RoPE at head_dim 64 + MQA append in the QKV epilogue, GELU in h_to_4h's, dense / 4h_to_h
    partials reduced together by the fused add+LayerNorm, tied lm_head) vs the cuBLAS step vs the oracle's eager forward;
    then one decode step each way."""
    from promptcache_b200.model.falcon import FalconConfigLite, FalconForCausalLM
    torch.manual_seed(13)
    model = FalconForCausalLM(FalconConfigLite(vocab_size=1000, num_hidden_layers=2), device=DEV)
    for l in range(2):  # non-trivial norms
        model.ln_w[l].copy_(1 + 0.1 * torch.randn(4544, device=DEV))
        model.ln_b[l].copy_(0.1 * torch.randn(4544, device=DEV))
    g = torch.Generator().manual_seed(5)
    pre_ids = torch.randint(3, 1000, (1, 300), generator=g).to(DEV)
    pre_pos = torch.arange(300, device=DEV).unsqueeze(0)
    suf_ids = torch.randint(3, 1000, (1, 9), generator=g).to(DEV)
    suf_pos = torch.tensor([[340, 341, 350, 351, 352, 310, 311, 353, 354]], device=DEV)  # parameter slots: non-monotone ids
    dec_ids, dec_pos = torch.randint(3, 1000, (1, 1), generator=g).to(DEV), torch.tensor([[355]], device=DEV)
    outs = {}
    model.linear_max_q_falcon = 16
    for impl in ("native", "cublas"):
        model.linear_impl = impl
        past = model(input_ids=pre_ids, position_ids=pre_pos, use_cache=True).past_key_values
        launches0 = native.launch_count()
        o = model(input_ids=suf_ids, position_ids=suf_pos, past_key_values=past, use_cache=True)
        outs[impl + "_launches"] = native.launch_count() - launches0
        outs[impl] = o.logits.clone()
        outs[impl + "_k"] = o.past_key_values[1][0][..., 300:, :].clone()
        outs[impl + "_dec"] = model(input_ids=dec_ids, position_ids=dec_pos, past_key_values=o.past_key_values, use_cache=True).logits.clone()
    assert _rel(outs["native"], outs["cublas"]) <= 2e-3
    assert _rel(outs["native_k"], outs["cublas_k"]) <= 2e-3
    assert _rel(outs["native_dec"], outs["cublas_dec"]) <= 2e-3
    assert outs["native_launches"] >= outs["cublas_launches"] + 2 * 3  # qkv replaces rope_append; dense, h_to_4h, 4h_to_h, lm_head
    sd = {"transformer.word_embeddings.weight": model.word_embeddings, "transformer.ln_f.weight": model.lnf_w,
          "transformer.ln_f.bias": model.lnf_b}
    for l in range(2):
        p = f"transformer.h.{l}."
        sd.update({p + "input_layernorm.weight": model.ln_w[l], p + "input_layernorm.bias": model.ln_b[l],
                   p + "self_attention.query_key_value.weight": model.wqkv[l], p + "self_attention.dense.weight": model.wdense[l],
                   p + "mlp.dense_h_to_4h.weight": model.w_h4h[l], p + "mlp.dense_4h_to_h.weight": model.w_4hh[l]})
    sd = {k: v.detach() for k, v in sd.items()}
    cfg = dict(n_layers=2, n_heads=71, eps=1e-5, theta=10000.0)
    r1, rp = O.falcon_forward(sd, cfg, pre_ids, pre_pos, cos=model._cos, sin=model._sin)
    r2, _ = O.falcon_forward(sd, cfg, suf_ids, suf_pos, rp, cos=model._cos, sin=model._sin)
    assert _rel(outs["native"], r2) <= 3e-3
