"""Virtual splice: the short-suffix attention kernel reads the cached segments in place from
the schema's arena through a per-request tile list instead of from a copy (reference copy loop:
promptcache/cache_engine.py:135-151).  Checked against (a) the copied path of the same kernel -- bit-identical whenever
the segment lengths are multiples of the 128-key tile, because the tiles are then the same -- and (b) the oracle's eager
attention over the oracle's spliced cache within the north-star tolerance; plus the engine-level behaviour: no splice
launch on the engine's own path, lazy materialisation when a caller looks at the reference-shaped views, staleness."""
import pytest
import torch

from oracle import pc_oracle as O
from tests.test_kernels_gpu import _oracle_llama, _rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _arena_case(H, Hkv, lens, order, q, seed, extra_rows=37, tail_old=0, std=1.28):
    """One layer of a schema arena [Hkv, R, D] holding len(lens) segments (in arena order) + junk rows between/after them,
    a request that uses the segments in `order`, and the equivalent contiguous buffer (what the splice would build)."""
    D = 128
    g = torch.Generator(device="cpu").manual_seed(seed)
    starts, r = [], 0
    for n in lens:
        starts.append(r)
        r += n
    R = r + extra_rows
    AK = (torch.randn(Hkv, R, D, generator=g) * std).half().to(DEV)
    AV = torch.randn(Hkv, R, D, generator=g).half().to(DEV)
    N_virtual = sum(lens[i] for i in order)
    N = N_virtual + tail_old
    max_ctx = ((N + q + 127) // 128) * 128 + 64
    K = torch.zeros(Hkv, max_ctx, D, dtype=torch.half, device=DEV)
    V = torch.zeros_like(K)
    row, tiles = 0, []
    for i in order:
        K[:, row:row + lens[i]] = AK[:, starts[i]:starts[i] + lens[i]]
        V[:, row:row + lens[i]] = AV[:, starts[i]:starts[i] + lens[i]]
        row += lens[i]
        for t in range(0, lens[i], 128):
            tiles.append((starts[i] + t, min(128, lens[i] - t)))
    K[:, row:N + q] = (torch.randn(Hkv, tail_old + q, D, generator=g) * std).half().to(DEV)
    V[:, row:N + q] = torch.randn(Hkv, tail_old + q, D, generator=g).half().to(DEV)
    Q = (torch.randn(H, q, D, generator=g) * std).half().to(DEV)
    tiles_dev = torch.tensor(tiles, dtype=torch.int32, device=DEV)
    return Q, K, V, AK, AV, tiles_dev, len(tiles), N_virtual, N


def _run(native, Q, K, V, N, q, H, Hkv, virtual=None, K_for_virtual=None, impl_flags=0, n_dev=None, n_upper=None):
    D = 128
    ws = torch.zeros(max(native.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=DEV)
    out = torch.full((q, H * D), float("nan"), dtype=torch.half, device=DEV)
    native.attention(Q, K, V, out, N if n_upper is None else n_upper, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA,
                     workspace=ws, decode=(q == 1), impl=native.PC_ATTN_TCGEN05 | impl_flags, virtual=virtual, n_cached_dev=n_dev)
    torch.cuda.synchronize()
    assert int(ws[:4096].view(torch.int32).abs().max()) == 0
    return out


def _hide_cached_rows(K, V, n_virtual):
    """The buffer as the engine leaves it in virtual mode: rows [0, n_virtual) were never written."""
    K2, V2 = K.clone(), V.clone()
    K2[:, :n_virtual] = 0
    V2[:, :n_virtual] = 0
    return K2, V2


@pytest.mark.parametrize("H,Hkv,lens,order,q", [
    (32, 32, [256] * 8, [5, 0, 7, 2], 32),               # C5-like: whole tiles -> same tiles as the copied layout
    (32, 32, [128, 384, 256, 128], [3, 1, 0], 16),
    (8, 2, [256, 128, 512], [2, 0], 4),                  # GQA fold
    (32, 32, [256] * 64, sorted(range(0, 64, 2)), 32),   # C5 full size: 32 of a 64-segment pool = 8192 rows
])
def test_virtual_equals_copied_bit_for_bit_on_whole_tile_segments(native, H, Hkv, lens, order, q):
    Q, K, V, AK, AV, tiles, nt, nv, N = _arena_case(H, Hkv, lens, order, q, seed=sum(lens) + q)
    copied = _run(native, Q, K, V, N, q, H, Hkv)
    Kh, Vh = _hide_cached_rows(K, V, nv)
    virt = _run(native, Q, Kh, Vh, N, q, H, Hkv, virtual=(AK, AV, tiles, nt, nv))
    assert torch.equal(virt, copied)
    assert _rel(virt, _oracle_llama(Q, K, V, N, q)) <= 1e-3
    # cluster merge and global merge agree in virtual mode too
    virt_g = _run(native, Q, Kh, Vh, N, q, H, Hkv, virtual=(AK, AV, tiles, nt, nv), impl_flags=native.PC_ATTN_GLOBAL_MERGE)
    assert torch.equal(virt_g, virt)


@pytest.mark.parametrize("H,Hkv,lens,order,q,tail_old", [
    (32, 32, [24, 400, 400, 400, 400, 400, 24], [0, 1, 2, 3, 4, 5, 6], 16, 0),   # C2: ragged segments, partial tiles
    (32, 32, [24, 400, 400, 400, 400, 400, 24], [6, 3, 0, 5], 9, 0),             # usage-reordered subset
    (32, 32, [1, 129, 3, 127, 255], [4, 2, 1, 0, 3], 1, 0),                      # decode over odd lengths
    (32, 32, [24, 92, 92, 92, 92, 92, 92, 24], list(range(8)), 16, 0),           # C1 (persona) lengths
    (16, 4, [300, 77, 500], [1, 2, 0], 5, 21),                                   # GQA + rows appended by earlier forwards
    (32, 32, [400, 400], [1, 0], 1, 49),                                         # decode step 49 after a 16-token prefill
])
def test_virtual_ragged_segments_match_copied_path_and_oracle(native, H, Hkv, lens, order, q, tail_old):
    Q, K, V, AK, AV, tiles, nt, nv, N = _arena_case(H, Hkv, lens, order, q, seed=len(lens) * 1000 + q, tail_old=tail_old)
    copied = _run(native, Q, K, V, N, q, H, Hkv)
    Kh, Vh = _hide_cached_rows(K, V, nv)
    virt = _run(native, Q, Kh, Vh, N, q, H, Hkv, virtual=(AK, AV, tiles, nt, nv))
    ref = _oracle_llama(Q, K, V, N, q)
    assert _rel(virt, ref) <= 1e-3 and _rel(copied, ref) <= 1e-3
    # different tile boundaries, same keys: the outputs differ by fp32 summation order, i.e. an fp16 ulp here and there
    assert _rel(virt, copied) <= 1e-3
    assert float((virt.float() - copied.float()).abs().mean() / copied.float().abs().mean()) <= 3e-4
    assert torch.isfinite(virt.float()).all()


def test_virtual_runtime_length_decode_graph_semantics(native):
    """Run-time N (decode graph): n_cached_dev counts ALL keys (virtual rows + appended rows); one launch shape serves a
    growing sequence."""
    H = Hkv = 32
    lens, order = [400, 24, 400], [2, 0, 1]
    for tail_old in (0, 16, 17, 130):
        Q, K, V, AK, AV, tiles, nt, nv, N = _arena_case(H, Hkv, lens, order, 1, seed=tail_old + 5, tail_old=tail_old)
        Kh, Vh = _hide_cached_rows(K, V, nv)
        n_dev = torch.tensor([N], dtype=torch.int32, device=DEV)
        upper = K.shape[1] - 1
        got = _run(native, Q, Kh, Vh, N, 1, H, Hkv, virtual=(AK, AV, tiles, nt, nv), n_dev=n_dev, n_upper=upper)
        want = _run(native, Q, Kh, Vh, N, 1, H, Hkv, virtual=(AK, AV, tiles, nt, nv))
        assert _rel(got, _oracle_llama(Q, K, V, N, 1)) <= 1e-3
        assert _rel(got, want) <= 1e-3


def test_virtual_rejects_what_it_cannot_serve(native):
    Q, K, V, AK, AV, tiles, nt, nv, N = _arena_case(32, 32, [256, 256], [0, 1], 64, seed=1)  # G*q = 64 > 32
    with pytest.raises(native.NativeError):
        _run(native, Q, K, V, N, 64, 32, 32, virtual=(AK, AV, tiles, nt, nv))


# ---------------------------------------------------------------------------------------------------------------------
# engine level
# ---------------------------------------------------------------------------------------------------------------------
def _engine(layers=2, heads=4, kv_heads=None, hidden=512, vocab=512, seg_lens=(24, 130, 77), suffix=9, seed=3, max_ctx=512):
    from promptcache_b200 import CacheEngine, Prompt
    from promptcache_b200.model import Llama2, LlamaConfigLite, LlamaForCausalLM
    from promptcache_b200.synthetic import IdTokenizer, flat_schema
    kv_heads = kv_heads or heads
    sd = O.random_llama_state_dict(vocab, hidden, 2 * hidden, layers, heads, kv_heads, seed=0, dtype=torch.float16, device=DEV)
    cfg = LlamaConfigLite(vocab_size=vocab, hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=layers,
                          num_attention_heads=heads, num_key_value_heads=kv_heads, max_position_embeddings=max_ctx)
    model = LlamaForCausalLM(cfg, device=DEV)
    model.load_hf_state_dict(sd)
    lm = Llama2("virt", model=model, tokenizer=IdTokenizer(vocab))
    case = flat_schema("virt", list(seg_lens), suffix, vocab, seed=seed)
    eng = CacheEngine(max_ctx, lm)
    eng.add_schema(case.schema_xml)
    return eng, lm, case, Prompt(case.prompt_xml)


def _greedy(lm, eng, prompt, n=6):
    from promptcache_b200 import GenerationEngine, GenerationParameters
    params = GenerationParameters(temperature=0.0, max_new_tokens=n, stop_token_ids=[], stop_str=[])
    ids, pos, _, cache = eng.process(prompt)
    outs = list(GenerationEngine(lm).generate(ids, pos, params, cache, stream_interval=1))
    return outs[-1].new_text


def test_engine_serves_requests_without_a_splice_launch_and_materialises_on_demand(native):
    eng, lm, case, prompt = _engine()
    # all token sequences of the schema share one arena allocation
    entries = list(eng.schemas["virt"].cache_l1.values())
    assert all(e.arena is eng.schemas["virt"].arena for e in entries) and len({e.row0 for e in entries}) == len(entries)
    ids, pos, _, cache = eng.process(prompt)
    pc = eng.prompt_cache
    assert pc.virtual is not None and pc.last_spliced_rows == 0 and len(cache) == 2
    assert int(pc.buffer[:, :, :, :len(pc)].abs().max()) == 0, "virtual mode must not have copied anything"
    out_v = lm(input_ids=torch.tensor([ids], device=DEV), position_ids=torch.tensor([pos], device=DEV),
               past_key_values=cache.batched(), use_cache=True)
    assert pc.last_spliced_rows == 0
    # a caller that looks at the reference-shaped views gets real rows, bit-exact against the oracle's splice
    segs = [m.cache for m in pc.staged]
    want, n = O.splice(segs, [0] * len(segs), 512)
    for l in range(2):
        assert torch.equal(cache[l][0], want[l][0]) and torch.equal(cache[l][1], want[l][1])
    assert pc.last_spliced_rows == n == case.n_cached
    # the copied path (what the reference does) gives the same logits within fp32 summation-order noise
    eng.virtual_splice = False
    ids2, pos2, _, cache2 = eng.process(prompt)
    assert eng.prompt_cache.virtual is None and eng.prompt_cache.last_spliced_rows == 0  # retained: the rows are already there
    eng.prompt_cache.invalidate()
    ids2, pos2, _, cache2 = eng.process(prompt)
    assert eng.prompt_cache.virtual is None and eng.prompt_cache.last_spliced_rows == case.n_cached
    out_c = lm(input_ids=torch.tensor([ids2], device=DEV), position_ids=torch.tensor([pos2], device=DEV),
               past_key_values=cache2.batched(), use_cache=True)
    assert _rel(out_v.logits, out_c.logits) <= 1e-3
    # and against the oracle's forward over the oracle's spliced cache
    sd = {k: v for k, v in O.random_llama_state_dict(512, 512, 1024, 2, 4, 4, seed=0, dtype=torch.float16, device=DEV).items()}
    ocfg = dict(n_layers=2, n_heads=4, n_kv_heads=4, eps=1e-5, theta=10000.0, max_pos=512)
    ref, _ = O.llama_forward(sd, ocfg, torch.tensor([ids], device=DEV), torch.tensor([pos], device=DEV),
                             [(k.clone().unsqueeze(0), v.clone().unsqueeze(0)) for k, v in want], cos=lm.hf_model._cos, sin=lm.hf_model._sin)
    assert _rel(out_v.logits, ref) <= 2e-3


def test_generation_through_the_virtual_splice_equals_the_copied_path(native):
    eng, lm, case, prompt = _engine(seg_lens=(24, 400, 130, 24), suffix=12, max_ctx=1024)
    launches0 = native.launch_count()
    text_v = _greedy(lm, eng, prompt)
    assert eng.prompt_cache.last_spliced_rows == 0
    text_v2 = _greedy(lm, eng, prompt)  # second request: graph replay of the same tile / row counts
    eng.virtual_splice = False
    text_c = _greedy(lm, eng, prompt)
    assert text_v == text_v2 == text_c and len(text_v.split()) == 6
    assert native.launch_count() > launches0


def test_gqa_engine_cache_shape_and_virtual_splice(native):
    """Hkv < H through the engine (ADVICE: get_cache_shape must size the prompt cache for the KV heads)."""
    eng, lm, case, prompt = _engine(heads=8, kv_heads=2, hidden=1024, seg_lens=(130, 24, 260), suffix=4, max_ctx=640)
    assert lm.get_cache_shape() == (2, 2, 128) and eng.prompt_cache.buffer.shape[2] == 2
    text_v = _greedy(lm, eng, prompt, n=4)
    eng.virtual_splice = False
    assert _greedy(lm, eng, prompt, n=4) == text_v


def test_stale_cache_raises_and_reused_past_forks(native):
    from promptcache_b200 import StaleKVError
    eng, lm, case, prompt = _engine()
    ids, pos, _, cache = eng.process(prompt)
    ids_t, pos_t = torch.tensor([ids], device=DEV), torch.tensor([pos], device=DEV)
    eng.process(prompt)  # the engine moved on: `cache` describes rows that are being reused
    with pytest.raises(StaleKVError):
        lm(input_ids=ids_t, position_ids=pos_t, past_key_values=cache.batched(), use_cache=True)
    # the same past used twice (two continuations of one prefix): the second call must not clobber the first one's rows
    ids, pos, _, cache = eng.process(prompt)
    past = cache.batched()
    o1 = lm(input_ids=ids_t, position_ids=pos_t, past_key_values=past, use_cache=True)
    k1 = o1.past_key_values[0][0].clone()
    other = (ids_t + 7) % 512
    o2 = lm(input_ids=other, position_ids=pos_t, past_key_values=past, use_cache=True)
    assert torch.equal(o1.past_key_values[0][0], k1), "the first continuation's KV rows were overwritten"
    assert not torch.equal(o2.past_key_values[0][0][..., -1, :], k1[..., -1, :])
    o1b = lm(input_ids=ids_t, position_ids=pos_t, past_key_values=past, use_cache=True)  # forked again: same answer
    assert _rel(o1b.logits, o1.logits) <= 1e-3


@pytest.mark.parametrize("lens,order,q,tail_old", [([256, 128, 384], [2, 0], 7, 0), ([100, 45, 300, 17], [3, 1, 0], 1, 5),
                                                   ([128] * 6, [4, 1, 5, 0], 32, 0)])
def test_virtual_splice_with_alibi_key_ids_in_tile_order(native, lens, order, q, tail_old):
    """MPT: ALiBi bias at the KEY's position id (mpt.py:97-110,172-175).  With a tile list the id table is laid out in tile
    order ([128 slots per virtual tile][appended rows]); result == the copied path with the ids in buffer order."""
    H = Hkv = 8
    D = 128
    Q, K, V, AK, AV, tiles, nt, nv, N = _arena_case(H, Hkv, lens, order, q, seed=11 + q, tail_old=tail_old, std=1.0)
    g = torch.Generator(device="cpu").manual_seed(5)
    pos = torch.randperm(4 * (N + q), generator=g)[:N + q].to(torch.int32)  # gapped, unordered ids of all keys (buffer order)
    key_pos = (pos - pos.max()).to(DEV)
    slopes = (2.0 ** (-torch.arange(1, H + 1, dtype=torch.float32))).to(DEV)
    # slot -> buffer row, as PromptCache._upload_tiles builds it (test tiles follow the request's order, one run per segment)
    slot_rows, row = [], 0
    for i in order:
        for t in range(0, lens[i], 128):
            n = min(128, lens[i] - t)
            slot_rows += list(range(row + t, row + t + n)) + [0] * (128 - n)
        row += lens[i]
    slot_rows = torch.tensor(slot_rows, dtype=torch.int64, device=DEV)
    key_pos_tiles = torch.cat([key_pos[slot_rows], key_pos[nv:]]).contiguous()
    ws = torch.zeros(native.attn_workspace_bytes(H, q, D), dtype=torch.uint8, device=DEV)

    def run(kp, virtual, Kb, Vb):
        out = torch.full((q, H * D), float("nan"), dtype=torch.half, device=DEV)
        native.attention(Q, Kb, Vb, out, N, q, H, Hkv, D, native.softmax_scale(D), native.PC_ROUND_LLAMA, workspace=ws,
                         alibi_slopes=slopes, key_pos=kp, alibi_max_pos=0, decode=(q == 1), impl=native.PC_ATTN_TCGEN05, virtual=virtual)
        torch.cuda.synchronize()
        return out

    copied = run(key_pos.contiguous(), None, K, V)
    Kh, Vh = _hide_cached_rows(K, V, nv)
    virt = run(key_pos_tiles, (AK, AV, tiles, nt, nv), Kh, Vh)
    assert torch.isfinite(virt.float()).all()
    assert _rel(virt, copied) <= 1e-3  # tile boundaries differ where segments are ragged: fp32 summation order only
    if all(n % 128 == 0 for n in lens):
        assert torch.equal(virt, copied)


def test_mpt_engine_virtual_splice_equals_copied_splice(native):
    """MPT through the engine with full position ids: the virtual splice (no copy; ids gathered into tile order through
    VirtualKV.slot_rows) against the copied splice -- same logits, same greedy continuation (incl. the graph-replayed decode),
    also after the usage counters have reordered the buffer layout (buffer order != arena order)."""
    from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
    from promptcache_b200.model import Mpt
    from promptcache_b200.model.mpt import MptConfigLite, MptForCausalLM
    from tests.helpers import WordTokenizer
    torch.manual_seed(3)
    model = MptForCausalLM(MptConfigLite(d_model=256, n_heads=2, n_layers=2, vocab_size=256), device=DEV)
    lm = Mpt("tiny-mpt", model=model, tokenizer=WordTokenizer(256))
    assert model.virtual_kv_max_q() == 32 and model.virtual_kv_full_ids
    eng = CacheEngine(1024, lm)
    words = lambda p, n: " ".join(f"{p}{i}" for i in range(n))
    eng.add_schema(f'<schema name="s"><module name="a">{words("a", 150)}</module><module name="b">{words("b", 70)}</module>'
                   f'<union><module name="c1">{words("c", 200)}</module><module name="c2">{words("d", 33)}</module></union>'
                   f'<module name="e">{words("e", 90)}</module></schema>')
    params = GenerationParameters(temperature=0.0, max_new_tokens=8, stop_token_ids=[], stop_str=[])
    gen = GenerationEngine(lm)

    def serve(xml, virtual):
        eng.virtual_splice = virtual
        before = native.launch_count()
        ids, pos, _, cache = eng.process(Prompt(xml), return_full_position_ids=True)
        assert (cache.virtual is not None) == virtual
        out = lm(input_ids=torch.tensor([ids], device=DEV), position_ids=torch.tensor([pos], device=DEV), past_key_values=cache,
                 use_cache=True)
        logits = out.logits.clone()
        ids, pos, _, cache = eng.process(Prompt(xml), return_full_position_ids=True)
        text = [o.new_text for o in gen.generate(ids, pos, params, cache, stream_interval=1, use_full_position_ids=True)][-1]
        return logits, text

    x1 = '<prompt schema="s"><a/><c2/><e/>what now my friend</prompt>'
    x2 = '<prompt schema="s"><b/><c1/><e/>and then</prompt>'
    for xml in (x1, x2, x1, x2, x2):  # repeated: graph capture + replay; the counters of b / c1 / e overtake a's on the way
        lv, tv = serve(xml, True)
        lc, tc = serve(xml, False)
        assert _rel(lv, lc) <= 2e-3
        assert tv == tc and len(tv.split()) == 8
    assert any(k[0] == "decode" for k in model._graphs)
