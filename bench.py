#!/usr/bin/env python
"""Headline benchmark: Llama-2-7B TTFT (ms) and prefill tok/s on the prompt-cache hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2]

One step = one request on BASELINE config C2 (code-generation-game shape: 7 cached segments [24,400x5,24] =
2048 tokens, 16-token uncached suffix): request plan -> splice of ALL segments (retained-prefix shortcut off) ->
32-layer prefill of the suffix over the spliced cache (per layer: QKV projection with RoPE at position ids + in-place
append in its epilogue, tcgen05 attention, o_proj / MLP as weight-streaming tcgen05 kernels) -> logits.  TTFT = ms_per_step; value = served prompt tokens/s
= n_gpus*(N+q)/TTFT.  Weights are seeded random fp16 at the real dims, token ids synthetic (no network).
N>1: one rank per GPU, every rank serves its own prompt against its own HBM-resident segment store (weak
scaling, no data-path collective); NCCL only for the barrier and the max-over-ranks time.
`--impl reference` times the reference's algorithm for the same step on the host cores (oracle port: the
reference itself cannot travel to the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def load_traffic(kernel, config):
    """DRAM traffic per launch measured once with `ncu --set full` (committed under profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh).get(kernel)
        return float(t["bytes"]) if t and t.get("config") == config else None
    except Exception:
        return None


def attention_alone(nat, lm, case, dev, peak):
    """The per-layer attention kernel timed ALONE (16 launches captured in a CUDA graph, rotating over 8 layers of
    K/V = 270 MB > L2) -- inside the step it runs within the forward graph and cannot be event-timed."""
    L, Hkv, D = lm.get_cache_shape()
    H, q, N = lm.hf_model.n_heads, case.q_len, case.n_cached
    rot = 8
    max_ctx = ((N + q + 127) // 128) * 128
    K = torch.randn(rot, Hkv, max_ctx, D, device=dev).half()
    V = torch.randn(rot, Hkv, max_ctx, D, device=dev).half()
    Q = torch.randn(H, q, D, device=dev).half()
    out = torch.empty(q, H * D, dtype=torch.half, device=dev)
    ws = torch.zeros(max(nat.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=dev)
    state = {"i": 0}

    def run():
        l = state["i"] % rot
        state["i"] += 1
        nat.attention(Q, K[l], V[l], out, N, q, H, Hkv, D, nat.softmax_scale(D), nat.PC_ROUND_LLAMA, workspace=ws)

    run()
    torch.cuda.synchronize()
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        run()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(16):
                run()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 160
    byts = 4.0 * Hkv * (N + q) * D + 4.0 * H * q * D
    ach = byts / (ms * 1e-3) / 1e9
    return {"kernel": "attn_small_kernel (pc_prefill_attn, one layer)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": load_traffic("attn_small_kernel", "C2"), "algorithmic_bytes_per_launch": byts,
            "avg_launch_ms": ms, "timed": "alone, CUDA-graph of 16 back-to-back launches over 8 layers of K/V (inputs > L2); "
            "launches carry the programmatic-dependent-launch attribute as in the step, so a launch's prologue overlaps its predecessor's tail",
            "note": "34 MB per launch = 5 us at peak: launch/prologue latency bound at this size; 0.66 of peak at C5 (135 MB)"}


def attention_tensor_alone(nat, dev, tflops_peak):
    """Long-suffix (tensor-bound) regime of the same entry point: 1024 uncached tokens over 8192 cached ones at
    Llama-2-7B head geometry, timed alone (CUDA-graph of 8 launches over 4 layers of K/V)."""
    H = Hkv = 32
    D, q, N, rot = 128, 1024, 8192, 4
    max_ctx = N + q
    K = torch.randn(rot, Hkv, max_ctx, D, device=dev).half()
    V = torch.randn(rot, Hkv, max_ctx, D, device=dev).half()
    Q = torch.randn(H, q, D, device=dev).half()
    out = torch.empty(q, H * D, dtype=torch.half, device=dev)
    ws = torch.zeros(max(nat.attn_workspace_bytes(H, q, D), 1), dtype=torch.uint8, device=dev)
    state = {"i": 0}

    def run():
        l = state["i"] % rot
        state["i"] += 1
        nat.attention(Q, K[l], V[l], out, N, q, H, Hkv, D, nat.softmax_scale(D), nat.PC_ROUND_LLAMA, workspace=ws)

    run()
    torch.cuda.synchronize()
    g, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        run()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(8):
                run()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 40
    flops = 4.0 * H * D * (q * N + q * (q + 1) / 2)  # causal-aware: key N+t is visible to rows r >= t
    ach = flops / (ms * 1e-3) / 1e12
    return {"kernel": "attn_pp_kernel (pc_prefill_attn, one layer, q=1024 over N=8192)", "bound": "tensor", "achieved": ach,
            "peak": tflops_peak, "unit": "TFLOP/s", "frac": ach / tflops_peak, "traffic": None, "algorithmic_flops_per_launch": flops,
            "avg_launch_ms": ms, "timed": "alone, CUDA-graph of 8 launches (burst peak applies)",
            "ncu": "sm__pipe_tensor_cycles_active 47% of active cycles (profiles/r01_ncu_full_attn_pp_v4.txt)"}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured", float(p.get("bf16_tflops", 1590.0))
    except Exception:
        return 6650.0, "fallback", 1590.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def mark(self):
        """Samples before this point (warm-up) are not reported."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """TTFT of the reference's CPU path = splice (cache_engine.py:116-156, fp16 cache) + prefill forward
    (llama2.py, fp32 weights as `Llama2(name, device_map=None)` loads them on CPU, eval_sys.py:45-51).
    One step times the FULL splice and `sample_layers` of the L decoder layers (+ embedding, final norm,
    lm_head) and scales the layer time by L/sample_layers."""

    def __init__(self, dims, n_cached, q_len, seg_lens, sample_layers, threads):
        from oracle import pc_oracle as O
        self.O = O
        torch.set_num_threads(threads)
        self.L, H, hid = dims["L"], dims["H"], dims["hidden"]
        self.sample_layers, self.threads = sample_layers, threads
        D = hid // H
        g = torch.Generator().manual_seed(0)
        self.segs = [[(torch.randn(H, n, D, generator=g).half(), torch.randn(H, n, D, generator=g).half())
                      for _ in range(self.L)] for n in seg_lens]
        self.buf = O.new_prompt_cache(self.L, H, n_cached + 256, D)  # preallocated once, like PromptCache.__init__
        self.sd = O.random_llama_state_dict(dims["vocab"], hid, dims["inter"], sample_layers, H, H, seed=0, dtype=torch.float32)
        rope_rows = max(4096, ((n_cached + q_len + 1023) // 1024) * 1024)  # grown on demand like llama2.py:141-142
        self.cfg = dict(n_layers=sample_layers, n_heads=H, n_kv_heads=H, eps=1e-5, theta=10000.0, max_pos=rope_rows)
        self.ids = torch.randint(3, dims["vocab"], (1, q_len), generator=g)
        self.pos = torch.arange(n_cached, n_cached + q_len).unsqueeze(0)
        self.cos, self.sin = O.rope_tables(D, rope_rows)

    def step(self):
        O = self.O
        with torch.inference_mode():
            t0 = time.perf_counter()
            n = O.splice_into(self.buf, self.segs, [0] * len(self.segs))
            t1 = time.perf_counter()
            past = [(self.buf[l][0][:, :n].unsqueeze(0), self.buf[l][1][:, :n].unsqueeze(0)) for l in range(self.sample_layers)]
            t2 = time.perf_counter()
            O.llama_forward(self.sd, self.cfg, self.ids, self.pos, past, cos=self.cos, sin=self.sin)
            t3 = time.perf_counter()
            # embedding + final norm + lm_head alone, to separate them from the per-layer time
            x = torch.nn.functional.embedding(self.ids, self.sd["model.embed_tokens.weight"])
            torch.nn.functional.linear(O.rms_norm(x, self.sd["model.norm.weight"], 1e-5), self.sd["lm_head.weight"]).float()
            t4 = time.perf_counter()
        head = t4 - t3
        layers = max((t3 - t2) - head, 0.0) * (self.L / self.sample_layers)
        return dict(ttft_s=(t1 - t0) + layers + head, splice_s=t1 - t0, layers_s=layers, head_s=head)

    def describe(self, n_cached):
        return (f"per step: full splice of {n_cached} rows x {self.L} layers (fp16 cache) + {self.sample_layers} of {self.L} "
                f"decoder layers fp32 (time x{self.L // self.sample_layers}) + embedding/norm/lm_head; oracle port of the reference")


def run_reference(args, case, dims, seg_lens):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = CpuReference(dims, case.n_cached, case.q_len, seg_lens, int(os.environ.get("PC_BENCH_CPU_LAYERS", "2")), threads)
    recs = []
    for i in range(args.warmup + args.steps):
        r = ref.step()
        if i >= args.warmup:
            recs.append(r)
    ttft = statistics.mean(r["ttft_s"] for r in recs)
    tokens = case.n_cached + case.q_len
    line = {"impl": "reference", "metric": METRIC, "value": tokens / ttft, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ttft * 1e3, "ttft_ms": ttft * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 weights / f16 cache (CPU)", "data": "synthetic",
            "config": config_dict(args, case),
            "cpu_baseline": {"value": tokens / ttft, "unit": "tok/s", "cores": threads, "kind": "port", "sample": ref.describe(case.n_cached),
                             "splice_ms": statistics.mean(r["splice_s"] for r in recs) * 1e3},
            "e2e": {"value": tokens / ttft, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


METRIC = "Llama-2-7B prefill tok/s (served prompt tokens / TTFT; TTFT = ms_per_step)"


def config_dict(args, case):
    from promptcache_b200.synthetic import CONFIGS
    return {"workload": f"{args.config}: {CONFIGS[args.config]['model']}, {case.n_segments} cached segments = {case.n_cached} tokens, "
                        f"{case.q_len}-token uncached suffix, one request per GPU per step",
            "n_cached": case.n_cached, "q_len": case.q_len, "segments": case.n_segments,
            "l2": "inputs larger than L2 (1 GiB of segments + 13.5 GB of weights stream per step)",
            "parallelism": f"batch partition x{args.gpus} (replicated weights and segment store)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3"])
    ap.add_argument("--layers", type=int, default=None, help="debug only: fewer layers (marks the line invalid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from promptcache_b200.synthetic import CONFIGS, MODEL_DIMS, IdTokenizer, make_case
    md = dict(MODEL_DIMS[CONFIGS[args.config]["model"]])
    case = make_case(args.config, vocab=md["vocab_size"])
    seg_lens = {"C1": [24] + [92] * 6 + [24], "C2": [24] + [400] * 5 + [24], "C3": [512] * 8}[args.config]
    dims = dict(L=md["num_hidden_layers"], H=md["num_attention_heads"], hidden=md["hidden_size"],
                inter=md["intermediate_size"], vocab=md["vocab_size"])
    if args.impl == "reference":
        run_reference(args, case, dims, seg_lens)
        return

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt, _native as nat
    from promptcache_b200 import build as pc_build
    from promptcache_b200.model import Llama2, LlamaConfigLite, LlamaForCausalLM
    from promptcache_b200.synthetic import flat_schema
    if rank == 0:
        pc_build.build()
    if world > 1:
        dist.barrier()
    nat.load()

    if args.layers:
        md["num_hidden_layers"] = args.layers
    torch.manual_seed(0)
    model = LlamaForCausalLM(LlamaConfigLite(**md), device=dev)
    lm = Llama2(CONFIGS[args.config]["model"], model=model, tokenizer=IdTokenizer(md["vocab_size"]))
    max_ctx = ((case.n_cached + case.q_len + 256 + 127) // 128) * 128
    eng = CacheEngine(max_ctx, lm)
    eng.prompt_cache.retain_prefix = False  # every step splices every segment
    from promptcache_b200 import parallel
    if world > 1:
        # rank 0 encodes the schema once; the segment store reaches the other GPUs over NCCL (NVLink)
        eng.add_schema(case.schema_xml, no_cache=(rank != 0))
        parallel.broadcast_schema(eng.schemas[case.name], src=0, device=dev)
    else:
        eng.add_schema(case.schema_xml)
    # each rank serves its own prompt (same modules, its own suffix tokens)
    if world > 1:
        mine = flat_schema(case.name, seg_lens, case.q_len, md["vocab_size"], seed=1002 + 17 * rank)
        prompt_xml = mine.prompt_xml
    else:
        prompt_xml = case.prompt_xml
    prompt = Prompt(prompt_xml)
    gen = GenerationEngine(lm)
    params = GenerationParameters(temperature=0.0, max_new_tokens=1, stop_token_ids=[], stop_str=[])
    tokens_per_step = case.n_cached + case.q_len

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ("value"): ids already in HBM, no host sync inside the step -------------------
    ids, pos, _, _ = eng.process(prompt)
    ids_dev = torch.tensor([ids], device=dev)
    pos_dev = torch.tensor([pos], device=dev)
    cached, used, _, _, _, _ = eng.plan(prompt, collect_original=False)
    seq_caches = [cached.get_cache_l1(s) for s in used]
    max_pos = max(pos)
    splice_ev = []

    def device_step(record):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        eng.prompt_cache.update(seq_caches)
        if record:
            e1.record()
            splice_ev.append((e0, e1))
        kv = [(k.unsqueeze(0), v.unsqueeze(0)) for k, v in eng.prompt_cache.cache]
        return lm(input_ids=ids_dev, position_ids=pos_dev, past_key_values=kv, use_cache=True, max_position=max_pos)

    # The clock sampler (nvidia-smi -lms) is started BEFORE the warm-up: its start-up enumerates every GPU of the box
    # and can hold the driver for tens of ms, which must not land in the 20-step timed region.  Only the samples
    # taken from the start of the timed region on are reported.
    sampler = ClockSampler(local)
    if rank == 0 and not args.profile:
        sampler.start()
    for _ in range(args.warmup):
        device_step(False)
    sync_all()
    if args.profile:
        torch.cuda.profiler.start()
        device_step(False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    if rank == 0:
        deadline = time.time() + 2.0  # give the sampler time to deliver its first line
        while not sampler.lines and time.time() < deadline:
            time.sleep(0.02)
    sync_all()
    device_step(False)  # one more untimed step after that wait, so that the timed region starts on a busy GPU
    sync_all()
    sampler.mark()
    launches0 = nat.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        out = device_step(True)
    t1.record()
    sync_all()
    dev_ms = t0.elapsed_time(t1)
    if os.environ.get("PC_BENCH_DEBUG"):
        print(f"[rank {rank}] device arm {dev_ms / args.steps:.3f} ms/step, splice {statistics.mean(a.elapsed_time(b) for a, b in splice_ev):.3f} ms, "
              f"graphs {len(model._graphs)}", file=sys.stderr, flush=True)
    launches = nat.launch_count() - launches0
    splice_ms = statistics.mean(a.elapsed_time(b) for a, b in splice_ev)

    # ---- end-to-end arm: the public API with host inputs (token lists -> device) and the first token read back ---
    def e2e_step():
        i2, p2, cache_ms, cache = eng.process(prompt)
        first = next(gen.generate(i2, p2, params, cache, stream_interval=1))
        return first

    for _ in range(args.warmup):
        e2e_step()
    sync_all()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    sync_all()
    e2e_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- decode after the prefill (informative; the headline metric is TTFT): 33 greedy tokens through the public API
    decode_ms = None
    try:
        dparams = GenerationParameters(temperature=0.0, max_new_tokens=33, stop_token_ids=[], stop_str=[])
        for _ in range(2):  # first pass warms up (captures the decode graph of this KV buffer), second is reported
            i2, p2, _, cache = eng.process(prompt)
            outs = list(gen.generate(i2, p2, dparams, cache))  # default stream_interval (2): greedy tokens stay on the device between yields
        decode_ms = (outs[-1].response_time - outs[0].response_time) / 32.0
    except Exception:
        pass

    times = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(times[0]), float(times[1])

    if rank == 0:
        peak, peak_kind, tflops_peak = load_peaks()
        ms_per_step = dev_ms / args.steps
        L, Hkv, D = lm.get_cache_shape()
        splice_bytes = 8.0 * L * Hkv * case.n_cached * D  # read + write of K and V, fp16
        achieved = splice_bytes / (splice_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": world * tokens_per_step / (ms_per_step * 1e-3), "unit": "tok/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "ttft_ms": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": config_dict(args, case),
                "e2e": {"value": world * tokens_per_step / (e2e_ms / args.steps * 1e-3), "unit": "tok/s",
                        "ttft_ms": e2e_ms / args.steps, "h2d_bytes_per_step": 2 * 8 * case.q_len, "d2h_bytes_per_step": 8,
                        "note": "CacheEngine.process + GenerationEngine.generate(max_new_tokens=1): token/position id lists go "
                                "host->device, the greedy token comes back; cached segments are HBM-resident by design"},
                "gpu_launches": int(launches), "decode_ms_per_token": decode_ms,
                "roofline": {"kernel": "splice_kernel (pc_splice_kv)", "bound": "hbm", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic("splice_kernel", args.config),
                             "peak_kind": peak_kind + " (burst copy)",
                             "algorithmic_bytes_per_launch": splice_bytes, "avg_launch_ms": splice_ms,
                             "share_of_step": splice_ms / ms_per_step},
                "clocks": clocks}
        try:
            line["roofline_attention"] = attention_alone(nat, lm, case, dev, peak)
        except Exception as ex:
            line["roofline_attention"] = {"error": str(ex)}
        try:
            line["roofline_attention_tensor"] = attention_tensor_alone(nat, dev, tflops_peak)
        except Exception as ex:
            line["roofline_attention_tensor"] = {"error": str(ex)}
        if args.layers:
            line["invalid"] = f"debug run with {args.layers} layers"
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            try:
                ref = CpuReference(dims, case.n_cached, case.q_len, seg_lens, 2, threads)
                ref.step()
                r = min((ref.step() for _ in range(3)), key=lambda d: d["ttft_s"])
                line["cpu_baseline"] = {"value": tokens_per_step / r["ttft_s"], "unit": "tok/s", "cores": threads, "kind": "port",
                                        "ttft_ms": r["ttft_s"] * 1e3, "splice_ms": r["splice_s"] * 1e3,
                                        "sample": "best of 3 after 1 warm-up; " + ref.describe(case.n_cached)}
            except Exception as ex:  # never lose the GPU line to a host-side OOM
                line["cpu_baseline"] = {"value": None, "unit": "tok/s", "cores": threads, "kind": "port", "sample": f"failed: {ex}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
