"""Wall-clock breakdown of the e2e TTFT path (C2): where the host time between the GPU work goes.
Each section is timed with perf_counter; GPU-side time comes from the engine's own events."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_lm
from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
from promptcache_b200.synthetic import make_case

dev = torch.device("cuda", 0)
lm, md = build_lm("llama2-7b", dev, int(sys.argv[1]) if len(sys.argv) > 1 else None)
case = make_case("C2", vocab=md["vocab_size"])
eng = CacheEngine(2432, lm)
eng.prompt_cache.retain_prefix = False
eng.add_schema(case.schema_xml)
prompt = Prompt(case.prompt_xml)
gen = GenerationEngine(lm)
params = GenerationParameters(temperature=0.0, max_new_tokens=1, stop_token_ids=[], stop_str=[])
rec = {"plan": [], "update": [], "process_total": [], "generate_total": [], "cache_ms": [], "prefill_ms": [], "wall": []}
for i in range(60):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.plan(prompt, collect_original=False)
    t1 = time.perf_counter()
    ids, pos, ms, cache = eng.process(prompt)
    t2 = time.perf_counter()
    out = next(gen.generate(ids, pos, params, cache, stream_interval=1))
    t3 = time.perf_counter()
    if i >= 10:
        rec["plan"].append((t1 - t0) * 1e3)
        rec["process_total"].append((t2 - t1) * 1e3)
        rec["generate_total"].append((t3 - t2) * 1e3)
        rec["cache_ms"].append(ms)
        rec["prefill_ms"].append(out.elapsed_time)
        rec["wall"].append((t3 - t1) * 1e3)
for k, v in rec.items():
    if v:
        print(f"{k:16s} mean {statistics.mean(v):.3f} ms  min {min(v):.3f}")
print(f"host-only in generate (generate_total - prefill_ms): {statistics.mean(rec['generate_total']) - statistics.mean(rec['prefill_ms']):.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    ids, pos, ms, cache = eng.process(prompt)
    next(gen.generate(ids, pos, params, cache, stream_interval=1))
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
