"""Where the host time of the e2e TTFT goes (C2, full Llama-2-7B): cProfile of CacheEngine.process + GenerationEngine.generate."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
from promptcache_b200.model import Llama2, LlamaConfigLite, LlamaForCausalLM
from promptcache_b200.synthetic import CONFIGS, MODEL_DIMS, IdTokenizer, make_case
md = dict(MODEL_DIMS[CONFIGS["C2"]["model"]])
case = make_case("C2", vocab=md["vocab_size"])
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = LlamaForCausalLM(LlamaConfigLite(**md), device=dev)
lm = Llama2("llama2-7b", model=model, tokenizer=IdTokenizer(md["vocab_size"]))
eng = CacheEngine(2432, lm)
eng.prompt_cache.retain_prefix = False
eng.add_schema(case.schema_xml)
prompt = Prompt(case.prompt_xml)
gen = GenerationEngine(lm)
params = GenerationParameters(temperature=0.0, max_new_tokens=1, stop_token_ids=[], stop_str=[])

def step():
    ids, pos, ms, cache = eng.process(prompt)
    return next(gen.generate(ids, pos, params, cache, stream_interval=1)), ms

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 50
acc_cache = acc_resp = 0.0
for _ in range(N):
    out, ms = step()
    acc_cache += ms
    acc_resp += out.response_time
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N * 1e3
print(f"wall per request {wall:.3f} ms; cache_time (splice, GPU events) {acc_cache/N:.3f} ms; response_time (prefill, GPU events) {acc_resp/N:.3f} ms; "
      f"host-only remainder {wall - acc_cache/N - acc_resp/N:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    step()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
