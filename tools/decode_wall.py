"""Wall-clock and GPU-event time per decoded token (C2, full Llama-2-7B) with and without device-side greedy sampling."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from promptcache_b200 import CacheEngine, GenerationEngine, GenerationParameters, Prompt
from promptcache_b200.model import Llama2, LlamaConfigLite, LlamaForCausalLM
from promptcache_b200.synthetic import CONFIGS, MODEL_DIMS, IdTokenizer, make_case
md = dict(MODEL_DIMS[CONFIGS["C2"]["model"]])
case = make_case("C2", vocab=md["vocab_size"])
torch.manual_seed(0)
model = LlamaForCausalLM(LlamaConfigLite(**md), device="cuda")
lm = Llama2("llama2-7b", model=model, tokenizer=IdTokenizer(md["vocab_size"]))
eng = CacheEngine(2432, lm)
eng.add_schema(case.schema_xml)
params = GenerationParameters(temperature=0.0, max_new_tokens=65, stop_token_ids=[], stop_str=[])
for ds in (True, False, True, False):
    gen = GenerationEngine(lm)
    gen.device_sampling = ds
    for rep in range(2):
        ids, pos, _, cache = eng.process(Prompt(case.prompt_xml))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = list(gen.generate(ids, pos, params, cache))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    print(f"device_sampling={ds}: 65 tokens wall {wall*1e3:.1f} ms -> {(wall*1e3 - outs[0].response_time)/64:.3f} ms/token wall after TTFT; "
          f"GPU-event {(outs[-1].response_time-outs[0].response_time)/64:.3f} ms/token", flush=True)
