#!/usr/bin/env python
"""eval_sys-style system evaluation on the B200 path (reference eval_sys.py:88-215, :300-375).

    python tools/eval_sys.py critical_point [--model llama2-7b] [--out gpurun_out/eval_sys_critical_point.json]
    python tools/eval_sys.py latency        [--model llama2-7b] [--out gpurun_out/eval_sys_latency.json]

`critical_point` is the reference's "recomputation overhead vs memory transfer overhead" sweep (run_critical_point): for
each sequence length, (1) the time to make a cached KV of that length attendable and (2) the time to recompute it with a
no-cache prefill.  The reference's (1) is a host->device upload of L x 2 pageable tensors per request (its segments live
in host memory, cache_engine.py:290-292); here segments are HBM-resident, so (1) has three columns: the same upload from
pinned host memory (what a host-resident store would cost on this box), the on-device splice copy (`PromptCache.update`,
one pc_splice_kv launch) and the virtual splice (a tile list, no copy).
`latency` is run_latency_eval on synthetic schemas (no LongBench on disk): TTFT with the prompt cache
(cache_time + response_time, exactly the two stamps eval_sys.py:327-359 records) against no_cache=True full prefill, over
a sweep of cached lengths.
Everything goes through the public API (CacheEngine / GenerationEngine-compatible lm() calls), synthetic weights/tokens.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SEQ_LENS = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 640, 768, 896, 1024, 1280, 1536, 1792, 2048, 2540, 3052, 3564, 4096]  # eval_sys.py:101-131


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _lm(model_key, dev, layers=None):
    from bench import build_lm
    return build_lm(model_key, dev, layers)


def critical_point(args):
    from promptcache_b200.cache_engine import PromptCache, TokenSequenceCache
    dev = torch.device("cuda", 0)
    lm, md = _lm(args.model, dev, args.layers)
    L, Hkv, D = lm.get_cache_shape()
    rows = []
    repeat = args.repeat
    for n in (args.seq_lens or SEQ_LENS):
        rec = {"seq_len": n}
        # ---- (1a) upload from (pinned) host memory: L x (K, V) tensors [Hkv, n, D] fp16, as the reference stores them
        host = [(torch.rand(Hkv, n, D, dtype=torch.float16).pin_memory(), torch.rand(Hkv, n, D, dtype=torch.float16).pin_memory())
                for _ in range(L)]
        ts = []
        for _ in range(repeat):
            torch.cuda.synchronize()
            e0, e1 = _events()
            e0.record()
            gpu = [(k.to(dev, non_blocking=True), v.to(dev, non_blocking=True)) for k, v in host]
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            del gpu
        rec["upload_ms"] = min(ts)
        del host
        # ---- (1b) on-device splice of one segment of n rows, and (1c) the virtual splice of the same segment
        seg = TokenSequenceCache(type("S", (), {"offset": 0, "__len__": lambda s: n})(), torch.rand(L, 2, Hkv, n, D, device=dev).half())
        pc = PromptCache(max_ctx_length=((n + 127) // 128) * 128 + 128, num_layers=L, num_head=Hkv, head_dim=D, target_device=dev)
        pc.retain_prefix = False
        for virtual, key in ((False, "splice_copy_ms"), (True, "virtual_splice_ms")):
            ts = []
            for _ in range(repeat + 1):
                torch.cuda.synchronize()
                e0, e1 = _events()
                e0.record()
                pc.update([seg], virtual=virtual)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            rec[key] = min(ts[1:])
        del pc, seg
        # ---- (2) recomputation: no-cache prefill of n tokens (eval_sys.py:178-203)
        ids = torch.full((1, n), 100, device=dev, dtype=torch.long)
        pos = torch.arange(n, device=dev).unsqueeze(0)
        ts = []
        for _ in range(repeat + 1):
            e0, e1 = _events()
            e0.record()
            out = lm(input_ids=ids, position_ids=pos, past_key_values=None, use_cache=False, max_position=n - 1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            del out
        rec["recompute_ms"] = min(ts[1:])
        rows.append(rec)
        print(rec, flush=True)
        torch.cuda.empty_cache()
    return {"model": args.model, "layers": md["num_hidden_layers"], "what": "critical point: make n cached tokens attendable vs recompute them",
            "results": rows}


def latency(args):
    from promptcache_b200 import CacheEngine, Prompt
    from promptcache_b200.synthetic import flat_schema
    dev = torch.device("cuda", 0)
    lm, md = _lm(args.model, dev, args.layers)
    vocab = md["vocab_size"]
    rows = []
    for n_cached in (args.seq_lens or (256, 512, 1024, 2048, 3072, 4096, 8192)):
        seg = 256 if n_cached >= 1024 else 128
        case = flat_schema(f"s{n_cached}", [seg] * (n_cached // seg), 16, vocab, seed=7000 + n_cached)
        eng = CacheEngine(((n_cached + 16 + 256 + 127) // 128) * 128, lm)
        eng.prompt_cache.retain_prefix = False
        rec = {"n_cached": n_cached, "q_len": 16, "segments": n_cached // seg}
        for do_cache in (True, False):
            eng.add_schema(case.schema_xml, no_cache=not do_cache)  # eval_sys.py:314-315
            prompt = Prompt(case.prompt_xml)
            res = []
            for _ in range(args.repeat + 2):
                token_ids, position_ids, cache_time, cache = eng.process(prompt, no_cache=not do_cache)
                input_ids = torch.tensor([token_ids], device=dev, dtype=torch.long)
                pos = torch.tensor([position_ids], device=dev, dtype=torch.long)
                past = cache.batched() if cache is not None else None  # == [(k.unsqueeze(0), v.unsqueeze(0)) ...] (eval_sys.py:330-331)
                e0, e1 = _events()
                e0.record()
                out = lm(input_ids=input_ids, position_ids=pos, past_key_values=past, use_cache=True, max_position=max(position_ids))
                e1.record()
                torch.cuda.synchronize()
                res.append((cache_time, e0.elapsed_time(e1)))
                del out
            res = res[2:]
            tag = "cache" if do_cache else "no_cache"
            rec[f"{tag}_cache_time_ms"] = statistics.mean(r[0] for r in res)
            rec[f"{tag}_response_time_ms"] = statistics.mean(r[1] for r in res)
            eng.remove_all_schemas()
        rec["speedup"] = (rec["no_cache_cache_time_ms"] + rec["no_cache_response_time_ms"]) / (rec["cache_cache_time_ms"] + rec["cache_response_time_ms"])
        rows.append(rec)
        print(rec, flush=True)
        del eng
        torch.cuda.empty_cache()
    return {"model": args.model, "layers": md["num_hidden_layers"], "what": "TTFT with the prompt cache vs no_cache full prefill (eval_sys.run_latency_eval)",
            "results": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["critical_point", "latency"])
    ap.add_argument("--model", default="llama2-7b")
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--seq-lens", type=int, nargs="*", default=None, help="override the sweep (tests)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from promptcache_b200 import build as pc_build
    pc_build.build()
    with torch.inference_mode():
        res = critical_point(args) if args.mode == "critical_point" else latency(args)
    out = args.out or os.path.join(ROOT, "gpurun_out", f"eval_sys_{args.mode}.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as fh:
        json.dump(res, fh, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
