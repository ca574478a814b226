"""Prefill + decode loop with the reference's interface (promptcache/generation_engine.py:21-209).

`GenerationEngine(lm).generate(token_ids, position_ids, params, cache, stream_interval,
use_full_position_ids)` yields `Output`s exactly like the reference, including its quirks:
  * decode positions start at max(position_ids)+1+i with i >= 1, i.e. one position is skipped (:82,132);
  * `Output` is built positionally as (output, new_output, inference_time, response_time), so
    `.response_time` is the cumulative model time and `.elapsed_time` the prefill time (:53-58,201).
What changes is underneath: the KV is appended in place (no torch.cat per step) and every forward runs
pc_rope_append + pc_prefill_attn / pc_decode_attn.  Greedy decoding without a repetition penalty additionally
samples ON THE DEVICE: the argmax of a step feeds the next step's graph without going
through the host, and the host reads the pending token ids only where the reference loop would yield
(`i % stream_interval == 0`, the last step) -- same tokens, same texts, same per-step event times, fewer syncs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Generator, List, Optional

import torch

from .cache_engine import KVCache


@dataclass
class GenerationParameters:
    temperature: float = 1.0
    repetition_penalty: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    max_new_tokens: int = 256
    stop_token_ids: List[int] = field(default_factory=lambda: [])
    stop_str: List[str] = field(default_factory=lambda: [])
    echo: bool = True

    def get_logits_processor(self):
        from transformers.generation.logits_process import (LogitsProcessorList, RepetitionPenaltyLogitsProcessor,
                                                            TemperatureLogitsWarper, TopKLogitsWarper,
                                                            TopPLogitsWarper)
        procs = LogitsProcessorList()
        if self.temperature >= 1e-5 and self.temperature != 1.0:
            procs.append(TemperatureLogitsWarper(self.temperature))
        if self.repetition_penalty > 1.0:
            procs.append(RepetitionPenaltyLogitsProcessor(self.repetition_penalty))
        if 1e-8 <= self.top_p < 1.0:
            procs.append(TopPLogitsWarper(self.top_p))
        if self.top_k > 0:
            procs.append(TopKLogitsWarper(self.top_k))
        return procs


def is_partial_stop(output: str, stop_str: str) -> bool:
    """True when the tail of `output` could still grow into `stop_str` (generation_engine.py:45-50)."""
    for i in range(0, min(len(output), len(stop_str))):
        if stop_str.startswith(output[-i:]):
            return True
    return False


@dataclass
class Output:
    text: str
    new_text: str
    response_time: float = 0.0
    elapsed_time: float = 0.0


class GenerationEngine:
    def __init__(self, lm):
        self.lm = lm
        self.verbose = False
        self._events = None  # one (start, end) pair, reused: the loop is synchronous
        self._pool = []      # timing events of the device-sampling loop
        self.device_sampling = True  # greedy decoding keeps the sampled token on the device between yields

    def _forward(self, ids: List[List[int]], pos: List[List[int]], past, device):
        extra = {"max_position": max(pos[0])} if getattr(self.lm.hf_model, "accepts_max_position", False) else {}
        if len(ids[0]) == len(pos[0]):  # one host->device copy for both id lists
            both = torch.tensor([ids[0], pos[0]], device=device, dtype=torch.long)
            input_ids, position_ids = both[0:1], both[1:2]
        else:  # full position ids (MPT): cached + new positions
            input_ids = torch.tensor(ids, device=device, dtype=torch.long)
            position_ids = torch.tensor(pos, device=device, dtype=torch.long)
        if self._events is None:
            self._events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        start, end = self._events
        start.record()
        out = self.lm(input_ids=input_ids, position_ids=position_ids, past_key_values=past, use_cache=True, **extra)
        end.record()
        end.synchronize()
        return out, start.elapsed_time(end)

    def _device_sampling_ok(self, params: GenerationParameters, use_full_position_ids: bool) -> bool:
        greedy = params.temperature < 1e-5 or params.top_p < 1e-8
        # top-p / top-k warpers mask tokens but never move the argmax; a repetition penalty does
        return (self.device_sampling and greedy and params.repetition_penalty <= 1.0 and not use_full_position_ids
                and not self.verbose and getattr(self.lm.hf_model, "accepts_max_position", False))

    def _generate_greedy_on_device(self, token_ids, position_ids, params, cache, stream_interval):
        """The loop of `generate` for greedy decoding, with the sampled token kept on the device between yields."""
        device = self.lm.device
        output_ids = list(token_ids)
        new_output_ids: List[int] = []
        position_offset = max(position_ids) + 1
        n_max = params.max_new_tokens
        # decode step i >= 1 runs at position position_offset + i (one position is skipped, reference :82,132)
        pos_all = None  # built at the first decode step: a one-token request (the TTFT measurement) never needs it
        need = 2 * (max(1, stream_interval) + 1)
        if len(self._pool) < need:  # timing events are reused across requests: the loop is synchronous at every yield
            self._pool += [torch.cuda.Event(enable_timing=True) for _ in range(need - len(self._pool))]
        pool = self._pool
        pending = []  # (token on the device, start event, end event) of the steps the host has not looked at yet
        past = tok_dev = None
        inference_time = response_time = 0.0
        first = True
        for i in range(n_max):
            start, end = pool[2 * len(pending)], pool[2 * len(pending) + 1]
            if past is None:
                if cache is not None:
                    batched = getattr(cache, "batched", None)
                    cache = batched() if batched is not None else [(k.unsqueeze(0), v.unsqueeze(0)) for k, v in cache]
                both = torch.tensor([list(token_ids), list(position_ids)], device=device, dtype=torch.long)
                ids_t, pos_t, past_in, max_pos = both[0:1], both[1:2], cache, max(position_ids)
            else:
                if pos_all is None:
                    pos_all = torch.arange(position_offset, position_offset + n_max + 1, device=device, dtype=torch.long).view(1, -1)
                ids_t, pos_t, past_in, max_pos = tok_dev.view(1, 1), pos_all[:, i:i + 1], past, position_offset + i
            start.record()
            # the argmax below is taken before the next forward: a graph replay may return its static logits buffer
            out = self.lm(input_ids=ids_t, position_ids=pos_t, past_key_values=past_in, use_cache=True, max_position=max_pos,
                          static_logits_ok=True)
            end.record()
            past = out.past_key_values
            tok_dev = torch.argmax(out.logits[0, -1])
            pending.append((tok_dev, start, end))
            if not (i % stream_interval == 0 or i == n_max - 1):
                continue  # the reference loop would not yield here (unless this token stops it: found at the next flush)
            # one device->host read for all pending steps (it also waits for the last step's end event)
            ids = [int(pending[0][0])] if len(pending) == 1 else torch.stack([p[0] for p in pending]).tolist()
            stopped = False
            for (_, s_ev, e_ev), tid in zip(pending, ids):
                inference_time += s_ev.elapsed_time(e_ev)
                if first:
                    response_time, first = inference_time, False
                output_ids.append(tid)
                new_output_ids.append(tid)
                if tid in params.stop_token_ids:  # steps run past a stop token are dropped: their rows lie behind the sequence
                    stopped = True
                    break
            pending = []
            output = self.lm.decode(output_ids)
            new_output = self.lm.decode(new_output_ids)
            partially_stopped = False
            for each_stop in params.stop_str:
                pos_found = new_output.rfind(each_stop, 0)
                if pos_found != -1:
                    new_output = new_output[:pos_found]
                    stopped = True
                    break
                partially_stopped = is_partial_stop(output, each_stop)
                if partially_stopped:
                    break
            if not partially_stopped:
                yield Output(output, new_output, inference_time, response_time)
            if stopped:
                break
        del past  # the KV lives in the engine's buffers: nothing to collect (the reference runs gc.collect() here, :207-209)

    @torch.inference_mode()
    def generate(self, token_ids: List[int], position_ids: List[int], params: GenerationParameters,
                 cache: Optional[KVCache] = None, stream_interval: int = 2,
                 use_full_position_ids: bool = False) -> Generator[Output, None, None]:
        if self._device_sampling_ok(params, use_full_position_ids) and params.max_new_tokens > 0:
            yield from self._generate_greedy_on_device(token_ids, position_ids, params, cache, stream_interval)
            return
        logits_processor = params.get_logits_processor()
        output_ids = list(token_ids)
        new_output_ids: List[int] = []
        device = self.lm.device
        position_offset = max(position_ids) + 1
        position_ids_og = list(position_ids)
        past_key_values = None
        new_token_id = 0
        inference_time = 0.0
        response_time = 0.0

        for i in range(params.max_new_tokens):
            if past_key_values is None:  # prefill of the uncached suffix over the spliced cache
                if cache is not None:  # add the batch dimension (reference :101-102); O(1) for the engine's own views
                    batched = getattr(cache, "batched", None)
                    cache = batched() if batched is not None else [(k.unsqueeze(0), v.unsqueeze(0)) for k, v in cache]
                out, ms = self._forward([list(token_ids)], [position_ids_og], cache, device)
                inference_time += ms
                response_time = inference_time
                if self.verbose:
                    print(f"Prefill latency: {inference_time:.2f} ms")
            else:  # one-token decode
                if use_full_position_ids:
                    pos = position_ids_og + list(range(position_offset, position_offset + i))
                else:
                    pos = [position_offset + i]
                out, ms = self._forward([[new_token_id]], [pos], past_key_values, device)
                inference_time += ms
            logits = out.logits
            past_key_values = out.past_key_values

            tmp_output_ids = torch.as_tensor([output_ids], device=device) if params.repetition_penalty > 1.0 else None
            last_token_logits = logits_processor(tmp_output_ids, logits[:, -1, :])[0]
            if params.temperature < 1e-5 or params.top_p < 1e-8:  # greedy
                new_token_id = int(torch.argmax(last_token_logits))
            else:
                probs = torch.softmax(last_token_logits, dim=-1)
                new_token_id = int(torch.multinomial(probs, num_samples=1))
            output_ids.append(new_token_id)
            new_output_ids.append(new_token_id)

            stopped = new_token_id in params.stop_token_ids
            if i % stream_interval == 0 or i == params.max_new_tokens - 1 or stopped:
                output = self.lm.decode(output_ids)
                new_output = self.lm.decode(new_output_ids)
                partially_stopped = False
                for each_stop in params.stop_str:
                    pos_found = new_output.rfind(each_stop, 0)
                    if pos_found != -1:
                        new_output = new_output[:pos_found]
                        stopped = True
                        break
                    partially_stopped = is_partial_stop(output, each_stop)
                    if partially_stopped:
                        break
                if not partially_stopped:
                    yield Output(output, new_output, inference_time, response_time)
            if stopped:
                break

        del past_key_values  # no gc.collect(): the KV lives in the engine's buffers (reference :207-209 collects, ~45 ms)
