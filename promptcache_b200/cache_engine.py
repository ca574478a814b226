"""B200 prompt-cache engine behind the reference's API (promptcache/cache_engine.py).

  TokenSequenceCache   one cached segment; its K/V for ALL layers is ONE HBM allocation
                       [L, 2, Hkv, len, D] fp16 (reference :50-83 keeps L*2 host tensors and re-uploads them
                       on every request, :290-292)
  PromptCache          the contiguous cache, ONE allocation [L, 2, Hkv, max_ctx, D]; `update()` is the splice
                       (reference :116-156: 2*L*M Python-driven slice copies) = one pc_splice_kv launch
  SchemaCache          encodes a schema's scaffolds and slices per-TokenSequence segments (reference :186-307)
  CacheEngine          add_schema / process / remove_schema with the reference's signatures and ValueErrors
                       (reference :331-522)

Deliberate deviations from the reference, both called out in SURVEY.md section 7:
  * `PromptCache.staged` records the layout actually written (the usage-ordered list), so the retained-prefix
    shortcut can never hand back stale rows (the reference stores the unsorted list, :155);
  * zero-length token sequences are skipped instead of raising in `list.index` (:278).
"""
from __future__ import annotations

import gc
import itertools
from typing import Dict, List, Optional, Tuple, Union

import torch

from . import _native as nat
from .model._kv import KVBuffer, PastKV, StaleKVError, VirtualKV
from .prompt import ModuleRef, Prompt
from .schema import Module, Path, Schema, TokenSequence, UnionModule

KVCache = List[Tuple[torch.Tensor, torch.Tensor]]


class SplicedKV(list):
    """What `CacheEngine.process` returns as `KVCache`: a plain list of L `(K, V)` views `[Hkv, N, D]` into the
    engine-owned buffers (reference cache_engine.py:161-165) that also remembers those buffers, so that
    `GenerationEngine` can hand them to the model in O(1) instead of re-deriving 2L views and re-validating them on
    every request (that host work sat between the splice and the prefill with the GPU idle)."""

    kv_buffer: Optional[KVBuffer] = None
    rows: int = 0
    epoch: int = 0  # of the buffer when process() handed the views out: a later process() makes them stale
    virtual: Optional[VirtualKV] = None  # virtual splice: the rows the views show are copied in only when somebody looks

    def batched(self) -> PastKV:
        """The `[(k.unsqueeze(0), v.unsqueeze(0)) ...]` of generation_engine.py:101-102, materialised lazily."""
        return PastKV(self.kv_buffer, self.rows, self.epoch, virtual=self.virtual)

    def _touch(self) -> None:
        if self.virtual is not None and not self.virtual.done:
            self.virtual.materialise()

    # every way of reading the (K, V) views makes the rows real first; len() and batched() do not
    def __getitem__(self, i):
        self._touch()
        return list.__getitem__(self, i)

    def __iter__(self):
        self._touch()
        return list.__iter__(self)

    def __reversed__(self):
        self._touch()
        return list.__reversed__(self)

    def copy(self):
        self._touch()
        return list.copy(self)

    def __add__(self, other):
        self._touch()
        return list.__add__(self, other)


def pad_batch(batch_list: List[List[int]], pad_id: int) -> Tuple[List[List[int]], List[List[int]]]:
    """Right-pad id lists to a rectangle and return the matching 0/1 mask (cache_engine.py:38-47)."""
    width = max(map(len, batch_list))
    padded = [row + [pad_id] * (width - len(row)) for row in batch_list]
    mask = [[1] * len(row) + [0] * (width - len(row)) for row in batch_list]
    return padded, mask


class TokenSequenceCache:
    """KV of one TokenSequence, resident in HBM as a single packed allocation."""

    def __init__(self, seq: TokenSequence, packed: torch.Tensor, arena: Optional[torch.Tensor] = None, row0: int = 0):
        """`packed` is [L,2,Hkv,len,D]: a contiguous allocation of its own, or the row slice
        `arena[:, :, :, row0:row0+len, :]` of the schema's segment arena (SchemaCache)."""
        if packed.dim() != 5 or packed.shape[1] != 2:
            raise ValueError("a segment must be a [L,2,Hkv,len,D] tensor")
        if arena is None:
            if not packed.is_contiguous():
                raise ValueError("a stand-alone segment must be a contiguous [L,2,Hkv,len,D] tensor")
            arena, row0 = packed, 0
        elif packed.data_ptr() != arena[0, 0, 0, row0].data_ptr() or packed.stride() != arena.stride():
            raise ValueError("segment is not a row slice of its arena")
        self.token_sequence = seq
        self.packed = packed
        self.arena, self.row0 = arena, row0
        self.usage_counter = 0

    def inc_usage_counter(self) -> None:
        self.usage_counter += 1

    # kept for API compatibility: segments are always device-resident here
    def upload(self, device=None) -> None:
        return None

    def free(self) -> None:
        return None

    @property
    def cache(self) -> KVCache:
        """Per-layer (K, V) views [Hkv, len, D], the shape the reference stores (:287-288)."""
        return [(self.packed[l, 0], self.packed[l, 1]) for l in range(self.packed.shape[0])]

    host_cache = cache
    device_cache = cache

    def __len__(self) -> int:
        return int(self.packed.shape[3])


class PromptCache:
    """Contiguous per-layer KV buffers + the splice."""

    def __init__(self, max_ctx_length: int, num_layers: int, num_head: int, head_dim: int, target_device):
        device = torch.device(target_device)
        if device.type != "cuda":
            raise nat.NativeError("promptcache_b200.PromptCache lives in B200 HBM; target_device must be CUDA "
                                  "(there is no CPU path)")
        self.max_ctx_length = max_ctx_length
        self.num_layers = num_layers
        self.num_head = num_head
        self.head_dim = head_dim
        # zero-initialised: rows past `length` are read (and masked) by whole-tile attention loads
        self.buffer = torch.zeros(num_layers, 2, num_head, max_ctx_length, head_dim, dtype=torch.half, device=device)
        self._bases_k = [self.buffer[l, 0] for l in range(num_layers)]  # [Hkv, max_ctx, D] per layer, built once
        self._bases_v = [self.buffer[l, 1] for l in range(num_layers)]
        # append bookkeeping shared with the models: update() bumps its epoch, so a generation still running on the
        # previous splice (or a `cache` kept from an earlier process()) raises StaleKVError instead of reading rows the
        # new splice overwrote -- the reference's forward would hold a private torch.cat copy (llama2.py:361-364)
        self.kv = KVBuffer(self._bases_k, self._bases_v, max_ctx_length, storage=self.buffer)
        self._views: Dict[int, SplicedKV] = {}                          # spliced length -> the 2L views of that length
        self.staged: List[TokenSequenceCache] = []    # the layout of rows [0, length), in row order
        self._copied: List[TokenSequenceCache] = []   # the prefix of a layout the buffer physically holds
        self.length = 0
        self.retain_prefix = True
        self.last_spliced_rows = 0
        # virtual splice: per-request tile list (int32 [T,2]: first arena row, valid rows), fixed device address so that a
        # captured forward graph serves every request with the same tile / row counts
        cap = max_ctx_length // 128 + 1024
        # pinned staging slots, each guarded by the event of the copy that last read it (the host may run ahead of the GPU)
        self._tile_slots = [[torch.zeros(cap, 2, dtype=torch.int32).pin_memory(), None] for _ in range(4)]
        self._tile_slot = 0
        self._tiles_key, self._tiles_count = None, {}  # which layout the device tile list currently describes
        self._tiles_dev = torch.zeros(cap, 2, dtype=torch.int32, device=device)
        # buffer row of every tile slot (built only for models that carry per-key position ids: MPT's ALiBi)
        self._slot_rows_dev: Optional[torch.Tensor] = None
        self._slot_rows_pin: Optional[torch.Tensor] = None
        self._slots_key = None
        self.virtual: Optional[VirtualKV] = None

    @property
    def device_cache(self) -> KVCache:
        return [(self.buffer[l, 0], self.buffer[l, 1]) for l in range(self.num_layers)]

    def invalidate(self) -> None:
        """Forget the staged layout so the next update() moves every segment."""
        self.staged = []
        self._copied = []
        self.length = 0
        self.virtual = None
        self._tiles_key = None
        self._slots_key = None
        self.kv.rewritten(0)

    @torch.inference_mode()
    def update(self, modules: List[TokenSequenceCache], virtual: bool = False, slot_rows: bool = False) -> None:
        """Lay out `modules` as the cached rows [0, N) (reference :116-156).

        virtual=False: the rows are copied into the contiguous buffer now (one pc_splice_kv launch; with
        `retain_prefix` only the segments behind the longest common prefix with what the buffer already holds).
        virtual=True ("virtual splice"): nothing is copied.  When every segment lives in ONE arena the
        request's tile list {first arena row, valid rows} goes to the device (one small pinned copy) and the short-suffix
        attention kernel reads the segments in place; the copy is made later only if somebody asks for the contiguous
        rows (`materialise`).  Either way rows >= N of the buffer receive the suffix / decode rows in place."""
        # most-used segments first (stable), as the reference lays them out (:121)
        ordered = sorted(modules, key=lambda e: e.usage_counter, reverse=True)
        row = sum(len(m) for m in ordered)
        if row > self.max_ctx_length:
            raise ValueError(f"spliced prompt needs {row} rows but max_ctx_length is {self.max_ctx_length}")
        # what the buffer physically holds stays valid only as far as the new layout agrees with it (the rows behind are
        # about to receive another request's segments or appended tokens)
        kept = 0
        if self.retain_prefix:
            for cur, prev in zip(ordered, self._copied):
                if cur is not prev:
                    break
                kept += 1
        self._copied = self._copied[:kept]
        self.staged = ordered
        self.length = row
        self.kv.rewritten(row)  # rows >= `row` are scratch for the request being served; older pasts are stale now
        self.virtual = None
        self.last_spliced_rows = 0
        if virtual and ordered and self._upload_tiles(ordered, slot_rows):
            return
        self.materialise()

    def _upload_tiles(self, ordered: List[TokenSequenceCache], want_slot_rows: bool = False) -> bool:
        arena = ordered[0].arena
        if any(m.arena is not arena for m in ordered) or arena.device != self.buffer.device or arena.dtype != self.buffer.dtype \
                or tuple(arena.shape[:3]) != tuple(self.buffer.shape[:3]) or arena.shape[4] != self.head_dim or not arena.is_contiguous():
            return False  # segments from different allocations: the copy is made instead
        # Cached keys are visible to every query row, so their ORDER is irrelevant to attention: the tile list walks the
        # used segments in arena order and lets a tile run across segments that are neighbours in the arena (C2: the seven
        # segments of the schema are one contiguous run of 2048 rows = 16 tiles instead of 22 partly empty ones).
        runs: List[List[int]] = []
        for r0, n in sorted((m.row0, len(m)) for m in ordered):
            if runs and runs[-1][0] + runs[-1][1] == r0:
                runs[-1][1] += n
            else:
                runs.append([r0, n])
        key = (arena.data_ptr(),) + tuple(map(tuple, runs))  # what the tile list is a function of
        t = self._tiles_count.get(key) if key == self._tiles_key else None
        if t is None:  # the device list holds another request's tiles: build this layout's list and upload it
            tiles = []
            for r0, n in runs:
                for r in range(0, n, 128):
                    tiles.append((r0 + r, min(128, n - r)))
            if len(tiles) > self._tiles_dev.shape[0]:
                return False
            t = len(tiles)
            slot = self._tile_slots[self._tile_slot]
            self._tile_slot = (self._tile_slot + 1) % len(self._tile_slots)
            if slot[1] is not None:
                slot[1].synchronize()
            slot[0][:t] = torch.tensor(tiles, dtype=torch.int32)
            self._tiles_dev[:t].copy_(slot[0][:t], non_blocking=True)  # stream-ordered before the forward that reads it
            if slot[1] is None:
                slot[1] = torch.cuda.Event()
            slot[1].record()
            self._tiles_key, self._tiles_count = key, {key: t}
        slot_rows = None
        if want_slot_rows:
            # the buffer row each tile slot would occupy after the copy: full position ids come in buffer order (usage-ordered
            # segments), the kernel walks the keys in tile (arena) order
            skey = key + tuple((m.row0, len(m)) for m in ordered)
            if self._slot_rows_dev is None:
                cap = self._tiles_dev.shape[0] * 128
                self._slot_rows_dev = torch.zeros(cap, dtype=torch.int64, device=self.buffer.device)
                self._slot_rows_pin = torch.zeros(cap, dtype=torch.int64).pin_memory()
            if skey != self._slots_key:
                import numpy as np
                starts = np.array([m.row0 for m in ordered], dtype=np.int64)
                lens = np.array([len(m) for m in ordered], dtype=np.int64)
                offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
                order = np.argsort(starts)
                starts, lens, offs = starts[order], lens[order], offs[order]
                table = np.zeros(t * 128, dtype=np.int64)
                ti = 0
                for r0, n in runs:
                    for r in range(0, n, 128):
                        rows = np.arange(r0 + r, r0 + r + min(128, n - r), dtype=np.int64)
                        seg = np.searchsorted(starts, rows, side="right") - 1
                        table[ti * 128: ti * 128 + rows.shape[0]] = offs[seg] + rows - starts[seg]
                        ti += 1
                torch.cuda.current_stream().synchronize()  # the pinned table may still be read by the previous upload
                self._slot_rows_pin[:t * 128] = torch.from_numpy(table)
                self._slot_rows_dev[:t * 128].copy_(self._slot_rows_pin[:t * 128], non_blocking=True)
                self._slots_key = skey
            slot_rows = self._slot_rows_dev[:t * 128]
        epoch = self.kv.epoch

        def materialise():
            if self.kv.epoch != epoch:
                raise StaleKVError("the prompt cache was re-spliced after this cache was handed out")
            self.materialise()

        self.virtual = VirtualKV(arena, self._tiles_dev, t, self.length, materialise, slot_rows)
        return True

    @torch.inference_mode()
    def materialise(self) -> None:
        """Copy every staged segment the buffer does not hold yet (the splice proper)."""
        kept = len(self._copied)
        row = sum(len(m) for m in self.staged[:kept])
        plan = []
        for m in self.staged[kept:]:
            plan.append((m.packed, row))
            row += len(m)
        if plan:
            nat.splice_kv(plan, self.buffer)
        self.last_spliced_rows += sum(int(p.shape[3]) for p, _ in plan)
        self._copied = list(self.staged)
        if self.virtual is not None:
            self.virtual.done = True

    def __len__(self) -> int:
        return self.length

    @property
    def cache(self) -> KVCache:
        """Views [Hkv, N, D] of the spliced rows (reference :161-165).  The 2L view objects of a given length are
        built once and handed out as a fresh list each time (callers may mutate the list, not the engine's copy)."""
        n = self.length
        master = self._views.get(n)
        if master is None:
            if len(self._views) >= 64:
                self._views.pop(next(iter(self._views)))
            master = SplicedKV((self.buffer[l, 0, :, :n, :], self.buffer[l, 1, :, :n, :]) for l in range(self.num_layers))
            self._views[n] = master
        out = SplicedKV(master)
        out.kv_buffer, out.rows, out.epoch, out.virtual = self.kv, n, self.kv.epoch, self.virtual
        return out


class SchemaCache:
    """Encoded schema: id(TokenSequence) -> TokenSequenceCache."""

    def __init__(self, schema: Schema, lm, batch_size: int = 1, target_device=None, no_cache: bool = False):
        self.schema = schema
        self.lm = lm
        self.cache_l1: Dict[int, TokenSequenceCache] = {}
        self.cache_l2: Dict[Tuple[int, int], Tuple[TokenSequenceCache, TokenSequenceCache]] = {}
        self.target_device = lm.device if target_device is None else target_device
        self.arena: Optional[torch.Tensor] = None  # [L,2,Hkv,R,D]: every token sequence of the schema, one allocation
        if not no_cache:
            self._process(batch_size)

    def layout(self):
        """Encode plan + arena layout, derived from the schema alone (identical on every rank): the list of
        (path, scaffold, token sequences to store) per scaffold path, the first arena row of every sequence, and the
        total number of rows."""
        plan, rows, r = [], {}, 0
        for path in self.scaffold_paths():
            scaffold = self.schema.get_scaffold(path)
            target = [seq for seq in scaffold.select(path).all_token_sequences() if len(seq) > 0]
            for seq in target:
                if id(seq) not in rows:
                    rows[id(seq)] = r
                    r += len(seq)
            plan.append((path, scaffold, target))
        return plan, rows, r

    def allocate_arena(self, total_rows: int) -> torch.Tensor:
        L, hkv, d = self.lm.get_cache_shape()
        self.arena = torch.empty(L, 2, hkv, max(total_rows, 1), d, dtype=torch.half, device=self.target_device)
        return self.arena

    def adopt(self, seq: TokenSequence, row0: int) -> TokenSequenceCache:
        """Register rows [row0, row0+len) of the arena as the cache of `seq`."""
        view = self.arena[:, :, :, row0:row0 + len(seq), :]
        entry = TokenSequenceCache(seq, view, arena=self.arena, row0=row0)
        self.cache_l1[id(seq)] = entry
        return entry

    def scaffold_paths(self) -> List[Path]:
        """Default scaffold + one path per non-default union member, in the reference's order (:188-210)."""
        paths = [Path()]
        pending = [([], True, self.schema)] if self.schema.contains_union() else []
        while pending:
            prefix, parent_default, node = pending.pop()
            here = prefix + [node.name]
            for child in node.children:
                if type(child) is Module and child.contains_union():
                    pending.append((here, parent_default, child))
                elif type(child) is UnionModule:
                    for member in child.modules:
                        is_default = child.scaffold_name == member.name and parent_default
                        if member.contains_union():
                            pending.append((here, is_default, member))
                        if not is_default:
                            paths.append(Path(here + [member.name]).next)
        return paths

    reuse_prefix = True  # encode non-default union members over the default scaffold's KV prefix

    def _forward(self, token_ids: List[int], position_ids: List[int], past=None):
        device = self.lm.device
        extra = {"max_position": max(position_ids)} if getattr(self.lm.hf_model, "accepts_max_position", False) else {}
        return self.lm(input_ids=torch.tensor([token_ids], device=device, dtype=torch.long),
                       position_ids=torch.tensor([position_ids], device=device, dtype=torch.long),
                       attention_mask=torch.ones(1, len(token_ids), device=device, dtype=torch.float16),
                       past_key_values=past, use_cache=True, **extra)

    batch_token_budget = 4096  # tokens per batched member forward when `batch_size` leaves the choice to the engine

    def _store(self, kv, position_ids: List[int], target, rows) -> None:
        """Slice the per-TokenSequence segments out of a path's KV (row i holds scaffold token i) into the arena."""
        for seq in target:
            st = position_ids.index(seq.offset)  # offsets and row order can differ (:275-279)
            ed = st + len(seq)
            entry = self.adopt(seq, rows[id(seq)])  # straight into the schema's arena: no per-segment allocation
            for l in range(len(kv)):
                entry.packed[l, 0].copy_(self.lm.store_k_hook(kv[l][0])[0, :, st:ed, :])
                entry.packed[l, 1].copy_(self.lm.store_v_hook(kv[l][1])[0, :, st:ed, :])

    @torch.inference_mode()
    def _process(self, batch_size: int = 1) -> None:
        """Encode every scaffold path and store its token sequences (reference :212-304).

        The reference re-runs the WHOLE scaffold for every non-default union member, right-padded into batches of
        `batch_size` paths (pad_batch, :228-248).  Attention is causal, so the KV of a member only depends on the tokens
        before it: with `reuse_prefix` a member's path is encoded as a cached prefill over the longest token/position
        prefix it shares with the default scaffold and stops at the last token the path has to store -- prompt caching
        applied to its own encoder.  Those member prefills are BATCHED: the tokens of up to `batch_size`
        members (or `batch_token_budget` tokens when batch_size is 1) go through the model's linears as one
        concatenated sequence -- no padding -- and only RoPE+append / attention run per member, each over its own copy
        of the shared prefix (`forward_many`)."""
        full_ids = not getattr(self.lm, "use_full_position_ids", False)  # MPT wants cached + new ids: keep it simple
        many = getattr(self.lm.hf_model, "forward_many", None) if self.reuse_prefix and full_ids else None
        base_ids = base_pos = base_kv = None
        plan, rows, total = self.layout()
        self.allocate_arena(total)
        batch: List[tuple] = []  # (token ids, position ids, first, hi, target) of member paths waiting for one forward

        def flush():
            if not batch:
                return
            dev = self.lm.device
            items = [(torch.tensor(ids[first:hi], device=dev, dtype=torch.long), torch.tensor(pos[first:hi], device=dev, dtype=torch.long),
                      [(k[:, :, :first], v[:, :, :first]) for k, v in base_kv]) for ids, pos, first, hi, _ in batch]
            for (ids, pos, first, hi, target), kv in zip(batch, many(items, max_position=max(max(b[1]) for b in batch))):
                self._store(kv, pos, target, rows)
            batch.clear()

        for path, scaffold, target in plan:
            token_ids = scaffold.token_ids()
            position_ids = scaffold.position_ids()
            if len(token_ids) == 0:
                continue
            first = 0
            if self.reuse_prefix and full_ids and base_kv is not None and target:
                common = 0
                for a, b, c, d in zip(token_ids, base_ids, position_ids, base_pos):
                    if a != b or c != d:
                        break
                    common += 1
                starts = [position_ids.index(seq.offset) for seq in target]
                lo, hi = min(starts), max(st + len(seq) for st, seq in zip(starts, target))
                first = min(common, lo)
                if first > 0 and hi > first:
                    if many is not None:
                        batch.append((token_ids, position_ids, first, hi, target))
                        limit = batch_size if batch_size > 1 else 1 << 30
                        if len(batch) >= limit or sum(b[3] - b[2] for b in batch) >= self.batch_token_budget:
                            flush()
                        continue
                    past = [(k[:, :, :first].contiguous(), v[:, :, :first].contiguous()) for k, v in base_kv]
                    out = self._forward(token_ids[first:hi], position_ids[first:hi], past)
                else:
                    first = 0
                    out = self._forward(token_ids, position_ids)
            else:
                out = self._forward(token_ids, position_ids)
            kv = out.past_key_values  # rows [0, first) are the reused prefix, row i holds scaffold token i
            if base_kv is None:
                base_ids, base_pos, base_kv = token_ids, position_ids, kv
            self._store(kv, position_ids, target, rows)
            del out, kv
        flush()
        del base_kv
        gc.collect()

    def get_cache_l1(self, seq: TokenSequence) -> Optional[TokenSequenceCache]:
        return self.cache_l1.get(id(seq))

    def get_cache_l2(self, seq1: TokenSequence, seq2: TokenSequence):
        key = (max(id(seq1), id(seq2)), min(id(seq1), id(seq2)))
        return self.cache_l2.get(key)


class CacheEngine:
    def __init__(self, max_ctx_length: int, lm, target_device=None):
        self.lm = lm
        self.schemas: Dict[str, SchemaCache] = {}
        self.target_device = lm.device if target_device is None else target_device
        num_layers, num_head, head_dim = lm.get_cache_shape()
        self.prompt_cache = PromptCache(max_ctx_length=max_ctx_length, num_layers=num_layers, num_head=num_head,
                                        head_dim=head_dim, target_device=self.target_device)
        if hasattr(lm.hf_model, "ensure_rope"):
            lm.hf_model.ensure_rope(max_ctx_length)
        self._events = None
        self.virtual_splice = True  # False: PromptCache.update always copies (the reference's behaviour)
        from .model import LanguageModel
        self._identity_hooks = all(getattr(getattr(lm, name, None), "__func__", None) is getattr(LanguageModel, name)
                                   for name in ("read_k_hook", "read_v_hook"))

    def add_schema(self, schema: Union[str, Schema], batch_size: int = 1, max_tokens: Optional[int] = None,
                   no_cache: bool = False) -> None:
        if isinstance(schema, str):
            schema = Schema(schema, self.lm, max_tokens=max_tokens)
        if schema.name in self.schemas:
            raise ValueError(f"There is already a schema named {schema.name} in the cache")
        self.schemas[schema.name] = SchemaCache(schema, self.lm, batch_size, target_device=self.target_device,
                                                no_cache=no_cache)

    def get_schema(self, name: str) -> Optional[Schema]:
        entry = self.schemas.get(name)
        return None if entry is None else entry.schema

    def remove_schema(self, name: str) -> None:
        if name not in self.schemas:
            raise ValueError(f"There is no such schema named {name} in the cache")
        del self.schemas[name]
        self.prompt_cache.invalidate()
        gc.collect()
        torch.cuda.empty_cache()

    def remove_all_schemas(self) -> None:
        self.schemas = {}
        self.prompt_cache.invalidate()
        gc.collect()
        torch.cuda.empty_cache()

    def plan(self, prompt: Prompt, collect_original: bool):
        """Host-side request plan (reference :396-474): which cached sequences are used, in visit order, and
        the uncached suffix (parameter arguments at their slots, then trailing text after the schema)."""
        if prompt.schema not in self.schemas:
            raise ValueError(f"There is no such layout named {prompt.schema} in the cache")
        cached = self.schemas[prompt.schema]
        schema = cached.schema
        used: List[TokenSequence] = []
        orig_ids: List[List[int]] = []
        orig_pos: List[List[int]] = []
        arg_ids: List[List[int]] = []
        arg_pos: List[List[int]] = []
        # explicit LIFO stack => sibling modules are visited in reverse document order (:411-455)
        todo: List[Tuple[ModuleRef, Module]] = [(prompt, schema)]
        while todo:
            ref, module = todo.pop()
            for seq in module.token_sequences():
                used.append(seq)
                if collect_original:
                    orig_ids.append(seq.token_ids())
                    orig_pos.append(seq.position_ids())
            params = module.parameters()
            for arg in ref.args:
                param = next((p for p in params if p.name == arg.name), None)
                if param is None:
                    raise ValueError(f"There is no such parameter named {arg.name} in the module {module.name}")
                ids = self.lm.encode(arg.value)
                if len(ids) > param.length:
                    raise ValueError(f"The argument {arg.name} is too long. It should be at most {param.length} "
                                     f"characters long")
                arg_ids.append(ids)
                arg_pos.append(param.position_ids()[:len(ids)])
            for sub in ref.modules:
                target = module.select(sub.name)
                if target is None:
                    raise ValueError(f"There is no such module named @{sub.name} in the module @{module.name}")
                todo.append((sub, target))
        if len(prompt.text) > 0:
            text_ids = self.lm.encode(prompt.text)
            arg_ids.append(text_ids)
            arg_pos.append(list(range(len(schema), len(schema) + len(text_ids))))
        input_ids = list(itertools.chain(*arg_ids))
        position_ids = list(itertools.chain(*arg_pos))
        return cached, used, orig_ids, orig_pos, input_ids, position_ids

    def process(self, prompt: Prompt, no_cache: bool = False, return_full_position_ids: bool = False) -> Tuple[
            List[int], List[int], float, Optional[KVCache]]:
        if self._events is None:  # one pair, reused: process() is synchronous
            self._events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        start, end = self._events
        start.record()
        cached, used, orig_ids, orig_pos, input_ids, position_ids = self.plan(
            prompt, collect_original=no_cache or return_full_position_ids)

        if no_cache:
            all_pos = list(itertools.chain(*orig_pos)) + position_ids
            all_ids = list(itertools.chain(*orig_ids)) + input_ids
            pairs = sorted(zip(all_pos, all_ids))
            sorted_ids = tuple(tok for _, tok in pairs)
            end.record()
            end.synchronize()
            # positions are renumbered 0..n-1 (reference :491-493)
            return sorted_ids, list(range(len(pairs))), start.elapsed_time(end), None

        seq_caches: List[TokenSequenceCache] = []
        for seq in used:
            entry = cached.get_cache_l1(seq)
            if entry is None:
                if len(seq) == 0:
                    continue
                raise ValueError(f"token sequence at offset {seq.offset} of schema {prompt.schema} is not cached")
            entry.inc_usage_counter()
            seq_caches.append(entry)
        # virtual splice when the model's short-suffix attention can read the segments in place (decided per request:
        # the suffix length is known here); otherwise the rows are copied as the reference does
        max_q = getattr(self.lm.hf_model, "virtual_kv_max_q", lambda: 0)()
        # models that take the ids of ALL keys (MPT's ALiBi) can read the segments in place too: they get the tile-slot -> buffer
        # row table and put their ids in tile order themselves
        full_ok = bool(getattr(self.lm.hf_model, "virtual_kv_full_ids", False))
        virtual = self.virtual_splice and self._identity_hooks and (full_ok or not return_full_position_ids) \
            and 0 < len(input_ids) <= max_q
        self.prompt_cache.update(seq_caches, virtual=virtual, slot_rows=bool(return_full_position_ids and full_ok))
        cache = self.prompt_cache.cache
        end.record()
        end.synchronize()
        cache_time = start.elapsed_time(end)
        if not self._identity_hooks:  # the reference's hooks are identities (model/__init__.py:116-126)
            cache = [(self.lm.read_k_hook(k), self.lm.read_v_hook(v)) for k, v in cache]
        if return_full_position_ids:
            # ids of the cached rows IN BUFFER ORDER (equal to the reference's visit order, :517-519, whenever
            # the usage counters do not reorder the layout; the reference is wrong when they do)
            laid_out = [m.token_sequence.position_ids() for m in self.prompt_cache.staged]
            position_ids = list(itertools.chain(*laid_out)) + position_ids
        return input_ids, position_ids, cache_time, cache
