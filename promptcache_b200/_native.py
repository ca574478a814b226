"""ctypes binding of libpromptcache_b200.so (the C ABI in include/promptcache_b200.h).

This is the ONLY way the Python host reaches the kernels; there is no CPU or eager-PyTorch fallback.
If the shared library is missing or a call fails, a `NativeError` is raised -- loudly.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import os
from typing import Optional, Sequence

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpromptcache_b200.so")

PC_ROUND_QK = 1
PC_ROUND_SCALED = 2
PC_ROUND_LLAMA = 3
PC_ROUND_MPT = 3
PC_ROUND_FALCON = 0
PC_ATTN_AUTO = 0
PC_ATTN_SIMT = 1
PC_ATTN_TCGEN05 = 2
PC_ATTN_GLOBAL_MERGE = 0x100

EXPORTS = ("pc_version", "pc_last_error", "pc_device_info", "pc_splice_kv", "pc_splice_kv_strided", "pc_splice_kv_tensors",
           "pc_rope_append", "pc_rope_append_dyn", "pc_attn_workspace_bytes", "pc_prefill_attn", "pc_decode_attn", "pc_launch_count",
           "pc_add_rmsnorm", "pc_add_layernorm", "pc_add_layernorm_partials", "pc_silu_mul", "pc_linear_workspace_bytes", "pc_linear_small",
           "pc_linear_tiled_bytes", "pc_linear_tile_weights", "pc_linear_partial_splits", "pc_add_rmsnorm_partials", "pc_linear_chain")
PC_CHAIN_LINEAR = 0
PC_CHAIN_ADD_RMSNORM = 1
PC_CHAIN_ADD_LAYERNORM = 2
PC_CHAIN_MAX_OPS = 8
PC_LIN_PLAIN = 0
PC_LIN_SILU_MUL = 1
PC_LIN_ROPE_APPEND = 2
PC_LIN_PARTIALS = 3


class NativeError(RuntimeError):
    pass


class pc_segment(C.Structure):
    _fields_ = [("base", C.c_void_p), ("len", C.c_int32), ("dst_row", C.c_int32)]


class pc_copy2d(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_run_stride", C.c_int64),
                ("dst_run_stride", C.c_int64), ("run_bytes", C.c_int64), ("n_runs", C.c_int32),
                ("reserved", C.c_int32)]


class pc_attn_params(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
                ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32),
                ("n_cached", C.c_int32), ("q_len", C.c_int32), ("max_ctx", C.c_int32),
                ("scale", C.c_float), ("rounding", C.c_int32),
                ("alibi_slopes", C.c_void_p), ("key_pos", C.c_void_p),
                ("alibi_max_pos", C.c_int32), ("impl", C.c_int32), ("n_cached_dev", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("kv_arena_k", C.c_void_p), ("kv_arena_v", C.c_void_p), ("kv_tiles", C.c_void_p),
                ("kv_arena_rows", C.c_int32), ("kv_n_tiles", C.c_int32), ("kv_rows", C.c_int32), ("kv_reserved", C.c_int32),
                ("prefetch", C.c_void_p), ("prefetch_bytes", C.c_size_t)]


class pc_linear_params(C.Structure):
    _fields_ = [("x", C.c_void_p), ("w", C.c_void_p), ("x_stride", C.c_int64),
                ("q_len", C.c_int32), ("k", C.c_int32), ("n_rows", C.c_int32), ("epilogue", C.c_int32),
                ("out", C.c_void_p), ("out_stride", C.c_int64),
                ("position_ids", C.c_void_p), ("cos", C.c_void_p), ("sin", C.c_void_p),
                ("table_rows", C.c_int32), ("max_ctx", C.c_int32),
                ("q_out", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
                ("dst_row0", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("activation", C.c_int32),
                ("dst_row0_dev", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("w_tiled", C.c_void_p), ("prefetch_next", C.c_void_p), ("head_dim", C.c_int32), ("reserved", C.c_int32)]


class pc_chain_op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dep", C.c_int32), ("lin", pc_linear_params),
                ("x", C.c_void_p), ("part1", C.c_void_p), ("part2", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p),
                ("h_out", C.c_void_p), ("n_splits1", C.c_int32), ("n_splits2", C.c_int32), ("rows", C.c_int32), ("hidden", C.c_int32),
                ("eps", C.c_float), ("reserved", C.c_int32)]


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen the in-tree library (built by promptcache_b200.build) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python -m promptcache_b200.build` "
            "(promptcache_b200 has no fallback path; the CUDA extension is the product)")
    # make sure the CUDA runtime torch uses is initialised first so both share the primary context
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    lib.pc_version.restype = C.c_int
    lib.pc_last_error.restype = C.c_char_p
    lib.pc_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.pc_device_info.restype = C.c_int
    lib.pc_splice_kv.argtypes = [C.POINTER(pc_segment), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p]
    lib.pc_splice_kv.restype = C.c_int
    lib.pc_splice_kv_strided.argtypes = [C.POINTER(pc_segment), C.POINTER(C.c_int64), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p]
    lib.pc_splice_kv_strided.restype = C.c_int
    lib.pc_splice_kv_tensors.argtypes = [C.POINTER(pc_copy2d), C.c_int, C.c_void_p]
    lib.pc_splice_kv_tensors.restype = C.c_int
    lib.pc_rope_append.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.pc_rope_append.restype = C.c_int
    lib.pc_rope_append_dyn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.pc_rope_append_dyn.restype = C.c_int
    lib.pc_attn_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.pc_attn_workspace_bytes.restype = C.c_size_t
    lib.pc_prefill_attn.argtypes = [C.POINTER(pc_attn_params), C.c_void_p]
    lib.pc_prefill_attn.restype = C.c_int
    lib.pc_decode_attn.argtypes = [C.POINTER(pc_attn_params), C.c_void_p]
    lib.pc_decode_attn.restype = C.c_int
    lib.pc_launch_count.restype = C.c_int64
    lib.pc_add_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.pc_add_rmsnorm.restype = C.c_int
    lib.pc_add_layernorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                     C.c_void_p]
    lib.pc_add_layernorm.restype = C.c_int
    lib.pc_add_layernorm_partials.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int, C.c_int, C.c_float, C.c_void_p]
    lib.pc_add_layernorm_partials.restype = C.c_int
    lib.pc_linear_chain.argtypes = [C.POINTER(pc_chain_op), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.pc_linear_chain.restype = C.c_int
    lib.pc_silu_mul.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]
    lib.pc_silu_mul.restype = C.c_int
    lib.pc_linear_workspace_bytes.argtypes = []
    lib.pc_linear_workspace_bytes.restype = C.c_size_t
    lib.pc_linear_small.argtypes = [C.POINTER(pc_linear_params), C.c_void_p]
    lib.pc_linear_small.restype = C.c_int
    lib.pc_linear_partial_splits.argtypes = [C.c_int, C.c_int]
    lib.pc_linear_partial_splits.restype = C.c_int
    lib.pc_add_rmsnorm_partials.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                            C.c_void_p]
    lib.pc_add_rmsnorm_partials.restype = C.c_int
    lib.pc_linear_tiled_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.pc_linear_tiled_bytes.restype = C.c_size_t
    lib.pc_linear_tile_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.pc_linear_tile_weights.restype = C.c_int
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pc_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed ({rc}): {msg}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _stream_flush() -> int:
    """Stream handle for a launch that cannot be part of a chain: what the chain has recorded goes out first."""
    chain_flush()
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NativeError("promptcache_b200 kernels take CUDA tensors only (no CPU path exists)")


_replayed_launches = 0


def note_replayed_launches(n: int) -> None:
    """Kernels of this library re-launched by a CUDA-graph replay (they do not pass through the C ABI again)."""
    global _replayed_launches
    _replayed_launches += int(n)


def launch_count() -> int:
    """Kernels of this library launched by this process: direct C-ABI launches + graph-replayed ones."""
    return int(load().pc_launch_count()) + _replayed_launches


def splice_kv(segments: Sequence[tuple], cache: torch.Tensor) -> None:
    """segments: (segment tensor [L,2,Hkv,len,D], dst_row) pairs; cache: [L,2,Hkv,max_ctx,D].  A segment is either a
    packed (contiguous) allocation or a row slice `arena[:, :, :, r0:r0+len, :]` of a schema's segment arena."""
    lib = load()
    _require_cuda(cache)
    if cache.dim() != 5 or cache.shape[1] != 2 or not cache.is_contiguous():
        raise NativeError("splice_kv: cache must be a contiguous [L,2,Hkv,max_ctx,D] tensor")
    L, _, Hkv, max_ctx, D = cache.shape
    n = len(segments)
    arr = (pc_segment * max(n, 1))()
    run_rows = (C.c_int64 * max(n, 1))()
    strided = False
    for i, (seg, row) in enumerate(segments):
        _require_cuda(seg)
        if seg.dim() != 5 or tuple(seg.shape[:3]) != (L, 2, Hkv) or seg.shape[4] != D or seg.dtype != cache.dtype:
            raise NativeError(f"splice_kv: segment {i} has shape {tuple(seg.shape)}/{seg.dtype}, "
                              f"expected [{L},2,{Hkv},len,{D}] {cache.dtype}")
        s0, s1, s2, s3, s4 = seg.stride()
        ln = int(seg.shape[3])
        if s4 != 1 or s3 != D or s2 % D != 0 or s2 < ln * D or s1 != Hkv * s2 or s0 != 2 * s1:
            raise NativeError(f"splice_kv: segment {i} must be contiguous or a row slice of a [L,2,Hkv,R,D] arena "
                              f"(strides {seg.stride()})")
        arr[i].base = seg.data_ptr()
        arr[i].len = ln
        arr[i].dst_row = int(row)
        run_rows[i] = s2 // D
        strided |= run_rows[i] != ln
    if strided:
        _check(lib.pc_splice_kv_strided(arr, run_rows, n, cache.data_ptr(), L, Hkv, D, max_ctx, cache.element_size(), _stream_flush()),
               "pc_splice_kv_strided")
    else:
        _check(lib.pc_splice_kv(arr, n, cache.data_ptr(), L, Hkv, D, max_ctx, cache.element_size(), _stream_flush()),
               "pc_splice_kv")


def splice_kv_tensors(copies: Sequence[tuple]) -> None:
    """copies: (src[Hkv,len,D] contiguous, dst_view[Hkv,len,D] = tgt[:, st:ed, :]) pairs -- the reference layout."""
    lib = load()
    n = len(copies)
    arr = (pc_copy2d * max(n, 1))()
    for i, (src, dst) in enumerate(copies):
        _require_cuda(src, dst)
        if src.shape != dst.shape or src.dtype != dst.dtype or src.dim() != 3 or not src.is_contiguous():
            raise NativeError(f"splice_kv_tensors: copy {i}: src {tuple(src.shape)} vs dst {tuple(dst.shape)}")
        Hkv, ln, D = src.shape
        es = src.element_size()
        if dst.stride(2) != 1 or dst.stride(1) != D:
            raise NativeError(f"splice_kv_tensors: copy {i}: dst rows must be contiguous")
        arr[i].src = src.data_ptr()
        arr[i].dst = dst.data_ptr()
        arr[i].src_run_stride = ln * D * es
        arr[i].dst_run_stride = dst.stride(0) * es
        arr[i].run_bytes = ln * D * es
        arr[i].n_runs = Hkv
    _check(lib.pc_splice_kv_tensors(arr, n, _stream_flush()), "pc_splice_kv_tensors")


def rope_append(q_in: torch.Tensor, k_in: torch.Tensor, v_in: torch.Tensor, position_ids: Optional[torch.Tensor],
                cos: Optional[torch.Tensor], sin: Optional[torch.Tensor], q_out: torch.Tensor,
                k_cache: torch.Tensor, v_cache: torch.Tensor, dst_row0, n_heads: int, n_kv_heads: int,
                head_dim: int) -> None:
    """q_in/k_in/v_in: [q, H*D]-like 2-D views (row stride arbitrary, inner contiguous); caches [Hkv,max_ctx,D].
    `dst_row0` is a host int, or a CUDA int32 tensor read at run time (graph-replayable decode)."""
    lib = load()
    _require_cuda(q_in, k_in, v_in, q_out, k_cache, v_cache)
    q_len = q_in.shape[0]
    for t in (q_in, k_in, v_in):
        if t.dim() != 2 or t.stride(1) != 1 or t.dtype != torch.float16:
            raise NativeError("rope_append: inputs must be fp16 2-D with a contiguous inner dimension")
    if k_cache.stride(2) != 1 or k_cache.stride(1) != head_dim or k_cache.stride(0) != v_cache.stride(0):
        raise NativeError("rope_append: caches must be [Hkv,max_ctx,D] with contiguous rows")
    max_ctx = k_cache.stride(0) // head_dim
    if not q_out.is_contiguous():
        raise NativeError("rope_append: q_out must be contiguous")
    table_rows = 0
    if cos is not None:
        if cos.dtype != torch.float16 or sin.dtype != torch.float16 or not cos.is_contiguous() or not sin.is_contiguous():
            raise NativeError("rope_append: cos/sin must be contiguous fp16 [T,D] tables")
        table_rows = cos.shape[0]
        if position_ids.dtype != torch.int64 or not position_ids.is_contiguous():
            raise NativeError("rope_append: position_ids must be a contiguous int64 tensor")
    if isinstance(dst_row0, torch.Tensor):
        if dst_row0.dtype != torch.int32 or not dst_row0.is_cuda:
            raise NativeError("rope_append: a run-time dst_row0 must be a CUDA int32 tensor")
        _check(lib.pc_rope_append_dyn(q_in.data_ptr(), k_in.data_ptr(), v_in.data_ptr(), q_in.stride(0), k_in.stride(0),
                                      v_in.stride(0), position_ids.data_ptr() if position_ids is not None else None,
                                      cos.data_ptr() if cos is not None else None,
                                      sin.data_ptr() if sin is not None else None, table_rows, q_out.data_ptr(),
                                      k_cache.data_ptr(), v_cache.data_ptr(), max_ctx, dst_row0.data_ptr(), n_heads,
                                      n_kv_heads, head_dim, q_len, _stream_flush()), "pc_rope_append_dyn")
        return
    _check(lib.pc_rope_append(q_in.data_ptr(), k_in.data_ptr(), v_in.data_ptr(), q_in.stride(0), k_in.stride(0),
                              v_in.stride(0), position_ids.data_ptr() if position_ids is not None else None,
                              cos.data_ptr() if cos is not None else None,
                              sin.data_ptr() if sin is not None else None, table_rows, q_out.data_ptr(),
                              k_cache.data_ptr(), v_cache.data_ptr(), max_ctx, int(dst_row0), n_heads, n_kv_heads,
                              head_dim, q_len, _stream_flush()), "pc_rope_append")


def attn_workspace_bytes(n_heads: int, q_len: int, head_dim: int) -> int:
    return int(load().pc_attn_workspace_bytes(n_heads, q_len, head_dim))


def softmax_scale(head_dim: int) -> float:
    """fp32 reciprocal of sqrt(D): torch's CUDA `tensor / python_scalar` multiplies by the fp32 reciprocal."""
    return float(np.float32(1.0) / np.float32(math.sqrt(head_dim)))


def attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, out: torch.Tensor, n_cached: int,
              q_len: int, n_heads: int, n_kv_heads: int, head_dim: int, scale: float, rounding: int,
              workspace: Optional[torch.Tensor] = None, alibi_slopes: Optional[torch.Tensor] = None,
              key_pos: Optional[torch.Tensor] = None, alibi_max_pos: int = 0, impl: int = PC_ATTN_AUTO,
              decode: bool = False, n_cached_dev: Optional[torch.Tensor] = None, virtual: Optional[tuple] = None,
              prefetch: Optional[torch.Tensor] = None) -> None:
    """q: [H,q,D] fp16; caches [Hkv,max_ctx,D] (row-contiguous views allowed); out: [q,H*D] fp16.
    `virtual` = (arena_k [Hkv,R,D], arena_v [Hkv,R,D], tiles int32 [T,2] on the device, T, rows): the first `rows` cached
    keys are read in place from the segment arena through the tile list instead of from k_cache/v_cache (virtual splice)."""
    lib = load()
    _require_cuda(q, k_cache, v_cache, out)
    if k_cache.stride(2) != 1 or k_cache.stride(1) != head_dim or k_cache.stride(0) != v_cache.stride(0):
        raise NativeError("attention: caches must be [Hkv,max_ctx,D] with contiguous rows")
    if not q.is_contiguous() or not out.is_contiguous() or q.dtype != torch.float16 or out.dtype != torch.float16:
        raise NativeError("attention: q/out must be contiguous fp16")
    p = pc_attn_params()
    p.q, p.k, p.v, p.out = q.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), out.data_ptr()
    p.n_heads, p.n_kv_heads, p.head_dim = n_heads, n_kv_heads, head_dim
    p.n_cached, p.q_len = int(n_cached), int(q_len)
    p.max_ctx = k_cache.stride(0) // head_dim
    p.scale = scale
    p.rounding = rounding
    if alibi_slopes is not None:
        if alibi_slopes.dtype != torch.float32 or key_pos is None or key_pos.dtype != torch.int32:
            raise NativeError("attention: alibi needs fp32 slopes and int32 key positions")
        p.alibi_slopes, p.key_pos = alibi_slopes.data_ptr(), key_pos.data_ptr()
    p.alibi_max_pos = int(alibi_max_pos)
    p.impl = impl
    if n_cached_dev is not None:
        if n_cached_dev.dtype != torch.int32 or not n_cached_dev.is_cuda:
            raise NativeError("attention: n_cached_dev must be a CUDA int32 tensor")
        p.n_cached_dev = n_cached_dev.data_ptr()
    if workspace is not None:
        p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    if virtual is not None:
        ak, av, tiles, n_tiles, rows = virtual
        _require_cuda(ak, av, tiles)
        if ak.dtype != torch.float16 or ak.dim() != 3 or ak.shape != av.shape or ak.shape[0] != n_kv_heads or ak.shape[2] != head_dim \
                or not ak.is_contiguous() or not av.is_contiguous():
            raise NativeError("attention: arena planes must be contiguous fp16 [Hkv,R,D]")
        if tiles.dtype != torch.int32 or not tiles.is_contiguous() or tiles.numel() < 2 * n_tiles:
            raise NativeError("attention: the tile list must be a contiguous int32 [T,2] CUDA tensor")
        p.kv_arena_k, p.kv_arena_v, p.kv_tiles = ak.data_ptr(), av.data_ptr(), tiles.data_ptr()
        p.kv_arena_rows, p.kv_n_tiles, p.kv_rows = int(ak.shape[1]), int(n_tiles), int(rows)
    if prefetch is not None:  # hint: the next projection's pre-tiled weights go to L2 while this kernel merges its splits
        _require_cuda(prefetch)
        p.prefetch, p.prefetch_bytes = prefetch.data_ptr(), prefetch.numel() * prefetch.element_size()
    fn = lib.pc_decode_attn if decode else lib.pc_prefill_attn
    _check(fn(C.byref(p), _stream_flush()), "pc_decode_attn" if decode else "pc_prefill_attn")


# ---- chains: several linears / norms as ONE persistent launch (pc_linear_chain) -----------------------------------
class _Chain:
    """Ops recorded inside `with chain(workspace):`.  Dependencies are derived from the tensors: an op depends on the
    latest earlier op that writes something it reads (RAW) or reads / writes something it writes (WAR / WAW)."""

    def __init__(self, workspace: torch.Tensor):
        self.workspace = workspace
        self.ops = []  # (pc_chain_op, reads, writes, keep-alive)

    def add(self, op: pc_chain_op, reads, writes, keep) -> None:
        reads = {t.data_ptr() for t in reads if t is not None}
        writes = {t.data_ptr() for t in writes if t is not None}
        if len(self.ops) == PC_CHAIN_MAX_OPS:
            self.flush()
        dep = -1
        for i, (_, r, w, _) in enumerate(self.ops):
            if (w & reads) or (r & writes) or (w & writes):
                dep = i
        op.dep = dep
        self.ops.append((op, reads, writes, keep))

    def flush(self) -> None:
        if not self.ops:
            return
        arr = (pc_chain_op * len(self.ops))(*[o[0] for o in self.ops])
        n = len(self.ops)
        self.ops = []
        ws = self.workspace
        _check(load().pc_linear_chain(arr, n, ws.data_ptr(), ws.numel() * ws.element_size(), _stream()), "pc_linear_chain")


_chain: Optional[_Chain] = None


@contextlib.contextmanager
def chain(workspace: torch.Tensor):
    """Inside the block linear_small / linear_partials / linear_rope_append (pre-tiled weights) and add_rmsnorm /
    add_layernorm (partial-stack or no deltas) are recorded instead of launched; the block's end launches them as one
    pc_linear_chain (several, if more than PC_CHAIN_MAX_OPS were recorded).  Anything else that is called in between
    flushes the recorded ops first, so stream order is what the program order says."""
    global _chain
    if _chain is not None:
        raise NativeError("chain: already recording")
    _require_cuda(workspace)
    ch = _Chain(workspace)
    _chain = ch
    try:
        yield ch
    except BaseException:
        _chain = None
        raise
    _chain = None
    ch.flush()


def chain_flush() -> None:
    """Launch what the active chain has recorded so far (no-op outside a chain block)."""
    if _chain is not None:
        _chain.flush()


def _emit_linear(p: pc_linear_params, reads, writes) -> None:
    if _chain is not None and p.w_tiled:
        op = pc_chain_op()
        op.kind = PC_CHAIN_LINEAR
        op.lin = p
        _chain.add(op, reads, writes, (reads, writes))
        return
    chain_flush()
    _check(load().pc_linear_small(C.byref(p), _stream()), "pc_linear_small")


def _emit_norm(kind: int, x, part1, part2, weight, bias, h_out, eps) -> None:
    op = pc_chain_op()
    op.kind = kind
    op.x, op.weight, op.h_out = x.data_ptr(), weight.data_ptr(), h_out.data_ptr()
    op.bias = bias.data_ptr() if bias is not None else None
    op.part1 = part1.data_ptr() if part1 is not None else None
    op.part2 = part2.data_ptr() if part2 is not None else None
    op.n_splits1 = part1.shape[0] if part1 is not None else 0
    op.n_splits2 = part2.shape[0] if part2 is not None else 0
    op.rows, op.hidden, op.eps = x.shape[0], x.shape[1], float(eps)
    _chain.add(op, (x, part1, part2), (x, h_out), (x, part1, part2, weight, bias, h_out))


def add_rmsnorm(x: torch.Tensor, delta: Optional[torch.Tensor], weight: torch.Tensor, h_out: torch.Tensor,
                eps: float) -> None:
    """x[rows,hidden] += delta (in place, fp16), h_out = RMSNorm(x) * weight.  An fp32 3-D `delta` is a stack of split-K
    partials [splits, rows, hidden] (linear_partials): they are added in order and rounded to fp16 first."""
    _require_cuda(x, weight, h_out)
    if delta is not None and delta.dtype == torch.float32:
        rows, hidden = x.shape
        if not (x.is_contiguous() and h_out.is_contiguous() and delta.is_contiguous()) or x.dtype != torch.float16 \
                or delta.dim() != 3 or tuple(delta.shape[1:]) != (rows, hidden):
            raise NativeError("add_rmsnorm: partials must be contiguous fp32 [splits, rows, hidden]")
        if _chain is not None:
            return _emit_norm(PC_CHAIN_ADD_RMSNORM, x, delta, None, weight, None, h_out, eps)
        _check(load().pc_add_rmsnorm_partials(x.data_ptr(), delta.data_ptr(), delta.shape[0], weight.data_ptr(), h_out.data_ptr(),
                                              rows, hidden, float(eps), _stream_flush()), "pc_add_rmsnorm_partials")
        return
    if not (x.is_contiguous() and h_out.is_contiguous() and (delta is None or delta.is_contiguous())) or x.dtype != torch.float16:
        raise NativeError("add_rmsnorm: tensors must be contiguous fp16")
    if _chain is not None and delta is None:
        return _emit_norm(PC_CHAIN_ADD_RMSNORM, x, None, None, weight, None, h_out, eps)
    rows, hidden = x.shape
    _check(load().pc_add_rmsnorm(x.data_ptr(), delta.data_ptr() if delta is not None else None, weight.data_ptr(),
                                 h_out.data_ptr(), rows, hidden, float(eps), _stream_flush()), "pc_add_rmsnorm")


def add_layernorm(x: torch.Tensor, d1: Optional[torch.Tensor], d2: Optional[torch.Tensor], weight: torch.Tensor,
                  bias: Optional[torch.Tensor], h_out: torch.Tensor, eps: float) -> None:
    """x[rows,hidden] = half(x + half(d1 + d2)) in place (either delta may be None), h_out = LayerNorm(x) * weight + bias.
    fp32 3-D deltas are stacks of split-K partials [splits, rows, hidden] (linear_partials); both or only d1."""
    _require_cuda(x, weight, h_out, d1, d2, bias)
    if d1 is not None and d1.dtype == torch.float32:
        rows, hidden = x.shape
        for t in (d1, d2):
            if t is not None and (t.dtype != torch.float32 or t.dim() != 3 or tuple(t.shape[1:]) != (rows, hidden) or not t.is_contiguous()):
                raise NativeError("add_layernorm: partials must be contiguous fp32 [splits, rows, hidden] stacks")
        if not (x.is_contiguous() and h_out.is_contiguous()):
            raise NativeError("add_layernorm: x / h_out must be contiguous")
        if _chain is not None:
            return _emit_norm(PC_CHAIN_ADD_LAYERNORM, x, d1, d2, weight, bias, h_out, eps)
        _check(load().pc_add_layernorm_partials(x.data_ptr(), d1.data_ptr(), d1.shape[0], d2.data_ptr() if d2 is not None else None,
                                                d2.shape[0] if d2 is not None else 0, weight.data_ptr(),
                                                bias.data_ptr() if bias is not None else None, h_out.data_ptr(), rows, hidden, float(eps),
                                                _stream_flush()), "pc_add_layernorm_partials")
        return
    for t in (x, h_out, d1, d2):
        if t is not None and (not t.is_contiguous() or t.dtype != torch.float16 or t.shape != x.shape):
            raise NativeError("add_layernorm: x / deltas / h_out must be contiguous fp16 tensors of one shape")
    if _chain is not None and d1 is None and d2 is None:
        return _emit_norm(PC_CHAIN_ADD_LAYERNORM, x, None, None, weight, bias, h_out, eps)
    rows, hidden = x.shape
    _check(load().pc_add_layernorm(x.data_ptr(), d1.data_ptr() if d1 is not None else None, d2.data_ptr() if d2 is not None else None,
                                   weight.data_ptr(), bias.data_ptr() if bias is not None else None, h_out.data_ptr(), rows, hidden,
                                   float(eps), _stream_flush()), "pc_add_layernorm")


def silu_mul(gate_up: torch.Tensor, out: torch.Tensor, inter: int) -> None:
    """out[rows,inter] = silu(gate_up[:, :inter]) * gate_up[:, inter:]."""
    _require_cuda(gate_up, out)
    if gate_up.stride(1) != 1 or not out.is_contiguous() or gate_up.dtype != torch.float16:
        raise NativeError("silu_mul: tensors must be fp16 with contiguous rows")
    _check(load().pc_silu_mul(gate_up.data_ptr(), out.data_ptr(), gate_up.shape[0], inter, gate_up.stride(0), _stream_flush()),
           "pc_silu_mul")


def linear_workspace(device) -> torch.Tensor:
    """Zero-initialised scratch of pc_linear_small (tile arrival counters + split-K partial slots)."""
    return torch.zeros(int(load().pc_linear_workspace_bytes()), dtype=torch.uint8, device=device)


def tile_weights(w: torch.Tensor, epilogue: int = PC_LIN_PLAIN) -> Optional[torch.Tensor]:
    """Pre-tiled copy of an nn.Linear weight for pc_linear_small (None when k is not a multiple of 64)."""
    _require_cuda(w)
    if w.dim() != 2 or w.dtype != torch.float16 or not w.is_contiguous():
        raise NativeError("tile_weights: w must be a contiguous fp16 [n, k] tensor")
    n = int(load().pc_linear_tiled_bytes(w.shape[0], w.shape[1], epilogue))
    if n == 0:
        return None
    out = torch.empty(n, dtype=torch.uint8, device=w.device)
    _check(load().pc_linear_tile_weights(w.data_ptr(), w.shape[0], w.shape[1], epilogue, out.data_ptr(), _stream()),
           "pc_linear_tile_weights")
    return out


def linear_hint(w: torch.Tensor, epilogue: int, w_tiled: Optional[torch.Tensor], n_heads: int = 0, n_kv_heads: int = 0):
    """Description of a pc_linear_small launch for `prefetch_next` of the launch before it (partition fields only).
    None when that launch does not stream pre-tiled weights."""
    if w_tiled is None:
        return None
    h = pc_linear_params()
    h.k, h.n_rows, h.epilogue, h.n_heads, h.n_kv_heads = w.shape[1], w.shape[0], epilogue, n_heads, n_kv_heads
    h.w_tiled = w_tiled.data_ptr()
    h._keep = w_tiled
    return h


def _linear_params(x: torch.Tensor, w: torch.Tensor, workspace: torch.Tensor, epilogue: int,
                   w_tiled: Optional[torch.Tensor] = None, nxt=None) -> pc_linear_params:
    _require_cuda(x, w, workspace, w_tiled)
    if x.dim() != 2 or w.dim() != 2 or x.dtype != torch.float16 or w.dtype != torch.float16 or not x.is_contiguous() \
            or not w.is_contiguous() or x.shape[1] != w.shape[1]:
        raise NativeError(f"linear_small: x {tuple(x.shape)} / w {tuple(w.shape)} must be contiguous fp16 [q,k] / [n,k]")
    p = pc_linear_params()
    p.x, p.w, p.x_stride = x.data_ptr(), w.data_ptr(), x.stride(0)
    p.q_len, p.k, p.n_rows, p.epilogue = x.shape[0], x.shape[1], w.shape[0], epilogue
    p.workspace, p.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    if w_tiled is not None:
        if w_tiled.numel() * w_tiled.element_size() != int(load().pc_linear_tiled_bytes(w.shape[0], w.shape[1], epilogue)):
            raise NativeError("linear_small: w_tiled does not belong to this weight / epilogue")
        p.w_tiled = w_tiled.data_ptr()
    if nxt is not None:
        p.prefetch_next = C.addressof(nxt)
    return p


def linear_small(x: torch.Tensor, w: torch.Tensor, out: torch.Tensor, workspace: torch.Tensor, silu_mul: bool = False,
                 w_tiled: Optional[torch.Tensor] = None, nxt=None, gelu: bool = False) -> None:
    """out[q, n] = x[q, k] @ w[n, k]^T  (q <= 32), or with `silu_mul` out[q, n/2] = silu(x @ w[:n/2]^T) * (x @ w[n/2:]^T);
    `gelu`: exact GELU of the fp16-rounded product (the MLP activation of MPT / Falcon fused into the up projection)."""
    p = _linear_params(x, w, workspace, PC_LIN_SILU_MUL if silu_mul else PC_LIN_PLAIN, w_tiled, nxt)
    if gelu:
        if silu_mul:
            raise NativeError("linear_small: gelu and silu_mul are exclusive")
        p.activation = 1
    _require_cuda(out)
    cols = w.shape[0] // 2 if silu_mul else w.shape[0]
    if out.dim() != 2 or out.dtype != torch.float16 or out.stride(1) != 1 or out.shape[0] != x.shape[0] or out.shape[1] != cols:
        raise NativeError(f"linear_small: out {tuple(out.shape)} must be fp16 [{x.shape[0]},{cols}] with contiguous rows")
    p.out, p.out_stride = out.data_ptr(), out.stride(0)
    _emit_linear(p, (x,), (out,))


def linear_rope_append(x: torch.Tensor, wqkv: torch.Tensor, position_ids: Optional[torch.Tensor], cos: Optional[torch.Tensor],
                       sin: Optional[torch.Tensor], q_out: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                       dst_row0, n_heads: int, n_kv_heads: int, workspace: torch.Tensor,
                       w_tiled: Optional[torch.Tensor] = None, nxt=None, head_dim: int = 128) -> None:
    """Fused QKV projection + RoPE at `position_ids` + in-place KV append (head_dim 128 or 64).  Arguments as rope_append."""
    p = _linear_params(x, wqkv, workspace, PC_LIN_ROPE_APPEND, w_tiled, nxt)
    _require_cuda(q_out, k_cache, v_cache)
    D = int(head_dim)
    p.head_dim = D
    if k_cache.stride(2) != 1 or k_cache.stride(1) != D or k_cache.stride(0) != v_cache.stride(0) or not q_out.is_contiguous():
        raise NativeError("linear_rope_append: caches must be [Hkv,max_ctx,D] with contiguous rows, q_out contiguous")
    if cos is not None:
        if cos.dtype != torch.float16 or sin.dtype != torch.float16 or not cos.is_contiguous() or not sin.is_contiguous() \
                or cos.shape[1] != D:
            raise NativeError("linear_rope_append: cos/sin must be contiguous fp16 [T,D] tables")
        if position_ids.dtype != torch.int64 or not position_ids.is_contiguous():
            raise NativeError("linear_rope_append: position_ids must be a contiguous int64 tensor")
        p.cos, p.sin, p.table_rows, p.position_ids = cos.data_ptr(), sin.data_ptr(), cos.shape[0], position_ids.data_ptr()
    p.q_out, p.k_cache, p.v_cache = q_out.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr()
    p.max_ctx = k_cache.stride(0) // D
    p.n_heads, p.n_kv_heads = n_heads, n_kv_heads
    if isinstance(dst_row0, torch.Tensor):
        if dst_row0.dtype != torch.int32 or not dst_row0.is_cuda:
            raise NativeError("linear_rope_append: a run-time dst_row0 must be a CUDA int32 tensor")
        p.dst_row0, p.dst_row0_dev = 0, dst_row0.data_ptr()
    else:
        p.dst_row0 = int(dst_row0)
    _emit_linear(p, (x, position_ids, dst_row0 if isinstance(dst_row0, torch.Tensor) else None), (q_out, k_cache, v_cache))



def linear_partial_splits(n_rows: int, k: int) -> int:
    return int(load().pc_linear_partial_splits(int(n_rows), int(k)))


def linear_partials(x: torch.Tensor, w: torch.Tensor, part: torch.Tensor, workspace: torch.Tensor,
                    w_tiled: Optional[torch.Tensor] = None, nxt=None) -> int:
    """part[s, q, n] (fp32) = split-K partials of x @ w^T, s < linear_partial_splits(n, k); the consumer adds them
    (add_rmsnorm(..., partials=part)).  Returns the number of splits written."""
    p = _linear_params(x, w, workspace, PC_LIN_PARTIALS, w_tiled, nxt)
    _require_cuda(part)
    splits = linear_partial_splits(w.shape[0], w.shape[1])
    if part.dtype != torch.float32 or not part.is_contiguous() or tuple(part.shape) != (splits, x.shape[0], w.shape[0]):
        raise NativeError(f"linear_partials: part {tuple(part.shape)} must be contiguous fp32 [{splits},{x.shape[0]},{w.shape[0]}]")
    p.out, p.out_stride = part.data_ptr(), w.shape[0]
    _emit_linear(p, (x,), (part,))
    return splits
