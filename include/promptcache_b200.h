/*
 * promptcache_b200 -- C ABI of the B200 (sm_100a) prompt-cache prefill hot path.
 *
 * The reference (yale-sys/prompt-cache) has no FFI layer: its hot path is eager PyTorch.  This
 * header is the boundary a maintainer would bind (ctypes / cffi / pybind) to replace the torch
 * call sites cited on each entry point.  All `*_dev` / `void*` tensor arguments are DEVICE
 * pointers unless a comment says HOST; the caller owns every buffer and keeps it alive until the
 * stream has drained.  Every call is asynchronous on `stream` (a `cudaStream_t` passed as
 * `void*`; NULL = legacy default stream), re-entrant, keeps no state between calls other than
 * per-stream internal caches (pc_splice_kv_tensors stages its descriptor table in a pinned buffer
 * owned by the stream it is called on; concurrent calls must use different streams) and never throws.
 *
 * Return value: 0 on success, a negative `PC_ERR_*` otherwise; `pc_last_error()` returns the
 * thread-local message of the last failure.
 *
 * Element type of K/V/Q/out is IEEE fp16 (`elem_bytes == 2`) everywhere except `pc_splice_*`,
 * which moves opaque rows and accepts any `elem_bytes` whose row size is a multiple of 16 B.
 */
#ifndef PROMPTCACHE_B200_H_
#define PROMPTCACHE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PC_ABI_VERSION 6

enum {
  PC_OK = 0,
  PC_ERR_INVALID = -1,     /* bad argument (shape, alignment, null pointer)            */
  PC_ERR_UNSUPPORTED = -2, /* valid request this build has no kernel for               */
  PC_ERR_CUDA = -3,        /* a CUDA runtime / driver call failed (message has detail) */
  PC_ERR_WORKSPACE = -4    /* caller-provided workspace too small                      */
};

/* Rounding points of the reference attention that the kernels reproduce:
 *   PC_ROUND_QK     : Q.K^T is rounded to fp16 before scaling   (llama2.py:371, mpt.py:156 -- fp16 matmul output)
 *   PC_ROUND_SCALED : the scaled score is rounded to fp16       (llama2.py:371 "/ sqrt(d)", mpt.py:156 "* scale")
 * P is always rounded to fp16 before P.V (llama2.py:387, mpt.py:181) and the context to fp16 at the end.
 * Falcon goes through SDPA with fp32 scores (falcon.py:476-478): pass 0. */
#define PC_ROUND_QK 1
#define PC_ROUND_SCALED 2
#define PC_ROUND_LLAMA (PC_ROUND_QK | PC_ROUND_SCALED)
#define PC_ROUND_MPT (PC_ROUND_QK | PC_ROUND_SCALED)
#define PC_ROUND_FALCON 0

/* Which attention kernel runs.  AUTO = the tcgen05/TMEM kernel; SIMT = the plain CUDA-core
 * two-pass kernel kept as an on-device cross-check (never selected implicitly). */
#define PC_ATTN_AUTO 0
#define PC_ATTN_SIMT 1
#define PC_ATTN_TCGEN05 2
#define PC_ATTN_IMPL_MASK 0xff
/* Flag OR-ed into `impl` (A/B and tests): the short-suffix kernel merges its split-KV partials through the global
 * workspace + arrival counter instead of the thread-block cluster's shared memory (same arithmetic, same bits). */
#define PC_ATTN_GLOBAL_MERGE 0x100

int pc_version(void);
const char* pc_last_error(void);

/* Number of SMs / compute capability of the current device (diagnostics for the host side). */
int pc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Splice: PromptCache.update  (promptcache/cache_engine.py:116-156, the copy loop :135-151)
 *
 * Native layout.  One cached segment (= one TokenSequenceCache, cache_engine.py:50-83) is a single
 * allocation   seg[L][2][Hkv][len][D]   (2 = K then V); the contiguous prompt cache
 * (PromptCache.device_cache, cache_engine.py:104-107) is   cache[L][2][Hkv][max_ctx][D].
 * For every segment, layer, K/V and head the kernel copies the run  seg[l][kv][h][0:len][:]  to
 * cache[l][kv][h][dst_row : dst_row+len][:]  -- bit-exact, TMA bulk load -> smem -> TMA bulk store.
 * `segs` is a HOST array (it travels as a kernel parameter; no H2D copy); n_segs is unbounded
 * (launches are chunked internally). */
typedef struct pc_segment {
  const void* base; /* device: seg[L][2][Hkv][len][D]                         */
  int32_t len;      /* tokens in the segment (>0)                             */
  int32_t dst_row;  /* first row of the segment in the contiguous cache       */
} pc_segment;

int pc_splice_kv(const pc_segment* segs /*HOST*/, int n_segs, void* cache_dev, int n_layers,
                 int n_kv_heads, int head_dim, int max_ctx, int elem_bytes, void* stream);

/* Same, for segments that live in a shared ARENA  arena[L][2][Hkv][arena_rows][D]  (all token sequences of a schema in one
 * allocation; see "virtual splice" below): `segs[i].base` points at row `row0_i` of layer 0 / K / head 0 and consecutive
 * (layer, K|V, head) runs are `src_run_rows[i]` rows apart in the source (== len for a packed segment). */
int pc_splice_kv_strided(const pc_segment* segs /*HOST*/, const int64_t* src_run_rows /*HOST*/, int n_segs, void* cache_dev,
                         int n_layers, int n_kv_heads, int head_dim, int max_ctx, int elem_bytes, void* stream);

/* Reference layout.  The reference keeps one tensor per (segment, layer, K|V), each [Hkv][len][D]
 * contiguous, and one target per (layer, K|V), each [Hkv][max_ctx][D] (cache_engine.py:139-149).
 * One descriptor = one `tgt[:, st:ed, :].copy_(src)`:  n_runs (=Hkv) runs of run_bytes each. */
typedef struct pc_copy2d {
  const void* src;        /* device */
  void* dst;              /* device, already offset to row `st` */
  int64_t src_run_stride; /* bytes between runs in src (len*D*elem)     */
  int64_t dst_run_stride; /* bytes between runs in dst (max_ctx*D*elem) */
  int64_t run_bytes;      /* bytes per run (len*D*elem), multiple of 16 */
  int32_t n_runs;         /* Hkv */
  int32_t reserved;
} pc_copy2d;

int pc_splice_kv_tensors(const pc_copy2d* copies /*HOST*/, int n_copies, void* stream);

/* ------------------------------------------------------------------------------------------------
 * RoPE at explicit position ids + in-place KV append
 *   apply_rotary_pos_emb + rotate_half   (promptcache/model/llama2.py:195-210, call :357-359)
 *   FalconRotaryEmbedding.forward        (promptcache/model/falcon.py:120-154)
 *   torch.cat([past, new], dim=2)        (llama2.py:361-364, falcon.py:446-452, mpt.py:148-151)
 *
 * q_in/k_in/v_in are the projection outputs, token-major: row t of q_in holds H heads of D halfs,
 * rows `*_stride` ELEMENTS apart (so a fused QKV buffer works).  cos/sin are the [table_rows][D]
 * fp16 tables the reference model built (llama2.py:128-137); rows are gathered at position_ids[t]
 * (int64, device).  The rotation is evaluated as half(half(x*cos) + half(rot(x)*sin)) -- the three
 * fp16 roundings of the reference -- so Q and the new K are bit-exact.  cos == NULL skips the
 * rotation (MPT).  Rotated Q goes to q_out[H][q_len][D]; rotated K and V go straight into rows
 * [dst_row0, dst_row0+q_len) of k_cache/v_cache[Hkv][max_ctx][D] (no concatenation). */
int pc_rope_append(const void* q_in, const void* k_in, const void* v_in, int64_t q_stride,
                   int64_t k_stride, int64_t v_stride, const int64_t* position_ids_dev,
                   const void* cos_dev, const void* sin_dev, int table_rows, void* q_out,
                   void* k_cache, void* v_cache, int max_ctx, int dst_row0, int n_heads,
                   int n_kv_heads, int head_dim, int q_len, void* stream);
/* Same, but the destination row is read on the device from *dst_row0_dev at run time (clamped to
 * max_ctx - q_len), so that a captured CUDA graph of a decode step can be replayed while the
 * sequence grows (generation_engine.py:123-147). */
int pc_rope_append_dyn(const void* q_in, const void* k_in, const void* v_in, int64_t q_stride,
                       int64_t k_stride, int64_t v_stride, const int64_t* position_ids_dev,
                       const void* cos_dev, const void* sin_dev, int table_rows, void* q_out,
                       void* k_cache, void* v_cache, int max_ctx, const int32_t* dst_row0_dev,
                       int n_heads, int n_kv_heads, int head_dim, int q_len, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention of the uncached suffix over the spliced cache (prefill) and of one token (decode)
 *   mask construction   llama2.py:62-76,882-884 ; falcon.py:217-245 ; mpt.py:59-75
 *   repeat_kv           llama2.py:247-256,368-369
 *   scores/softmax/PV   llama2.py:371-398 ; falcon.py:460-491 ; mpt.py:156-185
 *
 * Keys [0, n_cached) are visible to every query row; key n_cached+t is visible to query row r iff
 * t <= r (sequence order, not position order).  With alibi_slopes != NULL the fp32 bias
 * slope[h] * (key_pos[j] - alibi_max_pos) is added after scaling (mpt.py:97-110,172-175).
 * The kernels never materialise the mask or the score matrix. */
typedef struct pc_attn_params {
  const void* q;             /* [H][q_len][D] fp16, rotated (q_out of pc_rope_append)        */
  const void* k;             /* [Hkv][max_ctx][D] fp16; rows [0, n_cached+q_len) are read    */
  const void* v;             /* same shape                                                   */
  void* out;                 /* [q_len][H*D] fp16, token-major (input of o_proj)             */
  int32_t n_heads;           /* H                                                            */
  int32_t n_kv_heads;        /* Hkv (H % Hkv == 0)                                           */
  int32_t head_dim;          /* D: 64 or 128                                                 */
  int32_t n_cached;          /* N                                                            */
  int32_t q_len;             /* q                                                            */
  int32_t max_ctx;           /* row stride of k/v heads                                      */
  float scale;               /* multiplies the raw score (1/sqrt(D))                         */
  int32_t rounding;          /* PC_ROUND_* flags                                             */
  const float* alibi_slopes; /* [H] or NULL                                                  */
  const int32_t* key_pos;    /* [n_cached+q_len] or NULL (required with alibi_slopes); with kv_tiles the table is in TILE  *
                              * order: [kv_n_tiles][128] slots of the virtual tiles, then the n_cached-kv_rows+q_len rows     *
                              * behind them                                                                                    */
  int32_t alibi_max_pos;     /* max(position_ids) (mpt.py:522)                               */
  int32_t impl;              /* PC_ATTN_*                                                    */
  const int32_t* n_cached_dev; /* optional: N is read from the device at run time (graph-replayable decode);  *
                              * `n_cached` is then an upper bound used for sizing, and all max_ctx rows of   *
                              * k/v must hold finite values (tcgen05 kernels only)                           */
  void* workspace;           /* split-KV scratch, >= pc_attn_workspace_bytes(...) or NULL;   *
                              * ZERO it once before first use (arrival counters; kernels     *
                              * leave them zeroed); one workspace per concurrent stream      */
  size_t workspace_bytes;
  /* ---- virtual splice: the copy of PromptCache.update (cache_engine.py:135-151) is
   * not made; the kernel reads the cached keys straight from the schema's segment ARENA through a per-request
   * tile list.  With kv_tiles != NULL the first `kv_rows` of the n_cached keys are NOT read from k/v: they are the
   * `kv_n_tiles` tiles listed in kv_tiles, each {first arena row, valid rows (1..128)} -- one tile never spans two
   * segments -- read from kv_arena_k / kv_arena_v = [Hkv][kv_arena_rows][D] (this layer's K and V planes of the
   * arena).  Rows [kv_rows, n_cached + q_len) of k/v (appended suffix / decode rows, in place as always) are read
   * as before.  Short-suffix kernel only (head_dim 128, G*q_len <= 32, no ALiBi); the tile list is DEVICE memory
   * that the caller may rewrite between launches (a captured CUDA graph serves every request with the same
   * kv_n_tiles / kv_rows).  Cached keys are all visible to all query rows, so their order does not matter. */
  const void* kv_arena_k;
  const void* kv_arena_v;
  const int32_t* kv_tiles;   /* device: int32[kv_n_tiles][2] */
  int32_t kv_arena_rows;
  int32_t kv_n_tiles;
  int32_t kv_rows;
  int32_t kv_reserved;
  /* ---- optional hint (short-suffix kernel): `prefetch_bytes` bytes at `prefetch` -- the pre-tiled weights of the projection
   * that follows (o_proj) -- are asked into L2 by the CTAs once their last K/V loads are out.  That kernel cannot start before
   * this one's CTAs leave their SMs (shared memory), so its weight stream would otherwise idle through the merge phase. */
  const void* prefetch;
  size_t prefetch_bytes;
} pc_attn_params;

size_t pc_attn_workspace_bytes(int n_heads, int q_len, int head_dim);
int pc_prefill_attn(const pc_attn_params* p, void* stream);
/* Same contract with q_len == 1 (generation_engine.py:123-147 -> llama2.py:353-398). */
int pc_decode_attn(const pc_attn_params* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Support kernels (not part of the parity scope; they replace ~25 eager
 * launches per layer so that the TTFT step is not launch-bound).  fp16 tensors, rounding as the reference:
 *   pc_add_rmsnorm: x[rows][hidden] += delta (fp16 add, written back; delta may be NULL), then
 *                   h_out = weight * half(x_f32 * rsqrt(mean(x_f32^2) + eps))   (llama2.py:104-109,636,644)
 *   pc_silu_mul   : out[r][i] = half(silu_f32(gate_up[r][i])) * gate_up[r][inter+i]   (llama2.py:241) */
int pc_add_rmsnorm(void* x, const void* delta, const void* weight, void* h_out, int rows, int hidden,
                   float eps, void* stream);
int pc_silu_mul(const void* gate_up, void* out, int rows, int inter, int64_t row_stride, void* stream);
/* LayerNorm families (Falcon, MPT): x[rows][hidden] = half(x + half(d1 + d2)) written back (d1 / d2 may be NULL: no add), then
 * h_out = half(((x - mean) * rstd) * weight + bias) with fp32 statistics -- F.layer_norm on fp16 input; bias may be NULL.
 *   Falcon  x + (mlp_out += attn_out), next input_layernorm   falcon.py:781,796-810
 *   MPT     x + attn_out -> norm_2 ; x + ffn_out -> norm_1    mpt.py:220-222,246-268 */
int pc_add_layernorm(void* x, const void* d1, const void* d2, const void* weight, const void* bias, void* h_out, int rows,
                     int hidden, float eps, void* stream);
/* Same with the deltas given as fp32 split-K partials [split][rows][hidden] (PC_LIN_PARTIALS output): each stack is added in
 * split order and rounded to fp16 once (= the fp16 output of its linear), then x = half(x + half(d1 + d2)).  The second
 * stack is optional (NULL, 0): MPT has one delta per norm, Falcon's parallel block two (attention dense + MLP). */
int pc_add_layernorm_partials(void* x, const void* delta_partials, int n_splits, const void* delta2_partials, int n_splits2,
                              const void* weight, const void* bias, void* h_out, int rows, int hidden, float eps, void* stream);
/* pc_add_rmsnorm whose delta is the sum of `n_splits` fp32 partials [split][rows][hidden] (PC_LIN_PARTIALS output),
 * added in split order and rounded to fp16 once before the fp16 residual add. */
int pc_add_rmsnorm_partials(void* x, const void* delta_partials, int n_splits, const void* weight, void* h_out, int rows,
                            int hidden, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Skinny linear layers of the step: y = x . W^T for q_len <= 32 tokens,
 * one weight-streaming stream-K tcgen05 kernel per layer with the following op fused into its epilogue.
 *   q/k/v projections + RoPE + KV append   llama2.py:345-351,357-364   (PC_LIN_ROPE_APPEND)
 *   o_proj, down_proj, lm_head             llama2.py:405,241,1050      (PC_LIN_PLAIN)
 *   gate_proj/up_proj + SiLU * up          llama2.py:241               (PC_LIN_SILU_MUL)
 * x is [q_len][k] fp16 contiguous; w is an nn.Linear weight [n_rows][k] fp16 row-major; accumulation is fp32
 * (split-K partials included, added in a fixed order: results are deterministic), one rounding to fp16.
 *   PLAIN        out[t][r] = half(sum)                                        out: [q_len][out_stride]
 *   SILU_MUL     w = [gate rows (n_rows/2); up rows (n_rows/2)];  out[t][i] = half(silu_f32(half(gate_i))) * half(up_i)
 *   ROPE_APPEND  w = [q heads; k heads; v heads] of head_dim 128 or 64; the fp16 projections get the same treatment as
 *                pc_rope_append (bit-identical to it on the same fp16 inputs): rotated Q -> q_out[H][q_len][128],
 *                rotated K and V -> rows [dst_row0, dst_row0+q_len) of k_cache/v_cache[Hkv][max_ctx][128];
 *                cos == NULL skips the rotation; dst_row0_dev != NULL reads the row on the device.
 * workspace: >= pc_linear_workspace_bytes() bytes, ZEROED once before first use (tile arrival counters; the
 * kernel leaves them zeroed); one workspace per concurrent stream. */
#define PC_LIN_PLAIN 0
#define PC_LIN_SILU_MUL 1
#define PC_LIN_ROPE_APPEND 2
/* PARTIALS: regular split-K without an in-kernel reduction.  `out` is fp32 [pc_linear_partial_splits(n_rows, k)][q_len][n_rows]
 * (out_stride == n_rows); the consumer adds the splits in order and rounds once -- pc_add_rmsnorm_partials does that
 * for the o_proj / down_proj outputs, whose only consumer is the residual add (llama2.py:636,644).  No workspace use. */
#define PC_LIN_PARTIALS 3
#define PC_LIN_ACT_GELU 1

typedef struct pc_linear_params {
  const void* x;
  const void* w;
  int64_t x_stride;  /* elements between rows of x; must equal k */
  int32_t q_len, k, n_rows, epilogue;
  void* out;         /* PLAIN / SILU_MUL */
  int64_t out_stride;
  const int64_t* position_ids; /* ROPE_APPEND: [q_len] int64 (device) */
  const void* cos;             /* [table_rows][128] fp16 or NULL */
  const void* sin;
  int32_t table_rows;
  int32_t max_ctx;
  void* q_out;
  void* k_cache;
  void* v_cache;
  int32_t dst_row0;
  int32_t n_heads;
  int32_t n_kv_heads;
  int32_t activation; /* PLAIN only: 0, or PC_LIN_ACT_GELU = exact GELU of the fp16-rounded output (MPT / Falcon MLP) */
  const int32_t* dst_row0_dev;
  void* workspace;
  size_t workspace_bytes;
  const void* w_tiled; /* optional: the same weights re-laid out by pc_linear_tile_weights (k % 128 == 0).  The kernel  *
                        * then streams them with one 32 KiB bulk copy per pipeline stage instead of 128-byte-row     *
                        * tensor-map boxes (an SM pulls ~2x more per second that way); `w` is still required.         */
  const struct pc_linear_params* prefetch_next; /* optional HOST pointer: the pc_linear_small launch that follows on this      *
                        * stream (only k, n_rows, epilogue, n_heads, n_kv_heads and w_tiled are read, during this call).     *
                        * Each CTA, once its own weight loads are all in flight, asks L2 for the first pipeline stages the   *
                        * CTAs of that launch will read, so its ring fills from L2 across the launch boundary.  A hint only. */
  int32_t head_dim;    /* ROPE_APPEND: 128 (0 means 128) or 64 (Falcon: two heads per 128-row tile, tables [T][64])          */
  int32_t reserved;
} pc_linear_params;

size_t pc_linear_workspace_bytes(void);
int pc_linear_partial_splits(int n_rows, int k);
int pc_linear_small(const pc_linear_params* p, void* stream);
/* One-off re-layout of w[n_rows][k] for `epilogue` into pc_linear_tiled_bytes(n_rows, k, epilogue) bytes at w_tiled:
 * the [tile][k-block] sequence of 16 KiB shared-memory images (128 rows x 64 columns, 128B-swizzled) the kernel consumes. */
/* k must be a multiple of 64; an odd number of 64-column blocks is padded with one block of zeros in the copy. */
size_t pc_linear_tiled_bytes(int n_rows, int k, int epilogue);
int pc_linear_tile_weights(const void* w, int n_rows, int k, int epilogue, void* w_tiled, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A chain of pc_linear_small launches and the norms between them as ONE persistent launch (one CTA per SM): the weight
 * stream of op k+1 starts while op k's epilogue, the grid barrier and the norm still run, so the launch-boundary dead
 * time of ~7 kernels per decoder layer is paid once.  Llama-2 (llama2.py:623-648):
 *   o_proj (PARTIALS) -> x += attn; h = RMSNorm(x) -> gate/up + SiLU*up -> down_proj (PARTIALS) -> x += mlp; h = RMSNorm(x)
 *   -> q/k/v projections + RoPE + KV append of the NEXT layer (or lm_head)
 * Ops run in order; `dep` names the earlier op of the chain whose output this op reads (-1: only kernels launched before
 * the chain).  Because every CTA finishes op j before op j+1, waiting for op `dep` covers all ops <= dep.  All linears
 * need pre-tiled weights and share q_len; norms take their deltas as split-K partial stacks (or none) and produce exactly
 * what pc_add_rmsnorm_partials / pc_add_layernorm_partials produce.  workspace: the pc_linear_small workspace (zeroed once). */
#define PC_CHAIN_LINEAR 0
#define PC_CHAIN_ADD_RMSNORM 1
#define PC_CHAIN_ADD_LAYERNORM 2
#define PC_CHAIN_MAX_OPS 8
typedef struct pc_chain_op {
  int32_t kind;
  int32_t dep;
  pc_linear_params lin;  /* PC_CHAIN_LINEAR (workspace / workspace_bytes / prefetch_next are ignored) */
  /* norms: x[rows][hidden] = half(x + half(sum part1) (+ half(sum part2))) in place; h_out = norm(x) * weight (+ bias) */
  void* x;
  const void* part1;     /* fp32 [n_splits1][rows][hidden] or NULL */
  const void* part2;     /* LayerNorm only (Falcon's parallel block), or NULL */
  const void* weight;
  const void* bias;      /* LayerNorm only, or NULL */
  void* h_out;
  int32_t n_splits1, n_splits2, rows, hidden;
  float eps;
  int32_t reserved;
} pc_chain_op;
int pc_linear_chain(const pc_chain_op* ops, int n_ops, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
int64_t pc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PROMPTCACHE_B200_H_ */
