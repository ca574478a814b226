// RoPE at explicit (discontiguous) position ids fused with the KV append.
//
// Replaces, per layer: cos[position_ids]/sin[position_ids] gathers, q*cos + rotate_half(q)*sin on Q
// and the new K (reference promptcache/model/llama2.py:195-210, falcon.py:120-154), the
// view/transpose to head-major (llama2.py:349-351) and torch.cat([past, new]) (llama2.py:361-364).
// One warp handles one (token, head) row of D halfs; Q rows land head-major in q_out, K/V rows land
// directly in the tail rows of the spliced cache.  Arithmetic is fp16 with a rounding after each
// of the two products and after the sum (__hmul2_rn/__hadd2_rn, never contracted to an FMA), which
// makes the result bit-identical to the reference's three eager fp16 kernels.
// Bound: launch latency (bytes = 2*elem*(H+2Hkv)*q*D, a few MiB at most).
#include "pc_common.cuh"

namespace {

struct RopeParams {
  const __half* q_in;
  const __half* k_in;
  const __half* v_in;
  long long q_stride, k_stride, v_stride;
  const long long* pos;
  const __half* cos;
  const __half* sin;
  int table_rows;
  __half* q_out;
  __half* k_cache;
  __half* v_cache;
  int max_ctx, dst_row0, n_heads, n_kv_heads, head_dim, q_len;
  const int* dst_row0_dev;  // optional run-time destination row (graph-replayable decode)
};

__global__ void __launch_bounds__(128) rope_append_kernel(const RopeParams p) {
  // let the attention kernel behind us start its prologue and stream the cached K/V tiles while we run
  pc::pdl_launch_dependents();
  pc::pdl_wait();  // the projections come from the previous kernel
  const int t = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vh = blockIdx.y * 4 + warp;  // virtual head: [0,H) Q, [H,H+Hkv) K, [H+Hkv,H+2Hkv) V
  const int H = p.n_heads, Hkv = p.n_kv_heads, D = p.head_dim;
  if (vh >= H + 2 * Hkv) return;
  const int half_d = D >> 1;
  int row0 = p.dst_row0;
  if (p.dst_row0_dev != nullptr) row0 = max(0, min(__ldg(p.dst_row0_dev), p.max_ctx - p.q_len));

  const __half* src;
  __half* dst;
  bool rotate;
  if (vh < H) {
    src = p.q_in + t * p.q_stride + static_cast<long long>(vh) * D;
    dst = p.q_out + (static_cast<long long>(vh) * p.q_len + t) * D;
    rotate = p.cos != nullptr;
  } else if (vh < H + Hkv) {
    const int h = vh - H;
    src = p.k_in + t * p.k_stride + static_cast<long long>(h) * D;
    dst = p.k_cache + (static_cast<long long>(h) * p.max_ctx + row0 + t) * D;
    rotate = p.cos != nullptr;
  } else {
    const int h = vh - H - Hkv;
    src = p.v_in + t * p.v_stride + static_cast<long long>(h) * D;
    dst = p.v_cache + (static_cast<long long>(h) * p.max_ctx + row0 + t) * D;
    rotate = false;
  }

  if (!rotate) {
    for (int i = lane * 2; i < D; i += 64)
      *reinterpret_cast<__half2*>(dst + i) = *reinterpret_cast<const __half2*>(src + i);
    return;
  }
  long long pos = p.pos[t];
  // Out-of-table ids are a host-side error (checked there); clamp so the kernel never faults.
  if (pos < 0) pos = 0;
  if (pos >= p.table_rows) pos = p.table_rows - 1;
  const __half* c = p.cos + pos * D;
  const __half* s = p.sin + pos * D;
  for (int i = lane * 2; i < half_d; i += 64) {
    const __half2 x1 = *reinterpret_cast<const __half2*>(src + i);
    const __half2 x2 = *reinterpret_cast<const __half2*>(src + i + half_d);
    const __half2 c1 = *reinterpret_cast<const __half2*>(c + i);
    const __half2 c2 = *reinterpret_cast<const __half2*>(c + i + half_d);
    const __half2 s1 = *reinterpret_cast<const __half2*>(s + i);
    const __half2 s2 = *reinterpret_cast<const __half2*>(s + i + half_d);
    // rotate_half(x) = cat(-x2, x1)
    const __half2 lo = __hadd2_rn(__hmul2_rn(x1, c1), __hmul2_rn(__hneg2(x2), s1));
    const __half2 hi = __hadd2_rn(__hmul2_rn(x2, c2), __hmul2_rn(x1, s2));
    *reinterpret_cast<__half2*>(dst + i) = lo;
    *reinterpret_cast<__half2*>(dst + i + half_d) = hi;
  }
}

}  // namespace

static int rope_append_impl(const void* q_in, const void* k_in, const void* v_in, int64_t q_stride,
                            int64_t k_stride, int64_t v_stride, const int64_t* position_ids_dev,
                            const void* cos_dev, const void* sin_dev, int table_rows, void* q_out,
                            void* k_cache, void* v_cache, int max_ctx, int dst_row0, const int32_t* dst_row0_dev,
                            int n_heads, int n_kv_heads, int head_dim, int q_len, void* stream) {
  PC_REQUIRE(q_len >= 0, "pc_rope_append: q_len < 0");
  if (q_len == 0) return PC_OK;
  PC_REQUIRE(q_in && k_in && v_in && q_out && k_cache && v_cache, "pc_rope_append: null tensor");
  PC_REQUIRE(n_heads > 0 && n_kv_heads > 0 && n_heads % n_kv_heads == 0,
             "pc_rope_append: bad head counts H=%d Hkv=%d", n_heads, n_kv_heads);
  PC_REQUIRE(head_dim > 0 && head_dim % 4 == 0, "pc_rope_append: head_dim %d must be a multiple of 4",
             head_dim);
  PC_REQUIRE((cos_dev == nullptr) == (sin_dev == nullptr), "pc_rope_append: cos/sin must both be set");
  PC_REQUIRE(cos_dev == nullptr || (position_ids_dev != nullptr && table_rows > 0),
             "pc_rope_append: rotary needs position ids and a table");
  PC_REQUIRE(dst_row0 >= 0 && static_cast<long long>(dst_row0) + q_len <= max_ctx,
             "pc_rope_append: rows [%d,%d) exceed max_ctx %d", dst_row0, dst_row0 + q_len, max_ctx);
  PC_REQUIRE(q_stride % 2 == 0 && k_stride % 2 == 0 && v_stride % 2 == 0,
             "pc_rope_append: row strides must be even");
  const uintptr_t align = reinterpret_cast<uintptr_t>(q_in) | reinterpret_cast<uintptr_t>(k_in) |
                          reinterpret_cast<uintptr_t>(v_in) | reinterpret_cast<uintptr_t>(q_out) |
                          reinterpret_cast<uintptr_t>(k_cache) | reinterpret_cast<uintptr_t>(v_cache) |
                          reinterpret_cast<uintptr_t>(cos_dev) | reinterpret_cast<uintptr_t>(sin_dev);
  PC_REQUIRE(align % 4 == 0, "pc_rope_append: tensors must be 4B aligned");
  RopeParams p;
  p.q_in = static_cast<const __half*>(q_in);
  p.k_in = static_cast<const __half*>(k_in);
  p.v_in = static_cast<const __half*>(v_in);
  p.q_stride = q_stride;
  p.k_stride = k_stride;
  p.v_stride = v_stride;
  p.pos = reinterpret_cast<const long long*>(position_ids_dev);
  p.cos = static_cast<const __half*>(cos_dev);
  p.sin = static_cast<const __half*>(sin_dev);
  p.table_rows = table_rows;
  p.q_out = static_cast<__half*>(q_out);
  p.k_cache = static_cast<__half*>(k_cache);
  p.v_cache = static_cast<__half*>(v_cache);
  p.max_ctx = max_ctx;
  p.dst_row0 = dst_row0;
  p.dst_row0_dev = dst_row0_dev;
  p.n_heads = n_heads;
  p.n_kv_heads = n_kv_heads;
  p.head_dim = head_dim;
  p.q_len = q_len;
  const int vheads = n_heads + 2 * n_kv_heads;
  dim3 grid(q_len, (vheads + 3) / 4);
  PC_CHECK_CUDA(pc::launch_pdl(rope_append_kernel, grid, dim3(128), 0, static_cast<cudaStream_t>(stream), p));
  pc::count_launch();
  return PC_OK;
}

PC_API int pc_rope_append(const void* q_in, const void* k_in, const void* v_in, int64_t q_stride, int64_t k_stride,
                          int64_t v_stride, const int64_t* position_ids_dev, const void* cos_dev, const void* sin_dev,
                          int table_rows, void* q_out, void* k_cache, void* v_cache, int max_ctx, int dst_row0,
                          int n_heads, int n_kv_heads, int head_dim, int q_len, void* stream) {
  return rope_append_impl(q_in, k_in, v_in, q_stride, k_stride, v_stride, position_ids_dev, cos_dev, sin_dev, table_rows,
                          q_out, k_cache, v_cache, max_ctx, dst_row0, nullptr, n_heads, n_kv_heads, head_dim, q_len, stream);
}

PC_API int pc_rope_append_dyn(const void* q_in, const void* k_in, const void* v_in, int64_t q_stride, int64_t k_stride,
                              int64_t v_stride, const int64_t* position_ids_dev, const void* cos_dev,
                              const void* sin_dev, int table_rows, void* q_out, void* k_cache, void* v_cache, int max_ctx,
                              const int32_t* dst_row0_dev, int n_heads, int n_kv_heads, int head_dim, int q_len,
                              void* stream) {
  if (dst_row0_dev == nullptr) return pc::set_error(PC_ERR_INVALID, "pc_rope_append_dyn: null dst_row0_dev");
  return rope_append_impl(q_in, k_in, v_in, q_stride, k_stride, v_stride, position_ids_dev, cos_dev, sin_dev, table_rows,
                          q_out, k_cache, v_cache, max_ctx, 0, dst_row0_dev, n_heads, n_kv_heads, head_dim, q_len, stream);
}
