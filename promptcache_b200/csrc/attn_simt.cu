// CUDA-core attention kernel: a literal two-pass restatement of the reference's eager attention
// (promptcache/model/llama2.py:371-398, mpt.py:156-185) kept as an ON-DEVICE cross-check for the
// tcgen05 kernel and for shapes the tensor-core kernel does not cover.  It reproduces every fp16
// rounding point including the normalised P = half(exp(s-m)/l).  It is never selected implicitly
// (PC_ATTN_SIMT must be requested).  One CTA per (query row, head); scores live in shared memory.
#include <math.h>

#include "pc_common.cuh"

namespace pc {

__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }

__global__ void __launch_bounds__(128) attn_simt_kernel(const pc_attn_params p) {
  extern __shared__ float sc[];  // [n_vis] scores, then probabilities
  __shared__ float red[128];
  const int r = blockIdx.x, h = blockIdx.y;
  const int D = p.head_dim, G = p.n_heads / p.n_kv_heads, hk = h / G;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_vis = p.n_cached + r + 1;  // masked keys get probability exactly 0 in the reference
  const __half* q = static_cast<const __half*>(p.q) + (static_cast<long long>(h) * p.q_len + r) * D;
  const __half* K = static_cast<const __half*>(p.k) + static_cast<long long>(hk) * p.max_ctx * D;
  const __half* V = static_cast<const __half*>(p.v) + static_cast<long long>(hk) * p.max_ctx * D;
  const float slope = p.alibi_slopes ? p.alibi_slopes[h] : 0.f;

  // pass 1: scores
  float qv[4];
  const int per_lane = D / 32;  // 2 or 4
  for (int i = 0; i < per_lane; ++i) qv[i] = __half2float(q[lane * per_lane + i]);
  float m = -INFINITY;
  for (int j = warp; j < n_vis; j += 4) {
    const __half* kr = K + static_cast<long long>(j) * D + lane * per_lane;
    float acc = 0.f;
    for (int i = 0; i < per_lane; ++i) acc = fmaf(qv[i], __half2float(kr[i]), acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (p.rounding & PC_ROUND_QK) acc = round_h(acc);
    float s = acc * p.scale;
    if (p.rounding & PC_ROUND_SCALED) s = round_h(s);
    if (p.alibi_slopes) s += slope * static_cast<float>(p.key_pos[j] - p.alibi_max_pos);
    if (lane == 0) sc[j] = s;
    m = fmaxf(m, s);
  }
  red[tid] = m;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) red[tid] = fmaxf(red[tid], red[tid + o]);
    __syncthreads();
  }
  m = red[0];
  __syncthreads();
  // pass 2: softmax in fp32, P rounded to fp16
  float l = 0.f;
  for (int j = tid; j < n_vis; j += 128) {
    const float e = expf(sc[j] - m);
    sc[j] = e;
    l += e;
  }
  red[tid] = l;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  l = red[0];
  __syncthreads();
  for (int j = tid; j < n_vis; j += 128) sc[j] = round_h(sc[j] / l);
  __syncthreads();
  // pass 3: context
  const int slices = 128 / D;  // 1 (D=128) or 2 (D=64)
  const int d = tid % D, sl = tid / D;
  float o = 0.f;
  for (int j = sl; j < n_vis; j += slices) o = fmaf(sc[j], __half2float(V[static_cast<long long>(j) * D + d]), o);
  red[tid] = o;
  __syncthreads();
  if (sl == 0) {
    for (int s2 = 1; s2 < slices; ++s2) o += red[s2 * D + d];
    static_cast<__half*>(p.out)[(static_cast<long long>(r) * p.n_heads + h) * D + d] = __float2half_rn(o);
  }
}

int launch_attn_simt(const pc_attn_params& p, cudaStream_t st) {
  if (p.head_dim != 64 && p.head_dim != 128)
    return set_error(PC_ERR_UNSUPPORTED, "attn(simt): head_dim %d not in {64,128}", p.head_dim);
  const size_t smem = static_cast<size_t>(p.n_cached + p.q_len) * sizeof(float);
  if (smem > 200 * 1024)
    return set_error(PC_ERR_UNSUPPORTED, "attn(simt): %d keys exceed the shared-memory score buffer",
                     p.n_cached + p.q_len);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
    configured = 200 * 1024;
  }
  dim3 grid(p.q_len, p.n_heads);
  attn_simt_kernel<<<grid, 128, smem, st>>>(p);
  PC_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return PC_OK;
}

}  // namespace pc
