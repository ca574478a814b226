// Pieces shared by the attention kernels: the split-KV combine kernel of the long-suffix path, the 3-D tensor-map
// builder and the workspace sizing.  (They used to live beside the first, single-tile tcgen05 kernel, which the
// ping-pong kernel of attn_tc2.cu superseded.)
#include <cuda.h>
#include <math.h>

#include <mutex>

#include "pc_common.cuh"

namespace pc {

namespace {

constexpr int BM = 128;
constexpr int kMaxSplits = 32;

// Merge split-KV partials: out = sum_s 2^(m_s-m*) O_s / sum_s 2^(m_s-m*) l_s.
template <int D>
__global__ void __launch_bounds__(D) attn_combine_kernel(const float* __restrict__ part_o,
                                                         const float* __restrict__ part_ml,
                                                         __half* __restrict__ out, int n_heads, int q_len,
                                                         int n_splits) {
  const int r = blockIdx.x, h = blockIdx.y, d = threadIdx.x;
#ifndef PC_PP_NO_PDL
  // launched programmatically behind the attention kernel (hides its launch latency); it does NOT trigger its own dependents
  // early: a register-heavy dependent waiting on the SMs would leave this many-CTA grid a fraction of its occupancy
  pdl_wait();  // the partials come from the attention kernel launched just before
#endif
  float m_star = -INFINITY;
  for (int s = 0; s < n_splits; ++s)
    m_star = fmaxf(m_star, part_ml[((static_cast<long long>(s) * n_heads + h) * q_len + r) * 2]);
  float acc = 0.f, l = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const long long idx = (static_cast<long long>(s) * n_heads + h) * q_len + r;
    const float m = part_ml[idx * 2];
    if (m == -INFINITY) continue;
    const float w = exp2f(m - m_star);
    l += w * part_ml[idx * 2 + 1];
    acc += w * part_o[idx * D + d];
  }
  out[(static_cast<long long>(r) * n_heads + h) * D + d] = __float2half_rn(l > 0.f ? acc / l : 0.f);
}

// ---- host: TMA descriptors ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

}  // namespace

// 3-D fp16 tensor [heads][rows][D] (row stride D, head stride head_stride elements), box 64 x box_rows x 1,
// 128B swizzle, out-of-bounds rows read as zeros.
int make_tensor_map_3d(CUtensorMap* map, const void* base, int D, long long rows, int heads, long long head_stride,
                       int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(PC_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(heads)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(D) * 2, static_cast<cuuint64_t>(head_stride) * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return set_error(PC_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", static_cast<int>(rc));
  return PC_OK;
}

int launch_attn_combine(const float* part_o, const float* part_ml, void* out, int n_heads, int q_len, int n_splits,
                        int head_dim, cudaStream_t st) {
  dim3 cgrid(q_len, n_heads);
#ifdef PC_PP_NO_PDL
  if (head_dim == 128)
    attn_combine_kernel<128><<<cgrid, 128, 0, st>>>(part_o, part_ml, static_cast<__half*>(out), n_heads, q_len, n_splits);
  else
    attn_combine_kernel<64><<<cgrid, 64, 0, st>>>(part_o, part_ml, static_cast<__half*>(out), n_heads, q_len, n_splits);
  PC_CHECK_CUDA(cudaGetLastError());
#else
  if (head_dim == 128)
    PC_CHECK_CUDA(launch_pdl(attn_combine_kernel<128>, cgrid, dim3(128), 0, st, part_o, part_ml, static_cast<__half*>(out), n_heads,
                             q_len, n_splits));
  else
    PC_CHECK_CUDA(launch_pdl(attn_combine_kernel<64>, cgrid, dim3(64), 0, st, part_o, part_ml, static_cast<__half*>(out), n_heads,
                             q_len, n_splits));
#endif
  count_launch();
  return PC_OK;
}

// Upper bound of what the long-suffix kernels may use: the split count is largest when every query head shares
// one KV head (MQA), i.e. base = ceil(H*q/256) CTAs (row-tile pairs of attn_tc2.cu) before splitting.
size_t attn_tc_workspace_bytes(int n_heads, int q_len, int head_dim) {
  const long long base = (static_cast<long long>(n_heads) * q_len + 2 * BM - 1) / (2 * BM);
  long long splits = sm_count() / base;
  if (splits > kMaxSplits) splits = kMaxSplits;
  if (splits <= 1) return 0;
  return static_cast<size_t>(splits) * n_heads * q_len * (head_dim + 2) * sizeof(float);
}

}  // namespace pc
