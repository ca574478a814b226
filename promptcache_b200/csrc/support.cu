// Support kernels around the hot path (norms / activation are unchanged by
// prompt caching and stay outside the parity scope, but as ~25 eager launches per layer they made the
// TTFT step launch-bound).  Rounding points follow the reference so the residual stream stays within
// fp16 round-off of promptcache/model/llama2.py:
//   pc_add_rmsnorm : x = half(x + delta)  (llama2.py:636,644 residual adds), then LlamaRMSNorm.forward
//                    (llama2.py:104-109): fp32 statistics, round to fp16, multiply by the fp16 weight.
//   pc_silu_mul    : half(silu(gate)) * up in fp16 (llama2.py:241, act_fn = SiLU evaluated in fp32).
#include "pc_common.cuh"

namespace {

constexpr int kNormThreads = 512;  // one 8-column chunk per thread at hidden 4096: a single load wave per pass

// `part` != nullptr: the residual delta arrives as `n_splits` fp32 split-K partials [split][row][hidden] of the
// preceding linear layer (pc_linear_small, PC_LIN_PARTIALS); they are added in split order and rounded to fp16 once,
// which is the fp16 output an nn.Linear would have produced.
__global__ void __launch_bounds__(kNormThreads)
add_rmsnorm_kernel(__half* __restrict__ x, const __half* __restrict__ delta, const __half* __restrict__ weight,
                   __half* __restrict__ h_out, int hidden, float eps, const float* __restrict__ part, int n_splits,
                   long long split_stride) {
  __shared__ float red[kNormThreads / 32];
  pc::pdl_launch_dependents();  // the following GEMM may start streaming its weights now
  // the norm weight is static: its (first) chunk is fetched before the wait, off the dependent part of the kernel
  const int i0 = threadIdx.x * 8;
  uint4 w0 = make_uint4(0u, 0u, 0u, 0u);
  if (i0 < hidden) w0 = __ldg(reinterpret_cast<const uint4*>(weight + i0));
  pc::pdl_wait();               // x / delta come from the previous kernel
  const int row = blockIdx.x;
  __half* xr = x + static_cast<long long>(row) * hidden;
  const __half* dr = delta ? delta + static_cast<long long>(row) * hidden : nullptr;
  __half* hr = h_out + static_cast<long long>(row) * hidden;
  float ss = 0.f;
  uint4 x0 = make_uint4(0u, 0u, 0u, 0u);
  // pass 1: residual add (written back) + sum of squares; 8 halfs per thread per step
  for (int i = threadIdx.x * 8; i < hidden; i += kNormThreads * 8) {
    uint4 xv = *reinterpret_cast<const uint4*>(xr + i);
    __half2* xh = reinterpret_cast<__half2*>(&xv);
    if (part) {
      const float* pr = part + static_cast<long long>(row) * hidden + i;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int s0 = 0; s0 < n_splits; s0 += 4) {  // four splits' loads in flight together, added in split order
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = s0 + u < n_splits;
          const float4* q = reinterpret_cast<const float4*>(pr + (ok ? s0 + u : s0) * split_stride);
          a[u] = ok ? __ldcg(q) : make_float4(0.f, 0.f, 0.f, 0.f);
          b[u] = ok ? __ldcg(q + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (s0 + u < n_splits) {
            acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
            acc[4] += b[u].x; acc[5] += b[u].y; acc[6] += b[u].z; acc[7] += b[u].w;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) xh[k] = __hadd2_rn(xh[k], __floats2half2_rn(acc[2 * k], acc[2 * k + 1]));
      *reinterpret_cast<uint4*>(xr + i) = xv;
    } else if (dr) {
      uint4 dv = *reinterpret_cast<const uint4*>(dr + i);
      const __half2* dh = reinterpret_cast<const __half2*>(&dv);
#pragma unroll
      for (int k = 0; k < 4; ++k) xh[k] = __hadd2_rn(xh[k], dh[k]);
      *reinterpret_cast<uint4*>(xr + i) = xv;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(xh[k]);
      ss = fmaf(f.x, f.x, ss);
      ss = fmaf(f.y, f.y, ss);
    }
    if (i == i0) x0 = xv;  // the first chunk stays in registers for pass 2
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) tot += red[w];
  const float inv = rsqrtf(tot / static_cast<float>(hidden) + eps);
  // pass 2: normalise (fp32) -> fp16 -> * weight (fp16)
  for (int i = threadIdx.x * 8; i < hidden; i += kNormThreads * 8) {
    uint4 xv = i == i0 ? x0 : *reinterpret_cast<const uint4*>(xr + i);
    uint4 wv = i == i0 ? w0 : *reinterpret_cast<const uint4*>(weight + i);
    __half2* xh = reinterpret_cast<__half2*>(&xv);
    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(xh[k]);
      xh[k] = __hmul2_rn(wh[k], __floats2half2_rn(f.x * inv, f.y * inv));
    }
    *reinterpret_cast<uint4*>(hr + i) = xv;
  }
}

// x = half(x + half(d1 + d2)) (d2 optional; d1 optional), then LayerNorm over the row: fp32 mean / variance of the fp16
// values, (x - mean) * rstd * weight + bias evaluated in fp32 and rounded once -- the arithmetic of F.layer_norm on fp16
// input.  Falcon: x + (mlp_out += attn_out) then the next block's input_layernorm (falcon.py:796-810, :781);
// MPT: x + attn_out -> norm_2, x + ffn -> next norm_1 (mpt.py:246-268), bias == nullptr (mpt.py:220-222).
__global__ void __launch_bounds__(kNormThreads)
add_layernorm_kernel(__half* __restrict__ x, const __half* __restrict__ d1, const __half* __restrict__ d2,
                     const __half* __restrict__ weight, const __half* __restrict__ bias, __half* __restrict__ h_out, int hidden,
                     float eps, const float* __restrict__ part, int n_splits, const float* __restrict__ part2, int n_splits2,
                     long long split_stride) {
  __shared__ float red[2][kNormThreads / 32];
  pc::pdl_launch_dependents();
  // static norm weight / bias: the (first) chunk is fetched before the wait
  const int i0 = threadIdx.x * 8;
  uint4 w0 = make_uint4(0u, 0u, 0u, 0u), b0 = make_uint4(0u, 0u, 0u, 0u), x0 = make_uint4(0u, 0u, 0u, 0u);
  if (i0 < hidden) {
    w0 = __ldg(reinterpret_cast<const uint4*>(weight + i0));
    if (bias) b0 = __ldg(reinterpret_cast<const uint4*>(bias + i0));
  }
  pc::pdl_wait();
  const int row = blockIdx.x;
  const long long off = static_cast<long long>(row) * hidden;
  __half* xr = x + off;
  __half* hr = h_out + off;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x * 8; i < hidden; i += kNormThreads * 8) {
    uint4 xv = *reinterpret_cast<const uint4*>(xr + i);
    __half2* xh = reinterpret_cast<__half2*>(&xv);
    if (part) {  // d1 arrives as split-K partials of the preceding linear: added in split order, rounded to fp16 once
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* pr = part + off + i;
      for (int sp = 0; sp < n_splits; ++sp) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(pr + sp * split_stride));
        const float4 b = __ldcg(reinterpret_cast<const float4*>(pr + sp * split_stride) + 1);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
      }
      __half2 dsum[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) dsum[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
      if (part2) {  // a second delta, also as partials (Falcon: mlp_out += attn_out, each the fp16 output of its own linear)
        float acc2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* pr2 = part2 + off + i;
        for (int sp = 0; sp < n_splits2; ++sp) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(pr2 + sp * split_stride));
          const float4 b = __ldcg(reinterpret_cast<const float4*>(pr2 + sp * split_stride) + 1);
          acc2[0] += a.x; acc2[1] += a.y; acc2[2] += a.z; acc2[3] += a.w;
          acc2[4] += b.x; acc2[5] += b.y; acc2[6] += b.z; acc2[7] += b.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dsum[k] = __hadd2_rn(dsum[k], __floats2half2_rn(acc2[2 * k], acc2[2 * k + 1]));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) xh[k] = __hadd2_rn(xh[k], dsum[k]);
      *reinterpret_cast<uint4*>(xr + i) = xv;
    } else if (d1) {
      uint4 av = *reinterpret_cast<const uint4*>(d1 + off + i);
      __half2* ah = reinterpret_cast<__half2*>(&av);
      if (d2) {
        const uint4 bv = *reinterpret_cast<const uint4*>(d2 + off + i);
        const __half2* bh = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
        for (int k = 0; k < 4; ++k) ah[k] = __hadd2_rn(ah[k], bh[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) xh[k] = __hadd2_rn(xh[k], ah[k]);
      *reinterpret_cast<uint4*>(xr + i) = xv;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(xh[k]);
      s1 += f.x + f.y;
    }
    if (i == i0) x0 = xv;  // the first chunk stays in registers for the next two passes
  }
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s1;
  __syncthreads();
  float t1 = 0.f;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) t1 += red[0][w];
  const float mean = t1 / static_cast<float>(hidden);
  // second pass over the row (this thread's own stores: L1 hits) for the centred sum of squares -- no cancellation
  for (int i = threadIdx.x * 8; i < hidden; i += kNormThreads * 8) {
    const uint4 xv = i == i0 ? x0 : *reinterpret_cast<const uint4*>(xr + i);
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(xh[k]);
      s2 = fmaf(f.x - mean, f.x - mean, s2);
      s2 = fmaf(f.y - mean, f.y - mean, s2);
    }
  }
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  float t2 = 0.f;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) t2 += red[1][w];
  const float var = t2 / static_cast<float>(hidden);
  const float rstd = rsqrtf(var + eps);
  for (int i = threadIdx.x * 8; i < hidden; i += kNormThreads * 8) {
    uint4 xv = i == i0 ? x0 : *reinterpret_cast<const uint4*>(xr + i);
    const uint4 wv = i == i0 ? w0 : *reinterpret_cast<const uint4*>(weight + i);
    uint4 bv = b0;
    if (i != i0) bv = bias ? *reinterpret_cast<const uint4*>(bias + i) : make_uint4(0u, 0u, 0u, 0u);
    __half2* xh = reinterpret_cast<__half2*>(&xv);
    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
    const __half2* bh = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(xh[k]);
      const float2 w = __half22float2(wh[k]);
      const float2 b = __half22float2(bh[k]);
      xh[k] = __floats2half2_rn(fmaf((f.x - mean) * rstd, w.x, b.x), fmaf((f.y - mean) * rstd, w.y, b.y));
    }
    *reinterpret_cast<uint4*>(hr + i) = xv;
  }
}

__global__ void __launch_bounds__(256)
silu_mul_kernel(const __half* __restrict__ gate_up, __half* __restrict__ out, int inter, long long row_stride) {
  pc::pdl_launch_dependents();
  pc::pdl_wait();  // gate/up come from the previous kernel
  const int row = blockIdx.y;
  const int i = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (i >= inter) return;
  const __half* g = gate_up + row * row_stride + i;
  uint4 gv = *reinterpret_cast<const uint4*>(g);
  uint4 uv = *reinterpret_cast<const uint4*>(g + inter);
  __half2* gh = reinterpret_cast<__half2*>(&gv);
  const __half2* uh = reinterpret_cast<const __half2*>(&uv);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __half22float2(gh[k]);
    const float sx = f.x / (1.f + expf(-f.x));
    const float sy = f.y / (1.f + expf(-f.y));
    gh[k] = __hmul2_rn(__floats2half2_rn(sx, sy), uh[k]);
  }
  *reinterpret_cast<uint4*>(out + static_cast<long long>(row) * inter + i) = gv;
}

}  // namespace

PC_API int pc_add_rmsnorm(void* x, const void* delta, const void* weight, void* h_out, int rows, int hidden, float eps,
                          void* stream) {
  if (rows == 0) return PC_OK;
  PC_REQUIRE(x && weight && h_out, "pc_add_rmsnorm: null tensor");
  PC_REQUIRE(rows > 0 && hidden > 0 && hidden % 8 == 0, "pc_add_rmsnorm: hidden %d must be a positive multiple of 8", hidden);
  const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(delta) |
                          reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(h_out);
  PC_REQUIRE(align % 16 == 0, "pc_add_rmsnorm: tensors must be 16B aligned");
  PC_CHECK_CUDA(pc::launch_pdl(add_rmsnorm_kernel, dim3(rows), dim3(kNormThreads), 0, static_cast<cudaStream_t>(stream),
                               static_cast<__half*>(x), static_cast<const __half*>(delta), static_cast<const __half*>(weight),
                               static_cast<__half*>(h_out), hidden, eps, static_cast<const float*>(nullptr), 0, 0LL));
  pc::count_launch();
  return PC_OK;
}

PC_API int pc_add_rmsnorm_partials(void* x, const void* delta_partials, int n_splits, const void* weight, void* h_out, int rows,
                                   int hidden, float eps, void* stream) {
  if (rows == 0) return PC_OK;
  PC_REQUIRE(x && weight && h_out && delta_partials, "pc_add_rmsnorm_partials: null tensor");
  PC_REQUIRE(rows > 0 && hidden > 0 && hidden % 8 == 0 && n_splits >= 1, "pc_add_rmsnorm_partials: bad shape");
  const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(delta_partials) |
                          reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(h_out);
  PC_REQUIRE(align % 16 == 0, "pc_add_rmsnorm_partials: tensors must be 16B aligned");
  PC_CHECK_CUDA(pc::launch_pdl(add_rmsnorm_kernel, dim3(rows), dim3(kNormThreads), 0, static_cast<cudaStream_t>(stream),
                               static_cast<__half*>(x), static_cast<const __half*>(nullptr), static_cast<const __half*>(weight),
                               static_cast<__half*>(h_out), hidden, eps, static_cast<const float*>(delta_partials), n_splits,
                               static_cast<long long>(rows) * hidden));
  pc::count_launch();
  return PC_OK;
}

PC_API int pc_add_layernorm(void* x, const void* d1, const void* d2, const void* weight, const void* bias, void* h_out, int rows,
                            int hidden, float eps, void* stream) {
  if (rows == 0) return PC_OK;
  PC_REQUIRE(x && weight && h_out, "pc_add_layernorm: null tensor");
  PC_REQUIRE(d1 != nullptr || d2 == nullptr, "pc_add_layernorm: d2 without d1");
  PC_REQUIRE(rows > 0 && hidden > 0 && hidden % 8 == 0, "pc_add_layernorm: hidden %d must be a positive multiple of 8", hidden);
  const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(d1) | reinterpret_cast<uintptr_t>(d2) |
                          reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(h_out);
  PC_REQUIRE(align % 16 == 0, "pc_add_layernorm: tensors must be 16B aligned");
  PC_CHECK_CUDA(pc::launch_pdl(add_layernorm_kernel, dim3(rows), dim3(kNormThreads), 0, static_cast<cudaStream_t>(stream),
                               static_cast<__half*>(x), static_cast<const __half*>(d1), static_cast<const __half*>(d2),
                               static_cast<const __half*>(weight), static_cast<const __half*>(bias), static_cast<__half*>(h_out),
                               hidden, eps, static_cast<const float*>(nullptr), 0, static_cast<const float*>(nullptr), 0, 0LL));
  pc::count_launch();
  return PC_OK;
}

PC_API int pc_add_layernorm_partials(void* x, const void* delta_partials, int n_splits, const void* delta2_partials, int n_splits2,
                                     const void* weight, const void* bias, void* h_out, int rows, int hidden, float eps, void* stream) {
  if (rows == 0) return PC_OK;
  PC_REQUIRE(x && weight && h_out && delta_partials, "pc_add_layernorm_partials: null tensor");
  PC_REQUIRE(rows > 0 && hidden > 0 && hidden % 8 == 0 && n_splits >= 1, "pc_add_layernorm_partials: bad shape");
  PC_REQUIRE((delta2_partials == nullptr) == (n_splits2 == 0), "pc_add_layernorm_partials: second stack and its split count disagree");
  const uintptr_t align = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(delta_partials) | reinterpret_cast<uintptr_t>(weight) |
                          reinterpret_cast<uintptr_t>(delta2_partials) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(h_out);
  PC_REQUIRE(align % 16 == 0, "pc_add_layernorm_partials: tensors must be 16B aligned");
  PC_CHECK_CUDA(pc::launch_pdl(add_layernorm_kernel, dim3(rows), dim3(kNormThreads), 0, static_cast<cudaStream_t>(stream),
                               static_cast<__half*>(x), static_cast<const __half*>(nullptr), static_cast<const __half*>(nullptr),
                               static_cast<const __half*>(weight), static_cast<const __half*>(bias), static_cast<__half*>(h_out),
                               hidden, eps, static_cast<const float*>(delta_partials), n_splits,
                               static_cast<const float*>(delta2_partials), n_splits2, static_cast<long long>(rows) * hidden));
  pc::count_launch();
  return PC_OK;
}

PC_API int pc_silu_mul(const void* gate_up, void* out, int rows, int inter, int64_t row_stride, void* stream) {
  if (rows == 0) return PC_OK;
  PC_REQUIRE(gate_up && out, "pc_silu_mul: null tensor");
  PC_REQUIRE(rows > 0 && inter > 0 && inter % 8 == 0 && row_stride % 8 == 0 && row_stride >= 2LL * inter,
             "pc_silu_mul: inter %d / row_stride %lld must be multiples of 8 with row_stride >= 2*inter", inter,
             static_cast<long long>(row_stride));
  PC_REQUIRE((reinterpret_cast<uintptr_t>(gate_up) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
             "pc_silu_mul: tensors must be 16B aligned");
  PC_REQUIRE(rows <= 65535, "pc_silu_mul: too many rows");
  dim3 grid((inter / 8 + 255) / 256, rows);
  PC_CHECK_CUDA(pc::launch_pdl(silu_mul_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream),
                               static_cast<const __half*>(gate_up), static_cast<__half*>(out), inter,
                               static_cast<long long>(row_stride)));
  pc::count_launch();
  return PC_OK;
}
