// tcgen05 / TMEM attention of the uncached suffix over the spliced cache (prefill and decode).
//
// Replaces reference promptcache/model/llama2.py:368-398 (repeat_kv, QK^T/sqrt(d), +mask, fp32
// softmax -> fp16, PV, transpose), falcon.py:460-491 and mpt.py:156-185, without materialising the
// [H,q,N+q] score/mask tensors.
//
// Decomposition: one CTA = (kv head, 128-row block of folded query rows, KV split).  The G query
// heads that share a KV head are folded into the M dimension (row = g*q + r), so Falcon's MQA
// (G=71) fills the 128-row MMA while Llama MHA (G=1) with a short suffix is a KV stream.
// Warp roles (192 threads): warps 0-3 softmax/correction/epilogue (one TMEM lane = one row per
// thread), warp 4 TMA producer, warp 5 tcgen05.mma issuer (+ TMEM owner).
//   S_j  = Q . K_j^T     SS-MMA 128x128xD, operands 128B-swizzled K-major tiles staged by TMA,
//                        accumulator in TMEM (double buffered so S_{j+1} overlaps softmax_j)
//   P_j  = softmax piece written back to TMEM as packed fp16 over S_j
//   O   += P_j . V_j     TS-MMA (A = P from TMEM, B = V tile, MN-major), accumulator in TMEM
// Online softmax with lazy rescaling (O is only rescaled when the row max grows by > 2^8).
// Reference-rounding mode reproduces the fp16 roundings of the eager reference on every score.
// Split-KV partials (fp32 O, m, l) are merged by attn_combine_kernel.
// Bound: HBM (K,V stream) when G*q < ~254 FLOP/B, tensor pipe above.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <mutex>

#include "pc_common.cuh"

namespace pc {

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int kSoftmaxThreads = 128;
constexpr int kThreads = 192;
constexpr int kMaxSplits = 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;

template <int D>
struct Cfg {
  static constexpr int kTileBytes = BN * D * 2;  // one K or V tile == one Q tile
  static constexpr int kStages = D == 128 ? 3 : 4;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmem = 1024 + kTileBytes + kStages * 2 * kTileBytes + kBarBytes;
  static constexpr int kTmemCols = 512;  // S0[0,128) S1[128,256) O[256,256+D)
};

struct TcParams {
  __half* out;
  float* part_o;
  float* part_ml;
  const float* alibi_slopes;
  const int* key_pos;
  int n_heads, n_kv_heads, n_cached, q_len, n_splits, m_blocks, alibi_max_pos;
  float scale;
};

__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int D, int ROUND, bool ALIBI>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const TcParams p) {
  using C = Cfg<D>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  unsigned char* sQ = smem;
  unsigned char* sKV = smem + C::kTileBytes;  // stage s: K at 2*s, V at 2*s+1
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + C::kStages * 2 * C::kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + C::kStages;
  uint64_t* kv_empty = v_full + C::kStages;
  uint64_t* s_full = kv_empty + C::kStages;  // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  // indexed by tile parity: a parity wait is only unambiguous within one phase of its barrier, and the softmax
  // warps can be up to two P.V MMAs ahead
  uint64_t* o_full = p_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = p.n_heads / p.n_kv_heads;
  const int rows_total = G * p.q_len;

  int bid = blockIdx.x;
  const int split = bid % p.n_splits;
  bid /= p.n_splits;
  const int mb = bid % p.m_blocks;
  const int hk = bid / p.m_blocks;

  // Causal extent of this row block: rows [mb*128, mb*128+128) folded as g*q + r.
  const int row_lo = mb * BM;
  const int row_hi = min(row_lo + BM, rows_total) - 1;
  int r_max;
  if (row_lo / p.q_len == row_hi / p.q_len) r_max = row_hi % p.q_len;
  else r_max = p.q_len - 1;
  const int kv_end = p.n_cached + r_max + 1;
  const int tiles_blk = (kv_end + BN - 1) / BN;
  const int tps = (tiles_blk + p.n_splits - 1) / p.n_splits;
  const int tile0 = split * tps;
  const int T = max(0, min(tiles_blk, tile0 + tps) - tile0);

  // per-row identity for the softmax threads
  const int R = row_lo + tid;
  const bool row_valid = tid < kSoftmaxThreads && R < rows_total;
  const int g = row_valid ? R / p.q_len : 0;
  const int r = row_valid ? R % p.q_len : 0;
  const int h = hk * G + g;

  const int n_valid_warps = (min(BM, rows_total - row_lo) + 31) / 32;  // softmax warps that own rows

  if (T == 0) {  // empty split: publish a neutral partial and leave (uniform per CTA)
    if (row_valid && p.n_splits > 1) {
      float* ml = p.part_ml + ((static_cast<long long>(split) * p.n_heads + h) * p.q_len + r) * 2;
      ml[0] = -INFINITY;
      ml[1] = 0.f;
    }
    return;
  }

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 32 * n_valid_warps);
    }
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, C::kTileBytes);
      for (int hf = 0; hf < D / 64; ++hf)
        tma_load_3d(sQ + hf * 16384, &map_q, q_full, hf * 64, row_lo, hk);
      for (int j = 0; j < T; ++j) {
        const int s = j % C::kStages;
        if (j >= C::kStages) mbar_wait(&kv_empty[s], static_cast<uint32_t>(((j / C::kStages) - 1) & 1));
        unsigned char* dk = sKV + (2 * s) * C::kTileBytes;
        unsigned char* dv = dk + C::kTileBytes;
        const int key0 = (tile0 + j) * BN;
        mbar_expect_tx(&k_full[s], C::kTileBytes);
        for (int hf = 0; hf < D / 64; ++hf) tma_load_3d(dk + hf * 16384, &map_k, &k_full[s], hf * 64, key0, hk);
        mbar_expect_tx(&v_full[s], C::kTileBytes);
        for (int hf = 0; hf < D / 64; ++hf) tma_load_3d(dv + hf * 16384, &map_v, &v_full[s], hf * 64, key0, hk);
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(BM, D, 0, 1);
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_qk = [&](int j) {
        const int s = j % C::kStages;
        mbar_wait(&k_full[s], static_cast<uint32_t>((j / C::kStages) & 1));
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + (2 * s) * C::kTileBytes);
        const uint32_t d = tmem_base + (j & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_ss(d, make_smem_desc(q_addr + off, 16, 1024, 2), make_smem_desc(k_addr + off, 16, 1024, 2),
                  idesc_qk, kk > 0 ? 1u : 0u);
        }
        tc_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_qk(j + 1);
        const int s = j % C::kStages;
        mbar_wait(&p_full[j & 1], static_cast<uint32_t>((j >> 1) & 1));
        mbar_wait(&v_full[s], static_cast<uint32_t>((j / C::kStages) & 1));
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + (2 * s + 1) * C::kTileBytes);
        const uint32_t pa = tmem_base + (j & 1) * 128;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)
          umma_ts(tmem_o, pa + kk * 8, make_smem_desc(v_addr + kk * 2048, 16384, 1024, 2), idesc_pv,
                  (j > 0 || kk > 0) ? 1u : 0u);
        tc_commit(&kv_empty[s]);
        tc_commit(&o_full[j & 1]);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const bool warp_valid = (row_lo + warp * 32) < rows_total;
    const float slope = (ALIBI && row_valid) ? p.alibi_slopes[h] : 0.f;
    const int vis_end = row_valid ? p.n_cached + r + 1 : 0;  // keys [0, vis_end) are visible
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < T; ++j) {
      const int b = j & 1;
      if (warp_valid) {
        mbar_wait(&s_full[b], static_cast<uint32_t>((j >> 1) & 1));
        tc_fence_after();
        const uint32_t s_addr = tmem_base + lane_base + b * 128;
        float sreg[BN];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          tmem_ld32(s_addr + ch * 32, reinterpret_cast<uint32_t(&)[32]>(sreg[ch * 32]));
        tmem_ld_wait();

        const int key0 = (tile0 + j) * BN;
        const bool need_mask = key0 + BN > vis_end;
        float mt = -INFINITY;
#pragma unroll
        for (int c = 0; c < BN; ++c) {
          float x = sreg[c];
          if (ROUND & PC_ROUND_QK) x = round_h(x);
          x *= p.scale;
          if (ROUND & PC_ROUND_SCALED) x = round_h(x);
          if (ALIBI) {
            const int kidx = min(key0 + c, p.n_cached + p.q_len - 1);
            x += slope * static_cast<float>(__ldg(p.key_pos + kidx) - p.alibi_max_pos);
          }
          x *= kLog2e;
          if (need_mask && key0 + c >= vis_end) x = -INFINITY;
          sreg[c] = x;
          mt = fmaxf(mt, x);
        }
        // lazy rescale: keep the stale max unless it grew by more than 2^8
        const float m_new = fmaxf(m_run, mt);
        float alpha = 1.f;
        if (m_run == -INFINITY) {
          // nothing accumulated for this row yet: l_run == 0 and its O row is unwritten (j == 0)
          // or exact zeros (every earlier P was 0), so adopting the new max needs no rescale
          m_run = m_new;
        } else if (m_new - m_run > kRescaleThreshold) {
          alpha = exp2f(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
        }
        const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
        float lt = 0.f;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t preg[32];
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            const float p0 = exp2f(sreg[hf * 64 + c] - m_use);
            const float p1 = exp2f(sreg[hf * 64 + c + 1] - m_use);
            lt += p0 + p1;
            preg[c >> 1] = pack_h2(p0, p1);
          }
          tmem_st32(s_addr + hf * 32, preg);
        }
        l_run += lt;

        if (__any_sync(0xffffffffu, alpha != 1.f)) {
          // O += P.V of tile j-1 must have landed before O is rescaled
          mbar_wait(&o_full[(j - 1) & 1], static_cast<uint32_t>(((j - 1) >> 1) & 1));
          tc_fence_after();
#pragma unroll
          for (int ch = 0; ch < D / 32; ++ch) {
            uint32_t oreg[32];
            tmem_ld32(tmem_o + lane_base + ch * 32, oreg);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) oreg[c] = __float_as_uint(__uint_as_float(oreg[c]) * alpha);
            tmem_st32(tmem_o + lane_base + ch * 32, oreg);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[b]);
      }
    }

    // ---- epilogue: O (TMEM) -> out / partials ----
    if (warp_valid) {
      mbar_wait(&o_full[(T - 1) & 1], static_cast<uint32_t>(((T - 1) >> 1) & 1));
      tc_fence_after();
      const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t oreg[32];
        tmem_ld32(tmem_o + lane_base + ch * 32, oreg);
        tmem_ld_wait();
        if (row_valid) {
          if (p.n_splits == 1) {
            __half* dst = p.out + (static_cast<long long>(r) * p.n_heads + h) * D + ch * 32;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint4 w;
              w.x = pack_h2(__uint_as_float(oreg[v * 8 + 0]) * inv_l, __uint_as_float(oreg[v * 8 + 1]) * inv_l);
              w.y = pack_h2(__uint_as_float(oreg[v * 8 + 2]) * inv_l, __uint_as_float(oreg[v * 8 + 3]) * inv_l);
              w.z = pack_h2(__uint_as_float(oreg[v * 8 + 4]) * inv_l, __uint_as_float(oreg[v * 8 + 5]) * inv_l);
              w.w = pack_h2(__uint_as_float(oreg[v * 8 + 6]) * inv_l, __uint_as_float(oreg[v * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(dst + v * 8) = w;
            }
          } else {
            float* dst = p.part_o + ((static_cast<long long>(split) * p.n_heads + h) * p.q_len + r) * D + ch * 32;
#pragma unroll
            for (int v = 0; v < 8; ++v)
              *reinterpret_cast<uint4*>(dst + v * 4) =
                  make_uint4(oreg[v * 4], oreg[v * 4 + 1], oreg[v * 4 + 2], oreg[v * 4 + 3]);
          }
        }
      }
      if (row_valid && p.n_splits > 1) {
        float* ml = p.part_ml + ((static_cast<long long>(split) * p.n_heads + h) * p.q_len + r) * 2;
        ml[0] = m_run;  // log2 domain
        ml[1] = l_run;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_free<C::kTmemCols>(tmem_base);
  }
}

// Merge split-KV partials: out = sum_s 2^(m_s-m*) O_s / sum_s 2^(m_s-m*) l_s.
template <int D>
__global__ void __launch_bounds__(D) attn_combine_kernel(const float* __restrict__ part_o,
                                                         const float* __restrict__ part_ml,
                                                         __half* __restrict__ out, int n_heads, int q_len,
                                                         int n_splits) {
  const int r = blockIdx.x, h = blockIdx.y, d = threadIdx.x;
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

namespace {

int make_map(CUtensorMap* map, const void* base, int D, long long rows, int heads, long long head_stride) {
  return make_tensor_map_3d(map, base, D, rows, heads, head_stride, BN);
}

template <int D, int ROUND, bool ALIBI>
int launch_one(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const TcParams& tp, int grid,
               cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<D, ROUND, ALIBI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<D>::kSmem));
    configured = true;
  }
  attn_tc_kernel<D, ROUND, ALIBI><<<grid, kThreads, Cfg<D>::kSmem, st>>>(mq, mk, mv, tp);
  PC_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return PC_OK;
}

template <int D>
int launch_d(const pc_attn_params& p, cudaStream_t st) {
  const int G = p.n_heads / p.n_kv_heads;
  const int rows_total = G * p.q_len;
  const int m_blocks = (rows_total + BM - 1) / BM;
  const int tiles = (p.n_cached + p.q_len + BN - 1) / BN;
  const int base = p.n_kv_heads * m_blocks;
  int n_splits = sm_count() / base;
  if (const char* e = getenv("PC_ATTN_SPLITS")) n_splits = atoi(e);
  if (n_splits > tiles) n_splits = tiles;
  if (n_splits > kMaxSplits) n_splits = kMaxSplits;
  if (n_splits < 1) n_splits = 1;

  float* part_o = nullptr;
  float* part_ml = nullptr;
  if (n_splits > 1) {
    const size_t need = static_cast<size_t>(n_splits) * p.n_heads * p.q_len * (D + 2) * sizeof(float);
    const size_t avail = (p.workspace && p.workspace_bytes > kAttnCounterBytes) ? p.workspace_bytes - kAttnCounterBytes : 0;
    if (avail < need) {
      // degrade to as many splits as the caller's workspace allows
      const size_t per = static_cast<size_t>(p.n_heads) * p.q_len * (D + 2) * sizeof(float);
      n_splits = static_cast<int>(avail / per);
      if (n_splits < 1) n_splits = 1;
    }
    if (n_splits > 1) {
      part_o = reinterpret_cast<float*>(static_cast<char*>(p.workspace) + kAttnCounterBytes);
      part_ml = part_o + static_cast<size_t>(n_splits) * p.n_heads * p.q_len * D;
    }
  }

  CUtensorMap mq, mk, mv;
  if (int rc = make_map(&mq, p.q, D, rows_total, p.n_kv_heads, static_cast<long long>(rows_total) * D)) return rc;
  const long long kv_rows = static_cast<long long>(p.n_cached) + p.q_len;
  if (int rc = make_map(&mk, p.k, D, kv_rows, p.n_kv_heads, static_cast<long long>(p.max_ctx) * D)) return rc;
  if (int rc = make_map(&mv, p.v, D, kv_rows, p.n_kv_heads, static_cast<long long>(p.max_ctx) * D)) return rc;

  TcParams tp;
  tp.out = static_cast<__half*>(p.out);
  tp.part_o = part_o;
  tp.part_ml = part_ml;
  tp.alibi_slopes = p.alibi_slopes;
  tp.key_pos = p.key_pos;
  tp.n_heads = p.n_heads;
  tp.n_kv_heads = p.n_kv_heads;
  tp.n_cached = p.n_cached;
  tp.q_len = p.q_len;
  tp.n_splits = n_splits;
  tp.m_blocks = m_blocks;
  tp.alibi_max_pos = p.alibi_max_pos;
  tp.scale = p.scale;
  const int grid = base * n_splits;

  int rc;
  const bool alibi = p.alibi_slopes != nullptr;
  if (p.rounding == 0)
    rc = alibi ? launch_one<D, 0, true>(mq, mk, mv, tp, grid, st) : launch_one<D, 0, false>(mq, mk, mv, tp, grid, st);
  else if (p.rounding == PC_ROUND_LLAMA)
    rc = alibi ? launch_one<D, 3, true>(mq, mk, mv, tp, grid, st) : launch_one<D, 3, false>(mq, mk, mv, tp, grid, st);
  else
    return set_error(PC_ERR_UNSUPPORTED, "attn(tcgen05): rounding flags %d (supported: 0, %d)", p.rounding,
                     PC_ROUND_LLAMA);
  if (rc) return rc;
  if (n_splits > 1) {
    dim3 cgrid(p.q_len, p.n_heads);
    attn_combine_kernel<D><<<cgrid, D, 0, st>>>(part_o, part_ml, static_cast<__half*>(p.out), p.n_heads, p.q_len,
                                               n_splits);
    PC_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return PC_OK;
}

}  // namespace

int launch_attn_combine(const float* part_o, const float* part_ml, void* out, int n_heads, int q_len, int n_splits,
                        int head_dim, cudaStream_t st) {
  dim3 cgrid(q_len, n_heads);
  if (head_dim == 128)
    attn_combine_kernel<128><<<cgrid, 128, 0, st>>>(part_o, part_ml, static_cast<__half*>(out), n_heads, q_len, n_splits);
  else
    attn_combine_kernel<64><<<cgrid, 64, 0, st>>>(part_o, part_ml, static_cast<__half*>(out), n_heads, q_len, n_splits);
  PC_CHECK_CUDA(cudaGetLastError());
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

int launch_attn_tc(const pc_attn_params& p, cudaStream_t st) {
  if (p.head_dim == 128) return launch_d<128>(p, st);
  if (p.head_dim == 64) return launch_d<64>(p, st);
  return set_error(PC_ERR_UNSUPPORTED, "attn(tcgen05): head_dim %d not in {64,128}", p.head_dim);
}

}  // namespace pc
