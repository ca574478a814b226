// tcgen05 / TMEM attention for LONG suffixes (more than 32 folded query rows per KV head): no-cache prefill, long
// uncached suffixes, Falcon's MQA (71 heads folded into M).  Same contract as attn_tc_small.cu; this is the
// tensor-bound regime (G*q FLOP per KV byte above the ~254 FLOP/B ridge).
//
// One CTA = (kv head, PAIR of 128-row query tiles A and B, KV split).  Both tiles share every K/V tile; each tile
// has its own softmax warpgroup and its own MMA-issuing thread (a single thread could not issue the 24 MMAs per
// 64 keys fast enough), so the tensor core always has the other tile's work queued while one softmax runs:
//
//   MMA thread :  S_A(0) S_B(0) S_A(1) S_B(1) | PV_A(0) S_A(2) | PV_B(0) S_B(2) | PV_A(1) S_A(3) | ...
//   WG0 (tile A):   softmax A(0) | softmax A(1) | softmax A(2) | ...      (S_A(j+1) is already waiting when
//   WG1 (tile B):   softmax B(0) | softmax B(1) | softmax B(2) | ...       softmax A(j) finishes)
//
//   S_X = Q_X . K_j^T   SS-MMA 128 x 64 keys x D, accumulator in TMEM, TWO buffers per query tile
//   P_X                 one TMEM lane = one query row per thread; packed fp16 written back over S_X
//   O_X += P_X . V_j    TS-MMA (A = P_X from TMEM, B = V tile MN-major), accumulator in TMEM
//
// KV tiles are 64 keys so that each query tile can double-buffer S within the 512 TMEM columns:
//   S_A0[0,64) S_A1[64,128) S_B0[128,192) S_B1[192,256) O_A[256,256+D) O_B[256+D,256+2D).
// Softmax reads a 64-key row of S from TMEM once (64 registers).  Pass 1 keeps only the row maximum of the RAW
// accumulators (the fp16-rounding + scale transform is monotone, so max(transform) = transform(max)); pass 2
// applies the reference's roundings in packed fp16, exponentiates (ex2.approx) and writes P.  Rescaling of O is
// lazy (only when the row maximum grew by more than 2^8).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "pc_common.cuh"

namespace pc {

int make_tensor_map_3d(CUtensorMap* map, const void* base, int D, long long rows, int heads, long long head_stride,
                       int box_rows);
int launch_attn_combine(const float* part_o, const float* part_ml, void* out, int n_heads, int q_len, int n_splits,
                        int head_dim, cudaStream_t st);

namespace {

constexpr int BM = 128;
#ifndef PC_PP_BN
#define PC_PP_BN 64
#endif
// Keys per K/V tile = N of the S MMA.  64 keys leave room for TWO S buffers inside a query tile's 128 TMEM columns, so
// S_x(j+1) executes while softmax_x(j) runs.  128 keys (-DPC_PP_BN=128: one S buffer, P written back over it, FA-style
// ping-pong between the two query tiles only, 2 K/V stages) halve the S MMA's shared-memory traffic per FLOP but
// serialise each tile's S -> softmax -> PV chain: measured SLOWER on the B200 (q=1024 over N=8192: 667 vs 864 TFLOP/s;
// without any softmax math 946 vs 1110), so 64 stays the default.  The softmax below handles both widths.
constexpr int BN = PC_PP_BN;
static_assert(BN == 64 || BN == 128, "PC_PP_BN must be 64 or 128");
constexpr int kSBufs = 128 / BN;  // S buffers per query tile
constexpr int kThreads = 352;  // WG0 (warps 0-3), WG1 (warps 4-7), TMA warp 8, MMA warps 9 (tile A) and 10 (tile B)
constexpr int kMaxSplits = 32;
constexpr float kLog2e = 1.4426950408889634f;
// Packed-fp16 exponentials in pass 2 (non-ALiBi): the softmax warps share their SM sub-partition's 4 MUFU lanes with the
// other query tile's warpgroup, and at one fp32 ex2 per score the MUFU pipe is busy exactly as long as the tensor pipe
// (8192 ex2 per 64-key tile and query tile = 512 cycles; the tile's MMAs: 512 cycles at D = 128, 256 at D = 64).
// ex2.approx.f16x2 halves that; P is rounded to fp16 for the P.V MMA anyway, and the row sum is taken from the same rounded
// values.  The running maximum is kept fp16-representable so that (score*log2e - m) is formed by one fp16 FMA per pair.
// Measured (B200, -DPC_PP_EX2_H2=1): q=1024 over N=8192 871 -> 888 TFLOP/s, Falcon MQA q=256 unchanged, and the llama-rounding
// parity case at q=1024 leaves the 1e-3 bar -- the softmax warps are bound by instruction issue, not by the MUFU pipe.  Off.
#ifndef PC_PP_EX2_H2
#define PC_PP_EX2_H2 0
#endif
constexpr float kRescaleThreshold = 8.0f;

template <int D>
struct Cfg2 {
  static constexpr int kQBytes = BM * D * 2;     // one query tile
  static constexpr int kTileBytes = BN * D * 2;  // one K or V tile
  static constexpr int kHalfStride = BN * 128;   // bytes between the two 64-column halves of a K/V tile (D=128)
  // as many (K, V) stages as fit beside the two query tiles (D=128: 4 at 64 keys, 2 at 128 keys; D=64: 4)
  static constexpr int kStagesFit = (220 * 1024 - 2 * kQBytes) / (2 * kTileBytes);
  static constexpr int kStages = kStagesFit < 4 ? kStagesFit : 4;
  static_assert(kStages >= 2, "the K/V ring needs two stages");
  static constexpr int kOffKV = 2 * kQBytes;     // Q_A, Q_B first
  static constexpr int kOffBars = kOffKV + kStages * 2 * kTileBytes;
  static constexpr int kSmem = 1024 + kOffBars + 256;
  static constexpr int kTmemCols = 512;
  static constexpr int kColO = 256;
};

struct PpParams {
  __half* out;
  float* part_o;
  float* part_ml;
  const float* alibi_slopes;
  const int* key_pos;
  const int* n_cached_dev;  // optional run-time N (graph-replayable decode): n_cached is then the sizing upper bound
  int n_heads, n_kv_heads, n_cached, q_len, n_splits, m_pairs, alibi_max_pos;
  float scale;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two exponentials per MUFU operation: 2^x of a packed fp16 pair, result packed fp16 (max relative error 2^-9.9)
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// round two fp32 values to fp16 and back (one packed cvt + two unpacks)
__device__ __forceinline__ void round_h2(float& a, float& b) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 f = __half22float2(h);
  a = f.x;
  b = f.y;
}
__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }

// Causal extent (exclusive key bound) of the 128-row tile starting at folded row `row_lo`.
__device__ __forceinline__ int tile_kv_end(int row_lo, int rows_total, int q_len, int n_cached) {
  if (row_lo >= rows_total) return 0;
  const int row_hi = min(row_lo + BM, rows_total) - 1;
  const int r_max = (row_lo / q_len == row_hi / q_len) ? row_hi % q_len : q_len - 1;
  return n_cached + r_max + 1;
}

template <int D, int ROUND, bool ALIBI>
__global__ void __launch_bounds__(kThreads, 1)
attn_pp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const PpParams pin) {
  using C = Cfg2<D>;
  PpParams p = pin;

  // run-time cached length: the grid, the split count and the tensor maps were sized for the upper bound; splits that end
  // up without tiles publish an empty partial (m = -inf, l = 0), rows past N + q are masked by the causal bound.  The value
  // is written by an earlier kernel of the stream: this variant waits for its predecessors first, the static one after its
  // barriers / TMEM / descriptors are set up (below).
  // (both are done after the set-up below: the run-time length is written by an earlier kernel of the stream)
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  unsigned char* sQ = smem;  // [2][kTileBytes]
  unsigned char* sKV = smem + C::kOffKV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + C::kStages;
  uint64_t* kv_empty = v_full + C::kStages;      // [2 query tiles][stages]: released by each tile's MMA thread
  uint64_t* s_full = kv_empty + 2 * C::kStages;  // [2 query tiles][2 S buffers]
  uint64_t* p_full = s_full + 4;             // [2][2]
  // o_full is also indexed by tile parity: a softmax warpgroup may run two tiles ahead of the P.V MMAs, and a
  // parity wait is only unambiguous within one phase of its barrier
  uint64_t* o_full = p_full + 4;             // [2][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = p.n_heads / p.n_kv_heads;
  const int rows_total = G * p.q_len;

  // valid 32-row warps of the two query tiles (static: needed for the barrier counts; kept in scalars -- arrays defined
  // before the inline-asm wait below would be forced into local memory by its memory clobber)
  const int mp_static = p.m_pairs - 1 - (static_cast<int>(blockIdx.x) / p.n_splits) / p.n_kv_heads;
  const int rl0 = mp_static * 2 * BM, rl1 = rl0 + BM;
  const int vw0 = rl0 < rows_total ? (min(BM, rows_total - rl0) + 31) / 32 : 0;
  const int vw1 = rl1 < rows_total ? (min(BM, rows_total - rl1) + 31) / 32 : 0;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&kv_empty[C::kStages + s], 1);
    }
    for (int x = 0; x < 2; ++x) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[x * 2 + b], 1);
        mbar_init(&p_full[x * 2 + b], max(32 * (x ? vw1 : vw0), 1));
        mbar_init(&o_full[x * 2 + b], 1);
      }
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything below reads what earlier kernels wrote (Q, the appended K/V rows, the run-time length) or writes what they read
#ifndef PC_PP_NO_PDL
  pdl_wait();
#endif
  if (pin.n_cached_dev != nullptr) p.n_cached = max(0, min(__ldg(pin.n_cached_dev), pin.n_cached));

  // Block order = longest work first: the causal extent grows with the row-pair index, so the last pairs of
  // every head are dispatched first and the short ones fill the tail of the grid.
  int bid = blockIdx.x;
  const int split = bid % p.n_splits;
  bid /= p.n_splits;
  const int hk = bid % p.n_kv_heads;
  const int mp = p.m_pairs - 1 - bid / p.n_kv_heads;

  // per-tile KV tile counts; both tiles walk the same global tile indices [t0, t1)
  const int row_lo[2] = {mp * 2 * BM, mp * 2 * BM + BM};
  int TT[2];
#pragma unroll
  for (int x = 0; x < 2; ++x) TT[x] = (tile_kv_end(row_lo[x], rows_total, p.q_len, p.n_cached) + BN - 1) / BN;
  const int TTmax = max(TT[0], TT[1]);
  const int t0 = static_cast<int>(static_cast<long long>(split) * TTmax / p.n_splits);
  const int t1 = static_cast<int>(static_cast<long long>(split + 1) * TTmax / p.n_splits);
  int T[2];
#pragma unroll
  for (int x = 0; x < 2; ++x) T[x] = max(0, min(t1, TT[x]) - t0);
  const int Tmax = max(T[0], T[1]);
  int valid_warps[2];
#pragma unroll
  for (int x = 0; x < 2; ++x) valid_warps[x] = row_lo[x] < rows_total ? (min(BM, rows_total - row_lo[x]) + 31) / 32 : 0;


  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0 && Tmax > 0) {
      const int nq = (T[0] > 0 ? 1 : 0) + (T[1] > 0 ? 1 : 0);
      mbar_expect_tx(q_full, nq * C::kQBytes);
      for (int x = 0; x < 2; ++x)
        if (T[x] > 0)
          for (int hf = 0; hf < D / 64; ++hf)
            tma_load_3d(sQ + x * C::kQBytes + hf * 16384, &map_q, q_full, hf * 64, row_lo[x], hk);
      for (int j = 0; j < Tmax; ++j) {
        const int s = j % C::kStages;
        if (j >= C::kStages) {  // both tiles' MMAs that read tile j-kStages must have completed
          for (int x = 0; x < 2; ++x)
            if (j - C::kStages < T[x])
              mbar_wait(&kv_empty[x * C::kStages + s], static_cast<uint32_t>(((j / C::kStages) - 1) & 1));
        }
        unsigned char* dk = sKV + (2 * s) * C::kTileBytes;
        unsigned char* dv = dk + C::kTileBytes;
        const int key0 = (t0 + j) * BN;
        mbar_expect_tx(&k_full[s], C::kTileBytes);
        for (int hf = 0; hf < D / 64; ++hf)
          tma_load_3d(dk + hf * C::kHalfStride, &map_k, &k_full[s], hf * 64, key0, hk);
        mbar_expect_tx(&v_full[s], C::kTileBytes);
        for (int hf = 0; hf < D / 64; ++hf)
          tma_load_3d(dv + hf * C::kHalfStride, &map_v, &v_full[s], hf * 64, key0, hk);
      }
    }
#ifndef PC_PP_NO_PDL
    // Programmatic launch of the dependents (the combine kernel, the next projection) once this CTA's last loads are out:
    // late enough that their CTAs do not sit on the SMs for the whole kernel (with the trigger at the top the 18 k-CTA
    // combine grid of the Falcon MQA shape cost 6.5 us), early enough to hide their launch latency.
    if (lane == 0) pdl_launch_dependents();
#endif
  } else if (warp == 9 || warp == 10) {
    // ===================== MMA issuers: warp 9 drives query tile A, warp 10 tile B =====================
    const int x = warp - 9;
    const int Tx = T[x];
    if (lane == 0 && Tx > 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(BM, D, 0, 1);
      // descriptors: everything but the start address is loop-invariant; offsets are added to the low word
      const uint64_t q_desc = make_smem_desc(smem_u32(sQ + x * C::kQBytes), 16, 1024, 2);
      const uint64_t k_desc0 = make_smem_desc(smem_u32(sKV), 16, 1024, 2);
      const uint64_t v_desc0 = make_smem_desc(smem_u32(sKV + C::kTileBytes), C::kHalfStride, 1024, 2);
      const uint32_t tmem_sx = tmem_base + x * 128;
      const uint32_t tmem_ox = tmem_base + C::kColO + x * D;
      auto wait_k = [&](int j) {
        mbar_wait(&k_full[j % C::kStages], static_cast<uint32_t>((j / C::kStages) & 1));
        tc_fence_after();
      };
      auto issue_s = [&](int j) {  // S_x(j) = Q_x . K_j^T into S buffer (x, j % kSBufs); k_full(j) already waited
        const uint64_t k_desc = k_desc0 + static_cast<uint64_t>(((j % C::kStages) * 2 * C::kTileBytes) >> 4);
        const uint32_t d = tmem_sx + (j % kSBufs) * BN;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_ss(d, q_desc + static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4),
                  k_desc + static_cast<uint64_t>(((kk >> 2) * C::kHalfStride + (kk & 3) * 32) >> 4), idesc_qk,
                  kk > 0 ? 1u : 0u);
        tc_commit(&s_full[x * 2 + (j % kSBufs)]);
      };
      auto issue_pv = [&](int j) {  // O_x += P_x(j) . V_j
        const uint64_t v_desc = v_desc0 + static_cast<uint64_t>(((j % C::kStages) * 2 * C::kTileBytes) >> 4);
        const uint32_t pa = tmem_sx + (j % kSBufs) * BN;
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)
          umma_ts(tmem_ox, pa + kk * 8, v_desc + static_cast<uint64_t>((kk * 2048) >> 4), idesc_pv,
                  (j > 0 || kk > 0) ? 1u : 0u);
        tc_commit(&o_full[x * 2 + (j & 1)]);  // by tile parity (see o_full)
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < kSBufs && j < Tx; ++j) {  // kSBufs S tiles in flight per query tile
        wait_k(j);
        issue_s(j);
      }
      for (int j = 0; j < Tx; ++j) {
        const int s = j % C::kStages;
        const int jn = j + kSBufs;  // the S tile that reuses P_x(j)'s buffer
        // operands first (they landed long ago), so that nothing but the MMA issue follows the arrival of P_x(j)
        mbar_wait(&v_full[s], static_cast<uint32_t>((j / C::kStages) & 1));
        if (jn < Tx) mbar_wait(&k_full[jn % C::kStages], static_cast<uint32_t>((jn / C::kStages) & 1));
        mbar_wait(&p_full[x * 2 + (j % kSBufs)], static_cast<uint32_t>((j / kSBufs) & 1));
        tc_fence_after();
        issue_pv(j);
        if (jn < Tx) issue_s(jn);  // the S buffer is free once PV_x(j) has been issued (in-order execution)
        tc_commit(&kv_empty[x * C::kStages + s]);  // this tile's MMAs on stage s have all been issued
      }
    }
  } else {
    // ===================== softmax warpgroups (WG x owns query tile x) =====================
    const int x = warp >> 2;
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_sx = tmem_base + lane_base + x * 128;
    const uint32_t tmem_o = tmem_base + lane_base + C::kColO + x * D;
    const int R = row_lo[x] + quad * 32 + lane;
    const bool warp_valid = quad < valid_warps[x];
    const bool row_valid = R < rows_total;
    const int g = row_valid ? R / p.q_len : 0;
    const int r = row_valid ? R % p.q_len : 0;
    const int h = hk * G + g;
    const int Tx = T[x];
    const float slope = (ALIBI && row_valid) ? p.alibi_slopes[h] : 0.f;
    const int vis_end = row_valid ? p.n_cached + r + 1 : 0;  // keys [0, vis_end) are visible to this row
    const float scale_l2 = p.scale * kLog2e;
    // fp32 scale split into two fp16 terms: half(half(acc)*scale) == hfma2(a16, s_hi, a16*s_lo) up to a
    // double-rounding event of probability ~2e-4 per score (1 fp16 ulp when it happens)
    const __half s_hi = __float2half_rn(p.scale);
    const __half2 scale_hi = __half2half2(s_hi);
    const __half2 scale_lo = __half2half2(__float2half_rn(p.scale - __half2float(s_hi)));
    float m_run = -INFINITY, l_run = 0.f;  // m in the log2 domain

    if (warp_valid) {
      constexpr int CH = 64;           // columns of S handled at a time (64 registers)
      constexpr int kChunks = BN / CH;  // 1 at 64-key tiles, 2 at 128-key tiles
      for (int j = 0; j < Tx; ++j) {
        const int b = j % kSBufs;
        const uint32_t tmem_s = tmem_sx + b * BN;
        mbar_wait(&s_full[x * 2 + b], static_cast<uint32_t>((j / kSBufs) & 1));
        tc_fence_after();
        const int key0 = (t0 + j) * BN;
        const bool need_mask = key0 + BN > vis_end;

#if defined(PC_EXP_NO_SOFTMAX)  // experiment: MMA / barrier chain alone (P = garbage, no math)
        {
          uint32_t z[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) z[c] = 0x3c003c00u;
#pragma unroll
          for (int ck = 0; ck < kChunks; ++ck) tmem_st32(tmem_s + ck * 32, z);
          l_run += 1.f;
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[x * 2 + b]);
          continue;
        }
#endif
        const bool warp_mask = __any_sync(0xffffffffu, need_mask);  // warp-uniform: mask code only near the diagonal
        // a 64-column chunk of the S row in registers (one TMEM read; with 128-key tiles each chunk is read in both passes)
        float sv[CH];
        auto load_chunk = [&](int ck) {
          tmem_ld32(tmem_s + ck * CH, reinterpret_cast<uint32_t(&)[32]>(sv[0]));
          tmem_ld32(tmem_s + ck * CH + 32, reinterpret_cast<uint32_t(&)[32]>(sv[32]));
          tmem_ld_wait();
        };
        // the reference's transform of one raw score (ALiBi: natural-log domain incl. the key's bias)
        auto alibi_score = [&](float t, int key) {
          if (ROUND & PC_ROUND_QK) t = round_h(t);
          t *= p.scale;
          if (ROUND & PC_ROUND_SCALED) t = round_h(t);
          return t + slope * static_cast<float>(__ldg(p.key_pos + min(key, p.n_cached + p.q_len - 1)) - p.alibi_max_pos);
        };

        // ---- pass 1: row maximum (log2 domain) ----
        float mt = -INFINITY;
        if (!ALIBI) {
          // max over the RAW accumulators; the fp16-rounding + scale transform is monotone, so it commutes with max
          float raw_max = -INFINITY;
#pragma unroll
          for (int ck = 0; ck < kChunks; ++ck) {
            load_chunk(ck);
            const int kc = key0 + ck * CH;
            if (warp_mask) {
#pragma unroll
              for (int c = 0; c < CH; ++c)
                if (kc + c < vis_end) raw_max = fmaxf(raw_max, sv[c]);
            } else {
#pragma unroll
              for (int c = 0; c < CH; c += 2) raw_max = fmaxf(raw_max, fmaxf(sv[c], sv[c + 1]));
            }
          }
          float t = raw_max;
          if (ROUND == PC_ROUND_LLAMA) {
            const __half2 a16 = __floats2half2_rn(t, t);
            t = __low2float(__hfma2(a16, scale_hi, __hmul2(a16, scale_lo)));
            if (raw_max == -INFINITY) t = -INFINITY;  // (-inf)*s_lo may be +inf -> NaN in the fused form
          } else {
            if (ROUND & PC_ROUND_QK) t = round_h(t);
            t *= p.scale;
            if (ROUND & PC_ROUND_SCALED) t = round_h(t);
          }
          mt = t * kLog2e;
        } else {
          // ALiBi adds a per-key bias: every score is transformed before the max (and again in pass 2 at 128-key tiles)
#pragma unroll
          for (int ck = 0; ck < kChunks; ++ck) {
            load_chunk(ck);
            const int kc = key0 + ck * CH;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              const float t = alibi_score(sv[c], kc + c);
              if (kChunks == 1) sv[c] = t;  // single chunk: pass 2 reuses the transformed row
              if (kc + c < vis_end) mt = fmaxf(mt, t * kLog2e);
            }
          }
        }

        // ---- running maximum, lazy rescale of O ----
        if (PC_PP_EX2_H2 && !ALIBI) mt = round_h(mt);  // fp16-representable maximum (a common shift of the row: cancels in p / l)
        const float m_new = fmaxf(m_run, mt);
        const bool grow = (m_new - m_run) > kRescaleThreshold;  // first finite max: +inf; both -inf: NaN -> false
        const float alpha = grow ? ex2(m_run - m_new) : 1.f;
        if (grow) {
          m_run = m_new;
          l_run *= alpha;
        }
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // PV_x(j-1) must have landed (with two S buffers S_x(j) was issued two tiles ahead: not implied by s_full)
          mbar_wait(&o_full[x * 2 + ((j - 1) & 1)], static_cast<uint32_t>(((j - 1) >> 1) & 1));
          tc_fence_after();
#pragma unroll
          for (int ch = 0; ch < D / 32; ++ch) {
            uint32_t o[32];
            tmem_ld32(tmem_o + ch * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st32(tmem_o + ch * 32, o);
          }
        }
        const float m_use = (m_run == -INFINITY) ? 0.f : m_run;

        // ---- pass 2: P = 2^(score - m), written over S as packed fp16 (32 TMEM columns per 64 keys) ----
        float lt = 0.f;
#pragma unroll
        for (int ck = 0; ck < kChunks; ++ck) {
          if (kChunks > 1) load_chunk(ck);  // (single chunk: sv still holds the row)
          const int kc = key0 + ck * CH;
          uint32_t pk[CH / 2];
          auto pass2 = [&](auto masked_tag) {
            constexpr bool kMasked = decltype(masked_tag)::value;
            if constexpr (PC_PP_EX2_H2 && !ALIBI && (ROUND == 0 || ROUND == PC_ROUND_LLAMA)) {
              const __half2 neg_m = __float2half2_rn(-m_use);  // exact: m is fp16-representable
              const __half2 l2e_hi = __float2half2_rn(kLog2e);
              const __half2 l2e_lo = __float2half2_rn(kLog2e - __half2float(__float2half_rn(kLog2e)));
              const __half2 ninf = __float2half2_rn(-INFINITY);
#pragma unroll
              for (int c = 0; c < CH; c += 2) {
                __half2 y;
                if (ROUND == PC_ROUND_LLAMA) {
                  // half(half(acc) * scale), then (t * log2e - m) with log2e = hi + lo: the first FMA is exact up to its one
                  // rounding at |y| (small where p matters), the second adds the t * lo correction
                  const __half2 a16 = __floats2half2_rn(sv[c], sv[c + 1]);
                  const __half2 t = __hfma2(a16, scale_hi, __hmul2(a16, scale_lo));
                  y = __hfma2(t, l2e_lo, __hfma2(t, l2e_hi, neg_m));
                } else {
                  y = __floats2half2_rn(fmaf(sv[c], scale_l2, -m_use), fmaf(sv[c + 1], scale_l2, -m_use));
                }
                if (kMasked) {
                  if (kc + c >= vis_end) y = __halves2half2(__low2half(ninf), __high2half(y));
                  if (kc + c + 1 >= vis_end) y = __halves2half2(__low2half(y), __low2half(ninf));
                }
                const uint32_t pr = ex2_h2(*reinterpret_cast<const uint32_t*>(&y));
                const float2 pf = __half22float2(*reinterpret_cast<const __half2*>(&pr));
                lt += pf.x + pf.y;
                pk[c >> 1] = pr;
              }
              return;
            }
#pragma unroll
            for (int c = 0; c < CH; c += 2) {
              const float a = sv[c], b2 = sv[c + 1];
              float ya, yb;
              if (ALIBI) {
                const float ta = kChunks == 1 ? a : alibi_score(a, kc + c);
                const float tb = kChunks == 1 ? b2 : alibi_score(b2, kc + c + 1);
                ya = fmaf(ta, kLog2e, -m_use);
                yb = fmaf(tb, kLog2e, -m_use);
              } else if (ROUND == 0) {
                ya = fmaf(a, scale_l2, -m_use);
                yb = fmaf(b2, scale_l2, -m_use);
              } else if (ROUND == PC_ROUND_LLAMA) {
                // half(half(acc) * scale) in packed fp16: scale = s_hi + s_lo (both fp16), one fused rounding
                const __half2 a16 = __floats2half2_rn(a, b2);
                const float2 f = __half22float2(__hfma2(a16, scale_hi, __hmul2(a16, scale_lo)));
                ya = fmaf(f.x, kLog2e, -m_use);
                yb = fmaf(f.y, kLog2e, -m_use);
              } else {
                float ta = a, tb = b2;
                if (ROUND & PC_ROUND_QK) round_h2(ta, tb);
                ta *= p.scale;
                tb *= p.scale;
                if (ROUND & PC_ROUND_SCALED) round_h2(ta, tb);
                ya = fmaf(ta, kLog2e, -m_use);
                yb = fmaf(tb, kLog2e, -m_use);
              }
              if (kMasked) {
                if (kc + c >= vis_end) ya = -INFINITY;
                if (kc + c + 1 >= vis_end) yb = -INFINITY;
              }
#if defined(PC_EXP_NO_MUFU)  // experiment: how much of the time is the exponential?
              const float p0 = ya * 1e-3f, p1 = yb * 1e-3f;
#else
              const float p0 = ex2(ya), p1 = ex2(yb);
#endif
              lt += p0 + p1;
              pk[c >> 1] = pack_h2(p0, p1);
            }
          };
          if (warp_mask) pass2(std::true_type{});
          else pass2(std::false_type{});
          // P chunk ck covers the first half of S chunk ck's columns and, for ck = 1, S columns 32-63: every S value
          // that lands under it is already in registers or consumed
          tmem_st32(tmem_s + ck * (CH / 2), pk);
        }
        l_run += lt;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[x * 2 + b]);
      }

      // ---- epilogue ----
      const int n_heads = p.n_heads;
      if (Tx > 0) {
        mbar_wait(&o_full[x * 2 + ((Tx - 1) & 1)], static_cast<uint32_t>(((Tx - 1) >> 1) & 1));
        tc_fence_after();
      }
      const float inv_l = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t o[32];
        if (Tx > 0) {
          tmem_ld32(tmem_o + ch * 32, o);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = 0u;
        }
        if (row_valid) {
          if (p.n_splits == 1) {
            __half* dst = p.out + (static_cast<long long>(r) * n_heads + h) * D + ch * 32;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              uint4 w;
              w.x = pack_h2(__uint_as_float(o[v * 8 + 0]) * inv_l, __uint_as_float(o[v * 8 + 1]) * inv_l);
              w.y = pack_h2(__uint_as_float(o[v * 8 + 2]) * inv_l, __uint_as_float(o[v * 8 + 3]) * inv_l);
              w.z = pack_h2(__uint_as_float(o[v * 8 + 4]) * inv_l, __uint_as_float(o[v * 8 + 5]) * inv_l);
              w.w = pack_h2(__uint_as_float(o[v * 8 + 6]) * inv_l, __uint_as_float(o[v * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(dst + v * 8) = w;
            }
          } else {
            float* dst = p.part_o + ((static_cast<long long>(split) * n_heads + h) * p.q_len + r) * D + ch * 32;
#pragma unroll
            for (int v = 0; v < 8; ++v)
              *reinterpret_cast<uint4*>(dst + v * 4) = make_uint4(o[v * 4], o[v * 4 + 1], o[v * 4 + 2], o[v * 4 + 3]);
          }
        }
      }
      if (row_valid && p.n_splits > 1) {
        float* ml = p.part_ml + ((static_cast<long long>(split) * n_heads + h) * p.q_len + r) * 2;
        ml[0] = m_run;  // log2 domain; -inf with l = 0 when this split saw no visible key
        ml[1] = l_run;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_free<C::kTmemCols>(tmem_base);
  }
}

template <int D, int ROUND, bool ALIBI>
int launch_pp_one(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const PpParams& pp, int grid,
                  cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(attn_pp_kernel<D, ROUND, ALIBI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg2<D>::kSmem));
    configured = true;
  }
#ifndef PC_PP_NO_PDL
  PC_CHECK_CUDA(launch_pdl(attn_pp_kernel<D, ROUND, ALIBI>, dim3(grid), dim3(kThreads), Cfg2<D>::kSmem, st, mq, mk, mv, pp));
#else
  attn_pp_kernel<D, ROUND, ALIBI><<<grid, kThreads, Cfg2<D>::kSmem, st>>>(mq, mk, mv, pp);
  PC_CHECK_CUDA(cudaGetLastError());
#endif
  count_launch();
  return PC_OK;
}

template <int D>
int launch_pp_d(const pc_attn_params& p, cudaStream_t st) {
  const int G = p.n_heads / p.n_kv_heads;
  const int rows_total = G * p.q_len;
  const int m_pairs = (rows_total + 2 * BM - 1) / (2 * BM);
  const int tiles = (p.n_cached + p.q_len + BN - 1) / BN;
  const int base = p.n_kv_heads * m_pairs;
  int n_splits = sm_count() / base;
  // at least three 64-key tiles per split: below that the per-CTA prologue and the combine pass over the partials cost more
  // than the extra CTAs bring (Falcon MQA q=16 over N=2048: 29 splits 22.4 us, 11 splits 13.4 us; decode at N=2048: 32 splits
  // 11.3 us, 11 splits 9.1 us; C4's N=277: 5 splits + combine 9.9 us, one split 9.0 us)
  // (a single row pair -- decode -- has a cheap combine and nothing else to fill the SMs with: at least 4 splits there)
  // Applied to folded heads (G > 1: MQA / GQA, where few row pairs meet many SMs); G = 1 keeps one tile per split as the floor.
  if (G > 1) {
    int by_tiles = tiles / 3;
    if (m_pairs == 1 && by_tiles < 4) by_tiles = tiles < 4 ? tiles : 4;
    if (n_splits > by_tiles) n_splits = by_tiles;
  }
  if (const char* e = dbg_env("PC_ATTN_SPLITS")) n_splits = atoi(e);
  if (n_splits > tiles) n_splits = tiles;
  if (n_splits > kMaxSplits) n_splits = kMaxSplits;
  if (n_splits < 1) n_splits = 1;
  float* part_o = nullptr;
  float* part_ml = nullptr;
  if (n_splits > 1) {
    const size_t per = static_cast<size_t>(p.n_heads) * p.q_len * (D + 2) * sizeof(float);
    const size_t avail = (p.workspace && p.workspace_bytes > kAttnCounterBytes) ? p.workspace_bytes - kAttnCounterBytes : 0;
    if (avail < per * n_splits) n_splits = static_cast<int>(avail / per);
    if (n_splits < 1) n_splits = 1;
    if (n_splits > 1) {
      part_o = reinterpret_cast<float*>(static_cast<char*>(p.workspace) + kAttnCounterBytes);
      part_ml = part_o + static_cast<size_t>(n_splits) * p.n_heads * p.q_len * D;
    }
  }
  CUtensorMap mq, mk, mv;
  if (int rc = make_tensor_map_3d(&mq, p.q, D, rows_total, p.n_kv_heads, static_cast<long long>(rows_total) * D, BM)) return rc;
  // run-time N: the maps cover the whole buffer (rows past N+q are masked; they must hold finite values)
  const long long kv_rows = p.n_cached_dev ? static_cast<long long>(p.max_ctx) : static_cast<long long>(p.n_cached) + p.q_len;
  if (int rc = make_tensor_map_3d(&mk, p.k, D, kv_rows, p.n_kv_heads, static_cast<long long>(p.max_ctx) * D, BN)) return rc;
  if (int rc = make_tensor_map_3d(&mv, p.v, D, kv_rows, p.n_kv_heads, static_cast<long long>(p.max_ctx) * D, BN)) return rc;
  PpParams pp;
  pp.out = static_cast<__half*>(p.out);
  pp.part_o = part_o;
  pp.part_ml = part_ml;
  pp.alibi_slopes = p.alibi_slopes;
  pp.key_pos = p.key_pos;
  pp.n_cached_dev = p.n_cached_dev;
  pp.n_heads = p.n_heads;
  pp.n_kv_heads = p.n_kv_heads;
  pp.n_cached = p.n_cached;
  pp.q_len = p.q_len;
  pp.n_splits = n_splits;
  pp.m_pairs = m_pairs;
  pp.alibi_max_pos = p.alibi_max_pos;
  pp.scale = p.scale;
  const int grid = base * n_splits;
  const bool alibi = p.alibi_slopes != nullptr;
  int rc;
  if (p.rounding == 0)
    rc = alibi ? launch_pp_one<D, 0, true>(mq, mk, mv, pp, grid, st) : launch_pp_one<D, 0, false>(mq, mk, mv, pp, grid, st);
  else if (p.rounding == PC_ROUND_LLAMA)
    rc = alibi ? launch_pp_one<D, 3, true>(mq, mk, mv, pp, grid, st) : launch_pp_one<D, 3, false>(mq, mk, mv, pp, grid, st);
  else
    return set_error(PC_ERR_UNSUPPORTED, "attn(tcgen05): rounding flags %d (supported: 0, %d)", p.rounding, PC_ROUND_LLAMA);
  if (rc) return rc;
  if (n_splits > 1) return launch_attn_combine(part_o, part_ml, p.out, p.n_heads, p.q_len, n_splits, D, st);
  return PC_OK;
}

}  // namespace

int launch_attn_pp(const pc_attn_params& p, cudaStream_t st) {
  if (p.head_dim == 128) return launch_pp_d<128>(p, st);
  if (p.head_dim == 64) return launch_pp_d<64>(p, st);
  return set_error(PC_ERR_UNSUPPORTED, "attn(tcgen05): head_dim %d not in {64,128}", p.head_dim);
}

}  // namespace pc
