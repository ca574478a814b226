// tcgen05 attention for SHORT suffixes (G*q <= 32 folded query rows per KV head): prefill of a few
// uncached tokens over a long spliced cache, and single-token decode.  Same contract as attn_tc.cu.
//
// With 16 query rows a 128-row MMA tile would leave 112 TMEM lanes (and 3.5 of 4 softmax warps) idle and
// the per-tile softmax would serialise on half a warp.  This kernel swaps the operands instead:
//   S^T_j = K_j . Q^T      SS-MMA 128 keys x NPAD queries x D : TMEM lane = key, column = query row
//   P^T_j                  8 softmax warps: a thread owns one key and NPAD/2 query columns (two warps per
//                          TMEM lane quarter, i.e. two warps per scheduler to hide instruction latency); the
//                          per-query max / sum over the 128 keys of a tile is a halving warp butterfly
//                          (NPAD/2+1 shuffles) + a 4-way exchange through shared memory
//   O^T  += V_j^T . P^T_j  SS-MMA D x NPAD x 128 keys : A = V tile (MN-major), B = P^T from shared memory
// so every softmax thread has NPAD elements of work per tile, the K/V stream is the only large traffic and
// the kernel is HBM-bound.  Split-KV partials are merged by the LAST CTA of each KV head (arrival counter in
// the workspace), so there is no separate combine launch.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "pc_common.cuh"

namespace pc {

int make_tensor_map_3d(CUtensorMap* map, const void* base, int D, long long rows, int heads, long long head_stride,
                       int box_rows);

namespace {

constexpr int BK = 128;  // keys per tile
constexpr int D = 128;
constexpr int kSoftmaxWarps = 8;
constexpr int kSoftmaxThreads = kSoftmaxWarps * 32;
constexpr int kThreads = kSoftmaxThreads + 64;  // + TMA warp + MMA warp
constexpr int kStages = 3;
constexpr int kTileBytes = BK * D * 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRescaleThreshold = 8.0f;

template <int NPAD>
struct SCfg {
  static constexpr int kQBytes = NPAD * D * 2;
  static constexpr int kPBytes = NPAD * BK * 2;
  static constexpr int kOffQ = kStages * 2 * kTileBytes;
  static constexpr int kOffP = kOffQ + kQBytes;
  static constexpr int kOffRed = kOffP + 2 * kPBytes;           // float red[2][4][NPAD]
  static constexpr int kOffMisc = kOffRed + 2 * 4 * NPAD * 4;   // rtab, slopetab, otab, ptab [NPAD] each; int flag
  static constexpr int kOffBars = kOffMisc + 4 * NPAD * 4 + 16;
  static constexpr int kSmem = 1024 + kOffBars + 256;
  static constexpr int kTmemCols = NPAD <= 16 ? 64 : 128;       // S^T x2 + O^T = 3*NPAD
};

struct SmallParams {
  __half* out;
  float* part_o;
  float* part_ml;
  int* counters;
  const float* alibi_slopes;
  const int* key_pos;
  const int* n_cached_dev;  // optional run-time N (graph-replayable decode)
  int n_heads, n_kv_heads, n_cached, q_len, n_splits, alibi_max_pos;
  float scale;
  unsigned long long* trace;  // PC_DEBUG builds: %globaltimer stamps of CTA 0 (PC_ATTN_TRACE_PTR), else nullptr
  bool global_merge;          // PC_ATTN_GLOBAL_MERGE: merge split-KV partials through the workspace, not the cluster
  // virtual splice: the first v_rows keys are not in k/v but in n_vtiles tiles {arena row, valid rows} of the segment arena
  const int2* vtiles;         // device; nullptr = everything is read from k/v
  int n_vtiles, v_rows;
  // hint: bytes of the NEXT projection's pre-tiled weights, asked into L2 once this CTA's last K/V loads are out
  const unsigned char* pf;
  unsigned long long pf_bytes;
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef PC_DEBUG
#define PC_TRACE(slot)                                                   \
  do {                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[(slot)] = gtime(); \
  } while (0)
#else
#define PC_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ float ex2(float x) {  // 2^x, one MUFU; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t addr, float (&v)[N]) {
  if constexpr (N == 8) tmem_ld8(addr, reinterpret_cast<uint32_t(&)[8]>(v));
  else tmem_ld16(addr, reinterpret_cast<uint32_t(&)[16]>(v));
}
template <int N>
__device__ __forceinline__ void tmem_st_n(uint32_t addr, const float (&v)[N]) {
  if constexpr (N == 8) tmem_st8(addr, reinterpret_cast<const uint32_t(&)[8]>(v));
  else tmem_st16(addr, reinterpret_cast<const uint32_t(&)[16]>(v));
}
// Reduce N per-lane values across the 32 lanes of a warp with N/2 + N/4 + ... + 1 (+ log2(32/N)) shuffles: at each
// step a lane keeps one half of its vector and hands the other half to its partner.  Returns the reduction of
// column ((lane >> (5 - log2 N)) & (N-1)); v is destroyed.
template <int N, bool IS_MAX>
__device__ __forceinline__ float warp_reduce_vec(float (&v)[N], int lane) {
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      v[i] = IS_MAX ? fmaxf(keep, recv) : keep + recv;
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) {
    const float recv = __shfl_xor_sync(0xffffffffu, v[0], off);
    v[0] = IS_MAX ? fmaxf(v[0], recv) : v[0] + recv;
  }
  return v[0];
}
template <int N>
__device__ __forceinline__ int reduced_column(int lane) {
  constexpr int sh = N == 8 ? 2 : 1;  // 5 - log2(N)
  return (lane >> sh) & (N - 1);
}

// CLUSTER: the n_splits CTAs of a KV head form one thread-block cluster (n_splits <= 8) and the split-KV merge goes
// through the leader's shared memory (two cluster barriers + DSMEM stores, ~1.5 us) instead of global partials, an
// arrival counter and the last CTA's L2 round trips (~5 us of the 11 us kernel at the C2 shape).
template <int NPAD, int ROUND, bool ALIBI, bool CLUSTER>
__global__ void __launch_bounds__(kThreads, 1)
attn_small_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                  const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_ak,
                  const __grid_constant__ CUtensorMap map_av, const SmallParams pin) {
  using C = SCfg<NPAD>;
  pdl_launch_dependents();  // programmatic dependent launch: see pc_common.cuh
  SmallParams p = pin;
  if (pin.n_cached_dev != nullptr) pdl_wait();  // the run-time length is written by the previous kernels
  if (pin.n_cached_dev != nullptr) p.n_cached = max(pin.v_rows, min(__ldg(pin.n_cached_dev), pin.n_cached));
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  unsigned char* sKV = smem;
  unsigned char* sQ = smem + C::kOffQ;
  unsigned char* sP = smem + C::kOffP;
  float* red = reinterpret_cast<float*>(smem + C::kOffRed);
  int* rtab = reinterpret_cast<int*>(smem + C::kOffMisc);  // query row r of column c (c = g*q + r)
  float* slopetab = reinterpret_cast<float*>(rtab + NPAD);  // ALiBi slope of column c
  int* otab = reinterpret_cast<int*>(slopetab + NPAD);     // out row of column c:  r*H + h
  int* ptab = otab + NPAD;                                 // partial row of column c: h*q + r
  int* flag = ptab + NPAD;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + kStages;
  uint64_t* kv_empty = v_full + kStages;
  uint64_t* s_full = kv_empty + kStages;
  uint64_t* p_full = s_full + 2;
  // indexed by tile parity: a parity wait is only unambiguous within one phase of its barrier, and the softmax
  // warps can be up to two P.V MMAs ahead
  uint64_t* o_full = p_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = p.n_heads / p.n_kv_heads;
  const int cols_valid = G * p.q_len;  // <= NPAD
  const int split = blockIdx.x % pin.n_splits;
  const int hk = blockIdx.x / pin.n_splits;
  // Key order: [virtual tiles: cached segments read in place from the arena] [rows v_rows.. of k/v: rows appended by
  // earlier forwards (n_tail of them, all visible), then the q_len new rows (causal)].  Without a tile list vt = 0.
  const int vt = p.n_vtiles;
  const int n_tail = p.n_cached - p.v_rows;
  const int tiles_total = vt + (n_tail + p.q_len + BK - 1) / BK;
  // With a run-time N the grid was sized for an upper bound: only the first min(n_splits, tiles) splits work, the
  // other CTAs leave before touching any barrier, TMEM or the arrival counter.
  if (p.n_splits > tiles_total) p.n_splits = tiles_total;
  if (!CLUSTER && split >= p.n_splits) return;
  // in a cluster every CTA has to reach the cluster barriers: surplus splits stay, with no tiles and an empty partial
  const bool active = split < p.n_splits;
  const int tile0 = active ? static_cast<int>(static_cast<long long>(split) * tiles_total / p.n_splits) : 0;
  const int tile1 = active ? static_cast<int>(static_cast<long long>(split + 1) * tiles_total / p.n_splits) : 0;
  const int T = tile1 - tile0;  // >= 1 for active splits
  if (tid == 0) PC_TRACE(0);
  if (tid == 0 && p.trace != nullptr) p.trace[64 + 2 * blockIdx.x] = gtime();  // per-CTA start

  if (warp == kSoftmaxWarps && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], kSoftmaxThreads);
    }
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    mbar_fence_init();
  }
  if (warp == kSoftmaxWarps + 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  if (tid < NPAD) {
    rtab[tid] = tid < cols_valid ? tid % p.q_len : 0x3fffffff;  // padding columns: nothing is masked, never stored
    slopetab[tid] = (ALIBI && tid < cols_valid) ? p.alibi_slopes[hk * G + tid / p.q_len] : 0.f;
    const int h = hk * G + tid / p.q_len, r = tid % p.q_len;
    otab[tid] = r * p.n_heads + h;
    ptab[tid] = h * p.q_len + r;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * NPAD;
  if (tid == 0) PC_TRACE(1);  // setup done
  constexpr int NH = NPAD / 2;
  // per-thread result of this split (softmax warps only): O^T element d = kl of NH columns, their maxima and sums
  float o[NH], m_col[NH], l_col[NH];
#pragma unroll
  for (int i = 0; i < NH; ++i) {
    o[i] = 0.f;
    m_col[i] = -INFINITY;
    l_col[i] = 0.f;
  }

  if (warp == kSoftmaxWarps) {
    // ===================== TMA producer =====================
    if (lane == 0 && T > 0) {
      auto load_kv = [&](int j) {
        const int s = j % kStages;
        unsigned char* dk = sKV + (2 * s) * kTileBytes;
        unsigned char* dv = dk + kTileBytes;
        const int t = tile0 + j;
        const bool virt = t < vt;
        // a virtual tile may run past its segment into the next one's rows (masked by `valid`) or past the arena (zeros)
        const int key0 = virt ? __ldg(&p.vtiles[t].x) : p.v_rows + (t - vt) * BK;
        const CUtensorMap* mk = virt ? &map_ak : &map_k;
        const CUtensorMap* mv = virt ? &map_av : &map_v;
        mbar_expect_tx(&k_full[s], kTileBytes);
        for (int hf = 0; hf < 2; ++hf) tma_load_3d(dk + hf * 16384, mk, &k_full[s], hf * 64, key0, hk);
        mbar_expect_tx(&v_full[s], kTileBytes);
        for (int hf = 0; hf < 2; ++hf) tma_load_3d(dv + hf * 16384, mv, &v_full[s], hf * 64, key0, hk);
      };
      // Programmatic dependent launch: this CTA may be running while the previous kernels of the step (the QKV
      // projection, RoPE + append) are still in flight.  Rows [0, n_cached) of this layer's K/V were written before
      // the forward began (splice, earlier forwards) and nothing inside a forward touches them, so tiles that lie
      // entirely in them are streamed into the ring right away; Q and the tiles holding new rows wait.  (Virtual tiles
      // are cached rows by definition, and their tile list was uploaded before the forward began.)
      int pre = 0;
      while (pre < T && pre < kStages && (tile0 + pre < vt || (tile0 + pre - vt + 1) * BK <= n_tail)) load_kv(pre++);
      pdl_wait();
      mbar_expect_tx(q_full, C::kQBytes);
      for (int hf = 0; hf < 2; ++hf) tma_load_3d(sQ + hf * (NPAD * 128), &map_q, q_full, hf * 64, 0, hk);
      for (int j = pre; j < T; ++j) {
        if (j >= kStages) mbar_wait(&kv_empty[j % kStages], static_cast<uint32_t>(((j / kStages) - 1) & 1));
        load_kv(j);
      }
    }
    if (lane == 0 && p.pf != nullptr) {
      // This CTA's K/V stream is out.  The projection that follows cannot start before the CTAs of this kernel leave their
      // SMs (shared memory), so HBM would idle through the last tiles' softmax and the split merge: ask L2 for this CTA's
      // share of that projection's weight stream now (33.5 MB for Llama-2-7B's o_proj: it fits L2 many times over).
      constexpr unsigned long long kChunk = 32768;
      const unsigned long long chunks = (p.pf_bytes + kChunk - 1) / kChunk;
      for (unsigned long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const unsigned long long off = ch * kChunk;
        const unsigned long long n = min(kChunk, p.pf_bytes - off);
        bulk_prefetch_l2(p.pf + off, static_cast<uint32_t>(n & ~15ull));
      }
    }
  } else if (warp == kSoftmaxWarps + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && T > 0) {
      constexpr uint32_t idesc_st = make_idesc_f16(BK, NPAD, 0, 0);  // S^T = K . Q^T, both K-major
      constexpr uint32_t idesc_pv = make_idesc_f16(D, NPAD, 1, 1);   // O^T = V^T . P^T, both MN-major
      const uint32_t q_addr = smem_u32(sQ);
      auto issue_st = [&](int j) {
        const int s = j % kStages;
        mbar_wait(&k_full[s], static_cast<uint32_t>((j / kStages) & 1));
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sKV + (2 * s) * kTileBytes);
        const uint32_t d = tmem_base + (j & 1) * NPAD;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_ss(d, make_smem_desc(k_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024, 2),
                  make_smem_desc(q_addr + (kk >> 2) * (NPAD * 128) + (kk & 3) * 32, 16, 1024, 2), idesc_st,
                  kk > 0 ? 1u : 0u);
        tc_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      PC_TRACE(5);  // Q landed
      issue_st(0);
      PC_TRACE(6);  // first K landed and S^T_0 issued
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_st(j + 1);
        const int s = j % kStages;
        mbar_wait(&p_full[j & 1], static_cast<uint32_t>((j >> 1) & 1));
        mbar_wait(&v_full[s], static_cast<uint32_t>((j / kStages) & 1));
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sKV + (2 * s + 1) * kTileBytes);
        const uint32_t p_addr = smem_u32(sP + (j & 1) * C::kPBytes);
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk)
          umma_ss(tmem_o, make_smem_desc(v_addr + kk * 2048, 16384, 1024, 2),
                  make_smem_desc(p_addr + kk * 256, 128, 2048, 0), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        tc_commit(&kv_empty[s]);
        tc_commit(&o_full[j & 1]);
      }
    }
  } else {
    // ============ softmax (thread = key x half of the columns) / correction / epilogue (thread = d) ============
    pdl_wait();  // everything these warps read (key positions) or write (out, partials, counters) belongs to the step
    const int quad = warp & 3, ch = warp >> 2;           // TMEM lane quarter, column half
    const int kl = quad * 32 + lane;                      // key within the tile == TMEM lane == output d
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const int c0 = ch * NH;                               // first column owned by this thread
    const int my_col = c0 + reduced_column<NH>(lane);     // column whose warp-reduction this lane ends up holding
    float l_part[NH];
#pragma unroll
    for (int i = 0; i < NH; ++i) l_part[i] = 0.f;

    for (int j = 0; j < T; ++j) {
      const int b = j & 1;
      const int tj = tile0 + j;
      const int valid = tj < vt ? __ldg(&p.vtiles[tj].y) : BK;  // issued before the wait: off the critical path
      mbar_wait(&s_full[b], static_cast<uint32_t>((j >> 1) & 1));
      tc_fence_after();
      if (tid == 0 && j < 8) PC_TRACE(8 + 4 * j);  // S^T_j landed
      float x[NH];
      tmem_ld_n<NH>(tmem_base + lane_base + b * NPAD + c0, x);
      tmem_ld_wait();
      const int key = p.v_rows + (tj - vt) * BK + kl;  // index among all keys (tail tiles; virtual tiles never use it)
      float kp = 0.f;
      if (ALIBI) {
        // key ids: by key index; with a tile list the table is laid out in TILE order instead -- 128 slots per virtual tile,
        // then the appended rows -- so that slot tj * 128 + kl serves virtual and tail tiles alike
        const int kidx = vt > 0 ? min(tj * BK + kl, vt * BK + n_tail + p.q_len - 1) : min(key, p.n_cached + p.q_len - 1);
        kp = static_cast<float>(__ldg(p.key_pos + kidx) - p.alibi_max_pos);
      }
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        float v = x[i];
        if (ROUND & PC_ROUND_QK) v = round_h(v);
        v *= p.scale;
        if (ROUND & PC_ROUND_SCALED) v = round_h(v);
        if (ALIBI) v += slopetab[c0 + i] * kp;
        x[i] = v * kLog2e;
      }
      if (tj < vt) {  // virtual tile: rows past the segment's end belong to another segment (or are zero fill)
        if (kl >= valid) {
#pragma unroll
          for (int i = 0; i < NH; ++i) x[i] = -INFINITY;
        }
      } else if ((tj - vt + 1) * BK > n_tail) {  // tile touches the suffix: key n_cached+t is visible to row r iff t <= r
        const int t = key - p.n_cached;
#pragma unroll
        for (int i = 0; i < NH; ++i)
          if (t > rtab[c0 + i]) x[i] = -INFINITY;
      }
      // per-query max over the 128 keys of the tile
      float wm[NH];
#pragma unroll
      for (int i = 0; i < NH; ++i) wm[i] = x[i];
      const float wmax = warp_reduce_vec<NH, true>(wm, lane);
      float* rb = red + b * 4 * NPAD;
      if ((lane & (32 / NH - 1)) == 0) rb[quad * NPAD + my_col] = wmax;
      named_bar_sync(1, kSoftmaxThreads);
      if (tid == 0 && j < 8) PC_TRACE(9 + 4 * j);  // tile max exchanged
      bool any_grow = false;
      float alpha[NH];
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        const int c = c0 + i;
        const float mt = fmaxf(fmaxf(rb[c], rb[NPAD + c]), fmaxf(rb[2 * NPAD + c], rb[3 * NPAD + c]));
        const float m_new = fmaxf(m_col[i], mt);
        // first finite max of a column: (m_new - (-inf)) = +inf > threshold and alpha = 2^(-inf) = 0
        const bool grow = (m_new - m_col[i]) > kRescaleThreshold;  // NaN (both -inf) compares false
        alpha[i] = grow ? ex2(m_col[i] - m_new) : 1.f;
        m_col[i] = grow ? m_new : m_col[i];
        l_part[i] *= alpha[i];
        any_grow |= grow;
      }
      uint32_t pk[NH / 2];
#pragma unroll
      for (int i = 0; i < NH; i += 2) {
        const float p0 = ex2(x[i] - (m_col[i] == -INFINITY ? 0.f : m_col[i]));
        const float p1 = ex2(x[i + 1] - (m_col[i + 1] == -INFINITY ? 0.f : m_col[i + 1]));
        l_part[i] += p0;
        l_part[i + 1] += p1;
        pk[i >> 1] = pack_h2(p0, p1);
      }
      // P^T tile in shared memory, MN-major core matrices: 16 B (8 columns) per key, column blocks 2048 B apart
      unsigned char* pb = sP + b * C::kPBytes + (c0 / 8) * 2048 + kl * 16;
#pragma unroll
      for (int m8 = 0; m8 < NH / 8; ++m8)
        *reinterpret_cast<uint4*>(pb + m8 * 2048) = make_uint4(pk[m8 * 4], pk[m8 * 4 + 1], pk[m8 * 4 + 2], pk[m8 * 4 + 3]);
      fence_proxy_async_smem();
      // CTA-uniform per column half: every thread of a half derived the same maxima from the same smem values
      if (__any_sync(0xffffffffu, any_grow) && j > 0) {
        mbar_wait(&o_full[(j - 1) & 1], static_cast<uint32_t>(((j - 1) >> 1) & 1));
        tc_fence_after();
        float o[NH];
        tmem_ld_n<NH>(tmem_o + lane_base + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < NH; ++i) o[i] *= alpha[i];
        tmem_st_n<NH>(tmem_o + lane_base + c0, o);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(&p_full[b]);
      if (tid == 0 && j < 8) PC_TRACE(10 + 4 * j);  // P^T_j published
    }

    // ---- per-query sums over all keys of this split (independent of the last P.V, so done first) ----
    const float lsum = warp_reduce_vec<NH, false>(l_part, lane);
    named_bar_sync(1, kSoftmaxThreads);  // everyone is done reading `red` of the last tile
    if ((lane & (32 / NH - 1)) == 0) red[quad * NPAD + my_col] = lsum;
    named_bar_sync(1, kSoftmaxThreads);
#pragma unroll
    for (int i = 0; i < NH; ++i) {
      const int c = c0 + i;
      l_col[i] = (red[c] + red[NPAD + c]) + (red[2 * NPAD + c] + red[3 * NPAD + c]);
    }

    // ---- epilogue: O^T (TMEM lane = d, column = query row) ----
    if (T > 0) {
      mbar_wait(&o_full[(T - 1) & 1], static_cast<uint32_t>(((T - 1) >> 1) & 1));
      tc_fence_after();
      if (tid == 0) PC_TRACE(2);  // last PV landed
      tmem_ld_n<NH>(tmem_o + lane_base + c0, o);
      tmem_ld_wait();
    }
    const int d = kl;
    const int HQ = p.n_heads * p.q_len;
    if constexpr (CLUSTER) {
      // the merge happens after the role branches, where every thread of the CTA can reach the cluster barriers
    } else if (p.n_splits == 1) {
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        const int c = c0 + i;
        if (c < cols_valid) {
          const float inv = l_col[i] > 0.f ? 1.f / l_col[i] : 0.f;
          p.out[static_cast<long long>(otab[c]) * D + d] = __float2half_rn(o[i] * inv);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        const int c = c0 + i;
        if (c < cols_valid) {
          const long long idx = static_cast<long long>(split) * HQ + ptab[c];
          p.part_o[idx * D + d] = o[i];
          if (kl == i) *reinterpret_cast<float2*>(p.part_ml + idx * 2) = make_float2(m_col[i], l_col[i]);
        }
      }
      // The LAST CTA of this KV head merges the partials (no separate combine launch).  bar.sync orders every
      // thread's partial stores before thread 0's gpu-scope fence + arrival (fence cumulativity).
      named_bar_sync(1, kSoftmaxThreads);
      if (tid == 0) {
        // one acq_rel RMW: releases this CTA's partials (ordered before it by the bar.sync), acquires the other splits'
        int old;
        asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(p.counters + hk) : "memory");
        const int last = (old == p.n_splits - 1) ? 1 : 0;
        if (last) p.counters[hk] = 0;  // self-cleaning for the next launch (ordered by the kernel boundary)
        *flag = last;
      }
      named_bar_sync(1, kSoftmaxThreads);
      if (tid == 0) PC_TRACE(3);  // partials published
      if (*flag) {
        // stage (m, l) of every split x column in shared memory (reuses the P^T buffers: all MMAs are done)
        float2* ml_s = reinterpret_cast<float2*>(sP);  // [n_splits][NPAD]
        for (int e = tid; e < p.n_splits * cols_valid; e += kSoftmaxThreads) {
          const int s = e / cols_valid, c = e - s * cols_valid;
          ml_s[s * NPAD + c] = __ldcg(reinterpret_cast<const float2*>(p.part_ml) + static_cast<long long>(s) * HQ + ptab[c]);
        }
        named_bar_sync(1, kSoftmaxThreads);
        if (tid < cols_valid) {  // weights w_s = 2^(m_s - m*) / sum_s 2^(m_s - m*) l_s, written over m
          float m_star = -INFINITY;
          for (int s = 0; s < p.n_splits; ++s) m_star = fmaxf(m_star, ml_s[s * NPAD + tid].x);
          float l = 0.f;
          for (int s = 0; s < p.n_splits; ++s) {
            const float2 v = ml_s[s * NPAD + tid];
            const float w = v.x == -INFINITY ? 0.f : exp2f(v.x - m_star);
            l += w * v.y;
            ml_s[s * NPAD + tid].x = w;
          }
          const float inv = l > 0.f ? 1.f / l : 0.f;
          for (int s = 0; s < p.n_splits; ++s) ml_s[s * NPAD + tid].x *= inv;
        }
        named_bar_sync(1, kSoftmaxThreads);
        float acc[NH];
#pragma unroll
        for (int i = 0; i < NH; ++i) acc[i] = 0.f;
        for (int s0 = 0; s0 < p.n_splits; s0 += 4) {
          float v[NH][4];
#pragma unroll
          for (int i = 0; i < NH; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // NH*4 independent L2 loads in flight per thread
              const int s = s0 + u, c = c0 + i;
              v[i][u] = (s < p.n_splits && c < cols_valid)
                            ? __ldcg(p.part_o + (static_cast<long long>(s) * HQ + ptab[c]) * D + d) : 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < NH; ++i) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (s0 + u < p.n_splits) acc[i] = fmaf(ml_s[(s0 + u) * NPAD + c0 + i].x, v[i][u], acc[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < NH; ++i) {
          const int c = c0 + i;
          if (c < cols_valid) p.out[static_cast<long long>(otab[c]) * D + d] = __float2half_rn(acc[i]);
        }
      }
    }
  }

  if constexpr (CLUSTER) {
    // ---- split-KV merge through the leader's shared memory (rank 0 of the cluster = split 0 of this KV head) ----
    const int n_ranks = pin.n_splits;                       // launched cluster size; surplus ranks carry m = -inf
    const bool smx = warp < kSoftmaxWarps;
    const int quad = warp & 3, ch = (warp >> 2) & 1;
    const int kl = quad * 32 + lane, c0 = ch * NH, d = kl;
    // receive area in the K/V ring, free once a CTA's own MMAs are done: fp32 O^T [rank][column][d], (m, l) [rank][column]
    const uint32_t recv_o = smem_u32(sKV);
    const uint32_t recv_ml = recv_o + 8u * NPAD * D * 4u;
    cluster_sync_all();  // A: every CTA of the cluster has finished its tiles -> the leader's ring may be overwritten
    if (smx && split != 0) {
      const uint32_t ro = cluster_map_shared(recv_o + (static_cast<uint32_t>(split) * NPAD + c0) * D * 4u + d * 4u, 0);
#pragma unroll
      for (int i = 0; i < NH; ++i) st_cluster_f32(ro + i * D * 4u, o[i]);
#pragma unroll
      for (int i = 0; i < NH; ++i)
        if (kl == i) st_cluster_v2f32(cluster_map_shared(recv_ml + (static_cast<uint32_t>(split) * NPAD + c0 + i) * 8u, 0), m_col[i], l_col[i]);
    }
    if (smx && split == 0) {
      float2* ml_own = reinterpret_cast<float2*>(sKV + 8 * NPAD * D * 4);
#pragma unroll
      for (int i = 0; i < NH; ++i)
        if (kl == i) ml_own[c0 + i] = make_float2(m_col[i], l_col[i]);
    }
    cluster_sync_all();  // B: the partials have landed in the leader's shared memory
    if (tid == 0) PC_TRACE(3);
    if (smx && split == 0) {
      const float2* ml_s = reinterpret_cast<const float2*>(sKV + 8 * NPAD * D * 4);  // [rank][NPAD]
      const float* po = reinterpret_cast<const float*>(sKV);                           // [rank][NPAD][D]
      float* wtab = red;                                                               // [rank][NPAD], 8*NPAD floats
      if (tid < cols_valid) {  // weights w_s = 2^(m_s - m*) / sum_s 2^(m_s - m*) l_s  (same order as the global merge)
        float m_star = -INFINITY;
        for (int s = 0; s < n_ranks; ++s) m_star = fmaxf(m_star, ml_s[s * NPAD + tid].x);
        float l = 0.f;
        for (int s = 0; s < n_ranks; ++s) {
          const float2 v = ml_s[s * NPAD + tid];
          const float w = v.x == -INFINITY ? 0.f : exp2f(v.x - m_star);
          l += w * v.y;
          wtab[s * NPAD + tid] = w;
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        for (int s = 0; s < n_ranks; ++s) wtab[s * NPAD + tid] *= inv;
      }
      named_bar_sync(1, kSoftmaxThreads);
#pragma unroll
      for (int i = 0; i < NH; ++i) {
        const int c = c0 + i;
        float acc = 0.f;
        acc = fmaf(wtab[c], o[i], acc);
        for (int s = 1; s < n_ranks; ++s) acc = fmaf(wtab[s * NPAD + c], po[(s * NPAD + c) * D + d], acc);
        if (c < cols_valid) p.out[static_cast<long long>(otab[c]) * D + d] = __float2half_rn(acc);
      }
    }
  }

  if (tid == 0) PC_TRACE(4);  // softmax group done (incl. merge if last)
  if (tid == 0 && p.trace != nullptr) p.trace[64 + 2 * blockIdx.x + 1] = gtime();  // per-CTA end
  tc_fence_before();
  __syncthreads();
  if (warp == kSoftmaxWarps + 1) {
    tc_fence_after();
    tmem_free<C::kTmemCols>(tmem_base);
  }
}

// Can `clusters` clusters of `cluster_x` CTAs of this kernel be resident at once?  (A cluster lives inside one GPC; with
// one 211 KB CTA per SM that is a placement question the occupancy API answers.)  Cached per cluster size.
template <int NPAD, int ROUND, bool ALIBI>
bool clusters_fit(int cluster_x, int clusters) {
  static int max_active[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // 0 = not queried yet, -1 = query failed
  if (cluster_x < 2 || cluster_x > 8) return false;
  if (max_active[cluster_x] == 0) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    fill_launch_config(cfg, attr, dim3(cluster_x * 64), dim3(kThreads), SCfg<NPAD>::kSmem, nullptr, cluster_x);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, attn_small_kernel<NPAD, ROUND, ALIBI, true>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = -1;
    }
    max_active[cluster_x] = n > 0 ? n : -1;
  }
  return max_active[cluster_x] >= clusters;
}

template <int NPAD, int ROUND, bool ALIBI>
int launch_small_one(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mak,
                     const CUtensorMap& mav, const SmallParams& sp, int grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<NPAD, ROUND, ALIBI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SCfg<NPAD>::kSmem));
    PC_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<NPAD, ROUND, ALIBI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SCfg<NPAD>::kSmem));
    configured = true;
  }
  const bool want_cluster = sp.n_splits >= 2 && sp.n_splits <= 8 && !sp.global_merge;
  if (want_cluster && clusters_fit<NPAD, ROUND, ALIBI>(sp.n_splits, grid / sp.n_splits)) {
    PC_CHECK_CUDA(launch_pdl_cluster(attn_small_kernel<NPAD, ROUND, ALIBI, true>, sp.n_splits, dim3(grid), dim3(kThreads),
                                     SCfg<NPAD>::kSmem, st, mq, mk, mv, mak, mav, sp));
  } else {
    PC_CHECK_CUDA(launch_pdl(attn_small_kernel<NPAD, ROUND, ALIBI, false>, dim3(grid), dim3(kThreads), SCfg<NPAD>::kSmem, st, mq, mk,
                             mv, mak, mav, sp));
  }
  count_launch();
  return PC_OK;
}

template <int NPAD>
int launch_small_n(const pc_attn_params& p, cudaStream_t st) {
  const int G = p.n_heads / p.n_kv_heads;
  const bool virt = p.kv_tiles != nullptr && p.kv_n_tiles > 0;
  if (virt) {
    PC_REQUIRE(p.kv_arena_k && p.kv_arena_v && p.kv_arena_rows > 0 && p.kv_rows > 0 && p.kv_rows <= p.n_cached,
               "attn(tcgen05-small): bad virtual-splice arguments (arena rows %d, kv_rows %d, n_cached %d)", p.kv_arena_rows,
               p.kv_rows, p.n_cached);
    PC_REQUIRE((reinterpret_cast<uintptr_t>(p.kv_arena_k) | reinterpret_cast<uintptr_t>(p.kv_arena_v)) % 16 == 0,
               "attn(tcgen05-small): arena planes must be 16B aligned");
  }
  const int v_rows = virt ? p.kv_rows : 0, n_vtiles = virt ? p.kv_n_tiles : 0;
  const int tiles = n_vtiles + (p.n_cached - v_rows + p.q_len + BK - 1) / BK;  // with n_cached_dev: an upper bound
  int n_splits = sm_count() / p.n_kv_heads;
  if (const char* e = dbg_env("PC_ATTN_SPLITS")) n_splits = atoi(e);
  if (n_splits > tiles) n_splits = tiles;
  if (n_splits > 64) n_splits = 64;
  if (n_splits < 1) n_splits = 1;
  // workspace: [counters: 4096 B][part_o][part_ml]
  const size_t ctr_bytes = kAttnCounterBytes;
  if (static_cast<size_t>(p.n_kv_heads) * 4 > ctr_bytes)
    return set_error(PC_ERR_UNSUPPORTED, "attn(tcgen05-small): %d KV heads exceed the counter region", p.n_kv_heads);
  const size_t per_split = static_cast<size_t>(p.n_heads) * p.q_len * (D + 2) * sizeof(float);
  if (n_splits > 1) {
    const size_t avail = (p.workspace && p.workspace_bytes > ctr_bytes) ? p.workspace_bytes - ctr_bytes : 0;
    if (avail < per_split * n_splits) n_splits = static_cast<int>(avail / per_split);
    if (n_splits < 1) n_splits = 1;
  }
  SmallParams sp;
  sp.out = static_cast<__half*>(p.out);
  sp.counters = nullptr;
  sp.part_o = nullptr;
  sp.part_ml = nullptr;
  if (n_splits > 1) {
    sp.counters = static_cast<int*>(p.workspace);
    sp.part_o = reinterpret_cast<float*>(static_cast<char*>(p.workspace) + ctr_bytes);
    sp.part_ml = sp.part_o + static_cast<size_t>(n_splits) * p.n_heads * p.q_len * D;
  }
  sp.alibi_slopes = p.alibi_slopes;
  sp.key_pos = p.key_pos;
  sp.n_cached_dev = p.n_cached_dev;
  sp.n_heads = p.n_heads;
  sp.n_kv_heads = p.n_kv_heads;
  sp.n_cached = p.n_cached;
  sp.q_len = p.q_len;
  sp.n_splits = n_splits;
  sp.alibi_max_pos = p.alibi_max_pos;
  sp.scale = p.scale;
  sp.trace = nullptr;
  sp.global_merge = (p.impl & PC_ATTN_GLOBAL_MERGE) != 0;
  sp.pf = (p.prefetch != nullptr && p.prefetch_bytes >= 16 && reinterpret_cast<uintptr_t>(p.prefetch) % 16 == 0)
              ? static_cast<const unsigned char*>(p.prefetch) : nullptr;
  sp.pf_bytes = p.prefetch_bytes;
  sp.vtiles = virt ? reinterpret_cast<const int2*>(p.kv_tiles) : nullptr;
  sp.n_vtiles = n_vtiles;
  sp.v_rows = v_rows;
  if (const char* e = dbg_env("PC_ATTN_TRACE_PTR")) sp.trace = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));

  const int rows_total = G * p.q_len;
  CUtensorMap mq, mk, mv;
  if (int rc = make_tensor_map_3d(&mq, p.q, D, rows_total, p.n_kv_heads, static_cast<long long>(rows_total) * D, NPAD)) return rc;
  // run-time N: the maps cover the whole buffer (rows past N+q are masked; they must hold finite values)
  const long long kv_rows = p.n_cached_dev ? static_cast<long long>(p.max_ctx) : static_cast<long long>(p.n_cached) + p.q_len;
  if (int rc = make_tensor_map_3d(&mk, p.k, D, kv_rows, p.n_kv_heads, static_cast<long long>(p.max_ctx) * D, BK)) return rc;
  if (int rc = make_tensor_map_3d(&mv, p.v, D, kv_rows, p.n_kv_heads, static_cast<long long>(p.max_ctx) * D, BK)) return rc;
  CUtensorMap mak = mk, mav = mv;  // unused without a tile list
  if (virt) {
    const long long ar = p.kv_arena_rows;
    if (int rc = make_tensor_map_3d(&mak, p.kv_arena_k, D, ar, p.n_kv_heads, ar * D, BK)) return rc;
    if (int rc = make_tensor_map_3d(&mav, p.kv_arena_v, D, ar, p.n_kv_heads, ar * D, BK)) return rc;
  }
  const int grid = p.n_kv_heads * n_splits;
  const bool alibi = p.alibi_slopes != nullptr;
  if (p.rounding == 0)
    return alibi ? launch_small_one<NPAD, 0, true>(mq, mk, mv, mak, mav, sp, grid, st)
                 : launch_small_one<NPAD, 0, false>(mq, mk, mv, mak, mav, sp, grid, st);
  return alibi ? launch_small_one<NPAD, 3, true>(mq, mk, mv, mak, mav, sp, grid, st)
               : launch_small_one<NPAD, 3, false>(mq, mk, mv, mak, mav, sp, grid, st);
}

}  // namespace

bool attn_small_supported(const pc_attn_params& p) {
  const int G = p.n_heads / p.n_kv_heads;
  return p.head_dim == 128 && G * p.q_len <= 32 && (p.rounding == 0 || p.rounding == PC_ROUND_LLAMA);
}

// up to 64 splits of fp32 partials (the counter region is added by the caller)
size_t attn_small_partial_bytes(int n_heads, int q_len, int head_dim) {
  if (head_dim != 128 || q_len > 32) return 0;
  long long splits = sm_count();  // worst case: one KV head
  if (splits > 64) splits = 64;
  return static_cast<size_t>(splits) * n_heads * q_len * (head_dim + 2) * sizeof(float);
}

int launch_attn_small(const pc_attn_params& p, cudaStream_t st) {
  const int G = p.n_heads / p.n_kv_heads;
  if (G * p.q_len <= 16) return launch_small_n<16>(p, st);
  return launch_small_n<32>(p, st);
}

}  // namespace pc
