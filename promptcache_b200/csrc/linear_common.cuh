// Shared pieces of the weight-streaming linear kernels (linear_small.cu: one linear per launch; linear_chain.cu: a chain of
// linears and norms in ONE persistent launch): tile constants, per-CTA work iterators and the epilogues.
#pragma once
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "pc_common.cuh"

namespace pc {

int make_tensor_map_3d(CUtensorMap* map, const void* base, int D, long long rows, int heads, long long head_stride,
                       int box_rows);

namespace lin {

constexpr int BM = 128;                 // out-features per tile (MMA M)
constexpr int BKK = 64;                 // k per stage: one 128-byte swizzled row
constexpr int kWBytes = BM * BKK * 2;   // 16 KiB
constexpr int kEpiThreads = 128;
constexpr int kThreads = 64 + kEpiThreads + 32;  // W-TMA warp + MMA warp + 4 epilogue warps + X-TMA warp
constexpr size_t kLinCounterBytes = 16384;  // 4096 tile arrival counters
constexpr int kMaxTiles = 4096;
constexpr size_t kChainBarBytes = 256;     // grid-barrier counters of pc_linear_chain, behind the stream-K slots

// KPS = k-blocks per pipeline stage: 1 with tensor-map loads of the row-major weight, 2 with pre-tiled weights (one
// 32 KiB bulk copy per stage).
template <int NPAD, int KPS>
struct LCfg {
  static constexpr int kXBytes = NPAD * BKK * 2;
  static constexpr int kStageBytes = KPS * (kWBytes + kXBytes);
  // bytes in flight per SM bound its streaming rate (Little's law, ~2.5 us loaded latency): as many stages as fit
#ifndef PC_LIN_STAGES16
#define PC_LIN_STAGES16 6
#endif
  static constexpr int kStages = KPS == 2 ? (NPAD <= 16 ? PC_LIN_STAGES16 : 5) : (NPAD <= 16 ? 11 : 10);
  static constexpr int kOffExch = kStages * kStageBytes;
  static constexpr int kExchBytes = NPAD * BM * 2;
  static constexpr int kOffBars = kOffExch + kExchBytes;
  // 1 KiB of slack for the 1024-byte alignment of the swizzled tiles; the half-size ring (two CTAs per SM) has 768 B --
  // the kernel traps if the dynamic shared memory base ever needs more
  static constexpr int kAlignSlack = (KPS == 2 && NPAD <= 16 && PC_LIN_STAGES16 <= 3) ? 768 : 1024;
  static constexpr int kSmem = kAlignSlack + kOffBars + 256;
  static constexpr int kMinBlocks = (KPS == 2 && NPAD <= 16 && PC_LIN_STAGES16 <= 3) ? 2 : 1;
  static constexpr int kTmemCols = 2 * NPAD <= 32 ? 32 : 64;
};

// Table-driven (SM-weighted) stream-K boundaries: experiment only (tools/sm_weights_probe.py), compiled in with -DPC_LIN_BOUNDS.
#ifdef PC_LIN_BOUNDS
#define PC_LIN_HAS_BOUNDS(p) ((p).bounds != nullptr)
#else
#define PC_LIN_HAS_BOUNDS(p) false
#endif

struct LinParams {
  __half* out;
  long long out_stride;
  float* slots;
  int* counters;
  int n_out;             // valid output columns (PLAIN: weight rows; SILU: inter)
  int q_len, kb_total, tile_row_stride, second_half_off;
  int w_one_box;         // the weight tile is 128 consecutive rows: one TMA box instead of two
  const unsigned char* w_tiled;  // TILED: weights re-laid out as [tile][k-block] 16 KiB shared-memory images
  int grid;              // CTAs this linear is partitioned over (== gridDim.x of a single-linear launch)
  int cl_helper_units;   // > 0: clusters of cl_size CTAs = cl_size - 1 tile owners + 1 helper that streams the last
  int cl_size;           // cl_helper_units units of each owner's K range and hands the fp32 partials over through
                         // distributed shared memory (cl_size 3 or 4)
  // optional: unit boundaries of the stream-K region per partition index, [grid + 1] entries (bounds[0] = 0, bounds[grid] =
  // units_total) instead of the uniform split -- a partition weighted by how fast each SM streams (see sm_weights)
  const long long* bounds;
  int whole_per_cta;     // tiles [0, whole_per_cta * grid) are processed whole, the rest is the stream-K region
  int splits;            // PARTIALS: regular split-K, CTA = (tile, split); fp32 partials go to part_out[split][t][col]
  float* part_out;
  long long units_total;
  // RoPE + append epilogue
  const long long* pos;
  const __half* cos;
  const __half* sin;
  int table_rows;
  __half* q_out;
  __half* k_cache;
  __half* v_cache;
  int max_ctx, row0, n_heads, n_kv_heads;
  int head_dim;          // ROPE_APPEND: 128 (one head per tile) or 64 (two heads per tile: Falcon)
  const int* row0_dev;
  unsigned long long* trace;  // debug: 8 %globaltimer stamps per CTA (PC_LIN_TRACE_PTR), else nullptr
  // L2 prefetch of the NEXT linear's weight stream (pre-tiled): once this CTA has issued its last weight load it asks
  // L2 for the first pf_units pipeline stages of the CTAs of the next launch (their partition is replayed here), so
  // that the next kernel's ring fills from L2 instead of paying the HBM latency after every launch boundary
  int act;               // PLAIN: 1 = exact GELU applied to the fp16-rounded output (mpt.py:200-208 / falcon MLP)
  const unsigned char* pf_w;
  long long pf_units_total;
  int pf_kb, pf_grid, pf_whole, pf_splits, pf_units;
};

// units [u0, u1) of CTA c (same arithmetic as the kernel prologue; `splits` > 0 = PC_LIN_PARTIALS)
__device__ __forceinline__ void cta_units(int c, int G, long long U, int KB, int splits, long long& u0, long long& u1) {
  if (splits > 0) {
    const int tile_c = c / splits, sp = c - tile_c * splits;
    u0 = static_cast<long long>(tile_c) * KB + (static_cast<long long>(sp) * KB) / splits;
    u1 = static_cast<long long>(tile_c) * KB + (static_cast<long long>(sp + 1) * KB) / splits;
    return;
  }
  const int Gs = static_cast<int>(min(static_cast<long long>(G), U));
  u0 = c < Gs ? (static_cast<long long>(c) * U) / Gs : 0;
  u1 = c < Gs ? (static_cast<long long>(c + 1) * U) / Gs : 0;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef PC_DEBUG
#define LIN_TRACE(slot)                                                 \
  do {                                                                  \
    if (p.trace != nullptr) p.trace[blockIdx.x * 8 + (slot)] = gtime(); \
  } while (0)
#else
#define LIN_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// CTA whose unit range [c*U/G, (c+1)*U/G) holds unit u.
__device__ __forceinline__ int cta_of_unit(long long u, long long U, int G) {
  int c = static_cast<int>(((u + 1) * G + U - 1) / U) - 1;
  if (c < 0) c = 0;
  if (c > G - 1) c = G - 1;
  while (c + 1 < G && (static_cast<long long>(c + 1) * U) / G <= u) ++c;
  while (c > 0 && (static_cast<long long>(c) * U) / G > u) --c;
  return c;
}

template <int N>
__device__ __forceinline__ void tmem_ld_acc(uint32_t addr, float (&v)[N]) {
  if constexpr (N == 16) tmem_ld16(addr, reinterpret_cast<uint32_t(&)[16]>(v));
  else tmem_ld32(addr, reinterpret_cast<uint32_t(&)[32]>(v));
}

// One CTA's share of the (tile, k-block) space, cut at tile boundaries.  Segments that cover only part of a
// tile's K ("partial": at most the first and the last of the range) are visited FIRST, whole tiles last: the
// cross-CTA reduction of a partial (publish -> arrival counter -> slot reads) then overlaps the streaming of the
// whole tiles instead of sitting in the kernel's tail.
struct Seg {
  int tile, kb0, n;
  int rel;  // tile index inside the stream-K region (what the split bookkeeping counts in), -1 for a whole tile of phase 2
};
struct SegIter {
  long long u0, u1;
  int KB, t_first, t_last, phase, cur;
  bool first_partial, last_partial;
  __device__ SegIter(long long u0_, long long u1_, int KB_) : u0(u0_), u1(u1_), KB(KB_) {
    if (u0 >= u1) {  // empty range
      t_first = t_last = 0;
      first_partial = last_partial = false;
      phase = 2;
      cur = 1;
      return;
    }
    t_first = static_cast<int>(u0 / KB);
    t_last = static_cast<int>((u1 - 1) / KB);
    const long long first_end = min(u1, static_cast<long long>(t_first + 1) * KB);
    first_partial = (u0 != static_cast<long long>(t_first) * KB) || (first_end - u0 != KB);
    last_partial = (t_last > t_first) && (u1 != static_cast<long long>(t_last + 1) * KB);
    phase = 0;
    cur = first_partial ? t_first + 1 : t_first;
  }
  __device__ bool next(Seg& s) {
    if (phase == 0) {
      phase = 1;
      if (last_partial) {
        s.tile = t_last;
        s.kb0 = 0;
        s.n = static_cast<int>(u1 - static_cast<long long>(t_last) * KB);
        return true;
      }
    }
    if (phase == 1) {
      phase = 2;
      if (first_partial) {
        s.tile = t_first;
        s.kb0 = static_cast<int>(u0 - static_cast<long long>(t_first) * KB);
        s.n = static_cast<int>(min(u1, static_cast<long long>(t_first + 1) * KB) - u0);
        return true;
      }
    }
    const int end = last_partial ? t_last - 1 : t_last;
    if (cur <= end) {
      s.tile = cur++;
      s.kb0 = 0;
      s.n = KB;
      return true;
    }
    return false;
  }
};

// A CTA's whole work list.  With more tiles than CTAs (gate/up, lm_head) the tiles [0, W*G) are dealt out WHOLE, W per
// CTA, and only the remaining n_tiles - W*G tiles are shared stream-K fashion -- and those come first, so that their
// cross-CTA reduction overlaps the streaming of the whole tiles and the kernel ends with a plain epilogue.
struct WorkIter {
  SegIter it;
  int W, G, c, KB, rem_tile0, r;
  __device__ WorkIter(long long u0, long long u1, int KB_, int W_, int G_, int c_) : it(u0, u1, KB_), W(W_), G(G_), c(c_), KB(KB_) {
    rem_tile0 = W * G;
    r = 0;
  }
  __device__ bool next(Seg& s) {
    if (it.next(s)) {
      s.rel = s.tile;
      s.tile += rem_tile0;
      return true;
    }
    if (r < W) {
      s.tile = c + r * G;
      s.rel = -1;
      s.kb0 = 0;
      s.n = KB;
      ++r;
      return true;
    }
    return false;
  }
};


// CTA whose range of the table-driven partition holds unit u: the last i with bounds[i] <= u.
__device__ __forceinline__ int cta_of_unit_bounds(const long long* b, int G, long long u) {
  int lo = 0, hi = G - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(b + mid) <= u) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// This CTA's unit range [u0, u1) of the linear `p` (stream-K region or regular split-K; empty when c >= p.grid).
__device__ __forceinline__ void lin_cta_range(const LinParams& p, int epi, int c, int KB, long long& u0, long long& u1) {
  if (c >= p.grid) {
    u0 = u1 = 0;
    return;
  }
  if (epi == PC_LIN_PARTIALS) {  // regular split-K: every tile is cut into `splits` equal K ranges
    const int tile_c = c / p.splits, sp = c - tile_c * p.splits;
    u0 = static_cast<long long>(tile_c) * KB + (static_cast<long long>(sp) * KB) / p.splits;
    u1 = static_cast<long long>(tile_c) * KB + (static_cast<long long>(sp + 1) * KB) / p.splits;
    return;
  }
  if (PC_LIN_HAS_BOUNDS(p)) {
    u0 = __ldg(p.bounds + c);
    u1 = __ldg(p.bounds + c + 1);
    return;
  }
  // the stream-K region is shared by the first Gs CTAs (all of them unless it holds fewer units than there are CTAs)
  const long long U = p.units_total;
  const int Gs = static_cast<int>(min(static_cast<long long>(p.grid), U));
  u0 = c < Gs ? (static_cast<long long>(c) * U) / Gs : 0;
  u1 = c < Gs ? (static_cast<long long>(c + 1) * U) / Gs : 0;
}

// Epilogue of ONE accumulator segment `sg` of linear `p`, run by the 128 epilogue threads (f = row of the tile = TMEM lane,
// et = thread index in the group, c = this CTA's index in the linear's partition); v = the fp32 accumulator row.
//   exch: NPAD * BM halfs of shared memory, flag: 3 ints of shared memory; named barrier 1 is used.
// cos / sin of every token's position id at column d of the RoPE table (clamped ids, as pc_rope_append): independent of the
// accumulators, so a caller may issue these gathers while the MMAs are still running and hand the values to lin_epilogue
template <int NPAD>
__device__ __forceinline__ void rope_gather(const LinParams& p, int d, __half (&cs)[NPAD], __half (&sn)[NPAD]) {
  const int HD = p.head_dim;
#pragma unroll
  for (int t = 0; t < NPAD; ++t) {
    long long pos = t < p.q_len ? __ldg(p.pos + t) : 0;
    if (pos < 0) pos = 0;
    if (pos >= p.table_rows) pos = p.table_rows - 1;
    cs[t] = __ldg(p.cos + pos * HD + d);
    sn[t] = __ldg(p.sin + pos * HD + d);
  }
}

template <int NPAD, int EPI>
__device__ __forceinline__ void lin_epilogue(const LinParams& p, const Seg& sg, float (&v)[NPAD], int c, int KB, int t_first,
                                             __half* exch, int* flag, int f, int et, int row0, const __half* cs_pre = nullptr,
                                             const __half* sn_pre = nullptr) {
  const int tile = sg.tile;
  const long long U = p.units_total;
  const int Gs = static_cast<int>(min(static_cast<long long>(p.grid), U));
  bool finalize = true;
  if constexpr (EPI == PC_LIN_PARTIALS) {
    // no cross-CTA reduction here: the consumer (pc_add_rmsnorm_partials) adds the splits in a fixed order
    finalize = false;
    const int col = tile * BM + f;
    float* po = p.part_out + static_cast<long long>(c - (c / p.splits) * p.splits) * p.q_len * p.n_out + col;
    if (col < p.n_out) {
#pragma unroll
      for (int t = 0; t < NPAD; ++t)
        if (t < p.q_len) po[static_cast<long long>(t) * p.n_out] = v[t];
    }
  } else if (sg.n != KB) {
    // partial K range: park the fp32 partial; the last CTA to arrive at this tile adds them up
    const int rel = sg.rel;  // tile index inside the stream-K region
    const int slot = (rel == t_first) ? 0 : 1;
    float* sp = p.slots + (static_cast<size_t>(c) * 2 + slot) * NPAD * BM;
#pragma unroll
    for (int t = 0; t < NPAD; ++t)
      if (t < p.q_len) __stcg(sp + t * BM + f, v[t]);
    named_bar_sync(1, kEpiThreads);  // orders every thread's stores before thread 0's gpu-scope fence + arrival
    if (et == 0) {
      const long long tu0 = static_cast<long long>(rel) * KB;
      const int c_first = PC_LIN_HAS_BOUNDS(p) ? cta_of_unit_bounds(p.bounds, p.grid, tu0) : cta_of_unit(tu0, U, Gs);
      const int c_last = PC_LIN_HAS_BOUNDS(p) ? cta_of_unit_bounds(p.bounds, p.grid, tu0 + KB - 1) : cta_of_unit(tu0 + KB - 1, U, Gs);
      // one acq_rel RMW instead of fence.sc + atomic + fence.sc: the release half publishes the parked partial (ordered
      // before it by the bar.sync above), the acquire half makes the other contributors' partials visible to the reduction
      int old;
#ifdef PC_LIN_SC_FENCES
      __threadfence();
      old = atomicAdd(p.counters + tile, 1);
      __threadfence();
#else
      asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(p.counters + tile) : "memory");
#endif
      const int last = (old == c_last - c_first) ? 1 : 0;
      if (last) p.counters[tile] = 0;  // self-cleaning for the next launch (ordered by the kernel boundary)
      flag[0] = last;
      flag[1] = c_first;
      flag[2] = c_last;
    }
    named_bar_sync(1, kEpiThreads);
    finalize = flag[0] != 0;
    if (finalize) {
      const int c_first = flag[1], c_last = flag[2];
#pragma unroll
      for (int t = 0; t < NPAD; ++t) v[t] = 0.f;
      constexpr int CH = NPAD <= 16 ? 4 : 2;  // contributors whose loads are in flight together
      for (int cc = c_first; cc <= c_last; cc += CH) {
        float tmp[CH][NPAD];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int ccj = min(cc + j, c_last);
          const long long cu0 = PC_LIN_HAS_BOUNDS(p) ? __ldg(p.bounds + ccj) : (static_cast<long long>(ccj) * U) / Gs;
          const int slot_c = (cu0 / KB == sg.rel) ? 0 : 1;
          const float* rp = p.slots + (static_cast<size_t>(ccj) * 2 + slot_c) * NPAD * BM;
#pragma unroll
          for (int t = 0; t < NPAD; ++t) tmp[j][t] = (cc + j <= c_last && t < p.q_len) ? __ldcg(rp + t * BM + f) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {  // fixed order: deterministic sums
#pragma unroll
          for (int t = 0; t < NPAD; ++t) v[t] += tmp[j][t];
        }
      }
    }
  }
  if (!finalize) return;
  if constexpr (EPI == PC_LIN_PLAIN || EPI == PC_LIN_PARTIALS) {
    const int col = tile * BM + f;
    if (col < p.n_out) {
      if (EPI == PC_LIN_PLAIN && p.act == 1) {
        // nn.GELU() on the fp16 output of the projection: 0.5 x (1 + erf(x / sqrt 2)) in fp32, rounded once
#pragma unroll
        for (int t = 0; t < NPAD; ++t) {
          const float xg = __half2float(__float2half_rn(v[t]));
          v[t] = 0.5f * xg * (1.f + erff(xg * 0.70710678118654752440f));
        }
      }
#pragma unroll
      for (int t = 0; t < NPAD; ++t)
        if (t < p.q_len) p.out[t * p.out_stride + col] = __float2half_rn(v[t]);
    }
  } else if constexpr (EPI == PC_LIN_SILU_MUL) {
    // rows [0,64) of the tile are gate features, rows [64,128) the up features of the same columns.  The two threads of a
    // feature split the tokens: the gate thread takes tokens [0, NPAD/2) and needs the up values of those, the up thread takes
    // [NPAD/2, NPAD) and needs the gate values (all 128 threads share the expf / division work)
    constexpr int TH = NPAD / 2;
    const int ff = f & 63;
    if (f >= 64) {
#pragma unroll
      for (int t = 0; t < TH; ++t) exch[t * 64 + ff] = __float2half_rn(v[t]);            // up, first half of the tokens
    } else {
#pragma unroll
      for (int t = TH; t < NPAD; ++t) exch[t * 64 + ff] = __float2half_rn(v[t]);         // gate, second half
    }
    named_bar_sync(1, kEpiThreads);
    const int col = tile * 64 + ff;
    if (col < p.n_out) {
      const int t_lo = f < 64 ? 0 : TH;
#pragma unroll
      for (int tt = 0; tt < TH; ++tt) {
        const int t = t_lo + tt;
        if (t < p.q_len) {
          const __half other = exch[t * 64 + ff];
          const float g = __half2float(f < 64 ? __float2half_rn(v[tt]) : other);        // gate_proj output is fp16
          const __half u = f < 64 ? other : __float2half_rn(v[TH + tt]);
          const float s = g / (1.f + expf(-g));
          p.out[t * p.out_stride + col] = __hmul_rn(__float2half_rn(s), u);
        }
      }
    }
    named_bar_sync(1, kEpiThreads);
  } else {
    // a tile holds 128 / D heads: [0,H) query heads, [H,H+Hkv) key heads, then value heads (rows past them: padding)
#pragma unroll
    for (int t = 0; t < NPAD; ++t) exch[t * BM + f] = __float2half_rn(v[t]);
    named_bar_sync(1, kEpiThreads);
    const int H = p.n_heads, Hkv = p.n_kv_heads, HD = p.head_dim;
    const int head = (tile * BM + f) / HD, d = f & (HD - 1);
    const bool rotate = (head < H + Hkv) && p.cos != nullptr;
    __half* dst;
    if (head < H) dst = p.q_out + static_cast<long long>(head) * p.q_len * HD + d;
    else if (head < H + Hkv) dst = p.k_cache + (static_cast<long long>(head - H) * p.max_ctx + row0) * HD + d;
    else dst = p.v_cache + (static_cast<long long>(head - H - Hkv) * p.max_ctx + row0) * HD + d;
    if (head < H + 2 * Hkv) {
      if (rotate) {
        // all table gathers first (independent loads in flight together), then the arithmetic
        __half cs[NPAD], sn[NPAD];
        if (cs_pre != nullptr) {
#pragma unroll
          for (int t = 0; t < NPAD; ++t) {
            cs[t] = cs_pre[t];
            sn[t] = sn_pre[t];
          }
        } else {
          rope_gather<NPAD>(p, d, cs, sn);
        }
        const int half_d = HD >> 1;
#pragma unroll
        for (int t = 0; t < NPAD; ++t) {
          if (t < p.q_len) {
            const __half x = exch[t * BM + f];
            const __half partner = exch[t * BM + (f ^ half_d)];
            const __half rot = d < half_d ? __hneg(partner) : partner;  // rotate_half(x) = cat(-x2, x1)
            dst[t * HD] = __hadd_rn(__hmul_rn(x, cs[t]), __hmul_rn(rot, sn[t]));
          }
        }
      } else {
#pragma unroll
        for (int t = 0; t < NPAD; ++t)
          if (t < p.q_len) dst[t * HD] = exch[t * BM + f];
      }
    }
    named_bar_sync(1, kEpiThreads);
  }
}

// Host side (linear_small.cu): validation of one pc_linear_params + the partition of its (tile, k-block) space.
struct LinLaunch {
  LinParams lp;
  CUtensorMap map_w, map_x;
  int grid, npad, epilogue;
  bool tiled;
};
// allow_cluster: the single-launch kernel may pick the owner + helper cluster partition (lp.cl_helper_units > 0)
int prepare_linear(const pc_linear_params* p, LinLaunch* out, bool allow_cluster = false);
int tiled_k_blocks(int k);

}  // namespace lin
}  // namespace pc
