// A CHAIN of the step's skinny linears and the norms between them as ONE persistent launch (SURVEY.md section 8 row a13;
// reference promptcache/model/llama2.py:623-648 decoder layer: o_proj -> residual add + post_attention_layernorm -> gate/up
// -> SiLU*up -> down_proj -> residual add + next input_layernorm -> q/k/v projections + RoPE + KV append).
//
// Why: launched one by one (linear_small.cu) every linear pays, per SM, the dead time between its last weight load and
// the first load of the next kernel's CTA (epilogue, exit, launch, barrier init, ~2.5 us of HBM latency): ~18 us of a
// ~100 us layer although each kernel alone streams at the HBM rate.  Here one CTA per SM walks the whole op list:
//   * the weight producer thread never stops: weights are static, so it streams op k+1's tiles into the ring while op k's
//     epilogue, the grid barrier and the norm are still running (up to the ring's 192 KiB per SM ahead);
//   * the X producer waits for the op's input (a grid barrier: every CTA's epilogue group arrives at counter[op] after its
//     last store of that op) and then loads the activation tiles with TMA;
//   * norms are ops of their own: row r is handled by the epilogue group of CTA G-1-r (the CTAs the 128-CTA o/down
//     partitions leave idle), between two grid barriers;
//   * MMA issue, accumulator double-buffering, per-linear work partitions and epilogues are those of linear_small.cu
//     (linear_common.cuh), the epilogue kind is a run-time switch per op.
// All G <= #SM CTAs are co-resident (1 CTA per SM, launch_dependents at the top), so the grid barriers cannot deadlock.
#include "linear_common.cuh"

namespace pc {
namespace lin {
namespace {

constexpr int kMaxOps = PC_CHAIN_MAX_OPS;

struct NormParams {
  __half* x;
  const float* part1;
  const float* part2;
  const __half* weight;
  const __half* bias;
  __half* h_out;
  int n1, n2, rows, hidden;
  float eps;
};

struct ChainOp {
  alignas(64) CUtensorMap map_x;
  union {
    LinParams lin;    // kind == PC_CHAIN_LINEAR
    NormParams norm;  // the norms
  };
  int kind;  // PC_CHAIN_*
  int epi;   // PC_LIN_* of a linear op
  int dep;   // op whose grid barrier must be complete before this op reads its input; -1: only the kernels before the chain
};

struct alignas(64) ChainParams {
  ChainOp ops[kMaxOps];
  int n_ops;
  int flags;  // experiments (PC_DEBUG builds, PC_CHAIN_FLAGS): 1 no proxy fences, 2 fence.sc arrival, 4 no norm math, 8 sleepy poll, 16 no waits
  int* bars;  // [kMaxOps] arrival counters, [kMaxOps] = exit counter; all zero between launches
};

// Grid barrier wait.  The spin is bounded (~seconds): a protocol error must surface as a launch failure, not as a hung GPU.
__device__ __forceinline__ void chain_wait(const int* bar, int target, int flags) {
  if (flags & 16) return;
  int v;
  unsigned spins = 0;
  do {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    if (++spins > (1u << 26)) __trap();
    if ((flags & 8) && v < target) __nanosleep(64);
  } while (v < target);
}
// Arrival at a grid barrier: a release at gpu scope (cumulative over the stores the bar.sync before it has ordered).
__device__ __forceinline__ void chain_arrive(int* bar, int flags) {
  if (flags & 2) {
    __threadfence();
    atomicAdd(bar, 1);
  } else {
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(bar) : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// x[row] = half(x + half(sum of partials1) (+ half(sum of partials2))), then RMSNorm / LayerNorm of the row into h_out:
// the arithmetic of add_rmsnorm_kernel / add_layernorm_kernel (support.cu) on 128 threads.  `red`: 8 floats of shared memory.
__device__ __forceinline__ void chain_norm_row(const NormParams& np, int kind, int row, int et, float* red) {
  const int hidden = np.hidden;
  const long long off = static_cast<long long>(row) * hidden;
  const long long split_stride = static_cast<long long>(np.rows) * hidden;
  __half* xr = np.x + off;
  __half* hr = np.h_out + off;
  float s1 = 0.f;
#pragma unroll 2
  for (int i = et * 8; i < hidden; i += kEpiThreads * 8) {
    uint4 xv = __ldcg(reinterpret_cast<const uint4*>(xr + i));
    __half2* xh = reinterpret_cast<__half2*>(&xv);
    if (np.part1) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float* pr = np.part1 + off + i;
      for (int s0 = 0; s0 < np.n1; s0 += 4) {  // four splits' loads in flight together, added in split order
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = s0 + u < np.n1;
          const float4* q = reinterpret_cast<const float4*>(pr + (ok ? s0 + u : s0) * split_stride);
          a[u] = ok ? __ldcg(q) : make_float4(0.f, 0.f, 0.f, 0.f);
          b[u] = ok ? __ldcg(q + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (s0 + u < np.n1) {
            acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
            acc[4] += b[u].x; acc[5] += b[u].y; acc[6] += b[u].z; acc[7] += b[u].w;
          }
        }
      }
      __half2 dsum[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) dsum[k] = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
      if (np.part2) {
        float acc2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* pr2 = np.part2 + off + i;
        for (int sp = 0; sp < np.n2; ++sp) {
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
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(xh[k]);
      if (kind == PC_CHAIN_ADD_RMSNORM) {
        s1 = fmaf(f.x, f.x, s1);
        s1 = fmaf(f.y, f.y, s1);
      } else {
        s1 += f.x + f.y;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  if ((et & 31) == 0) red[et >> 5] = s1;
  named_bar_sync(1, kEpiThreads);
  const float t1 = red[0] + red[1] + red[2] + red[3];
  if (kind == PC_CHAIN_ADD_RMSNORM) {
    const float inv = rsqrtf(t1 / static_cast<float>(hidden) + np.eps);
#pragma unroll 2
    for (int i = et * 8; i < hidden; i += kEpiThreads * 8) {
      uint4 xv = *reinterpret_cast<const uint4*>(xr + i);  // this thread's own stores (or an unchanged row)
      const uint4 wv = __ldg(reinterpret_cast<const uint4*>(np.weight + i));
      __half2* xh = reinterpret_cast<__half2*>(&xv);
      const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(xh[k]);
        xh[k] = __hmul2_rn(wh[k], __floats2half2_rn(f.x * inv, f.y * inv));
      }
      *reinterpret_cast<uint4*>(hr + i) = xv;
    }
    named_bar_sync(1, kEpiThreads);  // `red` is reused by the next row
    return;
  }
  const float mean = t1 / static_cast<float>(hidden);
  float s2 = 0.f;
#pragma unroll 2
  for (int i = et * 8; i < hidden; i += kEpiThreads * 8) {
    const uint4 xv = *reinterpret_cast<const uint4*>(xr + i);
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(xh[k]);
      s2 = fmaf(f.x - mean, f.x - mean, s2);
      s2 = fmaf(f.y - mean, f.y - mean, s2);
    }
  }
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if ((et & 31) == 0) red[4 + (et >> 5)] = s2;
  named_bar_sync(1, kEpiThreads);
  const float var = (red[4] + red[5] + red[6] + red[7]) / static_cast<float>(hidden);
  const float rstd = rsqrtf(var + np.eps);
#pragma unroll 2
  for (int i = et * 8; i < hidden; i += kEpiThreads * 8) {
    uint4 xv = *reinterpret_cast<const uint4*>(xr + i);
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(np.weight + i));
    uint4 bv = make_uint4(0u, 0u, 0u, 0u);
    if (np.bias) bv = __ldg(reinterpret_cast<const uint4*>(np.bias + i));
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
  named_bar_sync(1, kEpiThreads);
}

// This CTA's weight stream over the whole chain: the address of every 32 KiB pipeline stage, in consumption order.
struct StageStream {
  const ChainParams* cp;
  int c, o, j, KB;
  WorkIter it;
  Seg sg;
  const unsigned char* w;
  __device__ StageStream(const ChainParams* cp_, int c_) : cp(cp_), c(c_), o(-1), j(0), KB(1), it(0, 0, 1, 0, 1, 0), w(nullptr) { sg.n = 0; }
  __device__ bool next(const unsigned char*& addr) {
    while (true) {
      if (j < sg.n) {
        addr = w + (static_cast<long long>(sg.tile) * KB + sg.kb0 + j) * (2 * kWBytes);
        ++j;
        return true;
      }
      j = 0;
      sg.n = 0;
      if (it.next(sg)) continue;
      do {
        ++o;
      } while (o < cp->n_ops && (cp->ops[o].kind != PC_CHAIN_LINEAR || c >= cp->ops[o].lin.grid));
      if (o >= cp->n_ops) return false;
      const LinParams& p = cp->ops[o].lin;
      KB = p.kb_total / 2;
      w = p.w_tiled;
      long long u0, u1;
      lin_cta_range(p, cp->ops[o].epi, c, KB, u0, u1);
      it = WorkIter(u0, u1, KB, p.whole_per_cta, p.grid, c);
    }
  }
};

// L2 lookahead behind the ring, in 32 KiB stages per CTA.  Measured (B200, Llama-2-7B layer, tools/chain_bench.py): 0 -> 104 us,
// 4 -> 105, 8 -> 111, 16 -> 138, 24 -> 144: prefetched lines are evicted before the ring reaches them (148 x 16 x 32 KiB = 77 MB
// against 2 x 63 MB of L2) and the stream is fetched twice.  Off.
#ifndef PC_CHAIN_LOOKAHEAD
#define PC_CHAIN_LOOKAHEAD 0
#endif

template <int NPAD>
__global__ void __launch_bounds__(kThreads, 1) linear_chain_kernel(const __grid_constant__ ChainParams cp) {
  constexpr int KPS = 2;  // pre-tiled weights only: one 32 KiB bulk copy per pipeline stage
  using C = LCfg<NPAD, KPS>;
  pdl_launch_dependents();
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  if (((1024u - (raw_addr & 1023u)) & 1023u) > static_cast<uint32_t>(C::kAlignSlack)) __trap();
  __half* exch = reinterpret_cast<__half*>(smem + C::kOffExch);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kOffBars);
  uint64_t* full = bars;
  uint64_t* empty = full + C::kStages;
  uint64_t* acc_full = empty + C::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, c = blockIdx.x;
  const int n_ops = cp.n_ops;

  if (warp == 0 && lane == 0) {
    for (int o = 0; o < n_ops; ++o)
      if (cp.ops[o].kind == PC_CHAIN_LINEAR) tma_prefetch_desc(&cp.ops[o].map_x);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 2);  // the weight producer and the X producer each arrive once (with their byte counts)
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiThreads);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== weight stream: never waits for anything but a free ring slot =====================
    if (lane == 0) {
      // Behind the ring, the next `look` stages are asked into L2: when the consumers stall at a grid barrier the ring
      // fills up and the smem loads stop, but HBM keeps streaming into L2, and the ring refills from L2 (short latency)
      // once the consumers move again.
      const int look = (cp.flags & 32) ? 0 : ((cp.flags >> 8) ? (cp.flags >> 8) : PC_CHAIN_LOOKAHEAD);
      StageStream cur(&cp, c), ahead(&cp, c);
      const unsigned char* addr;
      const unsigned char* pf;
      bool more = look > 0;
      for (int k = 0; k < look && more; ++k) {
        more = ahead.next(pf);
        if (more && k >= C::kStages) bulk_prefetch_l2(pf, KPS * kWBytes);
      }
      int i = 0;
      while (cur.next(addr)) {
        const int s = i % C::kStages;
        if (i >= C::kStages) mbar_wait(&empty[s], static_cast<uint32_t>(((i / C::kStages) - 1) & 1));
        mbar_expect_tx(&full[s], KPS * kWBytes);
        bulk_g2s(smem + s * C::kStageBytes, addr, KPS * kWBytes, &full[s]);
        ++i;
        if (more) {
          more = ahead.next(pf);
          if (more) bulk_prefetch_l2(pf, KPS * kWBytes);
        }
      }
      pdl_wait();  // every CTA waits at least once: completion stays transitive along the stream
    }
  } else if (warp == 6) {
    // ===================== activation tiles: wait for the op's input, then TMA =====================
    if (lane == 0) {
      pdl_wait();
      int i = 0;
      for (int o = 0; o < n_ops; ++o) {
        const ChainOp& op = cp.ops[o];
        if (op.kind != PC_CHAIN_LINEAR || c >= op.lin.grid) continue;
        const LinParams& p = op.lin;
        const int KB = p.kb_total / KPS;
        long long u0, u1;
        lin_cta_range(p, op.epi, c, KB, u0, u1);
        if (u0 >= u1 && p.whole_per_cta == 0) continue;
        if (op.dep >= 0) {
          chain_wait(cp.bars + op.dep, G, cp.flags);
          if (!(cp.flags & 1)) fence_proxy_async_global();  // producers' generic-proxy stores, then this thread's async-proxy loads
        }
        WorkIter it(u0, u1, KB, p.whole_per_cta, p.grid, c);
        Seg sg;
        while (it.next(sg)) {
          for (int j = 0; j < sg.n; ++j, ++i) {
            const int s = i % C::kStages;
            if (i >= C::kStages) mbar_wait(&empty[s], static_cast<uint32_t>(((i / C::kStages) - 1) & 1));
            mbar_expect_tx(&full[s], KPS * C::kXBytes);
#pragma unroll
            for (int sb = 0; sb < KPS; ++sb)
              tma_load_3d(smem + s * C::kStageBytes + KPS * kWBytes + sb * C::kXBytes, &op.map_x, &full[s],
                          ((sg.kb0 + j) * KPS + sb) * BKK, 0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, NPAD, 0, 0);  // both operands K-major
      int i = 0, seg = 0;
      for (int o = 0; o < n_ops; ++o) {
        const ChainOp& op = cp.ops[o];
        if (op.kind != PC_CHAIN_LINEAR || c >= op.lin.grid) continue;
        const LinParams& p = op.lin;
        const int KB = p.kb_total / KPS;
        long long u0, u1;
        lin_cta_range(p, op.epi, c, KB, u0, u1);
        WorkIter it(u0, u1, KB, p.whole_per_cta, p.grid, c);
        Seg sg;
        while (it.next(sg)) {
          const int buf = seg & 1;
          if (seg >= 2) {
            mbar_wait(&acc_empty[buf], static_cast<uint32_t>(((seg >> 1) - 1) & 1));
            tc_fence_after();
          }
          const uint32_t d = tmem_base + buf * NPAD;
          for (int j = 0; j < sg.n; ++j, ++i) {
            const int s = i % C::kStages;
            mbar_wait(&full[s], static_cast<uint32_t>((i / C::kStages) & 1));
            tc_fence_after();
            const uint32_t w_addr = smem_u32(smem + s * C::kStageBytes);
            const uint32_t x_addr = w_addr + KPS * kWBytes;
#pragma unroll
            for (int sb = 0; sb < KPS; ++sb) {
#pragma unroll
              for (int kk = 0; kk < BKK / 16; ++kk)
                umma_ss(d, make_smem_desc(w_addr + sb * kWBytes + kk * 32, 16, 1024, 2),
                        make_smem_desc(x_addr + sb * C::kXBytes + kk * 32, 16, 1024, 2), idesc, (j > 0 || sb > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit(&empty[s]);
          }
          tc_commit(&acc_full[buf]);
          ++seg;
        }
      }
    }
  } else {
    // ===================== epilogue group: linears' epilogues, norm rows, grid-barrier arrivals =====================
    pdl_wait();  // everything below touches buffers other kernels write or read
    const int et = tid - 64;
    const int quad = warp & 3;       // TMEM lane quarter this warp may read
    const int f = quad * 32 + lane;  // row of the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    int seg = 0;
    for (int o = 0; o < n_ops; ++o) {
      const ChainOp& op = cp.ops[o];
      if (op.kind == PC_CHAIN_LINEAR) {
        const LinParams& p = op.lin;
        if (c < p.grid) {
          const int KB = p.kb_total / KPS;
          long long u0, u1;
          lin_cta_range(p, op.epi, c, KB, u0, u1);
          int row0 = p.row0;
          if (op.epi == PC_LIN_ROPE_APPEND && p.row0_dev != nullptr) row0 = max(0, min(__ldg(p.row0_dev), p.max_ctx - p.q_len));
          WorkIter it(u0, u1, KB, p.whole_per_cta, p.grid, c);
          const int t_first = it.it.t_first;
          Seg sg;
          while (it.next(sg)) {
            const int buf = seg & 1;
            mbar_wait(&acc_full[buf], static_cast<uint32_t>((seg >> 1) & 1));
            tc_fence_after();
            float v[NPAD];
            tmem_ld_acc<NPAD>(tmem_base + lane_base + buf * NPAD, v);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
            switch (op.epi) {
              case PC_LIN_PLAIN: lin_epilogue<NPAD, PC_LIN_PLAIN>(p, sg, v, c, KB, t_first, exch, flag, f, et, row0); break;
              case PC_LIN_SILU_MUL: lin_epilogue<NPAD, PC_LIN_SILU_MUL>(p, sg, v, c, KB, t_first, exch, flag, f, et, row0); break;
              case PC_LIN_PARTIALS: lin_epilogue<NPAD, PC_LIN_PARTIALS>(p, sg, v, c, KB, t_first, exch, flag, f, et, row0); break;
              default: lin_epilogue<NPAD, PC_LIN_ROPE_APPEND>(p, sg, v, c, KB, t_first, exch, flag, f, et, row0); break;
            }
            ++seg;
          }
        }
      } else {
        const NormParams& np = op.norm;
        const int r0 = G - 1 - c;  // the CTAs at the end of the grid: idle in the 128-CTA o/down partitions
        if (r0 < np.rows) {
          if (op.dep >= 0) {
            if (et == 0) chain_wait(cp.bars + op.dep, G, cp.flags);
            named_bar_sync(1, kEpiThreads);
          }
          if (!(cp.flags & 4))
            for (int r = r0; r < np.rows; r += G) chain_norm_row(np, op.kind, r, et, reinterpret_cast<float*>(exch));
        }
      }
      // this CTA is done with op o: every store above is ordered before the arrival
      named_bar_sync(1, kEpiThreads);
      if (et == 0) {
        if (!(cp.flags & 1)) fence_proxy_async_global();
        chain_arrive(cp.bars + o, cp.flags);
      }
    }
    if (et == 0) {
      // every wait of this CTA is behind it; the last CTA to leave zeroes the counters for the next launch
      const int old = atomicAdd(cp.bars + kMaxOps, 1);
      if (old == G - 1) {
        for (int o = 0; o <= kMaxOps; ++o) cp.bars[o] = 0;
        __threadfence();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_free<C::kTmemCols>(tmem_base);
  }
}

template <int NPAD>
int launch_chain(const ChainParams& cp, int grid, cudaStream_t st) {
  using C = LCfg<NPAD, 2>;
  static bool configured = false;
  if (!configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(linear_chain_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured = true;
  }
  PC_CHECK_CUDA(launch_pdl(linear_chain_kernel<NPAD>, dim3(grid), dim3(kThreads), C::kSmem, st, cp));
  count_launch();
  return PC_OK;
}

}  // namespace
}  // namespace lin
}  // namespace pc

PC_API int pc_linear_chain(const pc_chain_op* ops, int n_ops, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace pc;
  using namespace pc::lin;
  PC_REQUIRE(ops != nullptr && n_ops >= 1 && n_ops <= kMaxOps, "pc_linear_chain: %d ops (1..%d)", n_ops, kMaxOps);
  PC_REQUIRE(workspace != nullptr && workspace_bytes >= pc_linear_workspace_bytes(), "pc_linear_chain: workspace of %zu bytes, need %zu",
             workspace_bytes, pc_linear_workspace_bytes());
  static_assert(sizeof(ChainParams) <= 4096, "kernel parameter space");
  ChainParams cp{};
  cp.n_ops = n_ops;
  if (const char* e = dbg_env("PC_CHAIN_FLAGS")) cp.flags = atoi(e);
  cp.bars = reinterpret_cast<int*>(static_cast<char*>(workspace) + pc_linear_workspace_bytes() - kChainBarBytes);
  static_assert((kMaxOps + 1) * sizeof(int) <= kChainBarBytes, "barrier counters");
  int grid = 1, npad = 0, rows = -1;
  for (int o = 0; o < n_ops; ++o) {
    const pc_chain_op& src = ops[o];
    ChainOp& op = cp.ops[o];
    PC_REQUIRE(src.dep >= -1 && src.dep < o, "pc_linear_chain: op %d depends on op %d (must be an earlier op or -1)", o, src.dep);
    op.kind = src.kind;
    op.dep = src.dep;
    if (src.kind == PC_CHAIN_LINEAR) {
      pc_linear_params lp = src.lin;
      lp.workspace = workspace;
      lp.workspace_bytes = workspace_bytes;
      lp.prefetch_next = nullptr;
      PC_REQUIRE(lp.w_tiled != nullptr, "pc_linear_chain: op %d needs pre-tiled weights (pc_linear_tile_weights)", o);
      PC_REQUIRE(lp.q_len > 0, "pc_linear_chain: op %d has q_len %d", o, lp.q_len);
      LinLaunch L;
      if (int rc = prepare_linear(&lp, &L)) return rc;
      PC_REQUIRE(npad == 0 || npad == L.npad, "pc_linear_chain: the linears of a chain must share the token-column padding (16 / 32)");
      PC_REQUIRE(rows < 0 || rows == lp.q_len, "pc_linear_chain: the ops of a chain must share q_len / rows");
      npad = L.npad;
      rows = lp.q_len;
      op.lin = L.lp;
      op.map_x = L.map_x;
      op.epi = L.epilogue;
      if (L.grid > grid) grid = L.grid;
    } else if (src.kind == PC_CHAIN_ADD_RMSNORM || src.kind == PC_CHAIN_ADD_LAYERNORM) {
      PC_REQUIRE(src.x && src.weight && src.h_out, "pc_linear_chain: op %d (norm): null tensor", o);
      PC_REQUIRE(src.rows > 0 && src.hidden > 0 && src.hidden % 8 == 0, "pc_linear_chain: op %d (norm): bad shape", o);
      PC_REQUIRE(rows < 0 || rows == src.rows, "pc_linear_chain: the ops of a chain must share q_len / rows");
      rows = src.rows;
      PC_REQUIRE((src.part1 == nullptr) == (src.n_splits1 == 0) && (src.part2 == nullptr) == (src.n_splits2 == 0) &&
                     (src.part2 == nullptr || src.part1 != nullptr),
                 "pc_linear_chain: op %d (norm): partial stacks and their split counts disagree", o);
      PC_REQUIRE(src.kind == PC_CHAIN_ADD_LAYERNORM || (src.part2 == nullptr && src.bias == nullptr),
                 "pc_linear_chain: op %d: RMSNorm takes one delta and no bias", o);
      const uintptr_t align = reinterpret_cast<uintptr_t>(src.x) | reinterpret_cast<uintptr_t>(src.part1) | reinterpret_cast<uintptr_t>(src.part2) |
                              reinterpret_cast<uintptr_t>(src.weight) | reinterpret_cast<uintptr_t>(src.bias) | reinterpret_cast<uintptr_t>(src.h_out);
      PC_REQUIRE(align % 16 == 0, "pc_linear_chain: op %d (norm): tensors must be 16B aligned", o);
      NormParams& np = op.norm;
      np.x = static_cast<__half*>(src.x);
      np.part1 = static_cast<const float*>(src.part1);
      np.part2 = static_cast<const float*>(src.part2);
      np.weight = static_cast<const __half*>(src.weight);
      np.bias = static_cast<const __half*>(src.bias);
      np.h_out = static_cast<__half*>(src.h_out);
      np.n1 = src.n_splits1;
      np.n2 = src.n_splits2;
      np.rows = src.rows;
      np.hidden = src.hidden;
      np.eps = src.eps;
    } else {
      return set_error(PC_ERR_INVALID, "pc_linear_chain: op %d has unknown kind %d", o, src.kind);
    }
  }
  if (npad == 0) npad = rows <= 16 ? 16 : 32;  // norms only
  PC_REQUIRE(grid <= sm_count(), "pc_linear_chain: %d CTAs exceed the SM count", grid);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return npad == 16 ? launch_chain<16>(cp, grid, st) : launch_chain<32>(cp, grid, st);
}
