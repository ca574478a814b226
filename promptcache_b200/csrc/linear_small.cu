// Skinny-M linear layers of the prefill / decode step as ONE weight-streaming tcgen05 kernel (SURVEY.md
// section 8 row a13: q/k/v/o projections, MLP and lm_head -- reference promptcache/model/llama2.py:345-351,
// :405, :213-244, :1050).  With q <= 32 uncached tokens these GEMMs are pure HBM streams of the fp16 weights
// (13.5 GB per Llama-2-7B forward, two thirds of the TTFT step) and every epilogue (RoPE + KV append, SiLU*up)
// used to cost another launch.
//
//   D^T[128 out-features x NPAD tokens] += W[128 x 64] . X[NPAD x 64]^T        (operands swapped so that the
//   weight tile fills the MMA's M = 128 and the few tokens are its N)
//
// * weights: either the row-major nn.Linear matrix through 128B-swizzled tensor-map boxes (one SM's TMA sustains
//   ~49 GB/s on such 128-byte rows), or -- the default of the models -- a pre-tiled copy (pc_linear_tile_weights:
//   the [tile][k-block] sequence of 16 KiB shared-memory images) streamed with ONE 32 KiB bulk copy per stage
//   (61 / 116 GB/s per SM with 16 / 32 KiB copies);
// * roles: a weight-producer thread (never waits for the previous kernel: weights are static), an X-producer
//   thread (waits: X is the previous kernel's output), an MMA-issuing thread with two TMEM accumulators, four
//   epilogue warps; 6 (pre-tiled) or 10-11 (tensor-map) ring stages, i.e. ~190 KiB of weights in flight per SM;
// * work partition (chosen on the host per shape, all expressed as a list of (tile, k-range) segments per CTA):
//     - whole tiles per CTA when the tile count divides a grid of >= 5/8 of the SMs; the QKV projection (96 head tiles) runs
//       as clusters of 3 tile owners + 1 helper CTA whose fp32 partials travel through distributed shared memory (ClIter);
//     - more tiles than SMs (gate/up, lm_head): whole tiles dealt out per CTA + the remainder shared stream-K
//       fashion and visited FIRST, so that its reduction overlaps the streaming of the whole tiles;
//     - PC_LIN_PARTIALS (o_proj, down_proj): regular split-K, NO reduction here -- the consumer
//       (pc_add_rmsnorm_partials) adds the fp32 partials in split order and rounds once;
//     - otherwise stream-K: equal contiguous ranges of the (tile, k-block) space; a range that covers only part of
//       a tile's K parks its fp32 partial in a per-CTA slot and the LAST CTA to arrive at the tile (arrival counter,
//       self-cleaning, ONE acq_rel RMW per arrival: gpu-scope fences do not complete while the SM's TMA stream keeps its
//       memory pipe full) adds the slots in a fixed order -- deterministic, no second launch;
// * epilogues: plain fp16 store; half(silu(gate)) * up (llama2.py:241) with gate and up rows of the same
//   features in one tile; RoPE at explicit position ids + in-place KV append with the reference's three fp16
//   roundings (llama2.py:195-210,357-364), a tile holding one head (D = 128) or two (D = 64, Falcon).
#include <type_traits>

#include "linear_common.cuh"

namespace pc {
namespace lin {
namespace {
// Work list of a CTA in the owner + helper cluster partition (CL): clusters of S (3 or 4) consecutive CTAs serve S - 1
// consecutive row tiles; ranks 0..S-2 own one tile each and stream units [0, KB - KH) of it, rank S-1 streams the remaining
// KH units of every owner's tile.  All CTAs of the cluster stream (nearly) the same number of bytes and nothing crosses HBM
// twice: the helper's fp32 partials travel through distributed shared memory (two cluster barriers at the very end).
struct ClIter {
  int tile0, rank, S, KB, KH, j;
  __device__ ClIter(int c, int S_, int KB_, int KH_) : tile0((c / S_) * (S_ - 1)), rank(c % S_), S(S_), KB(KB_), KH(KH_), j(0) {}
  __device__ bool next(Seg& s) {
    if (rank < S - 1) {
      if (j > 0) return false;
      s.tile = tile0 + rank;
      s.kb0 = 0;
      s.n = KB - KH;
    } else {
      if (j >= S - 1) return false;
      s.tile = tile0 + j;
      s.kb0 = KB - KH;
      s.n = KH;
    }
    s.rel = -1;
    ++j;
    return true;
  }
};

template <int NPAD, int EPI, bool TILED, bool CL = false>
__global__ void __launch_bounds__(kThreads, LCfg<NPAD, TILED ? 2 : 1>::kMinBlocks)
linear_small_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const LinParams p) {
  static_assert(!CL || (TILED && EPI == PC_LIN_ROPE_APPEND), "the cluster partition serves the pre-tiled QKV projection");
  constexpr int KPS = TILED ? 2 : 1;
  using C = LCfg<NPAD, KPS>;
  pdl_launch_dependents();  // the next kernel may start its prologue (and its own weight prefetch) right away
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
  int* flag = reinterpret_cast<int*>(tmem_slot + 1);  // [3]: last arriver?, first / last contributing CTA

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = p.kb_total / KPS;  // pipeline stages (units) per tile
  const int G = p.grid;
  int c = blockIdx.x;
#if defined(PC_DEBUG) && defined(PC_LIN_BOUNDS)
  if (p.bounds != nullptr) {  // experiment (PC_LIN_BOUNDS_PTR): partition index = the SM this CTA runs on; needs grid == #SMs on an idle GPU
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    c = static_cast<int>(smid);
  }
#endif
  long long u0, u1;
  lin_cta_range(p, EPI, c, KB, u0, u1);
  auto make_iter = [&]() {
    if constexpr (CL) return ClIter(c, p.cl_size, KB, p.cl_helper_units);
    else return WorkIter(u0, u1, KB, p.whole_per_cta, G, c);
  };
  // CL: the accumulator rows the epilogue threads carry across the two cluster barriers (owner: its tile; helper: three partials)
  float cl_v[CL ? 3 : 1][NPAD];
  __half cl_cs[CL ? NPAD : 1], cl_sn[CL ? NPAD : 1];  // owners: the RoPE table values of their rows, gathered under the MMAs
  if (tid == 0) LIN_TRACE(0);
  if (tid == 0 && p.trace != nullptr) {  // slot 7: the SM this CTA runs on
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.trace[blockIdx.x * 8 + 7] = smid;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_x);
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
  if (tid == 0) LIN_TRACE(1);

  if (warp == 0) {
    // ===================== TMA producer: the weight stream =====================
    if (lane == 0) {
      // One TMA instruction costs this thread ~110 ns whatever its size, so the weight tile travels as ONE 128-row box
      // when its rows are consecutive, and the X tiles are issued by a thread of their own (warp 6).  The weights are
      // static: the ring is filled with them without waiting for the previous kernel (whose output is X).
      std::conditional_t<CL, ClIter, WorkIter> it = make_iter();
      Seg sg;
      int i = 0;
      while (it.next(sg)) {
        for (int j = 0; j < sg.n; ++j, ++i) {
          const int s = i % C::kStages;
          if (i >= C::kStages) mbar_wait(&empty[s], static_cast<uint32_t>(((i / C::kStages) - 1) & 1));
          unsigned char* dw = smem + s * C::kStageBytes;
          const int row_a = sg.tile * p.tile_row_stride, k0 = (sg.kb0 + j) * BKK;
          mbar_expect_tx(&full[s], KPS * kWBytes);
          if constexpr (TILED) {
            // the stage's weights are ONE contiguous 32 KiB block of the pre-tiled copy: a single 1-D bulk load
            bulk_g2s(dw, p.w_tiled + (static_cast<long long>(sg.tile) * KB + sg.kb0 + j) * (KPS * kWBytes), KPS * kWBytes, &full[s]);
          } else if (p.w_one_box) {
            tma_load_3d(dw, &map_w, &full[s], k0, row_a, 0);
          } else {
            tma_load_3d(dw, &map_w, &full[s], k0, row_a, 0);
            tma_load_3d(dw + kWBytes / 2, &map_w, &full[s], k0, row_a + p.second_half_off, 0);
          }
        }
      }
      if (p.pf_w != nullptr) {
        // all of this CTA's weight loads are in flight: hand the memory system the head of the next launch's stream
        constexpr uint32_t kUnitBytes = 2 * kWBytes;  // a pre-tiled pipeline stage (two k-blocks)
        for (int c2 = c; c2 < p.pf_grid; c2 += G) {
          long long n0, n1;
          cta_units(c2, p.pf_grid, p.pf_units_total, p.pf_kb, p.pf_splits, n0, n1);
          WorkIter nit(n0, n1, p.pf_kb, p.pf_whole, p.pf_grid, c2);
          Seg ns;
          int left = p.pf_units;
          while (left > 0 && nit.next(ns)) {
            const int n = min(ns.n, left);
            for (int j = 0; j < n; ++j)
              bulk_prefetch_l2(p.pf_w + (static_cast<long long>(ns.tile) * p.pf_kb + ns.kb0 + j) * kUnitBytes, kUnitBytes);
            left -= n;
          }
        }
      }
      pdl_wait();  // every CTA waits at least once: completion stays transitive along the stream
    }
  } else if (warp == 6) {
    // ===================== TMA producer: the X tiles (output of the previous kernel) =====================
    if (lane == 0) {
      pdl_wait();
      std::conditional_t<CL, ClIter, WorkIter> it = make_iter();
      Seg sg;
      int i = 0;
      while (it.next(sg)) {
        for (int j = 0; j < sg.n; ++j, ++i) {
          const int s = i % C::kStages;
          if (i >= C::kStages) mbar_wait(&empty[s], static_cast<uint32_t>(((i / C::kStages) - 1) & 1));
          mbar_expect_tx(&full[s], KPS * C::kXBytes);
#pragma unroll
          for (int sb = 0; sb < KPS; ++sb)
            tma_load_3d(smem + s * C::kStageBytes + KPS * kWBytes + sb * C::kXBytes, &map_x, &full[s],
                        ((sg.kb0 + j) * KPS + sb) * BKK, 0, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(BM, NPAD, 0, 0);  // both operands K-major
      int i = 0, seg = 0;
      std::conditional_t<CL, ClIter, WorkIter> it = make_iter();
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
          if (i == 0) LIN_TRACE(2);
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
      LIN_TRACE(3);
    }
  } else {
    // ===================== epilogue: thread = out-feature row of the tile =====================
    pdl_wait();  // everything below touches buffers other kernels write or read
    const int et = tid - 64;
    const int quad = warp & 3;                 // TMEM lane quarter this warp may read
    const int f = quad * 32 + lane;            // row of the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    int row0 = p.row0;
    if (EPI == PC_LIN_ROPE_APPEND && p.row0_dev != nullptr) row0 = max(0, min(__ldg(p.row0_dev), p.max_ctx - p.q_len));
    int seg = 0;
    if constexpr (CL) {
      // read this CTA's accumulators (owner: one, helper: three); the exchange and the epilogue follow the role branches
      const int n_seg = (c % p.cl_size) < p.cl_size - 1 ? 1 : p.cl_size - 1;
      if (n_seg == 1 && p.cos != nullptr) {  // owner of a rotated (query / key) head: its table gathers do not wait for the sums
        const int own_tile = (c / p.cl_size) * (p.cl_size - 1) + c % p.cl_size;
        if ((own_tile * BM + f) / p.head_dim < p.n_heads + p.n_kv_heads) rope_gather<NPAD>(p, f & (p.head_dim - 1), cl_cs, cl_sn);
      }
#pragma unroll
      for (int sgi = 0; sgi < 3; ++sgi) {  // compile-time indices keep cl_v in registers
        if (sgi < n_seg) {
          const int buf = sgi & 1;
          mbar_wait(&acc_full[buf], static_cast<uint32_t>((sgi >> 1) & 1));
          tc_fence_after();
          tmem_ld_acc<NPAD>(tmem_base + lane_base + buf * NPAD, cl_v[sgi]);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&acc_empty[buf]);
        }
      }
      seg = n_seg;
    } else {
      WorkIter it(u0, u1, KB, p.whole_per_cta, G, c);
      const int t_first = it.it.t_first;
      Seg sg;
      while (it.next(sg)) {
        const int buf = seg & 1;
        mbar_wait(&acc_full[buf], static_cast<uint32_t>((seg >> 1) & 1));
        tc_fence_after();
        if (et == 0) LIN_TRACE(4);
        float v[NPAD];
        tmem_ld_acc<NPAD>(tmem_base + lane_base + buf * NPAD, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);
        lin_epilogue<NPAD, EPI>(p, sg, v, c, KB, t_first, exch, flag, f, et, row0);
        if (et == 0 && sg.n != KB) LIN_TRACE(5);  // end of a partial segment's epilogue (incl. the cross-CTA reduction if last)
        ++seg;
      }
    }
    if (et == 0) LIN_TRACE(6);
  }

  if constexpr (CL) {
    // ---- helper -> owners through distributed shared memory: fp32 [token][row] into the start of each owner's (drained) ring ----
    const bool epi_warp = warp >= 2 && warp < 6;
    const int S = p.cl_size, rank = c % S;
    const int f = (warp & 3) * 32 + lane;
    float* recv = reinterpret_cast<float*>(smem);  // NPAD * 128 floats
    cluster_sync_all();  // A: every CTA of the cluster has read its accumulators -> the owners' rings are free
    if (epi_warp && rank == S - 1) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (j < S - 1) {
          const uint32_t dst = cluster_map_shared(smem_u32(recv + f), static_cast<uint32_t>(j));
#pragma unroll
          for (int t = 0; t < NPAD; ++t) st_cluster_f32(dst + t * BM * 4, cl_v[j][t]);
        }
      }
    }
    cluster_sync_all();  // B: the partials have landed
    if (epi_warp && rank < S - 1) {
      const int et = tid - 64;
      int row0 = p.row0;
      if (p.row0_dev != nullptr) row0 = max(0, min(__ldg(p.row0_dev), p.max_ctx - p.q_len));
#pragma unroll
      for (int t = 0; t < NPAD; ++t) cl_v[0][t] += recv[t * BM + f];
      Seg whole;
      whole.tile = (c / S) * (S - 1) + rank;
      whole.kb0 = 0;
      whole.n = KB;  // complete sum: the plain RoPE + append epilogue
      whole.rel = -1;
      lin_epilogue<NPAD, EPI>(p, whole, cl_v[0], c, KB, 0, exch, flag, f, et, row0, p.cos != nullptr ? cl_cs : nullptr, cl_sn);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_free<C::kTmemCols>(tmem_base);
  }
}

// owner + helper clusters for the pre-tiled QKV projection (see ClIter)
template <int NPAD>
cudaError_t configure_cl() {
  static bool configured = false;
  if (configured) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(linear_small_kernel<NPAD, PC_LIN_ROPE_APPEND, true, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, LCfg<NPAD, 2>::kSmem);
  configured = e == cudaSuccess;
  return e;
}
template <int NPAD>
bool cl_clusters_fit(int cl_size, int n_clusters) {
  static int fit[5] = {-1, -1, -1, -1, -1};
  if (fit[cl_size] < 0) {
    if (configure_cl<NPAD>() != cudaSuccess) return false;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[2];
    fill_launch_config(cfg, attr, dim3(cl_size * 32), dim3(kThreads), LCfg<NPAD, 2>::kSmem, nullptr, cl_size);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, linear_small_kernel<NPAD, PC_LIN_ROPE_APPEND, true, true>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    fit[cl_size] = n;
  }
  return fit[cl_size] >= n_clusters;
}
template <int NPAD>
int launch_lin_cl(const CUtensorMap& mw, const CUtensorMap& mx, const LinParams& lp, int grid, cudaStream_t st) {
  PC_CHECK_CUDA(configure_cl<NPAD>());
  PC_CHECK_CUDA(launch_pdl_cluster(linear_small_kernel<NPAD, PC_LIN_ROPE_APPEND, true, true>, lp.cl_size, dim3(grid), dim3(kThreads),
                                   LCfg<NPAD, 2>::kSmem, st, mw, mx, lp));
  count_launch();
  return PC_OK;
}

template <int NPAD, int EPI, bool TILED>
int launch_lin(const CUtensorMap& mw, const CUtensorMap& mx, const LinParams& lp, int grid, cudaStream_t st) {
  using C = LCfg<NPAD, TILED ? 2 : 1>;
  static bool configured = false;
  if (!configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(linear_small_kernel<NPAD, EPI, TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured = true;
  }
  PC_CHECK_CUDA(launch_pdl(linear_small_kernel<NPAD, EPI, TILED>, dim3(grid), dim3(kThreads), C::kSmem, st, mw, mx, lp));
  count_launch();
  return PC_OK;
}

template <int NPAD, bool TILED>
int launch_lin_n(int epi, const CUtensorMap& mw, const CUtensorMap& mx, const LinParams& lp, int grid, cudaStream_t st) {
  switch (epi) {
    case PC_LIN_PLAIN: return launch_lin<NPAD, PC_LIN_PLAIN, TILED>(mw, mx, lp, grid, st);
    case PC_LIN_SILU_MUL: return launch_lin<NPAD, PC_LIN_SILU_MUL, TILED>(mw, mx, lp, grid, st);
    case PC_LIN_PARTIALS: return launch_lin<NPAD, PC_LIN_PARTIALS, TILED>(mw, mx, lp, grid, st);
    default: return launch_lin<NPAD, PC_LIN_ROPE_APPEND, TILED>(mw, mx, lp, grid, st);
  }
}

// One-off re-layout of a weight matrix into the [tile][k-block] sequence of 16 KiB shared-memory images the kernel's
// pipeline consumes (128 rows x 64 columns, 128B-swizzled K-major, exactly what the tensor-map load produces -- it IS
// produced by that load here and written back verbatim).
__global__ void __launch_bounds__(32) tile_weights_kernel(const __grid_constant__ CUtensorMap map_w, unsigned char* __restrict__ out,
                                                          long long units, int kb_total, int tile_row_stride, int second_half_off) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kWBytes);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    uint32_t phase = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x) {
      const int tile = static_cast<int>(u / kb_total), kb = static_cast<int>(u - static_cast<long long>(tile) * kb_total);
      const int row_a = tile * tile_row_stride;
      mbar_expect_tx(bar, kWBytes);
      tma_load_3d(smem, &map_w, bar, kb * BKK, row_a, 0);
      tma_load_3d(smem + kWBytes / 2, &map_w, bar, kb * BKK, row_a + second_half_off, 0);
      mbar_wait(bar, phase);
      phase ^= 1;
      bulk_s2g(out + u * kWBytes, smem, kWBytes);
      bulk_commit();
      bulk_wait_read<0>();  // the buffer is re-filled by the next load
    }
    bulk_wait_all<0>();
  }
}

struct TileGeom {
  int n_tiles, tile_row_stride, second_half_off;
};

int tile_geometry(int n_rows, int epilogue, TileGeom* g) {
  switch (epilogue) {
    case PC_LIN_PLAIN:
    case PC_LIN_PARTIALS:
    case PC_LIN_ROPE_APPEND:
      g->n_tiles = (n_rows + BM - 1) / BM;
      g->tile_row_stride = BM;
      g->second_half_off = 64;
      return PC_OK;
    case PC_LIN_SILU_MUL:
      if (n_rows % 128 != 0) return set_error(PC_ERR_INVALID, "pc_linear(silu_mul): n_rows %d must be 2 * a multiple of 64", n_rows);
      g->n_tiles = n_rows / 128;
      g->tile_row_stride = 64;
      g->second_half_off = n_rows / 2;
      return PC_OK;
    default:
      return set_error(PC_ERR_INVALID, "pc_linear: unknown epilogue %d", epilogue);
  }
}

}  // namespace
}  // namespace lin
}  // namespace pc

// k-blocks of the pre-tiled copy: a pipeline stage is two k-blocks, so an odd count (Falcon: k = 4544 = 71 blocks) is padded
// with one block of zeros (the tensor-map load that produces the copy reads zeros past k; so do the X loads of the kernel)
int pc::lin::tiled_k_blocks(int k) { return ((k / BKK) + 1) & ~1; }
using pc::lin::tiled_k_blocks;

PC_API size_t pc_linear_tiled_bytes(int n_rows, int k, int epilogue) {
  pc::lin::TileGeom g;
  if (n_rows <= 0 || k <= 0 || k % pc::lin::BKK != 0 || pc::lin::tile_geometry(n_rows, epilogue, &g) != PC_OK) return 0;
  return static_cast<size_t>(g.n_tiles) * tiled_k_blocks(k) * pc::lin::kWBytes;
}

PC_API int pc_linear_tile_weights(const void* w, int n_rows, int k, int epilogue, void* w_tiled, void* stream) {
  using namespace pc;
  using namespace pc::lin;
  PC_REQUIRE(w && w_tiled, "pc_linear_tile_weights: null pointer");
  PC_REQUIRE(n_rows > 0 && k > 0 && k % BKK == 0, "pc_linear_tile_weights: k %d must be a positive multiple of 64", k);
  PC_REQUIRE((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(w_tiled)) % 16 == 0, "pc_linear_tile_weights: 16B alignment");
  TileGeom g;
  if (int rc = tile_geometry(n_rows, epilogue, &g)) return rc;
  CUtensorMap mw;
  if (int rc = make_tensor_map_3d(&mw, w, k, n_rows, 1, static_cast<long long>(n_rows) * k, 64)) return rc;
  const int kb_total = tiled_k_blocks(k);
  const long long units = static_cast<long long>(g.n_tiles) * kb_total;
  static bool configured = false;
  if (!configured) {
    PC_CHECK_CUDA(cudaFuncSetAttribute(tile_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWBytes + 2048));
    configured = true;
  }
  int grid = sm_count() * 8;
  if (grid > units) grid = static_cast<int>(units);
  tile_weights_kernel<<<grid, 32, kWBytes + 2048, static_cast<cudaStream_t>(stream)>>>(mw, static_cast<unsigned char*>(w_tiled), units,
                                                                                       kb_total, g.tile_row_stride, g.second_half_off);
  PC_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return PC_OK;
}

PC_API int pc_linear_partial_splits(int n_rows, int k) {
  if (n_rows <= 0 || k <= 0) return 0;
  const int n_tiles = (n_rows + pc::lin::BM - 1) / pc::lin::BM;
  int s = pc::sm_count() / n_tiles;  // as many CTAs per row tile as fit at once ...
  if (s > 8) s = 8;
  if (s > k / 128) s = k / 128;      // ... with at least one pipeline stage (128 columns) each
  if (s < 1) s = 1;
  return s;
}

namespace pc {
namespace lin {
namespace {
#ifndef PC_LIN_PREFETCH_UNITS
#define PC_LIN_PREFETCH_UNITS 6
#endif
constexpr int kPrefetchUnits = PC_LIN_PREFETCH_UNITS;  // pipeline stages (32 KiB each) of the next launch asked from L2 per CTA

// Grid / partition of one launch (see the header comment).  An SM that is fed 16 KiB per TMA instruction pulls
// ~100 GB/s, so ~2/3 of the SMs saturate HBM.  When the tile count has a divisor in [sm/2, sm] every CTA gets WHOLE
// tiles (no split-K, no cross-CTA reduction: the tail is a plain epilogue); otherwise all SMs share the (tile, k-block)
// space evenly (stream-K, partial tiles reduced).  `units_total` comes in as n_tiles * units_per_tile.
void choose_grid(int epilogue, int n_tiles, int units_per_tile, bool tiled, int splits, int* grid_out, int* whole_per_cta,
                 long long* units_total) {
  int grid = sm_count();
  // row-major tensor-map loads need every SM's TMA; pre-tiled streams saturate with fewer SMs.  (Sharing the 96 head tiles of
  // the QKV projection stream-K fashion among all 148 SMs instead was measured: 3.51 vs 3.30 ms per C2 step -- the
  // cross-CTA reduction in front of the RoPE epilogue costs more than the 52 extra streams bring.)
  const int min_whole = tiled ? (sm_count() * 5) / 8 : sm_count();
  for (int g = sm_count(); g >= min_whole; --g)
    if (n_tiles % g == 0) {
      grid = g;
      break;
    }
  if (grid == sm_count() && n_tiles < sm_count()) {
    // few tiles (o_proj, down_proj): a regular split-K, the same number of CTAs per tile, keeps the reduction groups small
    const int per_tile = sm_count() / n_tiles;
    if (per_tile >= 2 && per_tile * n_tiles >= (sm_count() * 3) / 4) grid = per_tile * n_tiles;
  }
  const char* hy = dbg_env("PC_LIN_HYBRID");
  *whole_per_cta = 0;
  if (epilogue == PC_LIN_PARTIALS) grid = n_tiles * splits;
  else if (const char* e = dbg_env("PC_LIN_GRID")) grid = atoi(e);
  else if (n_tiles > sm_count() && !(hy && hy[0] == '0')) {
    // more tiles than SMs: whole tiles first-come per CTA, the remainder shared stream-K (and processed first)
    grid = sm_count();
    *whole_per_cta = n_tiles / grid;
    *units_total = static_cast<long long>(n_tiles - *whole_per_cta * grid) * units_per_tile;
  }
  if (grid > sm_count()) grid = sm_count();
  if (*whole_per_cta == 0 && grid > *units_total) grid = static_cast<int>(*units_total);
  if (grid < 1) grid = 1;
  *grid_out = grid;
}
}  // namespace
}  // namespace lin
}  // namespace pc

// [tile arrival counters][stream-K slots][chain barrier counters (linear_chain.cu)]
PC_API size_t pc_linear_workspace_bytes(void) {
  return pc::lin::kLinCounterBytes + static_cast<size_t>(pc::sm_count()) * 2 * 32 * pc::lin::BM * sizeof(float) + pc::lin::kChainBarBytes;
}

int pc::lin::prepare_linear(const pc_linear_params* p, LinLaunch* out, bool allow_cluster) {
  using namespace pc;
  using namespace pc::lin;
  if (!p) return set_error(PC_ERR_INVALID, "pc_linear_small: null params");
  PC_REQUIRE(p->x && p->w && p->workspace, "pc_linear_small: null tensor / workspace");
  PC_REQUIRE(p->q_len > 0 && p->q_len <= 32, "pc_linear_small: q_len %d must be in [1,32]", p->q_len);
  PC_REQUIRE(p->k > 0 && p->k % BKK == 0, "pc_linear_small: k %d must be a positive multiple of 64", p->k);
  PC_REQUIRE(p->n_rows > 0, "pc_linear_small: n_rows %d", p->n_rows);
  PC_REQUIRE(p->x_stride == p->k, "pc_linear_small: x must be contiguous (x_stride %lld != k %d)",
             static_cast<long long>(p->x_stride), p->k);
  PC_REQUIRE((reinterpret_cast<uintptr_t>(p->x) | reinterpret_cast<uintptr_t>(p->w) | reinterpret_cast<uintptr_t>(p->workspace)) % 16 == 0,
             "pc_linear_small: x / w / workspace must be 16B aligned");
  PC_REQUIRE(p->workspace_bytes >= pc_linear_workspace_bytes(), "pc_linear_small: workspace of %zu bytes, need %zu",
             p->workspace_bytes, pc_linear_workspace_bytes());
  LinParams& lp = out->lp;
  lp = LinParams{};
  lp.q_len = p->q_len;
  const bool tiled = p->w_tiled != nullptr;
  lp.kb_total = tiled ? tiled_k_blocks(p->k) : p->k / BKK;
  PC_REQUIRE(!tiled || reinterpret_cast<uintptr_t>(p->w_tiled) % 16 == 0, "pc_linear_small: w_tiled must be 16B aligned");
  int n_tiles = 0;
  switch (p->epilogue) {
    case PC_LIN_PLAIN:
      PC_REQUIRE(p->out && p->out_stride >= p->n_rows, "pc_linear_small: out / out_stride");
      n_tiles = (p->n_rows + BM - 1) / BM;
      lp.n_out = p->n_rows;
      lp.tile_row_stride = BM;
      lp.second_half_off = 64;
      break;
    case PC_LIN_PARTIALS:
      PC_REQUIRE(p->out && p->out_stride == p->n_rows, "pc_linear_small(partials): out must be fp32 [splits][q_len][n_rows]");
      n_tiles = (p->n_rows + BM - 1) / BM;
      lp.n_out = p->n_rows;
      lp.tile_row_stride = BM;
      lp.second_half_off = 64;
      lp.splits = pc_linear_partial_splits(p->n_rows, p->k);
      PC_REQUIRE(n_tiles * lp.splits <= sm_count(), "pc_linear_small(partials): %d row tiles exceed the SM count", n_tiles);
      break;
    case PC_LIN_SILU_MUL: {
      PC_REQUIRE(p->n_rows % 2 == 0, "pc_linear_small(silu_mul): n_rows %d must be even (gate rows, then up rows)", p->n_rows);
      const int inter = p->n_rows / 2;
      PC_REQUIRE(p->out && p->out_stride >= inter, "pc_linear_small(silu_mul): out / out_stride");
      n_tiles = (inter + 63) / 64;
      lp.n_out = inter;
      lp.tile_row_stride = 64;
      lp.second_half_off = inter;
      // a ragged last tile would pull up-rows into the gate half: keep it exact
      PC_REQUIRE(inter % 64 == 0, "pc_linear_small(silu_mul): intermediate size %d must be a multiple of 64", inter);
      break;
    }
    case PC_LIN_ROPE_APPEND: {
      const int hd = p->head_dim == 0 ? 128 : p->head_dim;
      PC_REQUIRE(hd == 64 || hd == 128, "pc_linear_small(rope_append): head_dim %d must be 64 or 128", hd);
      PC_REQUIRE(p->n_heads > 0 && p->n_kv_heads > 0 && p->n_rows == (p->n_heads + 2 * p->n_kv_heads) * hd,
                 "pc_linear_small(rope_append): n_rows %d must be (H + 2 Hkv) * head_dim", p->n_rows);
      lp.head_dim = hd;
      PC_REQUIRE(p->q_out && p->k_cache && p->v_cache, "pc_linear_small(rope_append): null q_out / cache");
      PC_REQUIRE((p->cos == nullptr) == (p->sin == nullptr), "pc_linear_small(rope_append): cos/sin must both be set");
      PC_REQUIRE(p->cos == nullptr || (p->position_ids != nullptr && p->table_rows > 0),
                 "pc_linear_small(rope_append): rotary needs position ids and a table");
      PC_REQUIRE(p->dst_row0 >= 0 && static_cast<long long>(p->dst_row0) + p->q_len <= p->max_ctx,
                 "pc_linear_small(rope_append): rows [%d,%d) exceed max_ctx %d", p->dst_row0, p->dst_row0 + p->q_len, p->max_ctx);
      n_tiles = (p->n_rows + BM - 1) / BM;
      lp.n_out = p->n_rows;
      lp.tile_row_stride = BM;
      lp.second_half_off = 64;
      break;
    }
    default:
      return set_error(PC_ERR_INVALID, "pc_linear_small: unknown epilogue %d", p->epilogue);
  }
  PC_REQUIRE(n_tiles <= kMaxTiles, "pc_linear_small: %d tiles exceed the counter region", n_tiles);
  {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      return set_error(PC_ERR_CUDA, "no CUDA device");
    if (major != 10) return set_error(PC_ERR_UNSUPPORTED, "this library is built for sm_100a only");
  }
  lp.part_out = p->epilogue == PC_LIN_PARTIALS ? static_cast<float*>(p->out) : nullptr;
  PC_REQUIRE(p->epilogue != PC_LIN_PARTIALS || (tiled ? lp.kb_total / 2 : lp.kb_total) >= lp.splits,
             "pc_linear_small(partials): k %d is too short for %d splits", p->k, lp.splits);
  lp.out = static_cast<__half*>(p->out);
  lp.out_stride = p->out_stride;
  lp.counters = static_cast<int*>(p->workspace);
  lp.slots = reinterpret_cast<float*>(static_cast<char*>(p->workspace) + kLinCounterBytes);
  const int units_per_tile = tiled ? lp.kb_total / 2 : lp.kb_total;
  lp.units_total = static_cast<long long>(n_tiles) * units_per_tile;
  lp.whole_per_cta = 0;
  lp.w_tiled = static_cast<const unsigned char*>(p->w_tiled);
  lp.pos = reinterpret_cast<const long long*>(p->position_ids);
  lp.cos = static_cast<const __half*>(p->cos);
  lp.sin = static_cast<const __half*>(p->sin);
  lp.table_rows = p->table_rows;
  lp.q_out = static_cast<__half*>(p->q_out);
  lp.k_cache = static_cast<__half*>(p->k_cache);
  lp.v_cache = static_cast<__half*>(p->v_cache);
  lp.max_ctx = p->max_ctx;
  lp.row0 = p->dst_row0;
  lp.row0_dev = p->dst_row0_dev;
  lp.n_heads = p->n_heads;
  lp.n_kv_heads = p->n_kv_heads;
  lp.act = (p->epilogue == PC_LIN_PLAIN && p->activation == PC_LIN_ACT_GELU) ? 1 : 0;
  PC_REQUIRE(p->activation == 0 || (p->epilogue == PC_LIN_PLAIN && p->activation == PC_LIN_ACT_GELU),
             "pc_linear_small: activation %d (only PC_LIN_ACT_GELU with PC_LIN_PLAIN)", p->activation);
  lp.trace = nullptr;
  if (const char* e = dbg_env("PC_LIN_TRACE_PTR")) lp.trace = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
  lp.bounds = nullptr;

  int grid = 1;
  choose_grid(p->epilogue, n_tiles, units_per_tile, tiled, lp.splits, &grid, &lp.whole_per_cta, &lp.units_total);
  // optional hint: the launch that follows on this stream (see LinParams::pf_w)
  lp.pf_w = nullptr;
  if (const pc_linear_params* nx = p->prefetch_next) {
    if (nx->w_tiled != nullptr && nx->k > 0 && nx->k % BKK == 0 && nx->n_rows > 0) {
      int nt = 0;
      switch (nx->epilogue) {
        case PC_LIN_SILU_MUL: nt = (nx->n_rows / 2 + 63) / 64; break;
        default: nt = (nx->n_rows + BM - 1) / BM; break;
      }
      const int upt = tiled_k_blocks(nx->k) / 2;
      const int nsplits = nx->epilogue == PC_LIN_PARTIALS ? pc_linear_partial_splits(nx->n_rows, nx->k) : 0;
      int ngrid = 1, nwhole = 0;
      long long nunits = static_cast<long long>(nt) * upt;
      choose_grid(nx->epilogue, nt, upt, true, nsplits, &ngrid, &nwhole, &nunits);
      lp.pf_w = static_cast<const unsigned char*>(nx->w_tiled);
      lp.pf_units_total = nunits;
      lp.pf_kb = upt;
      lp.pf_grid = ngrid;
      lp.pf_whole = nwhole;
      lp.pf_splits = nsplits;
      lp.pf_units = kPrefetchUnits;
    }
  }
  if (const char* e = dbg_env("PC_LIN_BOUNDS_PTR"))  // experiment: weighted stream-K boundaries, only for full-grid non-split launches
    if (p->epilogue != PC_LIN_PARTIALS && grid == sm_count()) lp.bounds = reinterpret_cast<const long long*>(strtoull(e, nullptr, 0));
  lp.w_one_box = (lp.second_half_off == 64) ? 1 : 0;
  if (const char* e = dbg_env("PC_LIN_ONEBOX")) lp.w_one_box = lp.w_one_box && atoi(e) != 0;
  lp.grid = grid;
  CUtensorMap& mw = out->map_w;
  CUtensorMap& mx = out->map_x;
  if (int rc = make_tensor_map_3d(&mw, p->w, p->k, p->n_rows, 1, static_cast<long long>(p->n_rows) * p->k, lp.w_one_box ? 128 : 64))
    return rc;
  const int npad = p->q_len <= 16 ? 16 : 32;
  // QKV projection on whole head tiles (96 for Llama-2-7B) leaves a third of the SMs idle, and sharing tiles stream-K fashion
  // costs a global reduction in front of the RoPE epilogue.  Clusters of 3 owners + 1 helper use 4/3 as many SMs with the
  // exchange in distributed shared memory (ClIter).
  lp.cl_helper_units = 0;
  lp.cl_size = 0;
#ifdef PC_LIN_NO_CL
  allow_cluster = false;
#endif
  if (allow_cluster && tiled && npad == 16 && p->epilogue == PC_LIN_ROPE_APPEND && lp.whole_per_cta == 0 && grid == n_tiles &&
      p->n_rows % BM == 0 && units_per_tile >= 8) {
    int cl_force = 0;
    if (const char* e = dbg_env("PC_LIN_CL_SIZE")) cl_force = atoi(e);
    // 3 owners + helper (4/3 as many SMs).  2 owners + helper (48 clusters of 3 = 144 SMs) was measured SLOWER than no clusters
    // at all (C2 step 3.36 vs 3.31 ms; clusters of 4: 3.25-3.27): kept selectable in debug builds only.
    for (int S = 4; S >= 3; --S) {
      if (cl_force == 0 ? S != 4 : cl_force != S) continue;
      if (n_tiles % (S - 1) != 0) continue;
      const int n_clusters = n_tiles / (S - 1);
      if (n_clusters * S > sm_count() || !cl_clusters_fit<16>(S, n_clusters)) continue;
      // the helper streams KH units of each of its S - 1 tiles, an owner KB - KH: balance (S - 1) * KH against KB - KH
      int KH = units_per_tile / S;
      lp.cl_helper_units = KH;
      lp.cl_size = S;
      grid = n_clusters * S;
      lp.grid = grid;
      break;
    }
  }
  // rows >= q_len of the X box read as zeros
  if (int rc = make_tensor_map_3d(&mx, p->x, p->k, p->q_len, 1, static_cast<long long>(p->q_len) * p->x_stride, npad)) return rc;
  out->grid = grid;
  out->npad = npad;
  out->epilogue = p->epilogue;
  out->tiled = tiled;
  return PC_OK;
}

PC_API int pc_linear_small(const pc_linear_params* p, void* stream) {
  using namespace pc;
  using namespace pc::lin;
  if (!p) return set_error(PC_ERR_INVALID, "pc_linear_small: null params");
  if (p->q_len == 0) return PC_OK;
  LinLaunch L;
  if (int rc = prepare_linear(p, &L, true)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (L.lp.cl_helper_units > 0) return launch_lin_cl<16>(L.map_w, L.map_x, L.lp, L.grid, st);
  if (L.tiled)
    return L.npad == 16 ? launch_lin_n<16, true>(L.epilogue, L.map_w, L.map_x, L.lp, L.grid, st)
                        : launch_lin_n<32, true>(L.epilogue, L.map_w, L.map_x, L.lp, L.grid, st);
  return L.npad == 16 ? launch_lin_n<16, false>(L.epilogue, L.map_w, L.map_x, L.lp, L.grid, st)
                      : launch_lin_n<32, false>(L.epilogue, L.map_w, L.map_x, L.lp, L.grid, st);
}
