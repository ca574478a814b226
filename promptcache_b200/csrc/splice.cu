// Splice kernel: PromptCache.update (reference promptcache/cache_engine.py:116-156).
//
// The splice is L*2*Hkv*M independent contiguous runs (SURVEY.md section 7 "Splice geometry").
// Each run is cut into <=32 KiB tiles; a persistent grid (one CTA per SM) walks the tiles
// round-robin.  A tile moves global -> shared with one 1-D TMA bulk load (UBLKCP) and shared ->
// global with one 1-D TMA bulk store: 128 B-coalesced on both sides, no register staging, and a
// ring of kStages tiles per SM keeps kAhead loads plus (kStages-kAhead) stores in flight
// (~96 KiB of reads in flight per SM, above the ~44 KiB Little's-law requirement at 6.5 TB/s).
// Bound: HBM.  Algorithmic bytes = 2 * payload (read + write).
#include <map>
#include <mutex>
#include <vector>

#include "pc_common.cuh"

namespace {

constexpr int kTileBytes = 32768;
constexpr int kStages = 6;
constexpr int kAhead = 3;
constexpr int kMaxInlineDescs = 48;

struct RunDesc {              // one 2-D copy: n_runs runs of run_bytes
  const char* src;
  char* dst;
  long long src_run_stride;
  long long dst_run_stride;
  long long run_bytes;
  long long tile_begin;       // first global tile index of this descriptor
  int tiles_per_run;
  int n_runs;
};
static_assert(sizeof(RunDesc) == 56, "RunDesc layout");

struct InlineTable {
  long long total_tiles;
  int n_descs;
  int pad;
  RunDesc d[kMaxInlineDescs];
};

struct TileRef {
  const char* src;
  char* dst;
  uint32_t bytes;
};

__device__ __forceinline__ TileRef locate(const RunDesc* __restrict__ descs, int n_descs, int& cursor,
                                          long long tile) {
  while (cursor + 1 < n_descs && descs[cursor + 1].tile_begin <= tile) ++cursor;
  const RunDesc& d = descs[cursor];
  const long long local = tile - d.tile_begin;
  const long long run = local / d.tiles_per_run;
  const long long off = (local - run * d.tiles_per_run) * static_cast<long long>(kTileBytes);
  const long long left = d.run_bytes - off;
  TileRef r;
  r.src = d.src + run * d.src_run_stride + off;
  r.dst = d.dst + run * d.dst_run_stride + off;
  r.bytes = static_cast<uint32_t>(left < kTileBytes ? left : kTileBytes);
  return r;
}

template <bool kInline>
__global__ void __launch_bounds__(32, 1)
splice_kernel(const __grid_constant__ InlineTable tab, const RunDesc* __restrict__ gdescs,
              int g_n_descs, long long g_total_tiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kTileBytes);
  char** dst_ring = reinterpret_cast<char**>(full + kStages);
  uint32_t* bytes_ring = reinterpret_cast<uint32_t*>(dst_ring + kStages);

  if (threadIdx.x != 0) return;  // one elected thread drives the TMA engine

  const RunDesc* descs = kInline ? tab.d : gdescs;
  const int n_descs = kInline ? tab.n_descs : g_n_descs;
  const long long total = kInline ? tab.total_tiles : g_total_tiles;

  for (int s = 0; s < kStages; ++s) pc::mbar_init(&full[s], 1);
  pc::mbar_fence_init();

  const long long first = blockIdx.x;
  const long long step = gridDim.x;
  if (first >= total) return;
  const long long n_mine = (total - first + step - 1) / step;

  int cursor = 0;
  auto issue_load = [&](long long i) {
    const int s = static_cast<int>(i % kStages);
    const TileRef t = locate(descs, n_descs, cursor, first + i * step);
    dst_ring[s] = t.dst;
    bytes_ring[s] = t.bytes;
    pc::mbar_expect_tx(&full[s], t.bytes);
    pc::bulk_g2s(smem + s * kTileBytes, t.src, t.bytes, &full[s]);
  };

  const long long pro = n_mine < kAhead ? n_mine : kAhead;
  for (long long i = 0; i < pro; ++i) issue_load(i);

  for (long long k = 0; k < n_mine; ++k) {
    const int s = static_cast<int>(k % kStages);
    pc::mbar_wait(&full[s], static_cast<uint32_t>((k / kStages) & 1));
    pc::bulk_s2g(dst_ring[s], smem + s * kTileBytes, bytes_ring[s]);
    pc::bulk_commit();
    const long long j = k + kAhead;
    if (j < n_mine) {
      // stage j%kStages was last used by tile j-kStages; its store must have drained smem.
      pc::bulk_wait_read<kStages - kAhead>();
      issue_load(j);
    }
  }
  pc::bulk_wait_all<0>();
}

constexpr size_t kSmemBytes = kStages * kTileBytes + kStages * (8 + 8 + 4) + 64;

int configure() {
  static std::once_flag once;
  static cudaError_t err = cudaSuccess;
  std::call_once(once, [] {
    err = cudaFuncSetAttribute(splice_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(kSmemBytes));
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(splice_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(kSmemBytes));
  });
  if (err != cudaSuccess)
    return pc::set_error(PC_ERR_CUDA, "cudaFuncSetAttribute(splice): %s", cudaGetErrorString(err));
  return PC_OK;
}

int grid_for(long long total_tiles) {
  const long long sms = pc::sm_count();
  return static_cast<int>(total_tiles < sms ? total_tiles : sms);
}

// Fill tile bookkeeping; returns total tiles.
long long finish_descs(RunDesc* d, int n) {
  long long t = 0;
  for (int i = 0; i < n; ++i) {
    d[i].tile_begin = t;
    d[i].tiles_per_run = static_cast<int>((d[i].run_bytes + kTileBytes - 1) / kTileBytes);
    t += static_cast<long long>(d[i].tiles_per_run) * d[i].n_runs;
  }
  return t;
}

// Device-side descriptor table for pc_splice_kv_tensors: one per stream (grown on demand), so calls on different
// streams never wait on each other; a call only waits for the previous table of ITS OWN stream to be consumed.
struct Staging {
  RunDesc* host = nullptr;  // pinned
  RunDesc* dev = nullptr;
  size_t cap = 0;
  cudaEvent_t done = nullptr;
  bool pending = false;
};
std::mutex g_stage_mu;  // guards the map only
std::map<cudaStream_t, Staging> g_stages;

}  // namespace

static int splice_kv_impl(const pc_segment* segs, const int64_t* src_run_rows, int n_segs, void* cache_dev, int n_layers,
                          int n_kv_heads, int head_dim, int max_ctx, int elem_bytes, void* stream) {
  PC_REQUIRE(n_segs >= 0, "pc_splice_kv: n_segs < 0");
  if (n_segs == 0) return PC_OK;
  PC_REQUIRE(segs && cache_dev, "pc_splice_kv: null pointer");
  PC_REQUIRE(n_layers > 0 && n_kv_heads > 0 && head_dim > 0 && max_ctx > 0 && elem_bytes > 0,
             "pc_splice_kv: non-positive dimension");
  const long long row_bytes = static_cast<long long>(head_dim) * elem_bytes;
  PC_REQUIRE(row_bytes % 16 == 0, "pc_splice_kv: head_dim*elem_bytes=%lld is not a multiple of 16",
             row_bytes);
  PC_REQUIRE(reinterpret_cast<uintptr_t>(cache_dev) % 16 == 0, "pc_splice_kv: cache not 16B aligned");
  for (int i = 0; i < n_segs; ++i) {
    PC_REQUIRE(segs[i].base != nullptr, "pc_splice_kv: segment %d has a null base", i);
    PC_REQUIRE(reinterpret_cast<uintptr_t>(segs[i].base) % 16 == 0,
               "pc_splice_kv: segment %d not 16B aligned", i);
    PC_REQUIRE(segs[i].len > 0, "pc_splice_kv: segment %d has len %d", i, segs[i].len);
    PC_REQUIRE(segs[i].dst_row >= 0 &&
                   static_cast<long long>(segs[i].dst_row) + segs[i].len <= max_ctx,
               "pc_splice_kv: segment %d rows [%d,%d) exceed max_ctx %d", i, segs[i].dst_row,
               segs[i].dst_row + segs[i].len, max_ctx);
  }
  if (int rc = configure()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_runs = n_layers * 2 * n_kv_heads;
  for (int base = 0; base < n_segs; base += kMaxInlineDescs) {
    InlineTable tab;
    tab.n_descs = n_segs - base < kMaxInlineDescs ? n_segs - base : kMaxInlineDescs;
    tab.pad = 0;
    for (int i = 0; i < tab.n_descs; ++i) {
      const pc_segment& s = segs[base + i];
      RunDesc& d = tab.d[i];
      d.src = static_cast<const char*>(s.base);
      d.dst = static_cast<char*>(cache_dev) + static_cast<long long>(s.dst_row) * row_bytes;
      d.run_bytes = static_cast<long long>(s.len) * row_bytes;
      d.src_run_stride = src_run_rows ? static_cast<long long>(src_run_rows[base + i]) * row_bytes : d.run_bytes;
      d.dst_run_stride = static_cast<long long>(max_ctx) * row_bytes;
      d.n_runs = n_runs;
    }
    tab.total_tiles = finish_descs(tab.d, tab.n_descs);
    splice_kernel<true><<<grid_for(tab.total_tiles), 32, kSmemBytes, st>>>(tab, nullptr, 0, 0);
    PC_CHECK_CUDA(cudaGetLastError());
    pc::count_launch();
  }
  return PC_OK;
}

PC_API int pc_splice_kv(const pc_segment* segs, int n_segs, void* cache_dev, int n_layers, int n_kv_heads, int head_dim,
                        int max_ctx, int elem_bytes, void* stream) {
  return splice_kv_impl(segs, nullptr, n_segs, cache_dev, n_layers, n_kv_heads, head_dim, max_ctx, elem_bytes, stream);
}

PC_API int pc_splice_kv_strided(const pc_segment* segs, const int64_t* src_run_rows, int n_segs, void* cache_dev, int n_layers,
                                int n_kv_heads, int head_dim, int max_ctx, int elem_bytes, void* stream) {
  PC_REQUIRE(n_segs <= 0 || src_run_rows != nullptr, "pc_splice_kv_strided: null src_run_rows");
  for (int i = 0; i < n_segs; ++i)
    PC_REQUIRE(src_run_rows[i] >= segs[i].len, "pc_splice_kv_strided: segment %d: run stride %lld rows < len %d", i,
               static_cast<long long>(src_run_rows[i]), segs[i].len);
  return splice_kv_impl(segs, src_run_rows, n_segs, cache_dev, n_layers, n_kv_heads, head_dim, max_ctx, elem_bytes, stream);
}

PC_API int pc_splice_kv_tensors(const pc_copy2d* copies, int n_copies, void* stream) {
  PC_REQUIRE(n_copies >= 0, "pc_splice_kv_tensors: n_copies < 0");
  if (n_copies == 0) return PC_OK;
  PC_REQUIRE(copies != nullptr, "pc_splice_kv_tensors: null table");
  for (int i = 0; i < n_copies; ++i) {
    const pc_copy2d& c = copies[i];
    PC_REQUIRE(c.src && c.dst, "pc_splice_kv_tensors: copy %d has a null pointer", i);
    PC_REQUIRE(c.n_runs > 0 && c.run_bytes > 0, "pc_splice_kv_tensors: copy %d is empty", i);
    PC_REQUIRE(c.run_bytes % 16 == 0 && c.src_run_stride % 16 == 0 && c.dst_run_stride % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(c.src) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(c.dst) % 16 == 0,
               "pc_splice_kv_tensors: copy %d is not 16B aligned", i);
  }
  if (int rc = configure()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Staging* stage_ptr;
  {
    std::lock_guard<std::mutex> lock(g_stage_mu);
    stage_ptr = &g_stages[st];  // std::map nodes are address-stable
  }
  Staging& g_stage = *stage_ptr;
  if (g_stage.pending) {  // the previous table of this stream may still be in flight
    PC_CHECK_CUDA(cudaEventSynchronize(g_stage.done));
    g_stage.pending = false;
  }
  if (static_cast<size_t>(n_copies) > g_stage.cap) {
    if (g_stage.host) cudaFreeHost(g_stage.host);
    if (g_stage.dev) cudaFree(g_stage.dev);
    g_stage.host = nullptr;
    g_stage.dev = nullptr;
    g_stage.cap = 0;
    size_t cap = 1024;
    while (cap < static_cast<size_t>(n_copies)) cap *= 2;
    PC_CHECK_CUDA(cudaMallocHost(&g_stage.host, cap * sizeof(RunDesc)));
    PC_CHECK_CUDA(cudaMalloc(&g_stage.dev, cap * sizeof(RunDesc)));
    g_stage.cap = cap;
  }
  if (!g_stage.done) PC_CHECK_CUDA(cudaEventCreateWithFlags(&g_stage.done, cudaEventDisableTiming));
  for (int i = 0; i < n_copies; ++i) {
    RunDesc& d = g_stage.host[i];
    d.src = static_cast<const char*>(copies[i].src);
    d.dst = static_cast<char*>(copies[i].dst);
    d.src_run_stride = copies[i].src_run_stride;
    d.dst_run_stride = copies[i].dst_run_stride;
    d.run_bytes = copies[i].run_bytes;
    d.n_runs = copies[i].n_runs;
  }
  const long long total = finish_descs(g_stage.host, n_copies);
  PC_CHECK_CUDA(cudaMemcpyAsync(g_stage.dev, g_stage.host, n_copies * sizeof(RunDesc),
                                cudaMemcpyHostToDevice, st));
  InlineTable dummy;
  dummy.total_tiles = 0;
  dummy.n_descs = 0;
  dummy.pad = 0;
  splice_kernel<false><<<grid_for(total), 32, kSmemBytes, st>>>(dummy, g_stage.dev, n_copies, total);
  PC_CHECK_CUDA(cudaGetLastError());
  PC_CHECK_CUDA(cudaEventRecord(g_stage.done, st));
  g_stage.pending = true;
  pc::count_launch();
  return PC_OK;
}
