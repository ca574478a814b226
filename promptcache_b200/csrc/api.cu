// C-ABI glue: error state, device info, attention dispatch.  See include/promptcache_b200.h.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>

#include "pc_common.cuh"

namespace pc {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = dbg_env("PC_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

int launch_attn_simt(const pc_attn_params& p, cudaStream_t st);
size_t attn_tc_workspace_bytes(int n_heads, int q_len, int head_dim);
int launch_attn_small(const pc_attn_params& p, cudaStream_t st);
int launch_attn_pp(const pc_attn_params& p, cudaStream_t st);
bool attn_small_supported(const pc_attn_params& p);
size_t attn_small_partial_bytes(int n_heads, int q_len, int head_dim);

static int check_sm100() {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      return set_error(PC_ERR_CUDA, "no CUDA device");
    ok = (major == 10) ? 1 : 0;
  }
  if (!ok) return set_error(PC_ERR_UNSUPPORTED, "this library is built for sm_100a only");
  return PC_OK;
}

static int attn_dispatch(const pc_attn_params* p, void* stream, const char* who) {
  if (!p) return set_error(PC_ERR_INVALID, "%s: null params", who);
  if (p->q_len == 0) return PC_OK;
  PC_REQUIRE(p->q && p->k && p->v && p->out, "%s: null tensor", who);
  PC_REQUIRE(p->n_heads > 0 && p->n_kv_heads > 0 && p->n_heads % p->n_kv_heads == 0,
             "%s: bad head counts H=%d Hkv=%d", who, p->n_heads, p->n_kv_heads);
  PC_REQUIRE(p->q_len > 0 && p->n_cached >= 0, "%s: bad lengths N=%d q=%d", who, p->n_cached, p->q_len);
  PC_REQUIRE(static_cast<long long>(p->n_cached) + p->q_len <= p->max_ctx,
             "%s: N+q=%lld exceeds max_ctx=%d", who, static_cast<long long>(p->n_cached) + p->q_len, p->max_ctx);
  PC_REQUIRE((p->alibi_slopes == nullptr) || (p->key_pos != nullptr), "%s: alibi needs key_pos", who);
  const uintptr_t align = reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->k) |
                          reinterpret_cast<uintptr_t>(p->v) | reinterpret_cast<uintptr_t>(p->out);
  PC_REQUIRE(align % 16 == 0, "%s: q/k/v/out must be 16B aligned", who);
  if (int rc = check_sm100()) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int impl = p->impl & PC_ATTN_IMPL_MASK;
  if (impl == PC_ATTN_AUTO) {
    const char* e = dbg_env("PC_ATTN_IMPL");
    if (e && e[0] == 's') impl = PC_ATTN_SIMT;
    else impl = PC_ATTN_TCGEN05;
  }
  if (p->kv_tiles != nullptr && p->kv_n_tiles > 0) {
    if (impl != PC_ATTN_TCGEN05 || !attn_small_supported(*p))
      return set_error(PC_ERR_UNSUPPORTED, "%s: the virtual splice (kv_tiles) needs the short-suffix tcgen05 kernel "
                       "(head_dim 128, G*q <= 32); materialise the cache with pc_splice_kv instead", who);
    return launch_attn_small(*p, st);
  }
  if (p->n_cached_dev != nullptr) {
    if (impl != PC_ATTN_TCGEN05) return set_error(PC_ERR_UNSUPPORTED, "%s: n_cached_dev needs the tcgen05 kernels", who);
    return attn_small_supported(*p) ? launch_attn_small(*p, st) : launch_attn_pp(*p, st);
  }
  if (impl == PC_ATTN_SIMT) return launch_attn_simt(*p, st);
  if (impl == PC_ATTN_TCGEN05) {
    const char* e = dbg_env("PC_ATTN_SMALL");
    if (attn_small_supported(*p) && !(e && e[0] == '0')) return launch_attn_small(*p, st);
    return launch_attn_pp(*p, st);
  }
  return set_error(PC_ERR_INVALID, "%s: unknown impl %d", who, p->impl);
}

}  // namespace pc

PC_API int pc_version(void) { return PC_ABI_VERSION; }

PC_API const char* pc_last_error(void) { return pc::g_err; }

PC_API int pc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  PC_CHECK_CUDA(cudaGetDevice(&dev));
  int v = 0;
  if (sm_count) {
    PC_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    *sm_count = v;
  }
  if (cc_major) {
    PC_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    *cc_major = v;
  }
  if (cc_minor) {
    PC_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    *cc_minor = v;
  }
  return PC_OK;
}

PC_API size_t pc_attn_workspace_bytes(int n_heads, int q_len, int head_dim) {
  if (n_heads <= 0 || q_len <= 0 || head_dim <= 0) return 0;
  const size_t a = pc::attn_tc_workspace_bytes(n_heads, q_len, head_dim);
  const size_t b = pc::attn_small_partial_bytes(n_heads, q_len, head_dim);
  return pc::kAttnCounterBytes + (a > b ? a : b);
}

PC_API int pc_prefill_attn(const pc_attn_params* p, void* stream) {
  return pc::attn_dispatch(p, stream, "pc_prefill_attn");
}

PC_API int pc_decode_attn(const pc_attn_params* p, void* stream) {
  if (p && p->q_len != 1) return pc::set_error(PC_ERR_INVALID, "pc_decode_attn: q_len must be 1 (got %d)", p->q_len);
  return pc::attn_dispatch(p, stream, "pc_decode_attn");
}

PC_API int64_t pc_launch_count(void) { return pc::g_launches.load(std::memory_order_relaxed); }
