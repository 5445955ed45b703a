// Library-wide plumbing of libnlsh_b200: version, thread-local last error, the host-side
// bit-pack twin of the reference's Cython helper, device properties cache.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_last_error[512] = "";

void nlsh_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int nlsh_check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return NLSH_OK;
  nlsh_set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return NLSH_ERR_CUDA;
}

int nlsh_num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}

static std::atomic<long long> g_launches{0};

cudaError_t nlsh_post_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

extern "C" long long nlsh_kernel_launch_count(void) { return g_launches.load(); }

// ---- scan-kernel timing: a ring of CUDA event pairs recorded on the launching stream ------
namespace {
constexpr int kProfileSlots = 256;
struct ProfileRing {
  cudaEvent_t begin[kProfileSlots];
  cudaEvent_t end[kProfileSlots];
  int created = 0;
  int count = 0;
  bool enabled = false;
};
thread_local ProfileRing g_prof;
}  // namespace

void nlsh_profile_mark(cudaStream_t st, bool begin) {
  if (!g_prof.enabled || g_prof.count >= kProfileSlots) return;
  if (g_prof.created <= g_prof.count) {
    cudaEventCreate(&g_prof.begin[g_prof.count]);
    cudaEventCreate(&g_prof.end[g_prof.count]);
    g_prof.created = g_prof.count + 1;
  }
  if (begin) {
    cudaEventRecord(g_prof.begin[g_prof.count], st);
  } else {
    cudaEventRecord(g_prof.end[g_prof.count], st);
    ++g_prof.count;
  }
}

extern "C" int nlsh_profile_enable(int on) {
  g_prof.enabled = on != 0;
  g_prof.count = 0;
  return NLSH_OK;
}

extern "C" int nlsh_profile_read(float* ms_out, int capacity) {
  int n = g_prof.count < capacity ? g_prof.count : capacity;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(g_prof.end[i]) != cudaSuccess) return NLSH_ERR_CUDA;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.begin[i], g_prof.end[i]) != cudaSuccess) return NLSH_ERR_CUDA;
    ms_out[i] = ms;
  }
  g_prof.count = 0;
  return n;
}

extern "C" int nlsh_version(void) { return NLSH_B200_VERSION; }

extern "C" const char* nlsh_last_error(void) { return g_last_error; }

// nlsh/utils.pyx:6-15: `out = (out << 1) | bit` over n_bits in an int32, returned as int16.
extern "C" int nlsh_pack_codes_host(const int32_t* bits, int64_t n, int64_t s, int64_t hs,
                                    int64_t stride_n, int64_t stride_s, int64_t stride_b,
                                    int16_t* out) {
  NLSH_REQUIRE(n >= 0 && s >= 0 && hs >= 0, "pack_codes: negative shape (%lld, %lld, %lld)",
               (long long)n, (long long)s, (long long)hs);
  NLSH_REQUIRE((bits != nullptr && out != nullptr) || n * s == 0, "pack_codes: null pointer");
  for (int64_t i = 0; i < n; ++i) {
    for (int64_t j = 0; j < s; ++j) {
      const int32_t* row = bits + i * stride_n + j * stride_s;
      uint32_t acc = 0;  // unsigned shift: same bits as the int32 accumulator, no UB
      for (int64_t b = 0; b < hs; ++b) acc = (acc << 1) | (uint32_t)row[b * stride_b];
      out[i * s + j] = (int16_t)(uint16_t)(acc & 0xffffu);
    }
  }
  return NLSH_OK;
}
