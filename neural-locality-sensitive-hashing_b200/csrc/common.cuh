// Shared device/host helpers for libnlsh_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nlsh_b200.h"

#define NLSH_FULL_MASK 0xffffffffu
#define NLSH_ID_SENTINEL 0x7fffffff

// ---- error plumbing (api.cu) -----------------------------------------------------------
void nlsh_set_error(const char* fmt, ...);
int nlsh_check_cuda(cudaError_t e, const char* what);
// cudaGetLastError() after a kernel launch; also counts the launch (nlsh_kernel_launch_count)
cudaError_t nlsh_post_launch();
// optional CUDA-event bracket around the scan kernel (nlsh_profile_*), api.cu
void nlsh_profile_mark(cudaStream_t st, bool begin);

#define NLSH_CUDA_TRY(expr)                                  \
  do {                                                       \
    int _rc = nlsh_check_cuda((expr), #expr);                \
    if (_rc != NLSH_OK) return _rc;                          \
  } while (0)

#define NLSH_REQUIRE(cond, ...)                              \
  do {                                                       \
    if (!(cond)) {                                           \
      nlsh_set_error(__VA_ARGS__);                           \
      return NLSH_ERR_INVALID;                               \
    }                                                        \
  } while (0)

static inline size_t nlsh_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carves 256-byte aligned regions out of a caller-supplied workspace.
struct WorkspaceCarver {
  char* base;
  size_t used;
  explicit WorkspaceCarver(void* p) : base(reinterpret_cast<char*>(p)), used(0) {}
  template <typename T>
  T* take(size_t count) {
    used = nlsh_align_up(used, 256);
    T* r = base ? reinterpret_cast<T*>(base + used) : nullptr;
    used += count * sizeof(T);
    return r;
  }
  size_t total() const { return nlsh_align_up(used, 256); }
};

int nlsh_num_sms();

// ---- device helpers --------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// (distance, id) lexicographic order: the total order every top-k in this library uses, so
// results do not depend on the order candidates are visited in.
template <typename IdT>
__device__ __forceinline__ bool lex_less(float d1, IdT i1, float d2, IdT i2) {
  return (d1 < d2) || (d1 == d2 && i1 < i2);
}

// Warp-resident sorted list of the 32*KPL smallest (distance, id) pairs seen so far.
// Position pos lives in lane (pos % 32), slot (pos / 32).  All lanes call every method
// with warp-uniform arguments.  tau / tau_id is the element at position k-1 (the current
// k-th best): candidates are filtered against it before insert() is attempted.
template <int KPL, typename IdT>
struct WarpTopK {
  float d[KPL];
  IdT id[KPL];
  float tau;
  IdT tau_id;

  __device__ __forceinline__ void init(IdT sentinel) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      d[j] = __int_as_float(0x7f800000);  // +inf
      id[j] = sentinel;
    }
    tau = __int_as_float(0x7f800000);
    tau_id = sentinel;
  }

  __device__ __forceinline__ void refresh_tau(int k) {
    const int slot = (k - 1) >> 5;
    const int src = (k - 1) & 31;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      if (j == slot) {  // warp-uniform
        tau = __shfl_sync(NLSH_FULL_MASK, d[j], src);
        tau_id = __shfl_sync(NLSH_FULL_MASK, id[j], src);
      }
    }
  }

  // Insert (nd, nid); the caller has checked lex_less(nd, nid, tau, tau_id).
  __device__ __forceinline__ void insert(float nd, IdT nid, int k) {
    const int lane = lane_id();
    int p = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j)
      p += __popc(__ballot_sync(NLSH_FULL_MASK, lex_less<IdT>(d[j], id[j], nd, nid)));
#pragma unroll
    for (int j = KPL - 1; j >= 0; --j) {
      float sd = __shfl_up_sync(NLSH_FULL_MASK, d[j], 1);
      IdT si = __shfl_up_sync(NLSH_FULL_MASK, id[j], 1);
      if (j > 0) {
        const float wd = __shfl_sync(NLSH_FULL_MASK, d[j - 1], 31);
        const IdT wi = __shfl_sync(NLSH_FULL_MASK, id[j - 1], 31);
        if (lane == 0) {
          sd = wd;
          si = wi;
        }
      }
      const int pos = j * 32 + lane;
      if (pos > p) {
        d[j] = sd;
        id[j] = si;
      } else if (pos == p) {
        d[j] = nd;
        id[j] = nid;
      }
    }
    refresh_tau(k);
  }

  // Offer one candidate per lane (valid = this lane has one).
  __device__ __forceinline__ void offer(float cd, IdT cid, bool valid, int k) {
    unsigned m = __ballot_sync(NLSH_FULL_MASK, valid && lex_less<IdT>(cd, cid, tau, tau_id));
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float nd = __shfl_sync(NLSH_FULL_MASK, cd, src);
      const IdT nid = __shfl_sync(NLSH_FULL_MASK, cid, src);
      if (lex_less<IdT>(nd, nid, tau, tau_id)) insert(nd, nid, k);
    }
  }

  // Replace the (empty) list by the sorted set of one candidate per lane: a 32-wide bitonic
  // sort over the lanes costs about four insert()s, where feeding the same 32 candidates
  // through offer() into an empty list costs k + k ln(32 / k) of them.
  __device__ __forceinline__ void seed32(float cd, IdT cid, bool valid, IdT sentinel, int k) {
    const int lane = lane_id();
    float sd = valid ? cd : __int_as_float(0x7f800000);
    IdT si = valid ? cid : sentinel;
    if (sd != sd) {  // NaN never enters a list
      sd = __int_as_float(0x7f800000);
      si = sentinel;
    }
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const float od = __shfl_xor_sync(NLSH_FULL_MASK, sd, stride);
        const IdT oi = __shfl_xor_sync(NLSH_FULL_MASK, si, stride);
        const bool up = (lane & size) == 0;          // ascending block?
        const bool lower = (lane & stride) == 0;     // this lane holds the lower index of the pair
        const bool other_less = lex_less<IdT>(od, oi, sd, si);
        // the lower lane of an ascending pair keeps the minimum, etc.
        if (other_less == (up == lower)) {
          sd = od;
          si = oi;
        }
      }
    }
    d[0] = sd;
    id[0] = si;
#pragma unroll
    for (int j = 1; j < KPL; ++j) {
      d[j] = __int_as_float(0x7f800000);
      id[j] = sentinel;
    }
    refresh_tau(k);
  }

  // Element at position pos (warp-uniform pos) broadcast to all lanes.
  __device__ __forceinline__ void get(int pos, float& od, IdT& oi) const {
    const int slot = pos >> 5;
    const int src = pos & 31;
    od = 0.f;
    oi = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      if (j == slot) {
        od = __shfl_sync(NLSH_FULL_MASK, d[j], src);
        oi = __shfl_sync(NLSH_FULL_MASK, id[j], src);
      }
    }
  }
};

// ---- mbarrier / bulk-async copy (TMA 1D) PTX wrappers ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware
      : "memory");
  return ok != 0;
}
// try_wait without a suspend-time hint: returns after the hardware's default (short) time slice
__device__ __forceinline__ bool mbar_try_wait_short(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t kMbarWatchdogNs = 30000000000ull;  // 30 s
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait suspends the thread in hardware (up to the hint) and returns early when the phase
  // completes.  A wait that lasts kMbarWatchdogNs means the pipeline is broken: trap (-> CUDA error on
  // the host) instead of hanging the device.  The bound is wall time (%globaltimer keeps running while
  // the context is time-sliced out or replayed under a profiler), hence generous; -DNLSH_NO_WATCHDOG
  // builds wait forever.
  if (mbar_try_wait(bar, parity)) return;
#ifdef NLSH_NO_WATCHDOG
  while (!mbar_try_wait(bar, parity)) {
  }
#else
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_timer_ns() - t0 > kMbarWatchdogNs) __trap();
  }
#endif
}
// Polling variant for deep pipelines (scan_tc.cu: five roles hand tiles to each other, so the
// wake-up latency of a long suspended wait would be paid several times per tile).
__device__ __forceinline__ void mbar_wait_poll(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_short(bar, parity)) return;
#ifdef NLSH_NO_WATCHDOG
  while (!mbar_try_wait_short(bar, parity)) {
  }
#else
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait_short(bar, parity)) {
    if (global_timer_ns() - t0 > kMbarWatchdogNs) __trap();
  }
#endif
}
// global -> shared bulk copy (SASS UBLKCP), completion counted in bytes on `bar`.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// 2-D TMA tensor copy global -> shared (SASS UTMALDG): box at (c0 = column, c1 = row) of the
// tensor described by `tmap`, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

#endif  // __CUDACC__
