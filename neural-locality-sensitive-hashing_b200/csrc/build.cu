// Index build: bucket codes -> CSR (offsets, ids ascending inside a bucket) + a
// bucket-contiguous copy of the vectors.  Replaces build_index (nlsh/indexer.py:6-24) as
// Indexer._build_index calls it (one code per row, indexer.py:36-38).
//
// Stable counting sort in four passes, all HBM/L2-atomic bound:
//   1. unit histograms   hist[bucket][unit]   (unit = one warp's contiguous row range;
//                                              warp-aggregated atomics via match.any)
//   2. exclusive scan of hist in memory order  (bucket-major => cursor[bucket][unit] is the
//                                              first output slot of that unit in that bucket)
//   3. stable scatter    ids[cursor++] = row   (a warp walks its rows in order, 32 at a time,
//                                              rank inside the step from the match mask)
//   4. row gather        x_sorted[i] = x[ids[i]]  (one warp per row, 128-bit lanes)
// Ascending row order inside a bucket (pinned by nlsh/tests/test_indexer.py:14-26) follows
// from units being ordered row ranges and every unit emitting its rows in order.
#include "common.cuh"

namespace {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096 entries per block
constexpr int64_t kMaxHistEntries = 1ll << 24;        // 64 MB of int32 counters
constexpr int kMinRowsPerUnit = 256;

struct BuildPlan {
  int64_t rows_per_unit;
  int64_t n_units;
  int64_t entries;         // n_buckets * n_units
  int64_t entries_padded;  // multiple of kScanTile
  int64_t n_scan_blocks;
};

BuildPlan build_plan(int64_t n, int32_t n_buckets) {
  BuildPlan p;
  int64_t max_units = kMaxHistEntries / (n_buckets > 0 ? n_buckets : 1);
  if (max_units < 1) max_units = 1;
  int64_t rpu = (n + max_units - 1) / max_units;
  if (rpu < kMinRowsPerUnit) rpu = kMinRowsPerUnit;
  rpu = (rpu + 31) / 32 * 32;
  p.rows_per_unit = rpu;
  p.n_units = n > 0 ? (n + rpu - 1) / rpu : 1;
  p.entries = (int64_t)n_buckets * p.n_units;
  p.entries_padded = (p.entries + kScanTile - 1) / kScanTile * kScanTile;
  p.n_scan_blocks = p.entries_padded / kScanTile;
  return p;
}

__global__ void __launch_bounds__(128)
    unit_hist_kernel(const int* __restrict__ codes, long long n, int n_buckets, long long n_units,
                     long long rows_per_unit, int* __restrict__ hist) {
  const long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= n_units) return;
  const int lane = lane_id();
  const long long r0 = unit * rows_per_unit;
  long long r1 = r0 + rows_per_unit;
  if (r1 > n) r1 = n;
  for (long long r = r0; r < r1; r += 32) {
    const long long row = r + lane;
    int code = -1;
    if (row < r1) {
      code = codes[row];
      if (code < 0 || code >= n_buckets) code = -1;  // dropped; host sees offsets[B] < n
    }
    const unsigned peers = __match_any_sync(NLSH_FULL_MASK, code);
    if (code >= 0 && lane == __ffs(peers) - 1)
      atomicAdd(&hist[(long long)code * n_units + unit], __popc(peers));
  }
}

__global__ void __launch_bounds__(128)
    unit_scatter_kernel(const int* __restrict__ codes, long long n, int n_buckets,
                        long long n_units, long long rows_per_unit, int* __restrict__ cursor,
                        int* __restrict__ ids_out) {
  const long long unit = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= n_units) return;
  const int lane = lane_id();
  const unsigned lt_mask = (1u << lane) - 1u;
  const long long r0 = unit * rows_per_unit;
  long long r1 = r0 + rows_per_unit;
  if (r1 > n) r1 = n;
  for (long long r = r0; r < r1; r += 32) {
    const long long row = r + lane;
    int code = -1;
    if (row < r1) {
      code = codes[row];
      if (code < 0 || code >= n_buckets) code = -1;
    }
    const unsigned peers = __match_any_sync(NLSH_FULL_MASK, code);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (code >= 0 && lane == leader)
      base = atomicAdd(&cursor[(long long)code * n_units + unit], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    if (code >= 0) ids_out[base + __popc(peers & lt_mask)] = (int)row;
  }
}

// ---- exclusive scan over int32 entries (padded to a multiple of kScanTile) -------------
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp, int& block_total) {
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(NLSH_FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    int w = lane < nw ? smem_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(NLSH_FULL_MASK, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) smem_warp[lane] = winc - w;  // exclusive warp offsets
    if (lane == nw - 1) smem_warp[32] = winc;   // block total
  }
  __syncthreads();
  block_total = smem_warp[32];
  return inc - v + smem_warp[warp];
}

__global__ void __launch_bounds__(kScanThreads)
    scan_reduce_kernel(const int* __restrict__ data, int* __restrict__ block_sums) {
  __shared__ int smem_warp[33];
  const int4* p =
      reinterpret_cast<const int4*>(data + (size_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems);
  const int4 a = p[0], b = p[1];
  int s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(NLSH_FULL_MASK, s, o);
  if (lane_id() == 0) smem_warp[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < (kScanThreads >> 5); ++w) t += smem_warp[w];
    block_sums[blockIdx.x] = t;
  }
}

// One block: exclusive scan of block_sums in place; total written to block_sums[n_blocks].
__global__ void __launch_bounds__(1024)
    scan_sums_kernel(int* __restrict__ block_sums, long long n_blocks) {
  __shared__ int smem_warp[33];
  int carry = 0;
  for (long long base = 0; base < n_blocks; base += blockDim.x) {
    const long long i = base + threadIdx.x;
    const int v = i < n_blocks ? block_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, smem_warp, total);
    if (i < n_blocks) block_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_sums[n_blocks] = carry;
}

__global__ void __launch_bounds__(kScanThreads)
    scan_apply_kernel(int* __restrict__ data, const int* __restrict__ block_sums) {
  __shared__ int smem_warp[33];
  int4* p = reinterpret_cast<int4*>(data + (size_t)blockIdx.x * kScanTile + threadIdx.x * kScanItems);
  int4 a = p[0], b = p[1];
  const int s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  int total;
  int run = block_exclusive_scan(s, smem_warp, total) + block_sums[blockIdx.x];
  int4 oa, ob;
  oa.x = run; run += a.x;
  oa.y = run; run += a.y;
  oa.z = run; run += a.z;
  oa.w = run; run += a.w;
  ob.x = run; run += b.x;
  ob.y = run; run += b.y;
  ob.z = run; run += b.z;
  ob.w = run;
  p[0] = oa;
  p[1] = ob;
}

__global__ void extract_offsets_kernel(const int* __restrict__ cursor, long long n_units,
                                       int n_buckets, const int* __restrict__ total,
                                       int* __restrict__ offsets) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_buckets) offsets[b] = cursor[(long long)b * n_units];
  if (b == n_buckets) offsets[b] = *total;
}

// x_sorted[i, :] = x[ids[i], :] (row stride rounded up to 4 floats); one warp per row, four
// rows in flight per warp.
template <bool VEC>
__global__ void __launch_bounds__(256)
    gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ ids, long long n, int d,
                       float* __restrict__ x_sorted, float* __restrict__ x_sqnorm) {
  constexpr int R = 4;
  const int lane = lane_id();
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long i0 = warp * R; i0 < n; i0 += n_warps * R) {
    long long src[R];
#pragma unroll
    for (int r = 0; r < R; ++r) src[r] = (i0 + r < n) ? (long long)ids[i0 + r] : -1;
    float ss[R];  // |row|^2 (the tensor-core scan's filter needs it; scan_tc.cu)
#pragma unroll
    for (int r = 0; r < R; ++r) ss[r] = 0.f;
    if (VEC) {
      const int nv = d >> 2;
      for (int v0 = 0; v0 < nv; v0 += 32) {
        const int v = v0 + lane;
        float4 val[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (src[r] >= 0 && v < nv)
            val[r] = __ldcs(reinterpret_cast<const float4*>(x + (size_t)src[r] * d) + v);
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (src[r] >= 0 && v < nv) {
            __stcs(reinterpret_cast<float4*>(x_sorted + (size_t)(i0 + r) * d) + v, val[r]);
            ss[r] = fmaf(val[r].x, val[r].x, ss[r]);
            ss[r] = fmaf(val[r].y, val[r].y, ss[r]);
            ss[r] = fmaf(val[r].z, val[r].z, ss[r]);
            ss[r] = fmaf(val[r].w, val[r].w, ss[r]);
          }
      }
    } else {
      const int d_pad = (d + 3) / 4 * 4;  // rows of x_sorted are zero-padded to 16 bytes
      for (int c0 = 0; c0 < d_pad; c0 += 32) {
        const int c = c0 + lane;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (src[r] >= 0 && c < d_pad) {
            const float v = c < d ? x[(size_t)src[r] * d + c] : 0.f;
            x_sorted[(size_t)(i0 + r) * d_pad + c] = v;
            ss[r] = fmaf(v, v, ss[r]);
          }
      }
    }
    if (x_sqnorm != nullptr) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float t = ss[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(NLSH_FULL_MASK, t, o);
        if (lane == 0 && src[r] >= 0) x_sqnorm[i0 + r] = t;
      }
    }
  }
}

}  // namespace

extern "C" size_t nlsh_build_workspace_bytes(int64_t n, int32_t n_buckets) {
  if (n < 0 || n_buckets < 1) return 0;
  const BuildPlan p = build_plan(n, n_buckets);
  WorkspaceCarver ws(nullptr);
  ws.take<int>((size_t)p.entries_padded);
  ws.take<int>((size_t)p.n_scan_blocks + 1);
  return ws.total();
}

extern "C" int nlsh_build_csr(const int32_t* codes, int64_t n, int32_t n_buckets, const float* x,
                              int32_t d, int32_t* offsets_out, int32_t* ids_out,
                              float* x_sorted_out, float* x_sqnorm_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  NLSH_REQUIRE(n >= 0 && n < (1ll << 31), "build: n=%lld outside [0, 2^31)", (long long)n);
  NLSH_REQUIRE(n_buckets >= 1 && n_buckets <= (1 << 20), "build: n_buckets=%d outside [1, 2^20]",
               n_buckets);
  NLSH_REQUIRE(offsets_out != nullptr, "build: offsets_out is NULL");
  NLSH_REQUIRE(n == 0 || (codes != nullptr && ids_out != nullptr), "build: null codes / ids_out");
  NLSH_REQUIRE(x_sorted_out == nullptr || n == 0 || (x != nullptr && d >= 1),
               "build: x_sorted_out given without x / d");
  NLSH_REQUIRE(x_sqnorm_out == nullptr || x_sorted_out != nullptr, "build: x_sqnorm_out needs x_sorted_out");
  const size_t need = nlsh_build_workspace_bytes(n, n_buckets);
  if (workspace == nullptr || workspace_bytes < need) {
    nlsh_set_error("build: workspace %zu bytes < required %zu", workspace_bytes, need);
    return NLSH_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BuildPlan p = build_plan(n, n_buckets);
  WorkspaceCarver ws(workspace);
  int* hist = ws.take<int>((size_t)p.entries_padded);
  int* block_sums = ws.take<int>((size_t)p.n_scan_blocks + 1);

  NLSH_CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)p.entries_padded * sizeof(int), st));
  const unsigned unit_blocks = (unsigned)((p.n_units + 3) / 4);
  if (n > 0) {
    unit_hist_kernel<<<unit_blocks, 128, 0, st>>>(codes, n, n_buckets, p.n_units, p.rows_per_unit,
                                                  hist);
    NLSH_CUDA_TRY(nlsh_post_launch());
  }
  scan_reduce_kernel<<<(unsigned)p.n_scan_blocks, kScanThreads, 0, st>>>(hist, block_sums);
  NLSH_CUDA_TRY(nlsh_post_launch());
  scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, p.n_scan_blocks);
  NLSH_CUDA_TRY(nlsh_post_launch());
  scan_apply_kernel<<<(unsigned)p.n_scan_blocks, kScanThreads, 0, st>>>(hist, block_sums);
  NLSH_CUDA_TRY(nlsh_post_launch());
  extract_offsets_kernel<<<(n_buckets + 1 + 255) / 256, 256, 0, st>>>(
      hist, p.n_units, n_buckets, block_sums + p.n_scan_blocks, offsets_out);
  NLSH_CUDA_TRY(nlsh_post_launch());
  if (n > 0) {
    unit_scatter_kernel<<<unit_blocks, 128, 0, st>>>(codes, n, n_buckets, p.n_units,
                                                     p.rows_per_unit, hist, ids_out);
    NLSH_CUDA_TRY(nlsh_post_launch());
    if (x_sorted_out != nullptr) {
      const int sms = nlsh_num_sms();
      const long long warps_needed = (n + 3) / 4;
      long long blocks = (warps_needed + 7) / 8;
      const long long max_blocks = (long long)sms * 8;
      if (blocks > max_blocks) blocks = max_blocks;
      const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                       ((reinterpret_cast<uintptr_t>(x_sorted_out) & 15) == 0);
      if (vec)
        gather_rows_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(x, ids_out, n, d, x_sorted_out, x_sqnorm_out);
      else
        gather_rows_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(x, ids_out, n, d, x_sorted_out, x_sqnorm_out);
      NLSH_CUDA_TRY(nlsh_post_launch());
    }
  }
  return NLSH_OK;
}
