// Interface between scan.cu (planning, merge) and scan_tc.cu (the tensor-core filtered scan).
#pragma once
#include "common.cuh"

constexpr int kTcNQ = 32;      // queries per item = UMMA N of the filter GEMM ...
constexpr int kTcNQMax = 128;  // ... or 128 when a bucket is probed by 64 or more queries of the batch
constexpr int kTcLadder = 16;  // threshold-ladder levels per query (scan_tc.cu, "threshold ladder")
constexpr int kTcLadderDen = 32;  // level l stands for the bound tau0 * (1 - l / kTcLadderDen)

// One work item of the tensor-core scan: a row chunk of one bucket against up to kTcNQ of the
// (query, probe) pairs that hit the bucket (pairs are stored grouped by bucket, so an item's
// queries are the contiguous range [pair_base, pair_base + nq) of the pair-ordered query copy).
struct __align__(16) TcItem {
  int row0, row1;  // rows of x_sorted
  int pair_base;
  int nq;          // 1..group size (kTcNQ or kTcNQMax); 0 = end-of-work sentinel
  int chunk;       // chunk index inside the bucket
  int pad[3];
};
static_assert(sizeof(TcItem) == 32, "TcItem must be 32 bytes");

// One scored candidate of a query: the exact distance (squared for L2) and the vector id.
struct __align__(8) TcCand {
  float d;
  int id;
};

struct TcScanArgs {
  const float* xs;       // [n_rows, d_pad] bucket-contiguous vectors
  const float* xnorm;    // [n_rows] |x|^2 of each row of xs
  const int* ids;        // [n_rows] vector id of each row of xs
  const float* qs;       // [n_pairs (+ kTcNQMax), d_pad] query vectors in pair order (pre-normalised for ANGULAR)
  const int* pq;         // [n_pairs] query index of each pair
  const float* pqn2;     // [n_pairs] |q|^2 of each pair's query
  const TcItem* items;
  const int* n_items;    // device scalar
  int max_items;
  int* item_counter;
  unsigned long long* stats;  // optional debug counters (or nullptr), see scan_tc.cu
  float* tau_g;          // [n_queries] best known upper bound of each query's final k-th distance
  const float* tau0;     // [n_queries] the seed's bound (fixed during the scan; the ladder's scale)
  int* ladder;           // [n_queries, kTcLadder] candidates counted per level, or nullptr (no tightening)
  int* cand_n;           // [n_queries] candidates appended so far (may exceed cap: overflow)
  TcCand* cand;          // [n_queries, cap]
  int cap;
  long long n_rows;
  long long n_pairs;
  int k, d, d_pad, kblocks, n_slots;
  int sets_log2;         // log2 of the TMEM accumulator sets in use (tiles the MMAs may run ahead of the filter)
  int nq_group;          // queries per item: kTcNQ or kTcNQMax (the plan's group size)
  int avg_item_rows;     // average rows per bucket (chooses the scorer's query source)
  int sm_reserve;        // SMs the persistent grid leaves free (flags bits 8-15 of nlsh_query_scan_topk)
  float l2_slack;        // 2.1e-6 * sqrt(d): bound of the eps cross term of F.pairwise_distance
};

bool nlsh_scan_tc_supported(int d, int k, int metric);
// nlsh_scan_tc_seed: tau_g[q] = tau0[q] = the exact k-th best distance of query q among the first rows of
// its probed buckets (+inf when they hold < k rows), inflated by the rounding bound of a different summation
// order.  nlsh_scan_tc_prepare: qs[i] = qn[pairs[i] / p], pq[i] = pairs[i] / p, pqn2[i] = |qs[i]|^2 for
// i < *n_valid, then the seed - or, with tau_seed != NULL, tau_g = tau0 = tau_seed (bounds from elsewhere).
int nlsh_scan_tc_seed(const float* qn, long long n_queries, const int* probes, int p, const int* offsets,
                      const float* xs, long long n_rows, int n_buckets, int d, int d_pad, int k, int metric,
                      float* tau_g, float* tau0, cudaStream_t st, int sample_rows = 0);
int nlsh_scan_tc_prepare(const float* qn, const int* pairs, const int* n_valid, long long n_pairs,
                         int p, int d_pad, float* qs, int* pq, float* pqn2, float* tau_g, float* tau0,
                         const float* tau_seed, long long n_queries, const int* probes, const int* offsets,
                         const float* xs, long long n_rows, int n_buckets, int d, int k, int metric,
                         cudaStream_t st);
int nlsh_scan_tc_launch(int metric, TcScanArgs a, cudaStream_t st);

#ifdef __CUDACC__
// Exact distance of one database row to one query by ONE thread, in exactly the arithmetic of
// scan.cu::consume_box (the fp32 SIMT scan): four partial sums over the columns = 0, 1, 2, 3 (mod 4) in
// ascending column order, the columns of a partial last float4 (d % 4 != 0) added to the first, combined
// as (s0 + s1) + (s2 + s3) - so both scan kernels produce the same bits for the same (q, x) and the
// parity tests can compare them with torch.equal.  L2 returns the squared distance in the reference's
// difference form sum((q - x + 1e-6)^2) (nlsh/data.py:201, the root is taken in the merge), ANGULAR
// 1 - <q, x> / max(|x|, 1e-8) for a pre-normalised q (nlsh/data.py:109).
// QLoad::load4(v) returns the query's columns 4 v .. 4 v + 3, QLoad::load1(c) column c.
// 32 bytes per load instruction (sm_100: LDG.256): a lane's 16-byte loads of one row are 32 different cache
// lines per warp instruction, so the scorer is bound by L1 wavefronts, and the wide load halves them.
// p must be 32-byte aligned.
__device__ __forceinline__ void tc_ldg256(const float* p, float4& a, float4& b) {
  // not volatile: a pure load the compiler may hoist and batch like __ldg; the rows are read once, so they are
  // kept out of L1 (what L1 there is beside the kernel's shared memory serves the query vectors)
  asm("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

__device__ __forceinline__ void tc_ldg256_keep(const float* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}

// V8: the row (and a global query) is read 32 bytes at a time; needs d_pad % 8 == 0 and 32-byte aligned bases.
template <int METRIC, int kBatch, bool V8, typename QLoad>
__device__ __forceinline__ float tc_thread_distance(const float* __restrict__ xrow, const QLoad& q, int d) {
  static_assert(kBatch % 2 == 0, "kBatch counts float4 loads, two per 32-byte load");
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
  const int nv = d >> 2, tail = d & 3;
  // kBatch row loads (16 bytes each) are in flight per step
  for (int v0 = 0; v0 < nv; v0 += kBatch) {
    float4 xv[kBatch], qv[kBatch];
    if (V8) {
#pragma unroll
      for (int i = 0; i < kBatch; i += 2) {
        if (v0 + i < nv) {  // the pair's second half may be the partial float4 (inside the padded row)
          tc_ldg256(xrow + 4 * (v0 + i), xv[i], xv[i + 1]);
        } else {
          xv[i] = xv[i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kBatch; ++i)
        xv[i] = (v0 + i < nv) ? __ldg(reinterpret_cast<const float4*>(xrow) + v0 + i)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      if (v0 + i < nv) {
        if (!V8) qv[i] = q.load4(v0 + i);
        else if ((i & 1) == 0) q.load8(v0 + i, qv[i], qv[i + 1]);  // the compiler hoists these as registers allow
        if (METRIC == NLSH_METRIC_L2) {
          const float t0 = __fadd_rn(__fsub_rn(qv[i].x, xv[i].x), 1e-6f);
          const float t1 = __fadd_rn(__fsub_rn(qv[i].y, xv[i].y), 1e-6f);
          const float t2 = __fadd_rn(__fsub_rn(qv[i].z, xv[i].z), 1e-6f);
          const float t3 = __fadd_rn(__fsub_rn(qv[i].w, xv[i].w), 1e-6f);
          s0 = fmaf(t0, t0, s0);
          s1 = fmaf(t1, t1, s1);
          s2 = fmaf(t2, t2, s2);
          s3 = fmaf(t3, t3, s3);
        } else {
          s0 = fmaf(qv[i].x, xv[i].x, s0);
          s1 = fmaf(qv[i].y, xv[i].y, s1);
          s2 = fmaf(qv[i].z, xv[i].z, s2);
          s3 = fmaf(qv[i].w, xv[i].w, s3);
          x0 = fmaf(xv[i].x, xv[i].x, x0);
          x1 = fmaf(xv[i].y, xv[i].y, x1);
          x2 = fmaf(xv[i].z, xv[i].z, x2);
          x3 = fmaf(xv[i].w, xv[i].w, x3);
        }
      }
    }
  }
  for (int c = 0; c < tail; ++c) {  // the partial float4 at vector index nv
    const float xc = __ldg(xrow + 4 * nv + c);
    const float qc = q.load1(4 * nv + c);
    if (METRIC == NLSH_METRIC_L2) {
      const float t = __fadd_rn(__fsub_rn(qc, xc), 1e-6f);
      s0 = fmaf(t, t, s0);
    } else {
      s0 = fmaf(qc, xc, s0);
      x0 = fmaf(xc, xc, x0);
    }
  }
  const float acc = __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3));
  if (METRIC == NLSH_METRIC_L2) return acc;
  const float xx = __fadd_rn(__fadd_rn(x0, x1), __fadd_rn(x2, x3));
  return 1.0f - acc / fmaxf(sqrtf(xx), 1e-8f);
}

// Query in global memory, row-major [.., d_pad] (the staged copy qn of scan.cu).
struct TcQueryGlobal {
  const float* q;
  __device__ __forceinline__ float4 load4(int v) const { return __ldg(reinterpret_cast<const float4*>(q) + v); }
  __device__ __forceinline__ void load8(int v, float4& a, float4& b) const { tc_ldg256_keep(q + 4 * v, a, b); }
  __device__ __forceinline__ float load1(int c) const { return __ldg(q + c); }
};
#endif
