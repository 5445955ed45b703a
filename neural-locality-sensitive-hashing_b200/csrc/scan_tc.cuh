// Interface between scan.cu (planning, merge) and scan_tc.cu (the tensor-core filtered scan).
#pragma once
#include "common.cuh"

constexpr int kTcNQ = 32;  // queries per item = UMMA N of the filter GEMM

// One work item of the tensor-core scan: a row chunk of one bucket against up to kTcNQ of the
// (query, probe) pairs that hit the bucket (pairs are stored grouped by bucket, so an item's
// queries are the contiguous range [pair_base, pair_base + nq) of the pair-ordered query copy).
struct __align__(16) TcItem {
  int row0, row1;  // rows of x_sorted
  int pair_base;
  int nq;          // 1..kTcNQ; 0 = end-of-work sentinel (shared memory only)
  int chunk;       // chunk index inside the bucket
  int pad[3];
};
static_assert(sizeof(TcItem) == 32, "TcItem must be 32 bytes");

struct TcScanArgs {
  const float* xs;       // [n_rows, d_pad] bucket-contiguous vectors
  const float* xnorm;    // [n_rows] |x|^2 of each row of xs
  const float* qs;       // [n_pairs, d_pad] query vectors in pair order (pre-normalised for ANGULAR)
  const float* qs_norm;  // [n_pairs] |q|^2 of each row of qs
  const int* pairs;      // [n_pairs] flat probe index f = q * p + slot of each pair
  const TcItem* items;
  const int* n_items;    // device scalar
  int max_items;
  int* item_counter;
  unsigned long long* stats;  // optional debug counters {survivors, re-rank batches} (or nullptr)
  float* tau_g;          // [n_queries] best known upper bound of each query's final k-th distance
  float* part_d;         // [n_queries * p, max_chunks, k]
  int* part_id;          // ROW indices into xs (merge_partials_kernel maps them through ids)
  long long n_rows;
  long long n_pairs;
  int p, k, d, d_pad, kblocks, max_chunks, n_slots;
  float l2_slack;        // 2.1e-6 * sqrt(d): bound of the eps cross term of F.pairwise_distance
};

bool nlsh_scan_tc_supported(int d, int k, int metric);
// qs[i] = qn[pairs[i] / p], qs_norm[i] = |qs[i]|^2 for i < *n_valid; tau_g[q] = the exact k-th best
// distance of query q among the first rows of its first probed bucket (+inf when it has < k rows)
int nlsh_scan_tc_prepare(const float* qn, const int* pairs, const int* n_valid, long long n_pairs,
                         int p, int d_pad, float* qs, float* qs_norm, float* tau_g,
                         long long n_queries, const int* probes, const int* offsets, const float* xs,
                         long long n_rows, int n_buckets, int d, int k, int metric, cudaStream_t st);
int nlsh_scan_tc_launch(int metric, TcScanArgs a, cudaStream_t st);
