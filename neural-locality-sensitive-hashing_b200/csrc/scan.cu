// Query hot path: multi-probe candidate scan + exact distance + top-k, batched over all
// queries.  Replaces the per-query Python loop of Indexer.query (nlsh/indexer.py:62-95):
// the index_select gather (77-82) disappears because buckets are contiguous in x_sorted,
// distance_func (84-87 -> nlsh/data.py:109, :201) and topk (90-91) are fused into the scan.
// The same kernel in "dense" mode (every query against every row) is the brute-force kNN
// of precompute.py:57-67.
//
// Work decomposition (bucket-major so a bucket tile fetched once serves every query that
// probes it):
//   plan_count / plan_scan / plan_scatter : (query, probe) pairs grouped by bucket
//   item = (bucket, row chunk, group of <= kG queries of that bucket)
//   scan kernel: persistent CTAs pull items from an atomic counter.  One elected producer
//     thread streams the chunk's rows HBM/L2 -> shared memory with 2-D TMA tensor copies
//     (cp.async.bulk.tensor, SASS UTMALDG; box = 128 rows x 32 floats, SWIZZLE_128B) into a
//     ring of stages guarded by full/empty mbarriers; four consumer warps (one row per
//     thread; the hardware swizzle makes the 128-bit row reads conflict-free while the query
//     values are broadcast) accumulate the distances of the row to the item's queries with
//     packed fp32x2 math; per row tile the 128 x kG scores are exchanged through shared
//     memory so that each warp owns the register-resident sorted top-k lists of two of the
//     item's queries (threshold-filtered by ballot, so inserts are rare).
//   merge kernel: one warp per query merges its partial lists -> ids/dists/n_candidates
//     (for L2 the lists hold squared distances; the square root is taken once, here).
// All top-k decisions use the (distance, id) lexicographic order, so the result does not
// depend on item scheduling or on how many GPUs the database is sharded over.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "scan_tc.cuh"

namespace {

constexpr int kConsumerWarps = 4;
constexpr int kTileRows = 32 * kConsumerWarps;  // rows per stage
constexpr int kG = 8;                           // queries per item
constexpr int kBoxCols = 32;                    // TMA box: 32 floats = one 128-byte swizzle row
constexpr int kBoxFloats = kTileRows * kBoxCols;  // 16 KB per box
constexpr int kMaxBoxesPerStage = 2;            // a stage holds up to 64 columns of 128 rows (32 KB)
constexpr int kMaxChunksPerBucket = 64;
constexpr int kListsPerWarp = kG / kConsumerWarps;  // queries whose top-k a warp owns
constexpr size_t kExchangeBytes = 2 * (kG + 1) * kTileRows * sizeof(float);
constexpr size_t kSmemPerSm = 227 * 1024;

// One work item as the scan kernel consumes it (64 bytes, written by plan_items_kernel).
struct __align__(16) ItemRec {
  int row0, row1;  // rows of x_sorted covered by this item
  int ng;          // queries in this item (1..kG); 0 = end-of-work sentinel (shared memory only)
  int chunk;       // chunk index inside the bucket / dense row block
  int f[kG];       // flat probe index q * p + j (probe mode) or query index (dense); -1 unused
  int pad[4];
};
static_assert(sizeof(ItemRec) == 64, "ItemRec must be 64 bytes");

struct ScanGeom {
  int d;         // real columns
  int d_pad;     // row stride of x_sorted (multiple of 4)
  int n_boxes;   // ceil(d_pad / 32) column boxes per row tile
  int bps;       // boxes per stage
  int n_chunks;  // ceil(n_boxes / bps) stages per row tile
  int stages;
  int nqb;       // item / query ring depth (2, or 1 when the queries are too wide)
  int ctas_per_sm;
  size_t smem_bytes;
};

// Occupancy plan: k <= 64 kernels are compiled for 3 CTAs / SM (<= 136 registers), the
// k <= 128 ones for 2.  NLSH_SCAN_TUNE="boxes_per_stage,stages,ctas_per_sm" overrides it
// (experiments only).
int scan_max_ctas(int k) { return k <= 64 ? 3 : 2; }

ScanGeom scan_geom(int d, int k, bool async) {
  ScanGeom g;
  g.d = d;
  g.d_pad = (d + 3) / 4 * 4;
  g.n_boxes = (g.d_pad + kBoxCols - 1) / kBoxCols;
  g.ctas_per_sm = scan_max_ctas(k);
  g.bps = 1;
  int want_stages = 3;
  if (const char* env = getenv("NLSH_SCAN_TUNE")) {
    int b = 0, st = 0, c = 0;
    if (sscanf(env, "%d,%d,%d", &b, &st, &c) == 3) {
      if (b >= 1 && b <= kMaxBoxesPerStage) g.bps = b;
      if (st >= 2 && st <= 4) want_stages = st;
      if (c >= 1 && c <= scan_max_ctas(k)) g.ctas_per_sm = c;
    }
  }
  if (g.bps > g.n_boxes) g.bps = g.n_boxes;
  g.n_chunks = (g.n_boxes + g.bps - 1) / g.bps;
  const size_t stage_bytes = (size_t)g.bps * kBoxFloats * sizeof(float);
  const size_t q_bytes = (size_t)kG * g.d_pad * sizeof(float);
  // score/id exchange (double buffered) + item records + barriers + 1024-byte alignment slack
  const size_t misc = kExchangeBytes + 2 * sizeof(ItemRec) + 256 + 1024;
  g.stages = 1;
  g.nqb = 1;
  if (async) {
    g.stages = want_stages;
    g.nqb = 2;
    // fit ctas_per_sm CTAs (each also pays 1 KB of driver-reserved shared memory): first give up
    // the second query buffer, then CTAs, then stages
    auto need = [&](int nqb, int stages) { return nqb * q_bytes + misc + stages * stage_bytes + 1024; };
    if (need(2, 2) * g.ctas_per_sm > kSmemPerSm) g.nqb = 1;
    while (g.ctas_per_sm > 1 && need(g.nqb, 2) * g.ctas_per_sm > kSmemPerSm) --g.ctas_per_sm;
    if (g.nqb == 1 && need(2, 2) * g.ctas_per_sm <= kSmemPerSm) g.nqb = 2;
    while (g.stages > 2 && need(g.nqb, g.stages) * g.ctas_per_sm > kSmemPerSm) --g.stages;
  }
  g.smem_bytes = g.nqb * q_bytes + misc + g.stages * stage_bytes;
  return g;
}

struct ScanPolicy {
  int rchunk;      // rows per item chunk (multiple of kTileRows)
  int max_chunks;  // chunks of the largest bucket
};

// row_bytes > 0 (the tensor-core scan): the query groups of one bucket chunk are consecutive items, picked up by
// different CTAs at about the same time, so a chunk streamed by the first of them is an L2 hit for the others
// only while the chunks in flight on all CTAs fit the L2 together; with several groups per bucket the chunk is
// bounded so that they take at most a quarter of it.
ScanPolicy scan_policy(int64_t n_queries, int p, int n_buckets, int64_t n_rows,
                       int64_t max_bucket_rows, int group = kG, int ctas_per_sm = 2, int64_t row_bytes = 0) {
  const int64_t grid = (int64_t)nlsh_num_sms() * ctas_per_sm;
  const int64_t target_items = grid * 8;
  const int64_t pairs = n_queries * p > 0 ? n_queries * p : 1;
  const int64_t distinct = pairs < n_buckets ? pairs : n_buckets;
  int64_t groups = pairs / group;
  if (groups < distinct) groups = distinct;
  if (groups < 1) groups = 1;
  const int64_t chunks_needed = (target_items + groups - 1) / groups;
  int64_t avg_bucket = n_rows / (n_buckets > 0 ? n_buckets : 1);
  if (avg_bucket < 1) avg_bucket = 1;
  int64_t rchunk = (avg_bucket + chunks_needed - 1) / chunks_needed;
  const int64_t groups_per_bucket = groups / (distinct > 0 ? distinct : 1);
  if (row_bytes > 0 && groups_per_bucket >= 2) {
    const int64_t l2_share = 32ll << 20;
    int64_t cap_rows = l2_share * groups_per_bucket / (grid * row_bytes);
    if (cap_rows < 2 * kTileRows) cap_rows = 2 * kTileRows;
    if (rchunk > cap_rows) rchunk = cap_rows;
  }
  if (max_bucket_rows < 1) max_bucket_rows = 1;
  const int64_t floor_rchunk = (max_bucket_rows + kMaxChunksPerBucket - 1) / kMaxChunksPerBucket;
  if (rchunk < floor_rchunk) rchunk = floor_rchunk;
  rchunk = (rchunk + kTileRows - 1) / kTileRows * kTileRows;
  ScanPolicy pol;
  pol.rchunk = (int)rchunk;
  pol.max_chunks = (int)((max_bucket_rows + rchunk - 1) / rchunk);
  if (pol.max_chunks < 1) pol.max_chunks = 1;
  return pol;
}

struct ScanArgs {
  const float* xs;    // [n_rows, d_pad]
  const int* ids;     // [n_rows] (probe mode) or nullptr (dense: id = row)
  const int* offsets;  // [n_buckets + 1] (probe mode)
  const float* q;     // [n_queries, d] (pre-normalised for the cosine metrics)
  const int* pair_off;
  const int* pairs;
  const int* item_off;
  const ItemRec* items;
  int* item_counter;
  float* part_d;
  int* part_id;
  long long n_rows;
  long long self_offset;
  int n_queries;
  int n_buckets;
  int p;
  int k;
  int rchunk;
  int max_chunks;
  int dense;
  int dense_qgroups;
  int dense_items;
  int exclude_self;
  int d, d_pad, n_boxes, bps, n_chunks, stages, nqb;
};

// ---- per-box distance accumulation --------------------------------------------------------
// xrow: this thread's 128-byte row inside a box whose 16-byte chunks are XOR-swizzled with
// (row & 7) (TMA SWIZZLE_128B); xr = row & 7; qs: query values of the box's first column
// (query g at qs + g * d_pad); nvec full float4 columns, tail = extra valid columns (0..3).
// Accumulators are packed float2 pairs so the arithmetic issues as Blackwell's two-wide fp32
// instructions (FADD2 / FFMA2): acc[g][0] sums columns 0,1 (mod 4), acc[g][1] columns 2,3.
struct Acc {
  float2 a[kG][2];
  float2 xx[2];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int g = 0; g < kG; ++g) a[g][0] = a[g][1] = make_float2(0.f, 0.f);
    xx[0] = xx[1] = make_float2(0.f, 0.f);
  }
  __device__ __forceinline__ float sum(int g) const {
    return (a[g][0].x + a[g][0].y) + (a[g][1].x + a[g][1].y);
  }
  __device__ __forceinline__ float sum_xx() const { return (xx[0].x + xx[0].y) + (xx[1].x + xx[1].y); }
};

template <int METRIC, int NG>
__device__ __forceinline__ void consume_box(Acc& acc, const float* __restrict__ xrow, int xr,
                                            const float* __restrict__ qs, int d_pad, int nvec,
                                            int tail) {
  const float2 eps2 = make_float2(1e-6f, 1e-6f);
#pragma unroll 8
  for (int v = 0; v < nvec; ++v) {
    const float4 xv = *reinterpret_cast<const float4*>(xrow + ((v ^ xr) << 2));
    const float2 x01 = make_float2(xv.x, xv.y), x23 = make_float2(xv.z, xv.w);
    const float2 n01 = make_float2(-xv.x, -xv.y), n23 = make_float2(-xv.z, -xv.w);
    if (METRIC == NLSH_METRIC_ANGULAR || METRIC == NLSH_METRIC_COSINE) {
      acc.xx[0] = __ffma2_rn(x01, x01, acc.xx[0]);
      acc.xx[1] = __ffma2_rn(x23, x23, acc.xx[1]);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float4 qv = *reinterpret_cast<const float4*>(qs + g * d_pad + 4 * v);
      const float2 q01 = make_float2(qv.x, qv.y), q23 = make_float2(qv.z, qv.w);
      if (METRIC == NLSH_METRIC_L2) {
        // F.pairwise_distance: (q - x) + eps, squared and summed (nlsh/data.py:201)
        const float2 t01 = __fadd2_rn(__fadd2_rn(q01, n01), eps2);
        const float2 t23 = __fadd2_rn(__fadd2_rn(q23, n23), eps2);
        acc.a[g][0] = __ffma2_rn(t01, t01, acc.a[g][0]);
        acc.a[g][1] = __ffma2_rn(t23, t23, acc.a[g][1]);
      } else if (METRIC == NLSH_METRIC_L2SQ) {
        const float2 t01 = __fadd2_rn(q01, n01);
        const float2 t23 = __fadd2_rn(q23, n23);
        acc.a[g][0] = __ffma2_rn(t01, t01, acc.a[g][0]);
        acc.a[g][1] = __ffma2_rn(t23, t23, acc.a[g][1]);
      } else {
        acc.a[g][0] = __ffma2_rn(q01, x01, acc.a[g][0]);
        acc.a[g][1] = __ffma2_rn(q23, x23, acc.a[g][1]);
      }
    }
  }
  // partial last vector (d not a multiple of 4): only the first `tail` components exist
  for (int c = 0; c < tail; ++c) {
    const float xv = xrow[((nvec ^ xr) << 2) + c];
    if (METRIC == NLSH_METRIC_ANGULAR || METRIC == NLSH_METRIC_COSINE)
      acc.xx[0].x = fmaf(xv, xv, acc.xx[0].x);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const float qv = qs[g * d_pad + 4 * nvec + c];
      if (METRIC == NLSH_METRIC_L2) {
        const float t = __fadd_rn(__fsub_rn(qv, xv), 1e-6f);
        acc.a[g][0].x = fmaf(t, t, acc.a[g][0].x);
      } else if (METRIC == NLSH_METRIC_L2SQ) {
        const float t = qv - xv;
        acc.a[g][0].x = fmaf(t, t, acc.a[g][0].x);
      } else {
        acc.a[g][0].x = fmaf(qv, xv, acc.a[g][0].x);
      }
    }
  }
}

template <int METRIC>
__device__ __forceinline__ void consume_dispatch(int ng, Acc& acc, const float* xrow, int xr,
                                                 const float* qs, int d_pad, int nvec, int tail) {
  if (ng <= 1)
    consume_box<METRIC, 1>(acc, xrow, xr, qs, d_pad, nvec, tail);
  else if (ng <= 2)
    consume_box<METRIC, 2>(acc, xrow, xr, qs, d_pad, nvec, tail);
  else if (ng <= 4)
    consume_box<METRIC, 4>(acc, xrow, xr, qs, d_pad, nvec, tail);
  else
    consume_box<METRIC, 8>(acc, xrow, xr, qs, d_pad, nvec, tail);
}

// Score kept in the top-k lists: the distance itself, except for L2 where it is the squared
// distance (same order; merge_partials_kernel takes the root once per result).
template <int METRIC>
__device__ __forceinline__ float finalize_distance(float acc, float xx) {
  if (METRIC == NLSH_METRIC_L2) return acc;
  if (METRIC == NLSH_METRIC_L2SQ) return acc;
  if (METRIC == NLSH_METRIC_ANGULAR) return 1.0f - acc / fmaxf(sqrtf(xx), 1e-8f);
  return 1.0f - acc / sqrtf(xx);  // precompute._cosine_distance: no clamp
}

// Rows of a bucket (or of the dense row range) are split into nch balanced chunks whose
// length is a multiple of the row tile: chunk c covers [c * rc, min(size, (c + 1) * rc)).
__host__ __device__ __forceinline__ int chunk_count(int size, int rchunk, int max_chunks) {
  int nch = (size + rchunk - 1) / rchunk;
  return nch > max_chunks ? max_chunks : nch;
}
__host__ __device__ __forceinline__ int chunk_rows(int size, int nch) {
  const int per = (size + nch - 1) / nch;
  return (per + kTileRows - 1) / kTileRows * kTileRows;
}

struct Item {
  long long row0, row1;  // rows of x_sorted covered by this item
  int ng;                // queries in this item (1..kG)
  int chunk;             // chunk index inside the bucket / dense row block
  int pair_base;         // probe mode: first entry of `pairs`; dense: first query index
};


__device__ __forceinline__ Item decode_item(const ScanArgs& a, int item) {
  Item it;
  if (a.dense) {
    const int blk = item / a.dense_qgroups;
    const int gq = item - blk * a.dense_qgroups;
    it.row0 = (long long)blk * a.rchunk;
    it.row1 = it.row0 + a.rchunk;
    if (it.row1 > a.n_rows || blk == a.max_chunks - 1) it.row1 = a.n_rows;
    it.chunk = blk;
    it.pair_base = gq * kG;
    const int left = a.n_queries - it.pair_base;
    it.ng = left < kG ? left : kG;
    return it;
  }
  // largest b with item_off[b] <= item
  int lo = 0, hi = a.n_buckets;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.item_off[mid] <= item) lo = mid; else hi = mid;
  }
  const int b = lo;
  const int local = item - a.item_off[b];
  const int p0 = a.pair_off[b];
  const int nq = a.pair_off[b + 1] - p0;
  const int ngroups = (nq + kG - 1) / kG;
  const int c = local / ngroups;
  const int gq = local - c * ngroups;
  const int r0 = a.offsets[b];
  const int size = a.offsets[b + 1] - r0;
  const int nch = chunk_count(size, a.rchunk, a.max_chunks);
  const int rc = chunk_rows(size, nch);
  long long c_lo = (long long)c * rc, c_hi = c_lo + rc;
  if (c_lo > size) c_lo = size;
  if (c_hi > size || c == nch - 1) c_hi = size;
  it.row0 = r0 + c_lo;
  it.row1 = r0 + c_hi;
  it.chunk = c;
  it.pair_base = p0 + gq * kG;
  const int left = nq - gq * kG;
  it.ng = left < kG ? left : kG;
  return it;
}

// Expands item index -> ItemRec (the binary search over item_off runs here, thousands of
// threads wide, instead of on the scan kernel's critical path).
__global__ void plan_items_kernel(const ScanArgs a, ItemRec* __restrict__ out, int max_items) {
  int total = a.dense ? a.dense_items : a.item_off[a.n_buckets];
  if (total > max_items) total = max_items;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const Item it = decode_item(a, i);
    ItemRec r;
    r.row0 = (int)it.row0;
    r.row1 = (int)it.row1;
    r.ng = it.ng;
    r.chunk = it.chunk;
#pragma unroll
    for (int g = 0; g < kG; ++g)
      r.f[g] = g < it.ng ? (a.dense ? it.pair_base + g : a.pairs[it.pair_base + g]) : -1;
    r.pad[0] = r.pad[1] = r.pad[2] = r.pad[3] = 0;
    out[i] = r;
  }
}

// Items of the tensor-core scan (scan_tc.cu): (bucket, row chunk, group of <= kTcNQ pairs).
__global__ void plan_tc_items_kernel(const int* __restrict__ item_off, const int* __restrict__ pair_off,
                                     const int* __restrict__ offsets, int n_buckets, int rchunk,
                                     int max_chunks, int group, TcItem* __restrict__ out, int max_items) {
  int total = item_off[n_buckets];
  if (total > max_items) total = max_items;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int lo = 0, hi = n_buckets;  // largest b with item_off[b] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (item_off[mid] <= i) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int local = i - item_off[b];
    const int p0 = pair_off[b];
    const int nq = pair_off[b + 1] - p0;
    const int ngroups = (nq + group - 1) / group;
    const int c = local / ngroups;
    const int gq = local - c * ngroups;
    const int r0 = offsets[b];
    const int size = offsets[b + 1] - r0;
    const int nch = chunk_count(size, rchunk, max_chunks);
    const int rc = chunk_rows(size, nch);
    long long c_lo = (long long)c * rc, c_hi = c_lo + rc;
    if (c_lo > size) c_lo = size;
    if (c_hi > size || c == nch - 1) c_hi = size;
    TcItem r;
    r.row0 = r0 + (int)c_lo;
    r.row1 = r0 + (int)c_hi;
    r.pair_base = p0 + gq * group;
    const int left = nq - gq * group;
    r.nq = left < group ? left : group;
    r.chunk = c;
    r.pad[0] = r.pad[1] = r.pad[2] = 0;
    out[i] = r;
  }
}

// ---- pieces shared by the two scan kernels --------------------------------------------------
template <int KPL>
struct ItemLists {
  WarpTopK<KPL, int> top[kListsPerWarp];  // lists of queries g = warp + kConsumerWarps * i
  int self_id[kListsPerWarp];

  __device__ __forceinline__ void init(const ScanArgs& a, const int* f, int ng, int warp) {
#pragma unroll
    for (int i = 0; i < kListsPerWarp; ++i) {
      top[i].init(NLSH_ID_SENTINEL);
      const int g = warp + kConsumerWarps * i;
      self_id[i] = (a.exclude_self && g < ng) ? (int)(a.self_offset + f[g]) : -1;
    }
  }

  // Row tile epilogue: scores of the 128 rows x ng queries go through shared memory to the
  // warp that owns each query's list (buffers double-buffered on `parity`: one named barrier
  // per tile is enough).
  // `first` = first tile of the item (lists still empty): its first 32 rows seed the list with
  // one bitonic sort instead of k + k ln(32/k) inserts.
  template <int METRIC>
  __device__ __forceinline__ void tile_epilogue(const Acc& acc, float* score_s, int* id_s, int parity,
                                                int r_local, int cand_id, bool valid, int ng,
                                                int warp, int lane, int k, bool first) {
    float* sc = score_s + parity * (kG * kTileRows);
    int* idb = id_s + parity * kTileRows;
    idb[r_local] = valid ? cand_id : NLSH_ID_SENTINEL;
    if (valid) {
      const float xx = acc.sum_xx();
#pragma unroll
      for (int g = 0; g < kG; ++g)
        if (g < ng) sc[g * kTileRows + r_local] = finalize_distance<METRIC>(acc.sum(g), xx);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kConsumerWarps) : "memory");
#pragma unroll
    for (int i = 0; i < kListsPerWarp; ++i) {
      const int g = warp + kConsumerWarps * i;
      if (g < ng) {
#pragma unroll
        for (int j = 0; j < kTileRows / 32; ++j) {
          const int r = lane + 32 * j;
          const int cid = idb[r];
          const bool ok = cid != NLSH_ID_SENTINEL && cid != self_id[i];
          const float cd = ok ? sc[g * kTileRows + r] : 0.f;
          if (first && j == 0)
            top[i].seed32(cd, cid, ok, NLSH_ID_SENTINEL, k);
          else
            top[i].offer(cd, cid, ok, k);
        }
      }
    }
  }

  // Partial lists: one per (query-probe, chunk).
  __device__ __forceinline__ void write_out(const ScanArgs& a, const int* f, int ng, int chunk,
                                            int warp, int lane) const {
#pragma unroll
    for (int i = 0; i < kListsPerWarp; ++i) {
      const int g = warp + kConsumerWarps * i;
      if (g < ng) {
        const size_t slot = (size_t)f[g] * a.max_chunks + chunk;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const int pos = j * 32 + lane;
          if (pos < a.k) {
            a.part_d[slot * a.k + pos] = top[i].d[j];
            a.part_id[slot * a.k + pos] = top[i].id[j];
          }
        }
      }
    }
  }
};

__device__ __forceinline__ unsigned char* align_smem_1024(unsigned char* p) {
  return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u);  // SWIZZLE_128B boxes need 1024-byte alignment
}

// Distances of this thread's row against the stage's boxes.
template <int METRIC>
__device__ __forceinline__ void consume_stage(const ScanArgs& a, Acc& acc, const float* stage, int box0,
                                              int nb, int r_local, int xr, const float* qs, int ng) {
  for (int b = 0; b < nb; ++b) {
    const int col0 = (box0 + b) * kBoxCols;
    int cols = a.d - col0;  // valid (unpadded) columns of this box
    if (cols > kBoxCols) cols = kBoxCols;
    if (cols <= 0) break;
    consume_dispatch<METRIC>(ng, acc, stage + b * kBoxFloats + r_local * kBoxCols, xr, qs + col0,
                             a.d_pad, cols >> 2, cols & 3);
  }
}

// ---- the scan kernel ------------------------------------------------------------------------
// Warps 0..3 consume; lane 0 of warp 4 is the producer.  The producer owns the work queue: it
// pulls item indices from the global atomic counter and, per item, publishes the 64-byte item
// record + the item's query vectors (1-D bulk copies) through a 2-deep item ring
// (q_full / q_empty mbarriers), then streams the item's row tiles with 2-D TMA tensor copies into
// the stage ring (full / empty mbarriers).  Both rings run across item boundaries, so the
// pipeline never drains and the consumers never touch global memory except for the row ids.
template <int METRIC, int KPL>
__global__ void __launch_bounds__(32 * (kConsumerWarps + 1), KPL <= 2 ? 3 : 2)
    scan_kernel(const ScanArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ unsigned char smem_raw[];
  float* stage_buf = reinterpret_cast<float*>(align_smem_1024(smem_raw));
  const size_t stage_floats = (size_t)a.bps * kBoxFloats;
  float* qbuf = stage_buf + (size_t)a.stages * stage_floats;               // [nqb][kG][d_pad]
  float* score_s = qbuf + (size_t)a.nqb * kG * a.d_pad;                    // [2][kG][kTileRows]
  int* id_s = reinterpret_cast<int*>(score_s + 2 * kG * kTileRows);        // [2][kTileRows]
  ItemRec* itm = reinterpret_cast<ItemRec*>(id_s + 2 * kTileRows);         // [nqb]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(itm + 2);               // [stages <= 4]
  uint64_t* empty_bar = full_bar + 4;                                      // [stages <= 4]
  uint64_t* q_full = empty_bar + 4;                                        // [nqb <= 2]
  uint64_t* q_empty = q_full + 2;                                          // [nqb <= 2]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    for (int s = 0; s < a.nqb; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], kConsumerWarps);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // =================================== producer =========================================
    if (lane != 0) return;
    const int total_items = a.dense ? a.dense_items : a.item_off[a.n_buckets];
    const unsigned q_bytes = (unsigned)a.d_pad * sizeof(float);
    unsigned ring = 0, icount = 0;
    while (true) {
      const int item = atomicAdd(a.item_counter, 1);
      const int islot = (int)(icount % (unsigned)a.nqb);
      mbar_wait(&q_empty[islot], ((icount / (unsigned)a.nqb) & 1u) ^ 1u);
      ItemRec* dst = &itm[islot];
      if (item >= total_items) {
        dst->ng = 0;  // end of work
        mbar_arrive(&q_full[islot]);
        break;
      }
      const ItemRec rec = a.items[item];
      *dst = rec;
      mbar_arrive_expect_tx(&q_full[islot], (unsigned)rec.ng * q_bytes);
      float* qdst = qbuf + (size_t)islot * kG * a.d_pad;
      for (int g = 0; g < rec.ng; ++g) {
        const int qidx = a.dense ? rec.f[g] : rec.f[g] / a.p;
        bulk_g2s(qdst + g * a.d_pad, a.q + (size_t)qidx * a.d_pad, q_bytes, &q_full[islot]);
      }
      const int n_tiles = (rec.row1 - rec.row0 + kTileRows - 1) / kTileRows;
      for (int t = 0; t < n_tiles; ++t) {
        const int tile_row0 = rec.row0 + t * kTileRows;
        for (int ch = 0; ch < a.n_chunks; ++ch) {
          const int box0 = ch * a.bps;
          const int nb = (a.n_boxes - box0) < a.bps ? (a.n_boxes - box0) : a.bps;
          const int s = (int)(ring % (unsigned)a.stages);
          mbar_wait(&empty_bar[s], ((ring / (unsigned)a.stages) & 1u) ^ 1u);
          // a box is always written in full (rows / columns past the tensor are zero filled)
          mbar_arrive_expect_tx(&full_bar[s], (unsigned)(nb * kBoxFloats * sizeof(float)));
          float* sdst = stage_buf + s * stage_floats;
          for (int b = 0; b < nb; ++b)
            tma_load_2d(sdst + b * kBoxFloats, &tmap, (box0 + b) * kBoxCols, tile_row0, &full_bar[s]);
          ++ring;
        }
      }
      ++icount;
    }
    return;
  }

  // ===================================== consumers ==========================================
  const int r_local = warp * 32 + lane;
  const int xr = r_local & 7;
  unsigned ring = 0, icount = 0, tcount = 0;
  while (true) {
    const int islot = (int)(icount % (unsigned)a.nqb);
    mbar_wait(&q_full[islot], (icount / (unsigned)a.nqb) & 1u);
    const ItemRec* rec = &itm[islot];
    const int ng = rec->ng;
    if (ng == 0) break;
    const int row0 = rec->row0, row1 = rec->row1;
    const float* qs = qbuf + (size_t)islot * kG * a.d_pad;
    ItemLists<KPL> lists;
    lists.init(a, rec->f, ng, warp);
    const int n_tiles = (row1 - row0 + kTileRows - 1) / kTileRows;
    for (int t = 0; t < n_tiles; ++t) {
      const int row = row0 + t * kTileRows + r_local;
      const bool valid = row < row1;
      const bool warp_has_rows = row0 + t * kTileRows + warp * 32 < row1;
      int cand_id = NLSH_ID_SENTINEL;
      if (valid) cand_id = a.ids ? a.ids[row] : row;  // latency hidden behind the chunk loop
      Acc acc;
      acc.clear();
      for (int ch = 0; ch < a.n_chunks; ++ch) {
        const int box0 = ch * a.bps;
        const int nb = (a.n_boxes - box0) < a.bps ? (a.n_boxes - box0) : a.bps;
        const int s = (int)(ring % (unsigned)a.stages);
        mbar_wait(&full_bar[s], (ring / (unsigned)a.stages) & 1u);
        if (warp_has_rows)  // a warp whose 32 rows all lie past the chunk end only keeps the ring in step
          consume_stage<METRIC>(a, acc, stage_buf + s * stage_floats, box0, nb, r_local, xr, qs, ng);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        ++ring;
      }
      lists.template tile_epilogue<METRIC>(acc, score_s, id_s, (int)(tcount & 1u), r_local, cand_id,
                                           valid, ng, warp, lane, a.k, t == 0);
      ++tcount;
    }
    lists.write_out(a, rec->f, ng, rec->chunk, warp, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&q_empty[islot]);
    ++icount;
  }
}

// Debug / A-B variant without TMA or mbarriers: 4 warps stage every chunk cooperatively with
// plain loads into the same swizzled layout (flags bit 0 of nlsh_query_scan_topk).
template <int METRIC, int KPL>
__global__ void __launch_bounds__(32 * kConsumerWarps, 2) scan_kernel_sync(const ScanArgs a) {
  extern __shared__ unsigned char smem_raw[];
  float* stage_buf = reinterpret_cast<float*>(align_smem_1024(smem_raw));
  const size_t stage_floats = (size_t)a.bps * kBoxFloats;
  float* qs = stage_buf + stage_floats;                               // [kG][d_pad]
  float* score_s = qs + (size_t)kG * a.d_pad;                         // [2][kG][kTileRows]
  int* id_s = reinterpret_cast<int*>(score_s + 2 * kG * kTileRows);   // [2][kTileRows]
  ItemRec* itm = reinterpret_cast<ItemRec*>(id_s + 2 * kTileRows);    // [1]
  int* s_item = reinterpret_cast<int*>(itm + 2);                      // [2]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int total_items = a.dense ? a.dense_items : a.item_off[a.n_buckets];
  if (tid == 0) s_item[0] = atomicAdd(a.item_counter, 1);
  __syncthreads();
  int round = 0;
  int item = s_item[0];
  const int r_local = warp * 32 + lane;
  const int xr = r_local & 7;
  unsigned tcount = 0;

  while (item < total_items) {
    if (tid == 0) {
      s_item[(round + 1) & 1] = atomicAdd(a.item_counter, 1);
      itm[0] = a.items[item];
    }
    __syncthreads();
    const ItemRec* rec = &itm[0];
    const int ng = rec->ng;
    const int row0 = rec->row0, row1 = rec->row1;
    for (int g = 0; g < kG; ++g) {
      int qidx = -1;
      if (g < ng) qidx = a.dense ? rec->f[g] : rec->f[g] / a.p;
      for (int c = tid; c < a.d_pad; c += 32 * kConsumerWarps)
        qs[g * a.d_pad + c] = qidx >= 0 ? a.q[(size_t)qidx * a.d_pad + c] : 0.f;
    }
    __syncthreads();
    ItemLists<KPL> lists;
    lists.init(a, rec->f, ng, warp);
    const int n_tiles = (row1 - row0 + kTileRows - 1) / kTileRows;
    for (int t = 0; t < n_tiles; ++t) {
      const int tile_row0 = row0 + t * kTileRows;
      const int row = tile_row0 + r_local;
      const bool valid = row < row1;
      int cand_id = NLSH_ID_SENTINEL;
      if (valid) cand_id = a.ids ? a.ids[row] : row;
      Acc acc;
      acc.clear();
      const int rows_in_tile = (row1 - tile_row0) < kTileRows ? (row1 - tile_row0) : kTileRows;
      for (int ch = 0; ch < a.n_chunks; ++ch) {
        const int box0 = ch * a.bps;
        const int nb = (a.n_boxes - box0) < a.bps ? (a.n_boxes - box0) : a.bps;
        // same layout as the TMA boxes: box-major, 128-byte rows, 16-byte chunk ^= (row & 7)
        const int cols = (a.d_pad - box0 * kBoxCols) < nb * kBoxCols ? (a.d_pad - box0 * kBoxCols)
                                                                   : nb * kBoxCols;
        const int cvec = cols >> 2;
        for (int idx = tid; idx < rows_in_tile * cvec; idx += 32 * kConsumerWarps) {
          const int r = idx / cvec, v = idx - r * cvec;
          const float4 val = __ldcs(reinterpret_cast<const float4*>(
                                        a.xs + (size_t)(tile_row0 + r) * a.d_pad + box0 * kBoxCols) + v);
          const int b = v >> 3, c = v & 7;
          *reinterpret_cast<float4*>(stage_buf + b * kBoxFloats + r * kBoxCols + ((c ^ (r & 7)) << 2)) = val;
        }
        __syncthreads();
        consume_stage<METRIC>(a, acc, stage_buf, box0, nb, r_local, xr, qs, ng);
        __syncthreads();
      }
      lists.template tile_epilogue<METRIC>(acc, score_s, id_s, (int)(tcount & 1u), r_local, cand_id,
                                           valid, ng, warp, lane, a.k, t == 0);
      ++tcount;
    }
    lists.write_out(a, rec->f, ng, rec->chunk, warp, lane);
    __syncthreads();
    ++round;
    item = s_item[round & 1];
  }
}

// ---- plan kernels ------------------------------------------------------------------------
__device__ __forceinline__ bool probe_valid(const int* __restrict__ probes, const int* offsets,
                                            int n_buckets, int p, long long f, int& b_out) {
  const int b = probes[f];
  b_out = b;
  if (b < 0 || b >= n_buckets) return false;
  if (offsets[b + 1] - offsets[b] <= 0) return false;
  const long long row_base = f - (f % p);
  for (long long e = row_base; e < f; ++e)
    if (probes[e] == b) return false;  // duplicate probe: the reference probes a set
  return true;
}

__global__ void plan_count_kernel(const int* __restrict__ probes, const int* __restrict__ offsets,
                                  int n_buckets, int p, long long n_pairs, int* __restrict__ cnt) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_pairs) return;
  int b;
  if (probe_valid(probes, offsets, n_buckets, p, f, b)) atomicAdd(&cnt[b], 1);
}

// One block: pair_off = exclusive scan of cnt, item_off = exclusive scan of items per bucket.
__global__ void __launch_bounds__(1024)
    plan_scan_kernel(const int* __restrict__ cnt, const int* __restrict__ offsets, int n_buckets,
                     int rchunk, int max_chunks, int group, int* __restrict__ pair_off,
                     int* __restrict__ item_off) {
  __shared__ int sw[2][33];
  int carry_p = 0, carry_i = 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n_buckets; base += blockDim.x) {
    const int b = base + threadIdx.x;
    int c = 0, items = 0;
    if (b < n_buckets) {
      c = cnt[b];
      const int size = offsets[b + 1] - offsets[b];
      items = ((c + group - 1) / group) * chunk_count(size, rchunk, max_chunks);
    }
    int ic = c, ii = items;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tc = __shfl_up_sync(NLSH_FULL_MASK, ic, o);
      const int ti = __shfl_up_sync(NLSH_FULL_MASK, ii, o);
      if (lane >= o) {
        ic += tc;
        ii += ti;
      }
    }
    if (lane == 31) {
      sw[0][warp] = ic;
      sw[1][warp] = ii;
    }
    __syncthreads();
    if (warp == 0) {
      int wc = sw[0][lane], wi = sw[1][lane];
      int sc = wc, si = wi;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int tc = __shfl_up_sync(NLSH_FULL_MASK, sc, o);
        const int ti = __shfl_up_sync(NLSH_FULL_MASK, si, o);
        if (lane >= o) {
          sc += tc;
          si += ti;
        }
      }
      sw[0][lane] = sc - wc;
      sw[1][lane] = si - wi;
      if (lane == 31) {
        sw[0][32] = sc;
        sw[1][32] = si;
      }
    }
    __syncthreads();
    if (b < n_buckets) {
      pair_off[b] = carry_p + sw[0][warp] + ic - c;
      item_off[b] = carry_i + sw[1][warp] + ii - items;
    }
    carry_p += sw[0][32];
    carry_i += sw[1][32];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pair_off[n_buckets] = carry_p;
    item_off[n_buckets] = carry_i;
  }
}

__global__ void plan_scatter_kernel(const int* __restrict__ probes, const int* __restrict__ offsets,
                                    int n_buckets, int p, long long n_pairs,
                                    const int* __restrict__ pair_off, int* __restrict__ cursor,
                                    int* __restrict__ pairs) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_pairs) return;
  int b;
  if (probe_valid(probes, offsets, n_buckets, p, f, b))
    pairs[pair_off[b] + atomicAdd(&cursor[b], 1)] = (int)f;
}

// Query staging copy: out[row, 0:d_pad] = q[row, 0:d] (zero padded), optionally scaled by
// 1 / max(|q|, eps) (mode 1: eps = 1e-8, the cosine_similarity clamp; mode 2: no clamp, as
// precompute._cosine_distance).  Rows of `out` are 16-byte aligned for the bulk copies.
__global__ void __launch_bounds__(128)
    prepare_queries_kernel(const float* __restrict__ q, long long n, int d, int d_pad, int mode,
                           float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = lane_id();
  float scale = 1.f;
  if (mode != 0) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float v = q[row * d + c];
      ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(NLSH_FULL_MASK, ss, o);
    float nrm = sqrtf(ss);
    if (mode == 1) nrm = fmaxf(nrm, 1e-8f);
    for (int c = lane; c < d_pad; c += 32) out[row * d_pad + c] = c < d ? q[row * d + c] / nrm : 0.f;
    return;
  }
  for (int c = lane; c < d_pad; c += 32) out[row * d_pad + c] = c < d ? q[row * d + c] * scale : 0.f;
}

// One warp per query: merge the partial lists of its probes (or of the dense row blocks).
template <int KPL>
__global__ void __launch_bounds__(128)
    merge_partials_kernel(const float* __restrict__ part_d, const int* __restrict__ part_id,
                          const int* __restrict__ probes, const int* __restrict__ offsets,
                          int n_buckets, int p, int k, int rchunk, int max_chunks, int dense,
                          int sqrt_scores, long long n_queries, long long id_offset,
                          long long* __restrict__ ids_out,
                          float* __restrict__ dists_out, int* __restrict__ ncand_out) {
  const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int lane = lane_id();
  WarpTopK<KPL, int> top;
  top.init(NLSH_ID_SENTINEL);
  int ncand = 0;
  const int n_outer = dense ? 1 : p;
  for (int j = 0; j < n_outer; ++j) {
    long long f;
    int nch;
    if (dense) {
      f = q;
      nch = max_chunks;
    } else {
      f = q * p + j;
      int b;
      if (!probe_valid(probes, offsets, n_buckets, p, f, b)) continue;
      const int size = offsets[b + 1] - offsets[b];
      ncand += size;
      nch = chunk_count(size, rchunk, max_chunks);
    }
    for (int c = 0; c < nch; ++c) {
      const size_t base = ((size_t)f * max_chunks + c) * k;
      for (int e0 = 0; e0 < k; e0 += 32) {
        const int e = e0 + lane;
        float cd = 0.f;
        int cid = NLSH_ID_SENTINEL;
        if (e < k) {
          cd = part_d[base + e];
          cid = part_id[base + e];
        }
        top.offer(cd, cid, cid != NLSH_ID_SENTINEL, k);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < k) {
      const int id = top.id[j];
      ids_out[q * k + pos] = (id == NLSH_ID_SENTINEL) ? -1ll : (long long)id + id_offset;
      dists_out[q * k + pos] = sqrt_scores ? sqrtf(top.d[j]) : top.d[j];
    }
  }
  if (lane == 0 && ncand_out) ncand_out[q] = ncand;
}

// Merge of the tensor-core scan (scan_tc.cu): one warp per query selects the k best (distance, id)
// pairs of its candidate buffer.  A query whose buffer overflowed (more than `cap` rows within its
// bound: duplicates, or a seed sample that was far too small for a skewed bucket) is re-scanned
// exactly afterwards by the whole block - every warp scores a share of the rows of the query's probed
// buckets, one row per lane, with the same thread-per-row arithmetic, and the warps' lists are merged
// through shared memory.  Slow, but complete.
constexpr int kMergeWarps = 8;

template <int KPL>
__global__ void __launch_bounds__(32 * kMergeWarps)
    merge_cands_kernel(const TcCand* __restrict__ cand, const int* __restrict__ cand_n, int cap,
                       const float* __restrict__ qn, const float* __restrict__ xs,
                       const int* __restrict__ row_ids, const int* __restrict__ probes,
                       const int* __restrict__ offsets, int n_buckets, int p, int k, int d, int d_pad,
                       int metric, int sqrt_out, long long n_queries, long long id_offset, const float* __restrict__ tau_g,
                       long long* __restrict__ ids_out, float* __restrict__ dists_out,
                       int* __restrict__ ncand_out, unsigned long long* __restrict__ stats) {
  __shared__ int ovf_q[kMergeWarps];
  __shared__ int ovf_count;
  __shared__ float sh_d[kMergeWarps][32 * KPL];
  __shared__ int sh_id[kMergeWarps][32 * KPL];
  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  if (threadIdx.x == 0) ovf_count = 0;
  __syncthreads();
  const long long q = (long long)blockIdx.x * kMergeWarps + warp;
  if (q < n_queries) {  // warp-uniform
    const int n = cand_n[q];
    int ncand = 0;
    for (int j = 0; j < p; ++j) {
      int b;
      if (probe_valid(probes, offsets, n_buckets, p, q * p + j, b)) ncand += offsets[b + 1] - offsets[b];
    }
    if (lane == 0 && ncand_out) ncand_out[q] = ncand;
    if (n <= cap) {
      WarpTopK<KPL, int> top;
      top.init(NLSH_ID_SENTINEL);
      const TcCand* cq = cand + (size_t)q * cap;
      const float bound = tau_g[q];  // the final bound: candidates appended under an earlier, looser one drop out
      for (int e0 = 0; e0 < n; e0 += 32) {
        const int e = e0 + lane;
        TcCand c;
        c.d = 0.f;
        c.id = NLSH_ID_SENTINEL;
        if (e < n) c = cq[e];
        const bool ok = e < n && c.d <= bound;
        if (e0 == 0)
          top.seed32(c.d, c.id, ok, NLSH_ID_SENTINEL, k);
        else
          top.offer(c.d, c.id, ok, k);
      }
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int pos = j * 32 + lane;
        if (pos < k) {
          const int id = top.id[j];
          ids_out[q * k + pos] = (id == NLSH_ID_SENTINEL) ? -1ll : (long long)id + id_offset;
          dists_out[q * k + pos] = sqrt_out ? sqrtf(top.d[j]) : top.d[j];
        }
      }
    } else if (lane == 0) {
      ovf_q[atomicAdd(&ovf_count, 1)] = (int)(q - (long long)blockIdx.x * kMergeWarps);
      if (stats != nullptr) atomicAdd(stats + 5, 1ull);  // [5] queries re-scanned after an overflow
    }
  }
  __syncthreads();
  const int n_ovf = ovf_count;  // block-uniform
  for (int o = 0; o < n_ovf; ++o) {
    const long long oq = (long long)blockIdx.x * kMergeWarps + ovf_q[o];
    TcQueryGlobal qg;
    qg.q = qn + (size_t)oq * d_pad;
    WarpTopK<KPL, int> top;
    top.init(NLSH_ID_SENTINEL);
    for (int j = 0; j < p; ++j) {
      int b;
      if (!probe_valid(probes, offsets, n_buckets, p, oq * p + j, b)) continue;
      const int r0 = offsets[b], r1 = offsets[b + 1];
      for (int base = r0 + 32 * warp; base < r1; base += 32 * kMergeWarps) {
        const int row = base + lane;
        float dist = 0.f;
        int id = NLSH_ID_SENTINEL;
        if (row < r1) {
          dist = metric == NLSH_METRIC_L2
                     ? tc_thread_distance<NLSH_METRIC_L2, 16, false>(xs + (size_t)row * d_pad, qg, d)
                     : tc_thread_distance<NLSH_METRIC_ANGULAR, 16, false>(xs + (size_t)row * d_pad, qg, d);
          id = row_ids[row];
        }
        top.offer(dist, id, row < r1, k);
      }
    }
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      sh_d[warp][j * 32 + lane] = top.d[j];
      sh_id[warp][j * 32 + lane] = top.id[j];
    }
    __syncthreads();
    if (warp == 0) {
      for (int w = 1; w < kMergeWarps; ++w) {
#pragma unroll
        for (int j = 0; j < KPL; ++j)
          top.offer(sh_d[w][j * 32 + lane], sh_id[w][j * 32 + lane], sh_id[w][j * 32 + lane] != NLSH_ID_SENTINEL, k);
      }
#pragma unroll
      for (int j = 0; j < KPL; ++j) {
        const int pos = j * 32 + lane;
        if (pos < k) {
          const int id = top.id[j];
          ids_out[oq * k + pos] = (id == NLSH_ID_SENTINEL) ? -1ll : (long long)id + id_offset;
          dists_out[oq * k + pos] = sqrt_out ? sqrtf(top.d[j]) : top.d[j];
        }
      }
    }
    __syncthreads();
  }
}

// The same merge with a whole block per query, for batches of few queries with long candidate buffers (config
// 5: 1000 queries x ~1700 candidates, k = 100 - one warp per query would leave most of the GPU idle and walk
// 50 batches of three-register-deep inserts in sequence): warp w takes the candidate batches w, w + 8, ...;
// the warps' lists are merged through shared memory.  An overflowed query is re-scanned the same way.
template <int KPL>
__global__ void __launch_bounds__(32 * kMergeWarps)
    merge_cands_block_kernel(const TcCand* __restrict__ cand, const int* __restrict__ cand_n, int cap,
                             const float* __restrict__ qn, const float* __restrict__ xs,
                             const int* __restrict__ row_ids, const int* __restrict__ probes,
                             const int* __restrict__ offsets, int n_buckets, int p, int k, int d, int d_pad,
                             int metric, int sqrt_out, long long n_queries, long long id_offset,
                             const float* __restrict__ tau_g, long long* __restrict__ ids_out,
                             float* __restrict__ dists_out, int* __restrict__ ncand_out,
                             unsigned long long* __restrict__ stats) {
  __shared__ float sh_d[kMergeWarps][32 * KPL];
  __shared__ int sh_id[kMergeWarps][32 * KPL];
  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  const long long q = blockIdx.x;
  if (q >= n_queries) return;
  const int n = cand_n[q];
  WarpTopK<KPL, int> top;
  top.init(NLSH_ID_SENTINEL);
  int ncand = 0;
  if (n <= cap) {
    const TcCand* cq = cand + (size_t)q * cap;
    const float bound = tau_g[q];
    for (int e0 = 32 * warp; e0 < n; e0 += 32 * kMergeWarps) {
      const int e = e0 + lane;
      TcCand c;
      c.d = 0.f;
      c.id = NLSH_ID_SENTINEL;
      if (e < n) c = cq[e];
      top.offer(c.d, c.id, e < n && c.d <= bound, k);
    }
  }
  for (int j = 0; j < p; ++j) {
    int b;
    if (!probe_valid(probes, offsets, n_buckets, p, q * p + j, b)) continue;
    const int r0 = offsets[b], r1 = offsets[b + 1];
    ncand += r1 - r0;
    if (n > cap) {  // block-uniform: the exact re-scan of an overflowed query
      TcQueryGlobal qg;
      qg.q = qn + (size_t)q * d_pad;
      for (int base = r0 + 32 * warp; base < r1; base += 32 * kMergeWarps) {
        const int row = base + lane;
        float dist = 0.f;
        int id = NLSH_ID_SENTINEL;
        if (row < r1) {
          dist = metric == NLSH_METRIC_L2
                     ? tc_thread_distance<NLSH_METRIC_L2, 16, false>(xs + (size_t)row * d_pad, qg, d)
                     : tc_thread_distance<NLSH_METRIC_ANGULAR, 16, false>(xs + (size_t)row * d_pad, qg, d);
          id = row_ids[row];
        }
        top.offer(dist, id, row < r1, k);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    sh_d[warp][j * 32 + lane] = top.d[j];
    sh_id[warp][j * 32 + lane] = top.id[j];
  }
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < kMergeWarps; ++w) {
#pragma unroll
    for (int j = 0; j < KPL; ++j)
      top.offer(sh_d[w][j * 32 + lane], sh_id[w][j * 32 + lane], sh_id[w][j * 32 + lane] != NLSH_ID_SENTINEL, k);
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < k) {
      const int id = top.id[j];
      ids_out[q * k + pos] = (id == NLSH_ID_SENTINEL) ? -1ll : (long long)id + id_offset;
      dists_out[q * k + pos] = sqrt_out ? sqrtf(top.d[j]) : top.d[j];
    }
  }
  if (lane == 0) {
    if (ncand_out) ncand_out[q] = ncand;
    if (stats != nullptr && n > cap) atomicAdd(stats + 5, 1ull);
  }
}

// Cross-shard merge (after the NCCL all-gather): lists [n_lists, n_queries, k].
template <int KPL>
__global__ void __launch_bounds__(128)
    merge_lists_kernel(const float* __restrict__ dists, const long long* __restrict__ ids,
                       long long dist_stride, long long id_stride, const int* __restrict__ ncand,
                       long long ncand_stride, int n_lists, long long n_queries, int k,
                       long long* __restrict__ ids_out, float* __restrict__ dists_out,
                       int* __restrict__ ncand_out) {
  const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int lane = lane_id();
  const long long sentinel = 0x7fffffffffffffffll;
  WarpTopK<KPL, long long> top;
  top.init(sentinel);
  int total = 0;
  for (int l = 0; l < n_lists; ++l) {
    const float* dl = dists + (size_t)l * dist_stride + (size_t)q * k;
    const long long* il = ids + (size_t)l * id_stride + (size_t)q * k;
    if (ncand) total += ncand[(size_t)l * ncand_stride + q];
    for (int e0 = 0; e0 < k; e0 += 32) {
      const int e = e0 + lane;
      float cd = 0.f;
      long long cid = -1;
      if (e < k) {
        cd = dl[e];
        cid = il[e];
      }
      top.offer(cd, cid, cid >= 0, k);
    }
  }
  if (lane == 0 && ncand_out) ncand_out[q] = total;
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < k) {
      const long long id = top.id[j];
      ids_out[q * k + pos] = (id == sentinel) ? -1ll : id;
      dists_out[q * k + pos] = top.d[j];
    }
  }
}

__global__ void fill_int_kernel(int* __restrict__ p, size_t n, int v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void recall_hits_kernel(const long long* __restrict__ gt, int k_gt,
                                   const long long* __restrict__ pred, int k_pred,
                                   long long n_queries, int* __restrict__ hits) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  // |set(gt) & set(pred)| (nlsh/metrics.py:4-7): count distinct gt ids present in pred
  int h = 0;
  for (int i = 0; i < k_gt; ++i) {
    const long long g = gt[q * k_gt + i];
    bool dup = false;
    for (int i2 = 0; i2 < i; ++i2) dup |= (gt[q * k_gt + i2] == g);
    if (dup) continue;
    bool found = false;
    for (int j = 0; j < k_pred; ++j) found |= (pred[q * k_pred + j] == g);
    h += found ? 1 : 0;
  }
  hits[q] = h;
}

// ---- launch helpers ------------------------------------------------------------------------
template <int METRIC, int KPL>
int launch_scan(const ScanArgs& a, const CUtensorMap& tmap, const ScanGeom& g, bool async, int grid,
                cudaStream_t st) {
  if (async) {
    auto kern = scan_kernel<METRIC, KPL>;
    NLSH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)g.smem_bytes));
    kern<<<grid, 32 * (kConsumerWarps + 1), g.smem_bytes, st>>>(a, tmap);
  } else {
    auto kern = scan_kernel_sync<METRIC, KPL>;
    NLSH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)g.smem_bytes));
    kern<<<grid, 32 * kConsumerWarps, g.smem_bytes, st>>>(a);
  }
  return nlsh_check_cuda(nlsh_post_launch(), "scan_kernel launch");
}

template <int METRIC>
int launch_scan_k(const ScanArgs& a, const CUtensorMap& tmap, const ScanGeom& g, bool async, int grid,
                  cudaStream_t st) {
  if (a.k <= 32) return launch_scan<METRIC, 1>(a, tmap, g, async, grid, st);
  if (a.k <= 64) return launch_scan<METRIC, 2>(a, tmap, g, async, grid, st);
  return launch_scan<METRIC, 4>(a, tmap, g, async, grid, st);
}

// The 2-D tensor map of the row-major [n_rows, d_pad] fp32 matrix the scan streams:
// box = kBoxCols floats (128 bytes) x kTileRows rows, SWIZZLE_128B, zero fill out of bounds.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_row_tensor_map(CUtensorMap* tmap, const float* base, long long n_rows, int d_pad) {
  static EncodeTiledFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NLSH_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      nlsh_set_error("cuTensorMapEncodeTiled is not available from this driver");
      return NLSH_ERR_CUDA;
    }
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)(n_rows > 0 ? n_rows : 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)d_pad * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kBoxCols, (cuuint32_t)kTileRows};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims,
                            strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    nlsh_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld, d_pad=%d)", (int)r,
                   n_rows, d_pad);
    return NLSH_ERR_CUDA;
  }
  return NLSH_OK;
}

int launch_scan_metric(int metric, const ScanArgs& a, const ScanGeom& g, bool async, int grid,
                       cudaStream_t st) {
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (async) {
    const int rc = make_row_tensor_map(&tmap, a.xs, a.n_rows, a.d_pad);
    if (rc != NLSH_OK) return rc;
  }
  switch (metric) {
    case NLSH_METRIC_L2: return launch_scan_k<NLSH_METRIC_L2>(a, tmap, g, async, grid, st);
    case NLSH_METRIC_ANGULAR: return launch_scan_k<NLSH_METRIC_ANGULAR>(a, tmap, g, async, grid, st);
    case NLSH_METRIC_L2SQ: return launch_scan_k<NLSH_METRIC_L2SQ>(a, tmap, g, async, grid, st);
    default: return launch_scan_k<NLSH_METRIC_COSINE>(a, tmap, g, async, grid, st);
  }
}

int launch_merge_partials(const float* part_d, const int* part_id, const int* probes,
                          const int* offsets, int n_buckets, int p, int k, int rchunk,
                          int max_chunks, int dense, int sqrt_scores, int64_t n_queries,
                          int64_t id_offset,
                          int64_t* ids_out, float* dists_out, int* ncand_out, cudaStream_t st) {
  const unsigned blocks = (unsigned)((n_queries + 3) / 4);
  long long* ids_ll = reinterpret_cast<long long*>(ids_out);
  if (k <= 32)
    merge_partials_kernel<1><<<blocks, 128, 0, st>>>(part_d, part_id, probes, offsets, n_buckets, p,
                                                     k, rchunk, max_chunks, dense, sqrt_scores,
                                                     n_queries, id_offset, ids_ll, dists_out, ncand_out);
  else if (k <= 64)
    merge_partials_kernel<2><<<blocks, 128, 0, st>>>(part_d, part_id, probes, offsets, n_buckets, p,
                                                     k, rchunk, max_chunks, dense, sqrt_scores,
                                                     n_queries, id_offset, ids_ll, dists_out, ncand_out);
  else
    merge_partials_kernel<4><<<blocks, 128, 0, st>>>(part_d, part_id, probes, offsets, n_buckets, p,
                                                     k, rchunk, max_chunks, dense, sqrt_scores,
                                                     n_queries, id_offset, ids_ll, dists_out, ncand_out);
  return nlsh_check_cuda(nlsh_post_launch(), "merge_partials_kernel launch");
}

struct QueryWorkspace {
  int* cnt;       // [B]   (cnt, cursor, counter, cand_n, ladder are zeroed with one memset)
  int* cursor;    // [B]
  int* counter;   // [64]  work-queue counter (+ the tensor-core scan's debug counters)
  int* cand_n;    // [Q]       tensor-core scan: candidates appended per query
  int* ladder;    // [Q * kTcLadder]  tensor-core scan: threshold ladder
  int* pair_off;  // [B+1]
  int* item_off;  // [B+1]
  int* pairs;     // [Q*p]
  float* qn;      // [Q*d_pad] staged queries
  ItemRec* items; // [max_items]
  int max_items;
  float* part_d;
  int* part_id;
  // tensor-core scan extras
  float* qs;        // [Q*p (+ kTcNQ), d_pad] queries in pair order
  int* pq;          // [Q*p] query index of each pair
  float* pqn2;      // [Q*p] |q|^2 of each pair's query
  float* tau_g;     // [Q]
  float* tau0;      // [Q]
  TcCand* cand;     // [Q * cand_cap]
  int cand_cap;
  TcItem* tc_items; // [max_tc_items]
  int max_tc_items;
  size_t zero_ints;
  size_t total;
};

// Candidate-buffer entries per query of the tensor-core scan.  The scan appends every row whose exact
// distance is within the query's bound at that moment.  Even a bound that always equals the running k-th
// best lets k (1 + ln(C / k)) of C candidates in random order through (the expected number of updates of a
// streaming top-k); the seed's bound starts looser (about k * first bucket rows / sample rows pass from the
// first bucket) and the threshold ladder tightens it in steps of 1/32.  The buffer holds 5x the streaming
// figure for C = 2 p * (average bucket), at least 512 entries, rounded up to a power of two: 512 for config 4
// (k = 10, C = 39k), 8192 for config 5 (k = 100, C = 500k).  A query that still overflows is re-scanned
// exactly by merge_cands_kernel.  NLSH_TC_CAND_CAP overrides (tests force overflows with it).
int tc_cand_cap(int k, int p, int64_t n_rows, int n_buckets) {
  double c = 2.0 * p * (double)(n_rows / (n_buckets > 0 ? n_buckets : 1) + 1);
  if (c < 2.0 * k) c = 2.0 * k;
  const double want = 5.0 * k * (1.0 + log(c / k));
  int cap = 512;
  while (cap < want && cap < (1 << 16)) cap *= 2;
  if (const char* env = getenv("NLSH_TC_CAND_CAP")) cap = atoi(env);
  if (cap < k) cap = k;
  if (cap > (1 << 16)) cap = 1 << 16;
  return cap;
}

// Upper bound of the work items of a batch: sum_b ceil(nq_b / group) * nch_b
// <= (pairs / group + #probed buckets) * max_chunks.  int64: the caller refuses what does not fit an int.
int64_t item_bound(int64_t pairs, int n_buckets, int group, int max_chunks) {
  return (pairs / group + (pairs < n_buckets ? pairs : n_buckets) + 1) * (int64_t)max_chunks;
}

// max_chunks / tc_max_chunks: chunks of the largest bucket under the SIMT / tensor-core policy
// (tc_max_chunks = 0: no tensor-core extras).
QueryWorkspace carve_query_ws(void* base, int64_t nq, int p, int k, int d, int n_buckets,
                              int max_chunks, int tc_max_chunks, int64_t n_rows) {
  QueryWorkspace w;
  WorkspaceCarver ws(base);
  const bool tc = tc_max_chunks > 0;
  const size_t head = ((size_t)2 * n_buckets + 64 + 3) / 4 * 4;  // ladder rows stay 16-byte aligned
  w.zero_ints = head + (tc ? (size_t)nq * (1 + kTcLadder) : 0);
  int* z = ws.take<int>(w.zero_ints);
  w.cnt = z;
  w.cursor = z ? z + n_buckets : nullptr;
  w.counter = z ? z + 2 * n_buckets : nullptr;
  w.ladder = (z && tc) ? z + head : nullptr;
  w.cand_n = (z && tc) ? z + head + (size_t)nq * kTcLadder : nullptr;
  w.pair_off = ws.take<int>((size_t)n_buckets + 1);
  w.item_off = ws.take<int>((size_t)n_buckets + 1);
  w.pairs = ws.take<int>((size_t)nq * p);
  w.qn = ws.take<float>((size_t)nq * ((d + 3) / 4 * 4));
  const int64_t pairs = nq * (int64_t)p;
  const int64_t item_cap = 1ll << 30;  // nlsh_query_scan_topk refuses batches beyond it
  int64_t mi = item_bound(pairs, n_buckets, kG, max_chunks);
  w.max_items = (int)(mi < item_cap ? mi : item_cap);
  w.items = ws.take<ItemRec>((size_t)w.max_items);
  const size_t lists = (size_t)nq * p * max_chunks * k;  // partial lists of the SIMT scan
  w.part_d = ws.take<float>(lists);
  w.part_id = ws.take<int>(lists);
  w.qs = w.pqn2 = w.tau_g = w.tau0 = nullptr;
  w.pq = nullptr;
  w.cand = nullptr;
  w.cand_cap = 0;
  w.tc_items = nullptr;
  w.max_tc_items = 0;
  if (tc) {
    const size_t d_pad = (size_t)((d + 3) / 4 * 4);
    w.qs = ws.take<float>((size_t)(pairs + kTcNQMax) * d_pad);
    w.pq = ws.take<int>((size_t)pairs + kTcNQMax);
    w.pqn2 = ws.take<float>((size_t)pairs + kTcNQMax);
    w.tau_g = ws.take<float>((size_t)nq);
    w.tau0 = ws.take<float>((size_t)nq);
    w.cand_cap = tc_cand_cap(k, p, n_rows, n_buckets);
    w.cand = ws.take<TcCand>((size_t)nq * w.cand_cap);
    mi = item_bound(pairs, n_buckets, kTcNQ, tc_max_chunks);
    w.max_tc_items = (int)(mi < item_cap ? mi : item_cap);
    w.tc_items = ws.take<TcItem>((size_t)w.max_tc_items);
  }
  w.total = ws.total();
  return w;
}

struct KnnPlan {
  int n_blocks;
  int rchunk;
  int qgroups;
};

KnnPlan knn_plan(int64_t n_queries, int64_t n_rows) {
  KnnPlan kp;
  kp.qgroups = (int)((n_queries + kG - 1) / kG);
  if (kp.qgroups < 1) kp.qgroups = 1;
  const int64_t grid = (int64_t)nlsh_num_sms() * 2;
  int64_t blocks = (grid * 8 + kp.qgroups - 1) / kp.qgroups;
  const int64_t max_blocks = (n_rows + kTileRows - 1) / kTileRows;
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks > 256) blocks = 256;
  if (blocks < 1) blocks = 1;
  int64_t rchunk = (n_rows + blocks - 1) / blocks;
  rchunk = (rchunk + kTileRows - 1) / kTileRows * kTileRows;
  if (rchunk < kTileRows) rchunk = kTileRows;
  kp.rchunk = (int)rchunk;
  kp.n_blocks = (int)((n_rows + rchunk - 1) / rchunk);
  if (kp.n_blocks < 1) kp.n_blocks = 1;
  return kp;
}

}  // namespace

namespace {
// NLSH_SCAN_IMPL=simt forces the fp32 SIMT scan kernel (A/B runs).
bool scan_use_tc(int d, int k, int metric) {
  const char* env = getenv("NLSH_SCAN_IMPL");
  if (env != nullptr && strcmp(env, "simt") == 0) return false;
  return nlsh_scan_tc_supported(d, k, metric);
}
}  // namespace

namespace {
// The tensor-core filter pays off when a bucket tile serves several queries (one GEMM tile scores
// 32 of them); with about one query per probed bucket the fp32 SIMT kernel already streams at the
// HBM rate (0.94 of the copy peak) and has no seed / re-rank stages, so it keeps those batches.
bool scan_batch_prefers_tc(int64_t n_queries, int32_t p, int32_t n_buckets) {
  const char* env = getenv("NLSH_SCAN_IMPL");
  if (env != nullptr && strcmp(env, "tc") == 0) return true;
  return n_queries * (int64_t)p >= 4 * (int64_t)n_buckets;
}
}  // namespace

namespace {
// Queries per item of the tensor-core scan: 128 for wide rows (d > 128) when the batch puts 64 or more
// (query, probe) pairs on a probed bucket (config 5: 6.3 against 7.0 ms), else 32 (narrow rows measured equal
// or better with 32: config 1 0.083 against 0.104 ms, config 2 equal).  NLSH_TC_NQ=32/128 overrides (A/B runs).
int tc_group_size(int64_t n_queries, int32_t p, int32_t n_buckets, int32_t d) {
  const int64_t pairs = n_queries * (int64_t)p;
  const int64_t distinct = pairs < n_buckets ? pairs : n_buckets;
  int group = (d > 128 && distinct > 0 && pairs >= 64 * distinct) ? kTcNQMax : kTcNQ;
  if (const char* env = getenv("NLSH_TC_NQ")) group = atoi(env) == kTcNQMax ? kTcNQMax : kTcNQ;
  return group;
}
}  // namespace

extern "C" int nlsh_query_scan_impl(int32_t d, int32_t k, int32_t metric, int32_t has_sqnorm,
                                    int64_t n_queries, int32_t p, int32_t n_buckets) {
  return (has_sqnorm && scan_use_tc(d, k, metric) && scan_batch_prefers_tc(n_queries, p, n_buckets)) ? 1 : 0;
}

extern "C" size_t nlsh_query_workspace_bytes(int64_t n_queries, int32_t p, int32_t k, int32_t d,
                                             int32_t n_buckets, int64_t n_rows,
                                             int64_t max_bucket_rows) {
  if (n_queries < 0 || p < 1 || k < 1 || d < 1 || n_buckets < 1) return 0;
  const ScanPolicy pol = scan_policy(n_queries, p, n_buckets, n_rows, max_bucket_rows);
  int tc_chunks = 0;  // the metric is not an argument: sized for either scan implementation
  if (scan_use_tc(d, k, NLSH_METRIC_L2))
    tc_chunks = scan_policy(n_queries, p, n_buckets, n_rows, max_bucket_rows, tc_group_size(n_queries, p, n_buckets, d), 1,
                            4ll * ((d + 3) / 4 * 4)).max_chunks;
  return carve_query_ws(nullptr, n_queries, p, k, d, n_buckets, pol.max_chunks, tc_chunks, n_rows).total;
}

extern "C" size_t nlsh_query_seed_workspace_bytes(int64_t n_queries, int32_t d) {
  if (n_queries < 0 || d < 1) return 0;
  WorkspaceCarver ws(nullptr);
  ws.take<float>((size_t)n_queries * ((d + 3) / 4 * 4));
  ws.take<float>((size_t)n_queries);
  return ws.total();
}

extern "C" int nlsh_query_seed_tau_rows(const float* xq, int64_t n_queries, int32_t d, const int32_t* probes,
                                        int32_t p, const int32_t* offsets, int32_t n_buckets,
                                   const float* x_sorted, int64_t n_rows, int32_t metric, int32_t k,
                                   int32_t sample_rows, float* tau_out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  NLSH_REQUIRE(n_queries >= 0 && d >= 1 && d <= 16384 && p >= 1 && p <= 1024, "seed: bad shape");
  NLSH_REQUIRE(k >= 1 && k <= NLSH_MAX_K, "seed: k=%d outside [1, %d]", k, NLSH_MAX_K);
  NLSH_REQUIRE(n_buckets >= 1 && n_buckets <= (1 << 20), "seed: n_buckets=%d outside [1, 2^20]", n_buckets);
  NLSH_REQUIRE(metric == NLSH_METRIC_L2 || metric == NLSH_METRIC_ANGULAR, "seed: metric %d is not a scan metric",
               metric);
  if (n_queries == 0) return NLSH_OK;
  NLSH_REQUIRE(xq && probes && offsets && tau_out && (n_rows == 0 || x_sorted), "seed: null pointer");
  NLSH_REQUIRE(nlsh_scan_tc_supported(d, k, metric), "seed: shape outside the tensor-core scan (d=%d, k=%d)", d, k);
  const size_t need = nlsh_query_seed_workspace_bytes(n_queries, d);
  if (workspace == nullptr || workspace_bytes < need) {
    nlsh_set_error("seed: workspace %zu bytes < required %zu", workspace_bytes, need);
    return NLSH_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d_pad = (d + 3) / 4 * 4;
  WorkspaceCarver ws(workspace);
  float* qn = ws.take<float>((size_t)n_queries * d_pad);
  float* tau_g = ws.take<float>((size_t)n_queries);
  prepare_queries_kernel<<<(unsigned)((n_queries + 3) / 4), 128, 0, st>>>(
      xq, n_queries, d, d_pad, metric == NLSH_METRIC_ANGULAR ? 1 : 0, qn);
  NLSH_CUDA_TRY(nlsh_post_launch());
  NLSH_REQUIRE(sample_rows >= 0 && sample_rows <= 4096, "seed: sample_rows=%d outside [0, 4096]", sample_rows);
  return nlsh_scan_tc_seed(qn, n_queries, probes, p, offsets, x_sorted, n_rows, n_buckets, d, d_pad, k, metric,
                           tau_g, tau_out, st, sample_rows);
}

extern "C" int nlsh_query_seed_tau(const float* xq, int64_t n_queries, int32_t d, const int32_t* probes,
                                   int32_t p, const int32_t* offsets, int32_t n_buckets,
                                   const float* x_sorted, int64_t n_rows, int32_t metric, int32_t k,
                                   float* tau_out, void* workspace, size_t workspace_bytes, void* stream) {
  return nlsh_query_seed_tau_rows(xq, n_queries, d, probes, p, offsets, n_buckets, x_sorted, n_rows, metric, k, 0,
                                  tau_out, workspace, workspace_bytes, stream);
}

extern "C" int nlsh_query_scan_topk(const float* xq, int64_t n_queries, int32_t d,
                                    const int32_t* probes, int32_t p, const int32_t* offsets,
                                    int32_t n_buckets, const int32_t* ids, const float* x_sorted,
                                    const float* x_sqnorm, int64_t n_rows, int64_t max_bucket_rows,
                                    int32_t metric,
                                    int32_t k, int64_t id_offset, int64_t* ids_out,
                                    float* dists_out, int32_t* ncand_out, void* workspace,
                                    size_t workspace_bytes, uint32_t flags, void* stream) {
  return nlsh_query_scan_topk_seeded(xq, n_queries, d, probes, p, offsets, n_buckets, ids, x_sorted, x_sqnorm,
                                     n_rows, max_bucket_rows, metric, k, id_offset, nullptr, ids_out, dists_out,
                                     ncand_out, workspace, workspace_bytes, flags, stream);
}

extern "C" int nlsh_query_scan_topk_seeded(const float* xq, int64_t n_queries, int32_t d,
                                           const int32_t* probes, int32_t p, const int32_t* offsets,
                                           int32_t n_buckets, const int32_t* ids, const float* x_sorted,
                                           const float* x_sqnorm, int64_t n_rows, int64_t max_bucket_rows,
                                           int32_t metric, int32_t k, int64_t id_offset,
                                           const float* tau_seed, int64_t* ids_out, float* dists_out,
                                           int32_t* ncand_out, void* workspace, size_t workspace_bytes,
                                           uint32_t flags, void* stream) {
  NLSH_REQUIRE(n_queries >= 0 && n_queries * (int64_t)p < (1ll << 31),
               "query: n_queries=%lld x p=%d outside [0, 2^31)", (long long)n_queries, p);
  NLSH_REQUIRE(d >= 1 && d <= 16384, "query: d=%d outside [1, 16384]", d);
  NLSH_REQUIRE(p >= 1 && p <= 1024, "query: p=%d outside [1, 1024]", p);
  NLSH_REQUIRE(k >= 1 && k <= NLSH_MAX_K, "query: k=%d outside [1, %d]", k, NLSH_MAX_K);
  NLSH_REQUIRE(n_buckets >= 1 && n_buckets <= (1 << 20), "query: n_buckets=%d outside [1, 2^20]",
               n_buckets);
  NLSH_REQUIRE(metric == NLSH_METRIC_L2 || metric == NLSH_METRIC_ANGULAR,
               "query: metric %d is not a scan metric (L2=0 / ANGULAR=1)", metric);
  NLSH_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "query: n_rows=%lld outside [0, 2^31)",
               (long long)n_rows);
  if (n_queries == 0) return NLSH_OK;
  NLSH_REQUIRE(xq && probes && offsets && ids_out && dists_out, "query: null pointer");
  NLSH_REQUIRE(n_rows == 0 || (ids && x_sorted), "query: null index arrays");
  NLSH_REQUIRE((reinterpret_cast<uintptr_t>(x_sorted) & 15) == 0, "query: x_sorted not 16-byte aligned");

  const ScanPolicy pol_simt = scan_policy(n_queries, p, n_buckets, n_rows, max_bucket_rows);
  const bool tc_sized = scan_use_tc(d, k, NLSH_METRIC_L2);
  ScanPolicy pol_tc = pol_simt;
  const int tc_group = tc_group_size(n_queries, p, n_buckets, d);
  if (tc_sized) pol_tc = scan_policy(n_queries, p, n_buckets, n_rows, max_bucket_rows, tc_group, 1, 4ll * ((d + 3) / 4 * 4));
  {
    const int64_t pairs64 = n_queries * (int64_t)p;
    const int64_t need_items = item_bound(pairs64, n_buckets, kG, pol_simt.max_chunks);
    const int64_t need_tc = tc_sized ? item_bound(pairs64, n_buckets, tc_group, pol_tc.max_chunks) : 0;
    NLSH_REQUIRE(need_items < (1ll << 30) && need_tc < (1ll << 30),
                 "query: batch of %lld queries x %d probes needs %lld work items (limit 2^30): split the batch",
                 (long long)n_queries, p, (long long)(need_items > need_tc ? need_items : need_tc));
  }
  const QueryWorkspace w = carve_query_ws(workspace, n_queries, p, k, d, n_buckets, pol_simt.max_chunks,
                                          tc_sized ? pol_tc.max_chunks : 0, n_rows);
  if (workspace == nullptr || workspace_bytes < w.total) {
    nlsh_set_error("query: workspace %zu bytes < required %zu", workspace_bytes, w.total);
    return NLSH_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool async = (flags & 1u) == 0;
  // L2 lists hold squared distances; the root is taken when a query's list is written - unless the caller
  // merges lists of several shards and takes it afterwards (flags bit 3), so that two candidates whose roots
  // round to the same float keep the order of their squared distances across the shard merge
  const int sqrt_scores = (metric == NLSH_METRIC_L2 && (flags & 8u) == 0) ? 1 : 0;
  // tensor-core filtered scan (scan_tc.cu) whenever the index carries the row norms
  const bool use_tc = tc_sized && async && (flags & 2u) == 0 && x_sqnorm != nullptr && n_rows > 0 &&
                      (reinterpret_cast<uintptr_t>(x_sqnorm) & 15) == 0 && scan_use_tc(d, k, metric) &&
                      scan_batch_prefers_tc(n_queries, p, n_buckets);
  const ScanPolicy pol = use_tc ? pol_tc : pol_simt;
  const ScanGeom geom = scan_geom(d, k, async);
  const long long n_pairs = (long long)n_queries * p;

  NLSH_CUDA_TRY(cudaMemsetAsync(w.cnt, 0, w.zero_ints * sizeof(int), st));
  plan_count_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(probes, offsets, n_buckets, p,
                                                                     n_pairs, w.cnt);
  NLSH_CUDA_TRY(nlsh_post_launch());
  plan_scan_kernel<<<1, 1024, 0, st>>>(w.cnt, offsets, n_buckets, pol.rchunk, pol.max_chunks,
                                       use_tc ? tc_group : kG, w.pair_off, w.item_off);
  NLSH_CUDA_TRY(nlsh_post_launch());
  plan_scatter_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(
      probes, offsets, n_buckets, p, n_pairs, w.pair_off, w.cursor, w.pairs);
  NLSH_CUDA_TRY(nlsh_post_launch());

  prepare_queries_kernel<<<(unsigned)((n_queries + 3) / 4), 128, 0, st>>>(
      xq, n_queries, d, geom.d_pad, metric == NLSH_METRIC_ANGULAR ? 1 : 0, w.qn);
  NLSH_CUDA_TRY(nlsh_post_launch());
  const float* q_used = w.qn;

  if (use_tc) {
    plan_tc_items_kernel<<<nlsh_num_sms() * 4, 256, 0, st>>>(w.item_off, w.pair_off, offsets, n_buckets,
                                                             pol.rchunk, pol.max_chunks, tc_group, w.tc_items,
                                                             w.max_tc_items);
    NLSH_CUDA_TRY(nlsh_post_launch());
    int rc = nlsh_scan_tc_prepare(w.qn, w.pairs, w.pair_off + n_buckets, n_pairs, p, geom.d_pad, w.qs, w.pq,
                                  w.pqn2, w.tau_g, w.tau0, tau_seed, n_queries, probes, offsets, x_sorted, n_rows,
                                  n_buckets, geom.d, k, metric, st);
    if (rc != NLSH_OK) return rc;
    unsigned long long* stats =
        getenv("NLSH_TC_STATS") ? reinterpret_cast<unsigned long long*>(w.counter + 2) : nullptr;
    const char* lad = getenv("NLSH_TC_LADDER");  // NLSH_TC_LADDER=0: seed bound only (A/B runs)
    int cap = w.cand_cap;
    if ((flags & 4u) != 0 && cap > k) cap = k;  // tests: tiny candidate buffers force the overflow re-scan
    TcScanArgs t{};
    t.xs = x_sorted;
    t.xnorm = x_sqnorm;
    t.ids = ids;
    t.qs = w.qs;
    t.pq = w.pq;
    t.pqn2 = w.pqn2;
    t.items = w.tc_items;
    t.n_items = w.item_off + n_buckets;
    t.max_items = w.max_tc_items;
    t.item_counter = w.counter;
    t.stats = stats;
    t.tau_g = w.tau_g;
    t.tau0 = w.tau0;
    t.ladder = (lad != nullptr && atoi(lad) == 0) ? nullptr : w.ladder;
    t.cand_n = w.cand_n;
    t.cand = w.cand;
    t.cap = cap;
    t.n_rows = n_rows;
    t.n_pairs = n_pairs;
    t.k = k;
    t.d = geom.d;
    t.d_pad = geom.d_pad;
    t.sm_reserve = (int)((flags >> 8) & 0xffu);
    t.avg_item_rows = (int)(n_rows / n_buckets);
    t.nq_group = tc_group;
    nlsh_profile_mark(st, true);
    rc = nlsh_scan_tc_launch(metric, t, st);
    nlsh_profile_mark(st, false);
    if (rc != NLSH_OK) return rc;
    const unsigned mblocks = (unsigned)((n_queries + kMergeWarps - 1) / kMergeWarps);
#define NLSH_MERGE_ARGS w.cand, w.cand_n, cap, w.qn, x_sorted, ids, probes, offsets, n_buckets, p, k, geom.d, geom.d_pad, \
                        metric, sqrt_scores, n_queries, id_offset, w.tau_g, reinterpret_cast<long long*>(ids_out), dists_out, ncand_out, stats
    // a block per query when the buffers are long (k > 32) or the batch is too small to fill the GPU with warps
    const bool per_block = k > 32 || n_queries <= 2048;
    if (per_block) {
      const unsigned qb = (unsigned)n_queries;
      if (k <= 32) merge_cands_block_kernel<1><<<qb, 32 * kMergeWarps, 0, st>>>(NLSH_MERGE_ARGS);
      else if (k <= 64) merge_cands_block_kernel<2><<<qb, 32 * kMergeWarps, 0, st>>>(NLSH_MERGE_ARGS);
      else merge_cands_block_kernel<4><<<qb, 32 * kMergeWarps, 0, st>>>(NLSH_MERGE_ARGS);
    } else if (k <= 32) {
      merge_cands_kernel<1><<<mblocks, 32 * kMergeWarps, 0, st>>>(NLSH_MERGE_ARGS);
    } else if (k <= 64) {
      merge_cands_kernel<2><<<mblocks, 32 * kMergeWarps, 0, st>>>(NLSH_MERGE_ARGS);
    } else {
      merge_cands_kernel<4><<<mblocks, 32 * kMergeWarps, 0, st>>>(NLSH_MERGE_ARGS);
    }
#undef NLSH_MERGE_ARGS
    return nlsh_check_cuda(nlsh_post_launch(), "merge_cands_kernel launch");
  }

  ScanArgs a{};
  a.xs = x_sorted;
  a.ids = ids;
  a.offsets = offsets;
  a.q = q_used;
  a.pair_off = w.pair_off;
  a.pairs = w.pairs;
  a.item_off = w.item_off;
  a.items = w.items;
  a.item_counter = w.counter;
  a.part_d = w.part_d;
  a.part_id = w.part_id;
  a.n_rows = n_rows;
  a.self_offset = 0;
  a.n_queries = (int)n_queries;
  a.n_buckets = n_buckets;
  a.p = p;
  a.k = k;
  a.rchunk = pol.rchunk;
  a.max_chunks = pol.max_chunks;
  a.dense = 0;
  a.exclude_self = 0;
  a.d = geom.d;
  a.d_pad = geom.d_pad;
  a.n_boxes = geom.n_boxes;
  a.bps = geom.bps;
  a.nqb = geom.nqb;
  a.n_chunks = geom.n_chunks;
  a.stages = geom.stages;
  const int grid = nlsh_num_sms() * geom.ctas_per_sm;
  plan_items_kernel<<<nlsh_num_sms() * 4, 256, 0, st>>>(a, w.items, w.max_items);
  NLSH_CUDA_TRY(nlsh_post_launch());
  nlsh_profile_mark(st, true);
  int rc = launch_scan_metric(metric, a, geom, async, grid, st);
  nlsh_profile_mark(st, false);
  if (rc != NLSH_OK) return rc;
  return launch_merge_partials(w.part_d, w.part_id, probes, offsets, n_buckets, p, k, pol.rchunk,
                               pol.max_chunks, 0, sqrt_scores, n_queries,
                               id_offset, ids_out, dists_out, ncand_out, st);
}

// ---- tensor-core kNN (tc_knn.cu) -----------------------------------------------------------------
bool nlsh_knn_tc_supported(int d, int metric, int k);
int nlsh_knn_tc_blocks(long long n_queries, long long n_rows, int k);
size_t nlsh_knn_tc_scratch_floats(long long n_queries, long long n_rows, int d);
int nlsh_knn_tc_run(const float* xq, long long n_queries, const float* xdb, long long n_rows, int d,
                    int metric, int k, int exclude_self, long long self_offset, float* scratch,
                    float* part_d, int* part_id, cudaStream_t st);

namespace {
// NLSH_KNN_IMPL=simt forces the fp32 SIMT scan kernel in dense mode (A/B runs); the default for the
// precompute.py metrics (L2SQ / COSINE, the expansion-form distances) is the tensor-core GEMM.
bool knn_use_tc(int d, int metric, int k, int64_t n_rows) {
  const char* env = getenv("NLSH_KNN_IMPL");
  if (env != nullptr && strcmp(env, "simt") == 0) return false;
  return n_rows > 0 && nlsh_knn_tc_supported(d, metric, k);
}

size_t knn_simt_ws(int64_t n_queries, int64_t n_rows, int32_t d, int32_t k) {
  const KnnPlan kp = knn_plan(n_queries, n_rows);
  WorkspaceCarver ws(nullptr);
  ws.take<int>(64);
  ws.take<float>((size_t)n_queries * d);
  ws.take<ItemRec>((size_t)kp.qgroups * kp.n_blocks);
  const size_t lists = (size_t)n_queries * kp.n_blocks * k;
  ws.take<float>(lists);
  ws.take<int>(lists);
  return ws.total();
}

size_t knn_tc_ws(int64_t n_queries, int64_t n_rows, int32_t d, int32_t k) {
  WorkspaceCarver ws(nullptr);
  ws.take<float>(nlsh_knn_tc_scratch_floats(n_queries, n_rows, d));
  const size_t lists = (size_t)n_queries * nlsh_knn_tc_blocks(n_queries, n_rows, k) * k;
  ws.take<float>(lists);
  ws.take<int>(lists);
  return ws.total();
}
}  // namespace

// The metric is not an argument: sized for whichever of the two implementations needs more.
extern "C" size_t nlsh_knn_workspace_bytes(int64_t n_queries, int64_t n_rows, int32_t d, int32_t k) {
  if (n_queries < 0 || n_rows < 0 || d < 1 || k < 1) return 0;
  size_t need = knn_simt_ws(n_queries, n_rows, d, k);
  if (knn_use_tc(d, NLSH_METRIC_L2SQ, k, n_rows)) {
    const size_t tc = knn_tc_ws(n_queries, n_rows, d, k);
    if (tc > need) need = tc;
  }
  return need;
}

extern "C" int nlsh_knn_bruteforce(const float* xq, int64_t n_queries, const float* xdb,
                                   int64_t n_rows, int32_t d, int32_t metric, int32_t k,
                                   int32_t exclude_self, int64_t self_offset, int64_t id_offset,
                                   int64_t* ids_out, float* dists_out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  NLSH_REQUIRE(n_queries >= 0 && n_queries < (1ll << 28), "knn: n_queries=%lld outside [0, 2^28)",
               (long long)n_queries);
  NLSH_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "knn: n_rows=%lld outside [0, 2^31)",
               (long long)n_rows);
  NLSH_REQUIRE(d >= 1 && d <= 16384, "knn: d=%d outside [1, 16384]", d);
  NLSH_REQUIRE(d % 4 == 0, "knn: d=%d must be a multiple of 4 (pad the rows with zeros)", d);
  NLSH_REQUIRE(k >= 1 && k <= NLSH_MAX_K, "knn: k=%d outside [1, %d]", k, NLSH_MAX_K);
  NLSH_REQUIRE(metric >= NLSH_METRIC_L2 && metric <= NLSH_METRIC_COSINE, "knn: unknown metric %d",
               metric);
  if (n_queries == 0) return NLSH_OK;
  NLSH_REQUIRE(xq && ids_out && dists_out, "knn: null pointer");
  NLSH_REQUIRE(n_rows == 0 || xdb, "knn: xdb is NULL");
  NLSH_REQUIRE((reinterpret_cast<uintptr_t>(xdb) & 15) == 0, "knn: xdb not 16-byte aligned");
  const size_t need = nlsh_knn_workspace_bytes(n_queries, n_rows, d, k);
  if (workspace == nullptr || workspace_bytes < need) {
    nlsh_set_error("knn: workspace %zu bytes < required %zu", workspace_bytes, need);
    return NLSH_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (knn_use_tc(d, metric, k, n_rows)) {
    WorkspaceCarver tws(workspace);
    float* scratch = tws.take<float>(nlsh_knn_tc_scratch_floats(n_queries, n_rows, d));
    const int n_blocks = nlsh_knn_tc_blocks(n_queries, n_rows, k);
    const size_t tlists = (size_t)n_queries * n_blocks * k;
    float* t_d = tws.take<float>(tlists);
    int* t_id = tws.take<int>(tlists);
    const int rc = nlsh_knn_tc_run(xq, n_queries, xdb, n_rows, d, metric, k, exclude_self ? 1 : 0,
                                   self_offset, scratch, t_d, t_id, st);
    if (rc != NLSH_OK) return rc;
    return launch_merge_partials(t_d, t_id, nullptr, nullptr, 1, 1, k, 0, n_blocks, 1, 0, n_queries,
                                 id_offset, ids_out, dists_out, nullptr, st);
  }
  const KnnPlan kp = knn_plan(n_queries, n_rows);
  WorkspaceCarver ws(workspace);
  int* counter = ws.take<int>(64);
  float* qn = ws.take<float>((size_t)n_queries * d);
  const int n_items = kp.qgroups * kp.n_blocks;
  ItemRec* items = ws.take<ItemRec>((size_t)n_items);
  const size_t lists = (size_t)n_queries * kp.n_blocks * k;
  float* part_d = ws.take<float>(lists);
  int* part_id = ws.take<int>(lists);

  NLSH_CUDA_TRY(cudaMemsetAsync(counter, 0, 64 * sizeof(int), st));
  const float* q_used = xq;
  const bool normalise = metric == NLSH_METRIC_ANGULAR || metric == NLSH_METRIC_COSINE;
  if (normalise || (reinterpret_cast<uintptr_t>(xq) & 15) != 0) {
    prepare_queries_kernel<<<(unsigned)((n_queries + 3) / 4), 128, 0, st>>>(
        xq, n_queries, d, d, metric == NLSH_METRIC_ANGULAR ? 1 : (metric == NLSH_METRIC_COSINE ? 2 : 0), qn);
    NLSH_CUDA_TRY(nlsh_post_launch());
    q_used = qn;
  }
  const ScanGeom geom = scan_geom(d, k, true);
  ScanArgs a{};
  a.xs = xdb;
  a.ids = nullptr;
  a.q = q_used;
  a.items = items;
  a.item_counter = counter;
  a.part_d = part_d;
  a.part_id = part_id;
  a.n_rows = n_rows;
  a.self_offset = self_offset;
  a.n_queries = (int)n_queries;
  a.n_buckets = 1;
  a.p = 1;
  a.k = k;
  a.rchunk = kp.rchunk;
  a.max_chunks = kp.n_blocks;
  a.dense = 1;
  a.dense_qgroups = kp.qgroups;
  a.dense_items = n_rows > 0 ? kp.qgroups * kp.n_blocks : 0;
  a.exclude_self = exclude_self ? 1 : 0;
  a.d = geom.d;
  a.d_pad = geom.d_pad;
  a.n_boxes = geom.n_boxes;
  a.bps = geom.bps;
  a.nqb = geom.nqb;
  a.n_chunks = geom.n_chunks;
  a.stages = geom.stages;
  const int grid = nlsh_num_sms() * geom.ctas_per_sm;
  if (n_rows > 0) {
    plan_items_kernel<<<nlsh_num_sms() * 4, 256, 0, st>>>(a, items, n_items);
    NLSH_CUDA_TRY(nlsh_post_launch());
    int rc = launch_scan_metric(metric, a, geom, true, grid, st);
    if (rc != NLSH_OK) return rc;
  } else {
    // no rows: every list is empty
    fill_int_kernel<<<(unsigned)((lists + 255) / 256), 256, 0, st>>>(part_id, lists, NLSH_ID_SENTINEL);
    NLSH_CUDA_TRY(nlsh_post_launch());
  }
  return launch_merge_partials(part_d, part_id, nullptr, nullptr, 1, 1, k, kp.rchunk, kp.n_blocks, 1,
                               metric == NLSH_METRIC_L2 ? 1 : 0, n_queries, id_offset, ids_out,
                               dists_out, nullptr, st);
}

extern "C" int nlsh_merge_topk(const float* dists, const int64_t* ids, int64_t dist_stride,
                               int64_t id_stride, const int32_t* ncand, int64_t ncand_stride,
                               int32_t n_lists, int64_t n_queries, int32_t k, int64_t* ids_out,
                               float* dists_out, int32_t* ncand_out, void* stream) {
  NLSH_REQUIRE(n_lists >= 1 && n_queries >= 0, "merge: bad shape n_lists=%d n_queries=%lld", n_lists,
               (long long)n_queries);
  NLSH_REQUIRE(k >= 1 && k <= NLSH_MAX_K, "merge: k=%d outside [1, %d]", k, NLSH_MAX_K);
  if (n_queries == 0) return NLSH_OK;
  NLSH_REQUIRE(dists && ids && ids_out && dists_out, "merge: null pointer");
  NLSH_REQUIRE((ncand == nullptr) == (ncand_out == nullptr), "merge: ncand and ncand_out go together");
  if (dist_stride == 0) dist_stride = n_queries * k;
  if (id_stride == 0) id_stride = n_queries * k;
  if (ncand_stride == 0) ncand_stride = n_queries;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((n_queries + 3) / 4);
  const long long* ids_ll = reinterpret_cast<const long long*>(ids);
  long long* out_ll = reinterpret_cast<long long*>(ids_out);
  if (k <= 32)
    merge_lists_kernel<1><<<blocks, 128, 0, st>>>(dists, ids_ll, dist_stride, id_stride, ncand, ncand_stride,
                                                  n_lists, n_queries, k, out_ll, dists_out, ncand_out);
  else if (k <= 64)
    merge_lists_kernel<2><<<blocks, 128, 0, st>>>(dists, ids_ll, dist_stride, id_stride, ncand, ncand_stride,
                                                  n_lists, n_queries, k, out_ll, dists_out, ncand_out);
  else
    merge_lists_kernel<4><<<blocks, 128, 0, st>>>(dists, ids_ll, dist_stride, id_stride, ncand, ncand_stride,
                                                  n_lists, n_queries, k, out_ll, dists_out, ncand_out);
  return nlsh_check_cuda(nlsh_post_launch(), "merge_lists_kernel launch");
}

extern "C" int nlsh_recall_hits(const int64_t* gt, int32_t k_gt, const int64_t* pred, int32_t k_pred,
                                int64_t n_queries, int32_t* hits_out, void* stream) {
  NLSH_REQUIRE(k_gt >= 1 && k_pred >= 1 && n_queries >= 0, "recall: bad shape");
  if (n_queries == 0) return NLSH_OK;
  NLSH_REQUIRE(gt && pred && hits_out, "recall: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  recall_hits_kernel<<<(unsigned)((n_queries + 127) / 128), 128, 0, st>>>(
      reinterpret_cast<const long long*>(gt), k_gt, reinterpret_cast<const long long*>(pred), k_pred,
      n_queries, hits_out);
  return nlsh_check_cuda(nlsh_post_launch(), "recall_hits_kernel launch");
}
