// Brute-force kNN on the tensor cores: replaces the `distance_func(target, all).topk(k+1)`
// batches of self_get_knn_pt (precompute.py:57-67) with distance_func = _l2 (precompute.py:
// 37-54: |q|^2 + |x|^2 - 2 q.x, squared, no eps) or _cosine_distance (precompute.py:22-34).
// The 512 x N distance matrix of the reference is never materialised: a GEMM tile's scores
// go straight from TMEM into per-query top-k heaps.
//
// One CTA = 128 queries (UMMA M, one query per TMEM lane) against a range of database rows,
// 64 rows (UMMA N) per tile.
//   * A = the 128 query rows, split hi/lo (3xTF32, see tc_linear.cu) and parked in TENSOR
//     MEMORY once per CTA (tcgen05.st): columns [0,128) hi, [128,256) lo.  The MMAs read A
//     from TMEM, so shared memory only carries B.
//   * B = database rows, hi/lo split copies streamed by TMA (box 64 rows x 32 fp32,
//     SWIZZLE_128B) through an 8-stage ring; the row term (|x|^2 or 1/|x|) of each tile rides
//     along as a 256-byte bulk copy.
//   * Accumulators: two sets of (main 64 + cross 64) TMEM columns, so the epilogue of tile j
//     overlaps the MMAs of tile j+1.
//   * Epilogue: 4 * cgroups warps (cgroups = 4, 2 or 1 for small / medium / large k); the warp
//     with TMEM lane quarter w % 4 and column group cg owns 64 / cgroups of the tile's columns
//     for its 32 queries.  Sixteen scores at a time: tcgen05.ld, form the distances, take their
//     minimum, and only when it beats the thread's current k-th best walk the sixteen through
//     the exact (distance, id) test into a per-thread max-heap (thread-local memory).  The ncu
//     capture of the first version (one warp per quarter, every score through the full test)
//     showed 2 000 warp instructions per tile at 21 % issue utilisation - a single warp per
//     scheduler cannot hide its own latencies - with the tensor pipe 16 % busy.
// Partial lists per (query, row range, column group) are merged by merge_partials_kernel (scan.cu).
#include <string.h>

#include "tc_common.cuh"

namespace {

constexpr int kKnBN = 64;         // database rows per tile (UMMA N)
constexpr int kKnMaxK = 128;      // columns of K held in TMEM per operand half (d <= 128)
constexpr int kKnStages = 8;      // B ring: one stage = one 32-column K block of a tile (hi + lo)
constexpr int kKnNormSlots = 16;  // row-term ring, one slot per tile: the producer runs at most
                                  // kKnStages + 2 tiles ahead of the epilogue
constexpr int kKnMaxCGroups = 4;
constexpr int kKnMaxThreads = 64 + 128 * kKnMaxCGroups;
constexpr int kKnHeapCap = NLSH_MAX_K;
constexpr int kKnLocalHeapK = 16;  // largest k whose heaps live in thread-local memory
constexpr int kKnCandBuf = 32;     // per-thread candidate buffer (shared-memory heaps only), flushed
                                   // warp-wide so the 32 lanes sift their heaps at the same time
constexpr uint32_t kKnStageBytes = 2u * kKnBN * kTcBK * sizeof(float);  // 16 KB

struct KnnTcArgs {
  const float* q;       // [n_queries, d]
  const float* xterm;   // [rows padded to 64] |x|^2 (L2SQ; +inf past the end) or 1/|x| (COSINE; 0 past the end)
  float* part_d;        // [n_queries, n_blocks, k]
  int* part_id;
  long long self_offset;  // global index of query 0 (exclude_self)
  long long row_base;     // global row index of this DB chunk's first row
  int n_queries, n_rows;  // rows in this DB chunk
  int d, kblocks;         // kblocks = ceil(d / 32) <= 4
  int k, metric, exclude_self;
  int n_blocks, block0;   // partial-list slots per query / first slot of this launch
  int tiles_per_split;
  int cgroups;            // epilogue column groups (1, 2 or 4)
  int stages;             // B ring depth (<= kKnStages; fewer when the heaps take shared memory)
};

// max-heap on (distance, id): root = current k-th best.  Entry i lives at d[i * stride]: for
// k <= kKnLocalHeapK the arrays are thread-local (stride 1; 512 threads x 16 entries stay in L1),
// above that they are columns of a shared-memory matrix [k][epilogue threads] - the ncu run of the
// thread-local version at k = 100 showed every sift level going out to L2 (the 128 heaps no longer
// fit beside the 128 KB stage ring), ~17 k cycles per push.
struct TopHeap {
  float* d;
  int* id;
  int stride;
  int n;

  __device__ __forceinline__ static bool less(float d1, int i1, float d2, int i2) {
    return d1 < d2 || (d1 == d2 && i1 < i2);
  }
  __device__ __noinline__ void push(float nd, int nid, int k) {
    if (n < k) {  // sift up
      int c = n++;
      while (c > 0) {
        const int p = (c - 1) >> 1;
        const float pd = d[p * stride];
        const int pi = id[p * stride];
        if (!less(pd, pi, nd, nid)) break;
        d[c * stride] = pd;
        id[c * stride] = pi;
        c = p;
      }
      d[c * stride] = nd;
      id[c * stride] = nid;
    } else {  // replace the root, sift down
      int p = 0;
      while (true) {
        int c = 2 * p + 1;
        if (c >= k) break;
        float cd = d[c * stride];
        int ci = id[c * stride];
        if (c + 1 < k) {
          const float rd = d[(c + 1) * stride];
          const int ri = id[(c + 1) * stride];
          if (less(cd, ci, rd, ri)) {
            ++c;
            cd = rd;
            ci = ri;
          }
        }
        if (!less(nd, nid, cd, ci)) break;
        d[p * stride] = cd;
        id[p * stride] = ci;
        p = c;
      }
      d[p * stride] = nd;
      id[p * stride] = nid;
    }
  }
};

template <int METRIC>
__global__ void __launch_bounds__(kKnMaxThreads, 1)
    tc_knn_kernel(const KnnTcArgs a, const __grid_constant__ CUtensorMap map_x_hi,
                  const __grid_constant__ CUtensorMap map_x_lo) {
  extern __shared__ unsigned char kn_smem_raw[];
  unsigned char* base = kn_smem_raw + ((1024u - (smem_u32(kn_smem_raw) & 1023u)) & 1023u);
  float* norm_ring = reinterpret_cast<float*>(base + (size_t)a.stages * kKnStageBytes);  // [slots][64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(norm_ring + kKnNormSlots * kKnBN);  // [kKnStages]
  uint64_t* empty_bar = full_bar + kKnStages;                                          // [kKnStages]
  uint64_t* acc_full = empty_bar + kKnStages;                                          // [2]
  uint64_t* acc_empty = acc_full + 2;                                                  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* heap_d_s = reinterpret_cast<float*>(tmem_slot + 4);  // [k][epilogue threads] when k > kKnLocalHeapK
  const int n_epi = 128 * a.cgroups;
  int* heap_id_s = reinterpret_cast<int*>(heap_d_s + (size_t)a.k * n_epi);
  float* cand_d_s = reinterpret_cast<float*>(heap_id_s + (size_t)a.k * n_epi);  // [kKnCandBuf][epilogue threads]
  int* cand_id_s = reinterpret_cast<int*>(cand_d_s + (size_t)kKnCandBuf * n_epi);

  // warp 0: TMA producer; warp 1: MMA issuer + TMEM owner; warps 2 .. 2 + 4 * cgroups: epilogue
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int q0 = blockIdx.x * kTcBM;
  const int split = blockIdx.y;
  const int n_tiles_total = (a.n_rows + kKnBN - 1) / kKnBN;
  const int tile0 = split * a.tiles_per_split;
  int n_tiles = n_tiles_total - tile0;
  if (n_tiles > a.tiles_per_split) n_tiles = a.tiles_per_split;
  if (n_tiles < 0) n_tiles = 0;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4 * a.cgroups);  // one arrive per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tc_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int quarter = warp & 3;  // the TMEM lanes a warp may touch: 32 * (warp % 4) ...
  const int cg = (warp - 2) >> 2;
  const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);

  // ---- A: the query rows -> TMEM (hi in columns [0,128), lo in [128,256)) by column group 0 -----
  float q_term = 0.f;  // |q|^2 (L2SQ) or 1 / |q| (COSINE)
  const int my_q = q0 + quarter * 32 + lane;
  const bool q_ok = my_q < a.n_queries;
  if (warp >= 2) {
    float ss = 0.f;
    for (int c0 = 0; c0 < a.kblocks * kTcBK; c0 += 16) {
      float hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = c0 + i;
        const float v = (q_ok && c < a.d) ? a.q[(size_t)my_q * a.d + c] : 0.f;
        ss = fmaf(v, v, ss);
        hi[i] = tf32_hi(v);
        lo[i] = v - hi[i];
      }
      if (cg == 0) {
        tc_st16(lane_base + (uint32_t)c0, hi);
        tc_st16(lane_base + (uint32_t)(kKnMaxK + c0), lo);
      }
    }
    if (cg == 0) tc_wait_st();
    q_term = (METRIC == NLSH_METRIC_COSINE) ? 1.0f / sqrtf(ss) : ss;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------
    if (lane == 0) {
      // running stage index / round parity: `ring % stages` with a run-time divisor is a 200-cycle dependent
      // chain per K block in this single-thread role (measured in scan_tc.cu's producer)
      int s = 0;
      unsigned par = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int row0 = (tile0 + j) * kKnBN;
        for (int kb = 0; kb < a.kblocks; ++kb) {
          mbar_wait_poll(&empty_bar[s], par ^ 1u);
          unsigned char* st = base + (size_t)s * kKnStageBytes;
          const unsigned extra = kb == 0 ? (unsigned)(kKnBN * sizeof(float)) : 0u;
          mbar_arrive_expect_tx(&full_bar[s], kKnStageBytes + extra);
          tma_load_2d(st, &map_x_hi, kb * kTcBK, row0, &full_bar[s]);
          tma_load_2d(st + kKnStageBytes / 2, &map_x_lo, kb * kTcBK, row0, &full_bar[s]);
          if (kb == 0)
            bulk_g2s(norm_ring + (j % kKnNormSlots) * kKnBN, a.xterm + row0, kKnBN * sizeof(float),
                     &full_bar[s]);
          if (++s == a.stages) {
            s = 0;
            par ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_tf32_idesc(kKnBN);
      int s = 0;
      unsigned par = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int set = j & 1;
        mbar_wait_poll(&acc_empty[set], (((unsigned)j >> 1) & 1u) ^ 1u);  // epilogue drained this set
        tc_fence_after();
        const uint32_t acc_main = tmem_base + 256u + (uint32_t)(set * 128);
        const uint32_t acc_cross = acc_main + 64u;
        for (int kb = 0; kb < a.kblocks; ++kb) {
          mbar_wait_poll(&full_bar[s], par);
          tc_fence_after();
          unsigned char* st = base + (size_t)s * kKnStageBytes;
          const uint64_t db_hi = make_kmajor_sw128_desc(st);
          const uint64_t db_lo = make_kmajor_sw128_desc(st + kKnStageBytes / 2);
#pragma unroll
          for (int k = 0; k < kTcBK / 8; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
            const uint32_t a_hi = tmem_base + (uint32_t)(kb * kTcBK + k * 8);
            const uint32_t a_lo = a_hi + (uint32_t)kKnMaxK;
            tc_mma_tf32_ts(acc_main, a_hi, db_hi + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_mma_tf32_ts(acc_cross, a_hi, db_lo + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_mma_tf32_ts(acc_cross, a_lo, db_hi + adv, idesc, 1u);
          }
          tc_commit(&empty_bar[s]);
          if (++s == a.stages) {
            s = 0;
            par ^= 1u;
          }
        }
        tc_commit(&acc_full[set]);
      }
    }
    __syncwarp();
  } else if (cg < a.cgroups) {
    // ------------------------------- epilogue --------------------------------------------
    float loc_d[kKnLocalHeapK];
    int loc_id[kKnLocalHeapK];
    TopHeap heap;
    heap.n = 0;
    if (a.k <= kKnLocalHeapK) {
      heap.d = loc_d;
      heap.id = loc_id;
      heap.stride = 1;
    } else {
      heap.d = heap_d_s + (tid - 64);
      heap.id = heap_id_s + (tid - 64);
      heap.stride = n_epi;
    }
    const bool buffered = a.k > kKnLocalHeapK;
    float* cand_d = cand_d_s + (tid - 64);
    int* cand_i = cand_id_s + (tid - 64);
    int n_cand = 0;
    // a thread without a query never passes the filter
    float tau = q_ok ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
    int tau_id = NLSH_ID_SENTINEL;
    const long long self_row = a.exclude_self ? a.self_offset + my_q - a.row_base : -1;
    const int ncols = kKnBN / a.cgroups;
    const int c_lo = cg * ncols;
    for (int j = 0; j < n_tiles; ++j) {
      const int set = j & 1;
      mbar_wait_poll(&acc_full[set], ((unsigned)j >> 1) & 1u);
      tc_fence_after();
      const uint32_t acc_main = lane_base + 256u + (uint32_t)(set * 128);
      const float* xt = norm_ring + (j % kKnNormSlots) * kKnBN;
      const int row0 = (tile0 + j) * kKnBN;
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_lo + ncols; c0 += 16) {
        uint32_t m[16], x[16];
        tc_ld16_nowait(acc_main + (uint32_t)c0, m);
        tc_ld16_nowait(acc_main + 64u + (uint32_t)c0, x);
        float xv[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 t = *reinterpret_cast<const float4*>(xt + c0 + i);
          xv[i] = t.x; xv[i + 1] = t.y; xv[i + 2] = t.z; xv[i + 3] = t.w;
        }
        tc_wait_ld2(m, x);
        float dist[16];
        float mn = __int_as_float(0x7f800000);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dot = __uint_as_float(m[i]) + __uint_as_float(x[i]);
          if (METRIC == NLSH_METRIC_COSINE)
            dist[i] = 1.0f - dot * q_term * xv[i];
          else
            dist[i] = fmaf(-2.0f, dot, xv[i]) + q_term;  // precompute._l2: addmm(|x|^2, q, x^T, alpha=-2) + |q|^2
          mn = fminf(mn, dist[i]);
        }
        if (mn <= tau) {  // rare once the heap is warm: the exact (distance, id) test
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (dist[i] <= tau) {
              const int row = row0 + c0 + i;
              if (row < a.n_rows && row != self_row && TopHeap::less(dist[i], row, tau, tau_id)) {
                if (buffered) {  // park it: the heap is updated by the warp-wide flush below
                  cand_d[n_cand * n_epi] = dist[i];
                  cand_i[n_cand * n_epi] = row;
                  ++n_cand;
                } else {
                  heap.push(dist[i], row, a.k);
                  if (heap.n == a.k) {
                    tau = heap.d[0];
                    tau_id = heap.id[0];
                  }
                }
              }
            }
          }
        }
        // Large k: a push is ~7 dependent sift levels through shared memory; one lane pushing while
        // 31 wait cost 850 cycles per candidate.  Parked candidates are pushed by all lanes together
        // whenever some lane's buffer could overflow on the next 16 scores.
        if (buffered && __any_sync(NLSH_FULL_MASK, n_cand > kKnCandBuf - 16)) {
          const int rounds = __reduce_max_sync(NLSH_FULL_MASK, n_cand);
          for (int e = 0; e < rounds; ++e) {
            if (e < n_cand) {
              const float cd = cand_d[e * n_epi];
              const int ci = cand_i[e * n_epi];
              if (TopHeap::less(cd, ci, tau, tau_id)) {
                heap.push(cd, ci, a.k);
                if (heap.n == a.k) {
                  tau = heap.d[0];
                  tau_id = heap.id[0];
                }
              }
            }
          }
          n_cand = 0;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[set]);
    }
    for (int e = 0; e < n_cand; ++e) {  // whatever is still parked
      const float cd = cand_d[e * n_epi];
      const int ci = cand_i[e * n_epi];
      if (TopHeap::less(cd, ci, tau, tau_id)) {
        heap.push(cd, ci, a.k);
        if (heap.n == a.k) {
          tau = heap.d[0];
          tau_id = heap.id[0];
        }
      }
    }
    if (q_ok) {
      const size_t slot = (((size_t)my_q * a.n_blocks + a.block0 + split) * a.cgroups + cg) * a.k;
      for (int i = 0; i < a.k; ++i) {
        const bool has = i < heap.n;
        a.part_d[slot + i] = has ? heap.d[i * heap.stride] : __int_as_float(0x7f800000);
        a.part_id[slot + i] = has ? (int)(a.row_base + heap.id[i * heap.stride]) : NLSH_ID_SENTINEL;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tc_dealloc(tmem_base, 512);
}

// Database chunk -> (hi, lo) split copies + the per-row term (|x|^2 or 1/|x|); one warp per row.
__global__ void __launch_bounds__(256)
    knn_split_rows_kernel(const float* __restrict__ x, long long n, int d, int metric,
                          float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ term,
                          long long n_term) {
  const int lane = lane_id();
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp; r < n_term; r += n_warps) {
    float ss = 0.f;
    if (r < n) {
      for (int c = lane; c < d; c += 32) {
        const float v = x[r * d + c];
        const float h = tf32_hi(v);
        hi[r * d + c] = h;
        lo[r * d + c] = v - h;
        ss = fmaf(v, v, ss);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(NLSH_FULL_MASK, ss, o);
    if (lane == 0)
      term[r] = r < n ? (metric == NLSH_METRIC_COSINE ? 1.0f / sqrtf(ss) : ss)
                      : (metric == NLSH_METRIC_COSINE ? 0.f : __int_as_float(0x7f800000));
  }
}

constexpr long long kKnChunkRows = 1ll << 20;  // database rows split per launch (bounds the hi/lo scratch)

struct KnnTcPlan {
  long long chunk_rows;
  int n_chunks;
  int splits;           // row ranges per chunk (grid.y)
  int tiles_per_split;
  int n_blocks;         // partial lists per query = n_chunks * splits
};

KnnTcPlan knn_tc_plan(long long n_queries, long long n_rows) {
  KnnTcPlan p;
  p.chunk_rows = n_rows < kKnChunkRows ? (n_rows > 0 ? n_rows : 1) : kKnChunkRows;
  p.n_chunks = (int)((n_rows + p.chunk_rows - 1) / p.chunk_rows);
  if (p.n_chunks < 1) p.n_chunks = 1;
  const long long q_tiles = (n_queries + kTcBM - 1) / kTcBM;
  const long long tiles = (p.chunk_rows + kKnBN - 1) / kKnBN;
  // row ranges per query tile: the split count whose CTA waves x tiles per CTA is smallest
  const long long sms = nlsh_num_sms();
  long long splits = 1, best_cost = -1;
  for (long long s = 1; s <= 64 && s <= tiles; ++s) {
    const long long waves = (q_tiles * s + sms - 1) / sms;
    const long long cost = waves * ((tiles + s - 1) / s + 8);  // + 8: a CTA's fixed set-up, in tiles
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      splits = s;
    }
  }
  p.tiles_per_split = (int)((tiles + splits - 1) / splits);
  p.splits = (int)((tiles + p.tiles_per_split - 1) / p.tiles_per_split);
  p.n_blocks = p.n_chunks * p.splits;
  return p;
}

}  // namespace

// ---- interface used by scan.cu (nlsh_knn_bruteforce) -------------------------------------------
bool nlsh_knn_tc_supported(int d, int metric, int k) {
  return d % 4 == 0 && d >= 4 && d <= kKnMaxK && k <= kKnHeapCap &&
         (metric == NLSH_METRIC_L2SQ || metric == NLSH_METRIC_COSINE);
}

// epilogue column groups: more threads per query while the per-thread heaps stay small
int knn_tc_cgroups(int k) { return k <= 16 ? 4 : (k <= 48 ? 2 : 1); }

int nlsh_knn_tc_blocks(long long n_queries, long long n_rows, int k) {
  return knn_tc_plan(n_queries, n_rows).n_blocks * knn_tc_cgroups(k);
}

size_t nlsh_knn_tc_scratch_floats(long long n_queries, long long n_rows, int d) {
  const KnnTcPlan p = knn_tc_plan(n_queries, n_rows);
  const size_t term = (size_t)((p.chunk_rows + kKnBN - 1) / kKnBN * kKnBN);
  return 2 * (size_t)p.chunk_rows * d + term + 64;
}

// Fills part_d / part_id [n_queries, nlsh_knn_tc_blocks(), k] (unsorted lists, sentinel padded).
int nlsh_knn_tc_run(const float* xq, long long n_queries, const float* xdb, long long n_rows, int d,
                    int metric, int k, int exclude_self, long long self_offset, float* scratch,
                    float* part_d, int* part_id, cudaStream_t st) {
  const KnnTcPlan p = knn_tc_plan(n_queries, n_rows);
  float* x_hi = scratch;
  float* x_lo = x_hi + (size_t)p.chunk_rows * d;
  float* term = x_lo + (size_t)p.chunk_rows * d;
  const int cgroups = knn_tc_cgroups(k);
  const size_t heap_bytes = k > kKnLocalHeapK ? (size_t)(k + kKnCandBuf) * 128 * cgroups * 8 : 0;
  const size_t fixed = kKnNormSlots * kKnBN * sizeof(float) + 512 + 1024 + heap_bytes;
  int stages = kKnStages;
  while (stages > 3 && stages * (size_t)kKnStageBytes + fixed > 222 * 1024) --stages;
  const size_t smem = (size_t)stages * kKnStageBytes + fixed;
  const bool cosine = metric == NLSH_METRIC_COSINE;
  auto kernel = cosine ? tc_knn_kernel<NLSH_METRIC_COSINE> : tc_knn_kernel<NLSH_METRIC_L2SQ>;
  NLSH_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int c = 0; c < p.n_chunks; ++c) {
    const long long r0 = (long long)c * p.chunk_rows;
    const long long rows = (n_rows - r0) < p.chunk_rows ? (n_rows - r0) : p.chunk_rows;
    const long long n_term = (rows + kKnBN - 1) / kKnBN * kKnBN;
    long long blocks = (n_term + 7) / 8;
    if (blocks > (long long)nlsh_num_sms() * 8) blocks = (long long)nlsh_num_sms() * 8;
    knn_split_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(xdb + (size_t)r0 * d, rows, d, metric, x_hi,
                                                            x_lo, term, n_term);
    NLSH_CUDA_TRY(nlsh_post_launch());
    CUtensorMap m_hi, m_lo;
    int rc;
    if ((rc = tc_make_map(&m_hi, x_hi, rows, d, kKnBN)) != NLSH_OK) return rc;
    if ((rc = tc_make_map(&m_lo, x_lo, rows, d, kKnBN)) != NLSH_OK) return rc;
    KnnTcArgs a{};
    a.q = xq;
    a.xterm = term;
    a.part_d = part_d;
    a.part_id = part_id;
    a.self_offset = self_offset;
    a.row_base = r0;
    a.n_queries = (int)n_queries;
    a.n_rows = (int)rows;
    a.d = d;
    a.kblocks = (d + kTcBK - 1) / kTcBK;
    a.k = k;
    a.metric = metric;
    a.exclude_self = exclude_self;
    a.n_blocks = p.n_blocks;  // (chunk, split) slots; each holds `cgroups` lists
    a.block0 = c * p.splits;
    a.cgroups = cgroups;
    a.stages = stages;
    a.tiles_per_split = p.tiles_per_split;
    dim3 grid((unsigned)((n_queries + kTcBM - 1) / kTcBM), (unsigned)p.splits);
    kernel<<<grid, 64 + 128 * cgroups, smem, st>>>(a, m_hi, m_lo);
    NLSH_CUDA_TRY(nlsh_post_launch());
  }
  return NLSH_OK;
}
