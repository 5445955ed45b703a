// Candidate scan with a tensor-core FILTER in front of the exact distance (the default path of
// nlsh_query_scan_topk for d <= 128, k <= 32).  Same job as scan.cu::scan_kernel - the per-query
// gather + distance_func + topk of Indexer.query (nlsh/indexer.py:62-95) - and the same results:
// every distance that enters a top-k list is computed in fp32 in the reference's difference form
// (nlsh/data.py:201 F.pairwise_distance, nlsh/data.py:109 1 - cosine_similarity), never from the
// GEMM expansion.  The GEMM only decides which (row, query) pairs cannot matter.
//
// Why: with the 10k-query batches of the BASELINE configs every bucket is probed by ~20 queries.
// The fp32 SIMT kernel fetches a bucket tile once for all of them but still pays 3 lane-ops per
// (row, query, column): ncu showed it fp32-issue bound (fma pipe 59 %, DRAM 43 %).  Here one
// tcgen05.mma (kind::tf32, raw fp32 bits as operands) produces the 128 x 32 dot products of a row
// tile against the item's queries into TMEM; an epilogue thread per row turns each into a LOWER
// BOUND of the exact distance and compares it with the query's current threshold.  Survivors
// (a few per tile once the lists are warm) are re-scored exactly from the tile that is still in
// shared memory, and inserted into register-resident sorted lists.  The kernel is then bound by
// streaming x_sorted from HBM, which is the roofline SURVEY 8(d) names for the scan.
//
// Exactness of the filter.  tf32 keeps 10 mantissa bits of each operand, so
// |dot_tc - <q, x>| <= 2^-9 * 1.02 * |q| |x|  (Cauchy-Schwarz; the 1.02 covers the fp32 accumulate).
//   L2:  d2 = |q|^2 + |x|^2 - 2 <q,x> >= (1 - c)(|q|^2 + |x|^2) - 2 dot_tc,  c = 2^-9 * 1.02 + 4e-5
//        (the 4e-5 covers fp32 rounding of the two norms).  The list holds the reference's
//        D = sum((q - x + 1e-6)^2) >= d2 - 2e-6 sqrt(dim) sqrt(d2) - 1e-5 D, so a pair whose bound
//        exceeds tau_eff = y + a sqrt(y) + a^2  (y = tau * 1.0001, a = 2.1e-6 sqrt(dim)) has D > tau.
//   angular: 1 - cos >= 1 - dot_tc / |x| - 2.1e-3.
// A pair is dropped only when its bound is above the threshold, so the top-k is the same set the
// SIMT kernel finds; ties are still broken by the (distance, id) order.
//
// The threshold of a query is min(k-th best of this item's list, tau_g[query]) where tau_g is a
// global per-query upper bound of the final k-th distance, lowered (atomicMin) whenever an item
// finishes with a full list: the 8 probes of a query tighten each other.  Any k-th-best of any
// subset of the candidates is such an upper bound, so the merged result does not depend on the
// order items happen to run in.
//
// Roles in a CTA (one persistent CTA per SM, 448 threads) - a four-stage pipeline in which no
// stage waits for the next one to finish a tile:
//   warp 0         producer: owns the work queue (atomic counter).  Per item the 32 lanes fetch the
//                  per-query state (flat probe index, |q|^2, tau_g -> initial thresholds) into
//                  shared memory, lane 0 TMA-loads the item's queries (box 32 rows x 32 fp32 per K
//                  block, SWIZZLE_128B; 4-deep item ring) and streams the row tiles (box 128 rows x
//                  32 fp32) into the slot ring;
//   warp 1 lane 0  MMA issuer: per tile and K block 4 x tcgen05.mma M128 N32 K8 into one of two
//                  TMEM accumulator sets; tcgen05.commit frees the slot, the last one of a tile
//                  signals acc_full - a slot lives from its TMA issue to the end of its MMAs;
//   warps 2-5      filter: thread = row; tcgen05.ld its 32 scores (then the TMEM set is free
//                  again), bound + compare, survivors appended to the list of the re-rank warp
//                  that owns the query (warp-aggregated shared-memory atomics; 4-deep list ring);
//   warps 6-13     re-rank: each warp owns 4 of the item's queries (register-resident sorted
//                  lists).  4 lanes per survivor re-read the row (an L2 hit: the tile has just
//                  streamed through) and compute the exact distance, WarpTopK::offer inserts, the
//                  thresholds in shared memory are lowered for the filter warps (who may be a few
//                  tiles ahead: stale thresholds only let more pairs through).
#include <stdlib.h>
#include <string.h>

#include "scan_tc.cuh"
#include "tc_common.cuh"

namespace {

#ifdef NLSH_TC_SUSPEND_WAIT
#define MBAR_WAIT mbar_wait
#else
#define MBAR_WAIT mbar_wait_poll
#endif

// Debug timing (NLSH_TC_STATS): cycles block 0 spends in each wait, summed into a.stats[slot].
#define TIMED_WAIT(slot, bar, par)                                  \
  do {                                                             \
    if (timing) {                                                  \
      const long long t0_ = clock64();                             \
      MBAR_WAIT(bar, par);                                         \
      tacc[slot] += clock64() - t0_;                               \
    } else {                                                       \
      MBAR_WAIT(bar, par);                                         \
    }                                                              \
  } while (0)

constexpr int kFilterWarps = 4;                 // one per TMEM lane quarter
constexpr int kRerankWarps = 8;                 // two per scheduler: their latencies overlap
constexpr int kRerankShift = 3;                 // log2(kRerankWarps): query j belongs to warp j & 7, list j >> 3
constexpr int kTile = 128;                      // rows per tile = UMMA M
constexpr int kThreads = 64 + 32 * (kFilterWarps + kRerankWarps);  // 320
constexpr int kOwn = kTcNQ / kRerankWarps;      // lists per re-rank warp
constexpr int kListCap = kOwn * kTile;          // survivors an owner warp can receive per tile
constexpr int kSurvBufs = 4;                    // survivor-list ring depth (tiles the filter may run ahead)
constexpr uint32_t kSlotBytes = kTile * kTcBK * sizeof(float);   // 16 KB: one K block of a row tile
constexpr uint32_t kQBoxBytes = kTcNQ * kTcBK * sizeof(float);   // 4 KB: one K block of the queries
constexpr int kMaxSlots = 16;
constexpr int kMetaFloats = kTile + 4;           // a tile's row norms, from the 16-byte boundary below its first row
constexpr uint32_t kMetaBytes = 640;            // kMetaFloats * 4 rounded up to 128
constexpr int kMetaBufs = 8;                    // row-norm ring depth (tiles)
constexpr int kItemBufs = 4;                    // item ring depth: buckets of a few tiles are shorter than the
                                                // pipeline, so several items must be in flight
constexpr int kMaxKBlocks = 4;                  // d_pad <= 128
constexpr float kFilterC = 0.001953125f * 1.02f + 4e-5f;  // see header comment
constexpr float kAngularC = 2.1e-3f;

__device__ __forceinline__ float pos_inf() { return __int_as_float(0x7f800000); }
__device__ __forceinline__ float neg_inf() { return __int_as_float(0xff800000); }

// Filter threshold of a query from the best known bound `eff` of its k-th distance.
template <int METRIC>
__device__ __forceinline__ float make_thr(float eff, float qn2, float l2_slack) {
  if (eff == pos_inf()) return pos_inf();
  if (METRIC == NLSH_METRIC_L2) {
    if (!(eff >= 0.f)) return neg_inf();
    const float y = eff * 1.0001f;
    const float te = (y + l2_slack * sqrtf(y) + l2_slack * l2_slack) * 1.0001f + 1e-30f;
    return te - (1.0f - kFilterC) * qn2;
  }
  return eff - 1.0f + kAngularC;
}

// Exact distance of database row `xrow` (global memory; an L2 hit, the tile has just streamed
// through) to query j of the item (shared memory, SWIZZLE_128B box: 16-byte chunk c of row j is
// stored at chunk position c ^ (j & 7)), 4 lanes per pair: lane l4 owns the columns = l4 (mod 4)
// and walks them in ascending order, the columns of a partial last float4 (d % 4 != 0) go to
// lane 0, and the four partial sums are combined as (s0 + s1) + (s2 + s3).  That is exactly the
// summation order of scan.cu::consume_box (packed float2 accumulators over the columns 0,1 / 2,3
// of each float4), so both scan kernels produce the same bits for the same (q, x) - the parity
// tests compare them with torch.equal.  All loads of the row are issued before the dependent sum.
// L2 returns the squared distance (the root is taken in merge_partials_kernel).
template <int METRIC>
__device__ __noinline__ float exact_distance(const float* __restrict__ xrow, const unsigned char* qsrc,
                                             int d, int j, int l4) {
  float acc = 0.f, xx = 0.f;
  const int nv = d >> 2, tail = d & 3;
  const int qrow = j * 128 + l4 * 4, qsw = j & 7;
  float xv[kMaxKBlocks * 8];
#pragma unroll
  for (int v = 0; v < kMaxKBlocks * 8; ++v) xv[v] = v < nv ? __ldg(xrow + 4 * v + l4) : 0.f;
  float xt[3] = {0.f, 0.f, 0.f};
  if (tail != 0 && l4 == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c < tail) xt[c] = __ldg(xrow + nv * 4 + c);
  }
#pragma unroll
  for (int kb = 0; kb < kMaxKBlocks; ++kb) {
    if (kb * 8 < nv) {
      const unsigned char* qs = qsrc + kb * kQBoxBytes + qrow;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        if (kb * 8 + c8 < nv) {
          const float x = xv[kb * 8 + c8];
          const float q = *reinterpret_cast<const float*>(qs + ((c8 ^ qsw) << 4));
          if (METRIC == NLSH_METRIC_L2) {
            // F.pairwise_distance: (q - x) + eps, squared and summed (nlsh/data.py:201)
            const float t = __fadd_rn(__fsub_rn(q, x), 1e-6f);
            acc = fmaf(t, t, acc);
          } else {
            acc = fmaf(q, x, acc);
            xx = fmaf(x, x, xx);
          }
        }
      }
    }
  }
  if (tail != 0 && l4 == 0) {  // the partial float4 at vector index nv: columns 4 nv .. d - 1
    const int kb = nv >> 3, c8 = nv & 7;
    const float* qs = reinterpret_cast<const float*>(qsrc + kb * kQBoxBytes + j * 128 + ((c8 ^ qsw) << 4));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < tail) {
        if (METRIC == NLSH_METRIC_L2) {
          const float t = __fadd_rn(__fsub_rn(qs[c], xt[c]), 1e-6f);
          acc = fmaf(t, t, acc);
        } else {
          acc = fmaf(qs[c], xt[c], acc);
          xx = fmaf(xt[c], xt[c], xx);
        }
      }
    }
  }
  acc += __shfl_xor_sync(NLSH_FULL_MASK, acc, 1);
  acc += __shfl_xor_sync(NLSH_FULL_MASK, acc, 2);
  if (METRIC == NLSH_METRIC_L2) return acc;
  xx += __shfl_xor_sync(NLSH_FULL_MASK, xx, 1);
  xx += __shfl_xor_sync(NLSH_FULL_MASK, xx, 2);
  return 1.0f - acc / fmaxf(sqrtf(xx), 1e-8f);  // nlsh/data.py:109, norms clamped at 1e-8
}

template <int METRIC>
__global__ void __launch_bounds__(kThreads, 1)
    scan_tc_kernel(const TcScanArgs a, const __grid_constant__ CUtensorMap map_x,
                   const __grid_constant__ CUtensorMap map_q) {
  extern __shared__ unsigned char stc_smem_raw[];
  unsigned char* base = stc_smem_raw + ((1024u - (smem_u32(stc_smem_raw) & 1023u)) & 1023u);
  unsigned char* slots = base;                                              // [n_slots][16 KB]
  unsigned char* qbuf = slots + (size_t)a.n_slots * kSlotBytes;             // [kItemBufs][kblocks][4 KB]
  unsigned char* meta = qbuf + (size_t)kItemBufs * a.kblocks * kQBoxBytes;          // [kMetaBufs][kMetaBytes] row norms
  uint16_t* surv = reinterpret_cast<uint16_t*>(meta + (size_t)kMetaBufs * kMetaBytes);  // [kSurvBufs][kRerankWarps][kListCap]
  float* thr_s = reinterpret_cast<float*>(surv + kSurvBufs * kRerankWarps * kListCap);      // [kItemBufs][kTcNQ]
  float* own_ext = thr_s + kItemBufs * kTcNQ;                               // [kItemBufs][kTcNQ]
  float* own_qn2 = own_ext + kItemBufs * kTcNQ;                             // [kItemBufs][kTcNQ]
  int* own_f = reinterpret_cast<int*>(own_qn2 + kItemBufs * kTcNQ);         // [kItemBufs][kTcNQ]
  int* cnt = own_f + kItemBufs * kTcNQ;                                             // [kSurvBufs][kRerankWarps]
  TcItem* itm = reinterpret_cast<TcItem*>(cnt + kSurvBufs * kRerankWarps);  // [kItemBufs]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(itm + kItemBufs);                // [kMaxSlots]
  uint64_t* empty_bar = full_bar + kMaxSlots;                               // [kMaxSlots]
  uint64_t* q_full = empty_bar + kMaxSlots;                                 // [kItemBufs]
  uint64_t* q_empty = q_full + kItemBufs;                                   // [kItemBufs]
  uint64_t* acc_full = q_empty + kItemBufs;                                 // [2]
  uint64_t* acc_empty = acc_full + 2;                                       // [2]
  uint64_t* surv_full = acc_empty + 2;                                      // [kSurvBufs]
  uint64_t* surv_empty = surv_full + kSurvBufs;                             // [kSurvBufs]
  uint64_t* meta_full = surv_empty + kSurvBufs;                             // [kMetaBufs]
  uint64_t* meta_empty = meta_full + kMetaBufs;                             // [kMetaBufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(meta_empty + kMetaBufs);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const unsigned n_slots = (unsigned)a.n_slots;
  const int kblocks = a.kblocks;

  if (tid == 0) {
    for (int s = 0; s < a.n_slots; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kMetaBufs; ++s) {
      mbar_init(&meta_full[s], 1);
      mbar_init(&meta_empty[s], kFilterWarps);
    }
    for (int s = 0; s < kItemBufs; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], kRerankWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kFilterWarps);
    }
    for (int s = 0; s < kSurvBufs; ++s) {
      mbar_init(&surv_full[s], kFilterWarps);
      mbar_init(&surv_empty[s], kRerankWarps);
    }
    mbar_fence_init();
  }
  if (tid < kSurvBufs * kRerankWarps) cnt[tid] = 0;
  if (warp == 1) tc_alloc(tmem_slot, 64);  // two accumulator sets of kTcNQ columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool timing = a.stats != nullptr && blockIdx.x == 0 && lane == 0;
  long long tacc[4] = {0, 0, 0, 0};
  const long long t_begin = clock64();

  if (warp == 0) {
    // =================================== producer =========================================
    int total = *a.n_items;
    if (total > a.max_items) total = a.max_items;
    unsigned ring = 0, icount = 0, tcount = 0;
    while (true) {
      int item = 0;
      if (lane == 0) item = atomicAdd(a.item_counter, 1);
      item = __shfl_sync(NLSH_FULL_MASK, item, 0);
      const int islot = (int)(icount % kItemBufs);
      TIMED_WAIT(2, &q_empty[islot], ((icount / kItemBufs) & 1u) ^ 1u);
      if (item >= total) {
        if (lane == 0) {
          itm[islot].nq = 0;  // end of work
          mbar_arrive(&q_full[islot]);
        }
        break;
      }
      const TcItem rec = a.items[item];
      {  // lane j: state of the item's query j
        float th = neg_inf(), ex = neg_inf(), qn = 0.f;
        int f = -1;
        if (lane < rec.nq) {
          f = a.pairs[rec.pair_base + lane];
          qn = a.qs_norm[rec.pair_base + lane];
          ex = __ldcg(a.tau_g + f / a.p);
          th = make_thr<METRIC>(ex, qn, a.l2_slack);
        }
        thr_s[islot * kTcNQ + lane] = th;
        own_ext[islot * kTcNQ + lane] = ex;
        own_qn2[islot * kTcNQ + lane] = qn;
        own_f[islot * kTcNQ + lane] = f;
      }
      if (lane == 0) itm[islot] = rec;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(&q_full[islot], (unsigned)kblocks * kQBoxBytes);
        unsigned char* qdst = qbuf + (size_t)islot * kblocks * kQBoxBytes;
        for (int kb = 0; kb < kblocks; ++kb)
          tma_load_2d(qdst + kb * kQBoxBytes, &map_q, kb * kTcBK, rec.pair_base, &q_full[islot]);
        const int n_tiles = (rec.row1 - rec.row0 + kTile - 1) / kTile;
        for (int t = 0; t < n_tiles; ++t, ++tcount) {
          const int trow0 = rec.row0 + t * kTile;
          {
            // The tile's row norms: a plain bulk copy needs a 16-byte aligned source, so it starts
            // at the 4-row boundary below the tile and stops at the last whole group of 4 rows of
            // the array (the filter reads the <= 3 rows after that directly).
            const unsigned mb = tcount % kMetaBufs;
            TIMED_WAIT(0, &meta_empty[mb], ((tcount / kMetaBufs) & 1u) ^ 1u);
            const long long m0 = trow0 & ~3ll;
            long long avail = (a.n_rows & ~3ll) - m0;
            if (avail > kMetaFloats) avail = kMetaFloats;
            if (avail > 0) {
              mbar_arrive_expect_tx(&meta_full[mb], (unsigned)avail * 4u);
              bulk_g2s(meta + (size_t)mb * kMetaBytes, a.xnorm + m0, (unsigned)avail * 4u, &meta_full[mb]);
            } else {
              mbar_arrive(&meta_full[mb]);
            }
          }
          for (int kb = 0; kb < kblocks; ++kb, ++ring) {
            const unsigned s = ring % n_slots;
            TIMED_WAIT(1, &empty_bar[s], ((ring / n_slots) & 1u) ^ 1u);
            // a box is always written in full (rows / columns past the tensor are zero filled)
            mbar_arrive_expect_tx(&full_bar[s], kSlotBytes);
            tma_load_2d(slots + (size_t)s * kSlotBytes, &map_x, kb * kTcBK, trow0, &full_bar[s]);
          }
        }
      }
      __syncwarp();
      ++icount;
    }
  } else if (warp == 1) {
    // =================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = make_tf32_idesc(kTcNQ);
      unsigned ring = 0, icount = 0, tcount = 0;
      while (true) {
        const int islot = (int)(icount % kItemBufs);
        MBAR_WAIT(&q_full[islot], (icount / kItemBufs) & 1u);
        const int nq = itm[islot].nq;
        if (nq == 0) break;
        const int n_tiles = (itm[islot].row1 - itm[islot].row0 + kTile - 1) / kTile;
        const unsigned char* qsrc = qbuf + (size_t)islot * kblocks * kQBoxBytes;
        for (int t = 0; t < n_tiles; ++t, ++tcount) {
          const unsigned set = tcount & 1u;
          TIMED_WAIT(0, &acc_empty[set], ((tcount >> 1) & 1u) ^ 1u);  // the filter drained this set
          tc_fence_after();
          const uint32_t acc = tmem_base + set * (uint32_t)kTcNQ;
          for (int kb = 0; kb < kblocks; ++kb, ++ring) {
            const unsigned s = ring % n_slots;
            TIMED_WAIT(1, &full_bar[s], (ring / n_slots) & 1u);
            tc_fence_after();
            const uint64_t da = make_kmajor_sw128_desc(slots + (size_t)s * kSlotBytes);
            const uint64_t db = make_kmajor_sw128_desc(qsrc + kb * kQBoxBytes);
#pragma unroll
            for (int k8 = 0; k8 < kTcBK / 8; ++k8)  // UMMA K = 8 tf32 = 32 bytes: +2 in (addr >> 4)
              tc_mma_tf32(acc, da + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc,
                          (kb > 0 || k8 > 0) ? 1u : 0u);
            tc_commit(&empty_bar[s]);  // the slot may be refilled once these MMAs have read it
          }
          tc_commit(&acc_full[set]);
        }
        ++icount;
      }
    }
    __syncwarp();
  } else if (warp < 2 + kFilterWarps) {
    // =================================== filter ============================================
    const int quarter = warp & 3;  // TMEM lanes this warp may read
    const int r_local = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    unsigned icount = 0, tcount = 0;
    while (true) {
      const int islot = (int)(icount % kItemBufs);
      MBAR_WAIT(&q_full[islot], (icount / kItemBufs) & 1u);
      const TcItem rec = itm[islot];
      if (rec.nq == 0) break;
      const float* th = thr_s + islot * kTcNQ;
      const int n_tiles = (rec.row1 - rec.row0 + kTile - 1) / kTile;
      for (int t = 0; t < n_tiles; ++t, ++tcount) {
        const int row = rec.row0 + t * kTile + r_local;
        const bool valid = row < rec.row1;
        const unsigned mb = tcount % kMetaBufs;
        TIMED_WAIT(0, &meta_full[mb], (tcount / kMetaBufs) & 1u);
        float xn = reinterpret_cast<const float*>(meta + (size_t)mb * kMetaBytes)[((rec.row0 + t * kTile) & 3) + r_local];
        if (row >= (a.n_rows & ~3ll)) xn = valid ? a.xnorm[row] : 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&meta_empty[mb]);
        const unsigned set = tcount & 1u;
        TIMED_WAIT(1, &acc_full[set], (tcount >> 1) & 1u);
        tc_fence_after();
        uint32_t v0[16], v1[16];
        tc_ld16_nowait(lane_base + set * (uint32_t)kTcNQ, v0);
        tc_ld16_nowait(lane_base + set * (uint32_t)kTcNQ + 16u, v1);
        tc_wait_ld2(v0, v1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[set]);  // TMEM set free for the tile after next
        float ra, rb;  // bound = ra * dot + rb
        if (METRIC == NLSH_METRIC_L2) {
          ra = -2.0f;
          rb = (1.0f - kFilterC) * xn;
        } else {
          ra = -1.0f / fmaxf(sqrtf(xn), 1e-8f);
          rb = 0.f;
        }
        unsigned mask = 0;
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const float4 t0 = *reinterpret_cast<const float4*>(th + j4);
          const float4 t1 = *reinterpret_cast<const float4*>(th + 16 + j4);
          mask |= (fmaf(ra, __uint_as_float(v0[j4 + 0]), rb) <= t0.x ? 1u : 0u) << (j4 + 0);
          mask |= (fmaf(ra, __uint_as_float(v0[j4 + 1]), rb) <= t0.y ? 1u : 0u) << (j4 + 1);
          mask |= (fmaf(ra, __uint_as_float(v0[j4 + 2]), rb) <= t0.z ? 1u : 0u) << (j4 + 2);
          mask |= (fmaf(ra, __uint_as_float(v0[j4 + 3]), rb) <= t0.w ? 1u : 0u) << (j4 + 3);
          mask |= (fmaf(ra, __uint_as_float(v1[j4 + 0]), rb) <= t1.x ? 1u : 0u) << (16 + j4 + 0);
          mask |= (fmaf(ra, __uint_as_float(v1[j4 + 1]), rb) <= t1.y ? 1u : 0u) << (16 + j4 + 1);
          mask |= (fmaf(ra, __uint_as_float(v1[j4 + 2]), rb) <= t1.z ? 1u : 0u) << (16 + j4 + 2);
          mask |= (fmaf(ra, __uint_as_float(v1[j4 + 3]), rb) <= t1.w ? 1u : 0u) << (16 + j4 + 3);
        }
        if (!valid) mask = 0;
        const unsigned sb = tcount % kSurvBufs, use = tcount / kSurvBufs;
        TIMED_WAIT(2, &surv_empty[sb], (use & 1u) ^ 1u);  // the re-rank warps are done with this buffer
        int* cn = cnt + sb * kRerankWarps;
        uint16_t* sv = surv + (size_t)sb * kRerankWarps * kListCap;
        unsigned any = __reduce_or_sync(NLSH_FULL_MASK, mask);
        while (any) {  // warp-uniform: hand this warp's survivors of query j to j's owner warp
          const int j = __ffs(any) - 1;
          any &= any - 1;
          const int owner = j & (kRerankWarps - 1);
          const bool mine = (mask >> j) & 1u;
          const unsigned b = __ballot_sync(NLSH_FULL_MASK, mine);
          int pos = 0;
          if (lane == 0) pos = atomicAdd(&cn[owner], __popc(b));
          pos = __shfl_sync(NLSH_FULL_MASK, pos, 0);
          if (mine)
            sv[owner * kListCap + pos + __popc(b & ((1u << lane) - 1u))] = (uint16_t)((j << 8) | r_local);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&surv_full[sb]);
      }
      ++icount;
    }
  } else {
    // =================================== re-rank ===========================================
    const int rw = warp - 2 - kFilterWarps;  // owner index: lists of queries j = rw + kRerankWarps * i
    const int l4 = lane & 3;
    unsigned icount = 0, tcount = 0;
    while (true) {
      const int islot = (int)(icount % kItemBufs);
      MBAR_WAIT(&q_full[islot], (icount / kItemBufs) & 1u);
      const TcItem rec = itm[islot];
      if (rec.nq == 0) break;
      const unsigned char* qsrc = qbuf + (size_t)islot * kblocks * kQBoxBytes;
      float* th = thr_s + islot * kTcNQ;

      WarpTopK<1, int> top[kOwn];
      float ext[kOwn], qn2[kOwn], tau_seen[kOwn];
      int fidx[kOwn];
#pragma unroll
      for (int i = 0; i < kOwn; ++i) {
        const int j = rw + kRerankWarps * i;
        top[i].init(NLSH_ID_SENTINEL);
        tau_seen[i] = pos_inf();
        fidx[i] = own_f[islot * kTcNQ + j];
        ext[i] = own_ext[islot * kTcNQ + j];
        qn2[i] = own_qn2[islot * kTcNQ + j];
      }

      const int n_tiles = (rec.row1 - rec.row0 + kTile - 1) / kTile;
      for (int t = 0; t < n_tiles; ++t, ++tcount) {
        const unsigned sb = tcount % kSurvBufs, use = tcount / kSurvBufs;
        TIMED_WAIT(0, &surv_full[sb], use & 1u);
        int* cn = cnt + sb * kRerankWarps;
        const int n = cn[rw];
        const long long tb0 = timing ? clock64() : 0;
        if (n > 0) {  // warp-uniform
          const uint16_t* mine_sv = surv + ((size_t)sb * kRerankWarps + rw) * kListCap;
          const int trow0 = rec.row0 + t * kTile;
          for (int b0 = 0; b0 < n; b0 += 8) {
            const int sidx = b0 + (lane >> 2);
            const bool has = sidx < n;
            const int e = has ? (int)mine_sv[sidx] : 0;  // (query 0, row 0 of the tile): valid addresses
            const int j = e >> 8, r = e & 255;
            // lists carry the ROW index: inside a bucket rows ascend with the ids (nlsh_build_csr), so
            // (distance, row) orders like (distance, id); merge_partials_kernel maps rows to ids
            const int cand = trow0 + r;
            const long long td0 = timing ? clock64() : 0;
            const float dist = exact_distance<METRIC>(a.xs + (size_t)cand * a.d_pad, qsrc, a.d, j, l4);
            if (timing) tacc[2] += clock64() - td0;
            const int own = j >> kRerankShift;  // j = rw + kRerankWarps * own
            bool pass = false;
#pragma unroll
            for (int i = 0; i < kOwn; ++i)
              pass |= own == i && dist <= ext[i] && lex_less<int>(dist, cand, top[i].tau, top[i].tau_id);
            pass &= has && l4 == 0;
            if (__any_sync(NLSH_FULL_MASK, pass)) {
#pragma unroll
              for (int i = 0; i < kOwn; ++i) top[i].offer(dist, cand, pass && own == i, a.k);
            }
          }
          if (lane == 0) {
            if (a.stats != nullptr) {  // debug counters: survivors, batches
              atomicAdd(a.stats, (unsigned long long)n);
              atomicAdd(a.stats + 1, (unsigned long long)((n + 7) / 8));
            }
            cn[rw] = 0;  // before the buffer is handed back
#pragma unroll
            for (int i = 0; i < kOwn; ++i) {
              if (top[i].tau < tau_seen[i]) {  // only a list that tightened moves its threshold
                tau_seen[i] = top[i].tau;
                th[rw + kRerankWarps * i] = make_thr<METRIC>(fminf(top[i].tau, ext[i]), qn2[i], a.l2_slack);
              }
            }
          }
        }
        if (timing && n > 0) tacc[1] += clock64() - tb0;
        __syncwarp();
        if (lane == 0) mbar_arrive(&surv_empty[sb]);
      }

      // ------------------------------- item done: partial lists ------------------------------
#pragma unroll
      for (int i = 0; i < kOwn; ++i) {
        if (fidx[i] >= 0) {
          const size_t slot = ((size_t)fidx[i] * a.max_chunks + rec.chunk) * a.k;
          if (lane < a.k) {
            a.part_d[slot + lane] = top[i].d[0];
            a.part_id[slot + lane] = top[i].id[0];
          }
          if (lane == 0 && top[i].tau < pos_inf())
            atomicMin(reinterpret_cast<int*>(a.tau_g + fidx[i] / a.p),
                      __float_as_int(fmaxf(top[i].tau, 0.f)));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_empty[islot]);
      ++icount;
    }
  }

  if (timing && (warp == 0 || warp == 1 || warp == 2 || warp == 2 + kFilterWarps)) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 2 ? 2 : 3));
    tacc[3] = clock64() - t_begin;
    for (int i = 0; i < 4; ++i) atomicAdd(a.stats + 2 + role * 4 + i, (unsigned long long)tacc[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tc_dealloc(tmem_base, 64);
}

// qs[i] = qn[pairs[i] / p] (pair order = grouped by bucket), qs_norm[i] = |qs[i]|^2; one warp per pair.
__global__ void __launch_bounds__(256)
    gather_pair_queries_kernel(const float* __restrict__ qn, const int* __restrict__ pairs,
                               const int* __restrict__ n_valid, long long n_pairs, int p, int d_pad,
                               float* __restrict__ qs, float* __restrict__ qs_norm,
                               float* __restrict__ tau_g, long long n_queries) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gtid < n_queries) tau_g[gtid] = __int_as_float(0x7f800000);
  const int lane = lane_id();
  const long long warp = gtid >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  long long nv = *n_valid;
  if (nv > n_pairs) nv = n_pairs;
  for (long long i = warp; i < nv; i += n_warps) {
    const long long q = pairs[i] / p;
    float ss = 0.f;
    for (int c = lane * 4; c < d_pad; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(qn + q * d_pad + c);
      *reinterpret_cast<float4*>(qs + i * d_pad + c) = v;
      ss = fmaf(v.x, v.x, ss);
      ss = fmaf(v.y, v.y, ss);
      ss = fmaf(v.z, v.z, ss);
      ss = fmaf(v.w, v.w, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(NLSH_FULL_MASK, ss, o);
    if (lane == 0) qs_norm[i] = ss;
  }
}

// Initial tau_g: the exact k-th best distance of each query among the first rows (32 to 384, by bucket
// size, see nlsh_scan_tc_prepare) of its first probed bucket (one warp per query, 8 lanes per row, four
// rows per step).  This is the one-step-in-flight version, kept for A/B (NLSH_SEED_PIPE=0); the
// library runs seed_tau_pipe_kernel below.  Any k candidates bound the
// final k-th distance from above, so this is a valid threshold from the very first tile; without it
// every (query, bucket) list starts empty and the first 128-row tile of each bucket survives the
// filter whole.  The bound is inflated by a few ulps-of-the-sum because the re-rank sums the same
// terms in a different order.
constexpr int kMaxSeedRows = 512;

template <int METRIC>
__global__ void __launch_bounds__(128)
    seed_tau_kernel(const float* __restrict__ qn, const int* __restrict__ probes,
                    const int* __restrict__ offsets, const float* __restrict__ xs, int n_buckets, int p,
                    int d, int d_pad, int k, int seed_rows, long long n_queries, float* __restrict__ tau_g) {
  const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_queries) return;
  const int lane = lane_id();
  const int b = probes[q * p];
  if (b < 0 || b >= n_buckets) return;
  const int r0 = offsets[b];
  int n = offsets[b + 1] - r0;
  if (n > seed_rows) n = seed_rows;
  if (n < k) return;
  // 8 lanes per row, four rows per step: lane l8 reads the 16-byte chunks l8, l8 + 8, ... of its row
  // (each group of 8 lanes reads whole 128-byte lines) against its own chunks of the query
  const int l8 = lane & 7, g = lane >> 3;
  const int nvec = d_pad >> 2;  // float4 chunks per row (<= 32 on this path)
  float4 qv[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = l8 + 8 * i;
    qv[i] = c < nvec ? *reinterpret_cast<const float4*>(qn + q * d_pad + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  WarpTopK<1, int> top;
  top.init(NLSH_ID_SENTINEL);
  for (int base = 0; base < n; base += 4) {
    const int r = base + g;
    float acc = 0.f, xx = 0.f;
    if (r < n) {
      const float* xr = xs + (size_t)(r0 + r) * d_pad;
      float4 xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = l8 + 8 * i;
        xv[i] = c < nvec ? *reinterpret_cast<const float4*>(xr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col0 = 4 * (l8 + 8 * i);
        const float xa[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        const float qa[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (col0 + e < d) {
            if (METRIC == NLSH_METRIC_L2) {
              const float t = (qa[e] - xa[e]) + 1e-6f;
              acc = fmaf(t, t, acc);
            } else {
              acc = fmaf(qa[e], xa[e], acc);
              xx = fmaf(xa[e], xa[e], xx);
            }
          }
        }
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      acc += __shfl_xor_sync(NLSH_FULL_MASK, acc, o);
      if (METRIC != NLSH_METRIC_L2) xx += __shfl_xor_sync(NLSH_FULL_MASK, xx, o);
    }
    const float dist = METRIC == NLSH_METRIC_L2 ? acc : 1.0f - acc / fmaxf(sqrtf(xx), 1e-8f);
    top.offer(dist, r, l8 == 0 && r < n, k);
  }
  if (lane == 0 && top.tau < pos_inf()) {
    const float t = top.tau;
    tau_g[q] = METRIC == NLSH_METRIC_L2 ? t * 1.00002f + 1e-30f : t + 4e-6f + 2e-5f * fabsf(t);
  }
}

// The same seed with three row steps in flight per warp.  seed_tau_kernel has one: a step is four
// 16-byte loads per lane, then the math that needs them, so a warp spends each of its 64 steps (256
// sample rows) waiting out a full memory latency - measured 0.25 ms for the 10k-query batch of
// config 4, about 1.3 us per step with 32 warps per SM.  Here the query sits in shared memory (one
// float4 chunk per lane), which frees the registers for a ring of three row buffers: the loads of
// steps i + 1 and i + 2 are in flight while step i is scored.  Per query the arithmetic (8 lanes per
// row, four rows per step, columns in ascending order, xor-shuffle sum, offers in row order) and
// therefore tau is the same as in seed_tau_kernel.
__device__ __forceinline__ void seed_load_rows(float4 (&dst)[4], const float* __restrict__ xb, int r, int n,
                                               int d_pad, int l8) {
  const int nvec = d_pad >> 2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = l8 + 8 * i;
    dst[i] = (r < n && c < nvec) ? *reinterpret_cast<const float4*>(xb + (size_t)r * d_pad + 4 * c)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int METRIC>
__device__ __forceinline__ void seed_step(const float4 (&xv)[4], const float4* __restrict__ qs, int r, int n,
                                          int d, int l8, int k, WarpTopK<1, int>& top) {
  float acc = 0.f, xx = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = l8 + 8 * i;
    const int col0 = 4 * c;
    if (col0 < d) {
      const float4 qq = qs[c];
      const float xa[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
      const float qa[4] = {qq.x, qq.y, qq.z, qq.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (col0 + e < d) {
          if (METRIC == NLSH_METRIC_L2) {
            const float t = (qa[e] - xa[e]) + 1e-6f;
            acc = fmaf(t, t, acc);
          } else {
            acc = fmaf(qa[e], xa[e], acc);
            xx = fmaf(xa[e], xa[e], xx);
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    acc += __shfl_xor_sync(NLSH_FULL_MASK, acc, o);
    if (METRIC != NLSH_METRIC_L2) xx += __shfl_xor_sync(NLSH_FULL_MASK, xx, o);
  }
  const float dist = METRIC == NLSH_METRIC_L2 ? acc : 1.0f - acc / fmaxf(sqrtf(xx), 1e-8f);
  top.offer(dist, r, l8 == 0 && r < n, k);
}

template <int METRIC>
__global__ void __launch_bounds__(128, 6)
    seed_tau_pipe_kernel(const float* __restrict__ qn, const int* __restrict__ probes,
                         const int* __restrict__ offsets, const float* __restrict__ xs, int n_buckets, int p,
                         int d, int d_pad, int k, int seed_rows, long long n_queries,
                         float* __restrict__ tau_g) {
  __shared__ float4 q_s[4][32];  // per warp: 16-byte chunk c of its query at [c]
  const int wb = threadIdx.x >> 5;
  const long long q = (long long)blockIdx.x * 4 + wb;
  if (q >= n_queries) return;
  const int lane = lane_id();
  const int b = probes[q * p];
  if (b < 0 || b >= n_buckets) return;
  const int r0 = offsets[b];
  int n = offsets[b + 1] - r0;
  if (n > seed_rows) n = seed_rows;
  if (n < k) return;
  const int l8 = lane & 7, g = lane >> 3;
  q_s[wb][lane] = lane < (d_pad >> 2) ? *reinterpret_cast<const float4*>(qn + q * d_pad + 4 * lane)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  const float* xb = xs + (size_t)r0 * d_pad;
  const float4* qs = q_s[wb];
  float4 x0[4], x1[4], x2[4];
  seed_load_rows(x0, xb, g, n, d_pad, l8);
  seed_load_rows(x1, xb, 4 + g, n, d_pad, l8);
  seed_load_rows(x2, xb, 8 + g, n, d_pad, l8);
  WarpTopK<1, int> top;
  top.init(NLSH_ID_SENTINEL);
  for (int base = 0; base < n; base += 12) {  // warp-uniform
    seed_step<METRIC>(x0, qs, base + g, n, d, l8, k, top);
    seed_load_rows(x0, xb, base + 12 + g, n, d_pad, l8);
    if (base + 4 < n) {
      seed_step<METRIC>(x1, qs, base + 4 + g, n, d, l8, k, top);
      seed_load_rows(x1, xb, base + 16 + g, n, d_pad, l8);
    }
    if (base + 8 < n) {
      seed_step<METRIC>(x2, qs, base + 8 + g, n, d, l8, k, top);
      seed_load_rows(x2, xb, base + 20 + g, n, d_pad, l8);
    }
  }
  if (lane == 0 && top.tau < pos_inf()) {
    const float t = top.tau;
    tau_g[q] = METRIC == NLSH_METRIC_L2 ? t * 1.00002f + 1e-30f : t + 4e-6f + 2e-5f * fabsf(t);
  }
}

size_t scan_tc_smem(int kblocks, int n_slots) {
  return (size_t)n_slots * kSlotBytes + (size_t)kMetaBufs * kMetaBytes + (size_t)kItemBufs * kblocks * kQBoxBytes +
         kSurvBufs * kRerankWarps * kListCap * sizeof(uint16_t) + 4 * kItemBufs * kTcNQ * sizeof(float) +
         kSurvBufs * kRerankWarps * sizeof(int) + kItemBufs * sizeof(TcItem) +
         (2 * kMaxSlots + 4 + 2 * kItemBufs + 2 * kSurvBufs + 2 * kMetaBufs) * sizeof(uint64_t) + 16 + 1024;
}

}  // namespace

bool nlsh_scan_tc_supported(int d, int k, int metric) {
  return d >= 1 && (d + 3) / 4 * 4 <= kMaxKBlocks * kTcBK && k <= 32 &&
         (metric == NLSH_METRIC_L2 || metric == NLSH_METRIC_ANGULAR);
}

int nlsh_scan_tc_prepare(const float* qn, const int* pairs, const int* n_valid, long long n_pairs,
                         int p, int d_pad, float* qs, float* qs_norm, float* tau_g,
                         long long n_queries, const int* probes, const int* offsets, const float* xs,
                         long long n_rows, int n_buckets, int d, int k, int metric, cudaStream_t st) {
  long long threads = n_pairs * 32;
  if (threads < n_queries) threads = n_queries;
  long long blocks = (threads + 255) / 256;
  const long long cap = (long long)nlsh_num_sms() * 16;
  const long long floor_blocks = (n_queries + 255) / 256;
  if (blocks > cap) blocks = cap;
  if (blocks < floor_blocks) blocks = floor_blocks;  // every query's tau_g is written by thread gtid
  if (blocks < 1) blocks = 1;
  gather_pair_queries_kernel<<<(unsigned)blocks, 256, 0, st>>>(qn, pairs, n_valid, n_pairs, p, d_pad, qs,
                                                             qs_norm, tau_g, n_queries);
  NLSH_CUDA_TRY(nlsh_post_launch());
  // rows sampled per bucket: half of the average bucket, between 32 and 128; more for larger buckets,
  // where a sample step costs the seed kernel less than it saves the scan (measured with the pipelined
  // seed kernel, 10k queries, p = 8: 305-row buckets 128 -> 192 rows: step 0.586 -> 0.551 ms, 256 rows
  // the same; 2441-row buckets 256 -> 384 rows: 1.688 -> 1.654 ms, 512 rows no further gain)
  // (NLSH_SCAN_SEED=<rows> overrides; 0 = no seeding, for A/B runs)
  const long long avg = n_buckets > 0 ? n_rows / n_buckets : 0;
  int seed_rows = (int)(avg / 2 / 32 * 32);
  if (seed_rows < 32) seed_rows = 32;
  if (seed_rows > 128) seed_rows = 128;
  if (avg >= 256) seed_rows = 192;
  if (avg >= 1024) seed_rows = 256;
  if (avg >= 2048) seed_rows = 384;
  if (const char* env = getenv("NLSH_SCAN_SEED")) seed_rows = atoi(env);
  if (seed_rows <= 0) return NLSH_OK;
  if (seed_rows > kMaxSeedRows) seed_rows = kMaxSeedRows;
  const unsigned sb = (unsigned)((n_queries + 3) / 4);
  // NLSH_SEED_PIPE=0: the one-step-in-flight kernel (A/B; tau does not depend on the choice)
  const char* pipe = getenv("NLSH_SEED_PIPE");
  if (pipe == nullptr || atoi(pipe) != 0) {
    if (metric == NLSH_METRIC_L2)
      seed_tau_pipe_kernel<NLSH_METRIC_L2><<<sb, 128, 0, st>>>(qn, probes, offsets, xs, n_buckets, p, d, d_pad, k,
                                                              seed_rows, n_queries, tau_g);
    else
      seed_tau_pipe_kernel<NLSH_METRIC_ANGULAR><<<sb, 128, 0, st>>>(qn, probes, offsets, xs, n_buckets, p, d, d_pad,
                                                                   k, seed_rows, n_queries, tau_g);
    return nlsh_check_cuda(nlsh_post_launch(), "seed_tau_pipe_kernel launch");
  }
  if (metric == NLSH_METRIC_L2)
    seed_tau_kernel<NLSH_METRIC_L2><<<sb, 128, 0, st>>>(qn, probes, offsets, xs, n_buckets, p, d, d_pad, k,
                                                       seed_rows, n_queries, tau_g);
  else
    seed_tau_kernel<NLSH_METRIC_ANGULAR><<<sb, 128, 0, st>>>(qn, probes, offsets, xs, n_buckets, p, d, d_pad,
                                                            k, seed_rows, n_queries, tau_g);
  return nlsh_check_cuda(nlsh_post_launch(), "seed_tau_kernel launch");
}

int nlsh_scan_tc_launch(int metric, TcScanArgs a, cudaStream_t st) {
  a.kblocks = (a.d_pad + kTcBK - 1) / kTcBK;
  a.l2_slack = 2.1e-6f * sqrtf((float)a.d);
  int n_slots = kMaxSlots;
  // each K block's slot is freed by its own tcgen05.commit, so any ring depth >= 2 makes progress
  while (n_slots > 3 && scan_tc_smem(a.kblocks, n_slots) > 224 * 1024) --n_slots;
  a.n_slots = n_slots;
  const size_t smem = scan_tc_smem(a.kblocks, n_slots);
  CUtensorMap map_x, map_q;
  int rc;
  if ((rc = tc_make_map(&map_x, a.xs, a.n_rows, a.d_pad, kTile)) != NLSH_OK) return rc;
  if ((rc = tc_make_map(&map_q, a.qs, a.n_pairs, a.d_pad, kTcNQ)) != NLSH_OK) return rc;
  const int grid = nlsh_num_sms();
  if (metric == NLSH_METRIC_L2) {
    auto kern = scan_tc_kernel<NLSH_METRIC_L2>;
    NLSH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(a, map_x, map_q);
  } else {
    auto kern = scan_tc_kernel<NLSH_METRIC_ANGULAR>;
    NLSH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(a, map_x, map_q);
  }
  return nlsh_check_cuda(nlsh_post_launch(), "scan_tc_kernel launch");
}
