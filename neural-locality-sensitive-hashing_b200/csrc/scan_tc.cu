// Candidate scan with a tensor-core FILTER in front of the exact distance (the default path of
// nlsh_query_scan_topk for k <= 128 and batches in which bucket tiles are shared).  Same
// job as scan.cu::scan_kernel - the per-query gather + distance_func + topk of Indexer.query
// (nlsh/indexer.py:62-95) - and the same results: every distance that reaches a top-k list is computed
// in fp32 in the reference's difference form (nlsh/data.py:201 F.pairwise_distance, nlsh/data.py:109
// 1 - cosine_similarity), never from the GEMM expansion.  The GEMM only decides which (row, query)
// pairs cannot matter.
//
// Why: with the 10k-query batches of the BASELINE configs every bucket is probed by ~20 queries.  The
// fp32 SIMT kernel fetches a bucket tile once for all of them but still pays 3 lane-ops per (row,
// query, column) and is fp32-issue bound.  Here one tcgen05.mma (kind::tf32, raw fp32 bits as operands)
// produces the 128 x 32 dot products of a row tile against the item's queries into TMEM; a thread per
// row turns each into a LOWER BOUND of the exact distance and compares it with the query's threshold.
// The few survivors (well under 1 % of the pairs) are scored exactly, one thread per survivor, and
// appended to the query's candidate buffer; merge_cands_kernel (scan.cu) selects the top k.
//
// Exactness of the filter.  tf32 keeps 10 mantissa bits of each operand, so
// |dot_tc - <q, x>| <= 2^-9 * 1.02 * |q| |x|  (Cauchy-Schwarz; the 1.02 covers the fp32 accumulate).
//   L2:  d2 = |q|^2 + |x|^2 - 2 <q,x> >= (1 - c)(|q|^2 + |x|^2) - 2 dot_tc,  c = 2^-9 * 1.02 + 4e-5
//        (the 4e-5 covers fp32 rounding of the two norms).  The reference's
//        D = sum((q - x + 1e-6)^2) >= d2 - 2e-6 sqrt(dim) sqrt(d2) - 1e-5 D, so a pair whose bound
//        exceeds tau_eff = y + a sqrt(y) + a^2  (y = tau * 1.0001, a = 2.1e-6 sqrt(dim)) has D > tau.
//   angular: 1 - cos >= 1 - dot_tc / |x| - 2.1e-3.
// A pair is dropped only when its bound is above the threshold, and the threshold of a query is always
// an upper bound tau_g[query] of its final k-th best distance, so the candidate buffer holds every row
// of the final top k; ties are broken by the (distance, id) order in the merge, as in the SIMT kernel.
//
// Thresholds.  tau_g starts as the seed (seed_tau_kernel: exact k-th best among the first rows of the
// query's probed buckets) and is lowered while the scan runs by the THRESHOLD LADDER: level l stands for
// the bound t_l = tau0 (1 - l/32); every scored candidate with distance <= t_l is counted in
// ladder[query][l] (one atomicAdd), and as soon as the levels >= l hold k candidates, t_l is an upper
// bound of the k-th best distance and goes into tau_g (atomicMin).  Lock-free, order-independent (any k
// candidates bound the k-th best), shared by all probes of a query on all SMs; an item reads tau_g
// when the planner picks it up and again whenever the planner waits.  There are no per-item lists and nothing in the kernel waits for a
// threshold: the stages below only hand tiles forward.
//
// Roles in a CTA (one persistent CTA per SM, 352 threads; mbarrier hand-offs only):
//   warp 0         planner (32 lanes): owns the work queue (one atom.global.add per item, its result consumed an
//                  iteration later), looks four items ahead (item index -> record -> per-pair query index / norm
//                  -> tau_g, one dependent load per iteration, so none of these latencies is exposed), publishes
//                  the per-query state (threshold, bound, query index) to the item-state ring, TMA-loads the item's
//                  queries (box 32 rows x 32 fp32 per K block, SWIZZLE_128B; the B operand) into the query-buffer
//                  ring and, whenever it has to wait for a ring slot, re-reads tau_g for the queries of the items in
//                  flight and rewrites the thresholds that got lower;
//   warp 10 lane 0 tile streamer: for every published item the row tiles (box 128 rows x 32 fp32 per K block;
//                  32-row boxes for the ragged last tile of a bucket) into the slot ring, row norms alongside
//                  (bulk copy).  Nothing but mbarrier waits and TMA issues, running slot / parity counters: the
//                  single producer warp that used to do the planner's work too was busy - not waiting - 90 % of
//                  the time (a run-time `ring % n_slots` alone is a 200-cycle dependent chain per K block);
//   warp 1 lane 0  MMA issuer: per tile and K block 4 x tcgen05.mma M128 N32 K8 into one of 16 TMEM accumulator
//                  sets (8 in use by default: with 16 the tiles between the TMA and the filter are more than the
//                  L2 holds, and the scorer re-reads rows from L2); tcgen05.commit frees the slot, the last one of
//                  a tile signals acc_full, the one after an item's last tile frees its query buffer;
//   warps 2-9      filter + score, two groups of four (one warp per TMEM lane quarter), group g takes
//                  the tiles with tile index = g (mod 2): tcgen05.ld the row's 32 scores (then the TMEM
//                  set is free again), bound + compare, survivors go to the warp's PRIVATE queue in
//                  shared memory (no atomics); whenever 32 are queued the warp scores them one per lane
//                  (score_batch): id, bound and ladder scale are requested first, the row is re-read from L2
//                  (it has just streamed through), candidates within the bound are appended to the query's
//                  buffer, the append's round trip overlaps the ladder's.
//
// Variants (template parameters): QGLOBAL - the scorer reads the query from the pair-ordered global copy; queue
// entries then do not refer to an item slot (buckets of a few tiles: 8-GPU shards; wide rows).  Otherwise the
// query comes from the item's shared-memory copy (half the scorer's L2 requests): a queue entry pins its item
// slot, which is released when the queue's front has passed the item's last entry - a partial batch is scored
// only when the warp would otherwise block.  WIDE (d_pad > 128, config 5's 960-wide rows, or 128 queries per
// item) - the queries do not fit shared memory, so the K block of the item's queries travels with each K block
// of a row tile (both by TMA; the queries come from L2), the accumulate runs over d_pad / 32 K blocks, and the
// scorer is QGLOBAL.  V8 - the scorer reads rows 32 bytes per load (LDG.256, not allocated in L1).
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

// -DNLSH_TC_ROLE_STATS (the `stats` target of the Makefile, lib/libnlsh_b200_stats.so): cycles every role
// spends in each of its waits, summed over all CTAs into stats[6..] (scripts/dbg_tc_roles.py prints them).
#ifdef NLSH_TC_ROLE_STATS
#define ROLE_T0() const long long role_t0 = clock64(); long long role_w[4] = {0, 0, 0, 0}
#define ROLE_WAIT(slot, bar, par)        \
  do {                                   \
    const long long t_ = clock64();      \
    MBAR_WAIT(bar, par);                 \
    role_w[slot] += clock64() - t_;      \
  } while (0)
#define ROLE_TIC() const long long role_tic = clock64()
#define ROLE_TOC(slot) role_w[slot] += clock64() - role_tic
#define ROLE_REPORT(base)                                                                   \
  do {                                                                                      \
    if (a.stats != nullptr && lane == 0) {                                                  \
      for (int i_ = 0; i_ < 4; ++i_) atomicAdd(a.stats + (base) + i_, (unsigned long long)role_w[i_]); \
      atomicAdd(a.stats + (base) + 4, (unsigned long long)(clock64() - role_t0));           \
    }                                                                                       \
  } while (0)
#else
#define ROLE_T0()
#define ROLE_WAIT(slot, bar, par) MBAR_WAIT(bar, par)
#define ROLE_TIC()
#define ROLE_TOC(slot)
#define ROLE_REPORT(base)
#endif

constexpr int kGroups = 2;                      // filter groups (tiles alternate between them)
constexpr int kFilterWarps = 4 * kGroups;       // one per TMEM lane quarter and group
constexpr int kTile = 128;                      // rows per tile = UMMA M
constexpr int kStreamWarp = 2 + kFilterWarps;    // warp 10: streams the row tiles
constexpr int kThreads = 96 + 32 * kFilterWarps;  // 352
constexpr int kTmemCols = 512;                  // TMEM accumulator ring: 16 sets of 32 columns or 4 sets of 128
constexpr int kMaxAccSets = 16;
constexpr uint32_t kSlotBytes = kTile * kTcBK * sizeof(float);   // 16 KB: one K block of a row tile
constexpr uint32_t kSubBoxBytes = 32 * kTcBK * sizeof(float);    // 4 KB: a 32-row box of the ragged last tile
constexpr uint32_t kQBoxBytes = kTcNQ * kTcBK * sizeof(float);   // 4 KB: one K block of the queries
constexpr int kMaxSlots = 16;
constexpr int kMetaFloats = kTile + 4;          // a tile's row norms, from the 16-byte boundary below its first row
constexpr uint32_t kMetaBytes = 528;            // kMetaFloats * 4 (a multiple of 16: bulk-copy destination)
constexpr int kMetaBufs = 16;                   // row-norm ring depth (tiles)
#ifndef NLSH_TC_STATE_DEEP  // A/B builds: make variant NAME=sb4 DEFS=-DNLSH_TC_STATE_DEEP=4
#define NLSH_TC_STATE_DEEP 4
#endif
constexpr int kStateBufsDeep = NLSH_TC_STATE_DEEP;              // item-state ring depth when the scorer does not use the item's queries
constexpr int kItemBufs = 4;                    // query-buffer ring depth: buckets of a few tiles are shorter than the
                                                // pipeline, so several items must be in flight
constexpr int kMaxKBlocks = 4;                  // d_pad <= 128: the item's queries stay in shared memory
constexpr int kMaxWideKBlocks = 128;            // WIDE: d_pad <= 4096
constexpr int kQueueCap = 64;                   // per-warp survivor queue (a power of two >= 2 * 32)
constexpr float kFilterC = 0.001953125f * 1.02f + 4e-5f;  // see header comment
constexpr float kAngularC = 2.1e-3f;

static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation: power of two <= 512");
static_assert(kMetaBufs % kGroups == 0, "a row-norm buffer must always belong to the same filter group");

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

// Query j of an item in shared memory: per K block a SWIZZLE_128B box of kTcNQ rows x 128 bytes, the
// 16-byte chunk c of row j stored at chunk position c ^ (j & 7).
struct TcQueryShared {
  const unsigned char* base;  // the item's query buffer + j * 128
  int sw;                     // j & 7
  __device__ __forceinline__ float4 load4(int v) const {
    return *reinterpret_cast<const float4*>(base + (v >> 3) * kQBoxBytes + (((v & 7) ^ sw) << 4));
  }
  __device__ __forceinline__ void load8(int v, float4& a, float4& b) const {
    a = load4(v);
    b = load4(v + 1);
  }
  __device__ __forceinline__ float load1(int c) const {
    const int v = c >> 2;
    return *reinterpret_cast<const float*>(base + (v >> 3) * kQBoxBytes + (((v & 7) ^ sw) << 4) + ((c & 3) << 2));
  }
};

struct SmemLayout {
  unsigned char* slots;   // [n_slots][16 KB (+ NQ * 128 bytes of queries, WIDE)]
  unsigned char* qbuf;    // [kItemBufs][kblocks][NQ * 128 bytes] (not WIDE)
  unsigned char* meta;    // [kMetaBufs][kMetaBytes] row norms
  // item state ring, [n_state] entries (state_bufs(nq)); the item's query buffer is ring slot icount % kItemBufs
  float* thr_s;           // [n_state][NQ] filter thresholds
  float* tau_s;           // [n_state][NQ] the bound each threshold was made from (planner only)
  float* qn2_s;           // [n_state][NQ] |q|^2 (planner only)
  int* qi_s;              // [n_state][NQ] query indices (-1 unused)
  TcItem* itm;            // [n_state]
  int* wq_row;            // [kFilterWarps][kQueueCap] survivor queues: row of x_sorted
  int* wq_meta;           // [kFilterWarps][kQueueCap] pair index (QGLOBAL) or (item slot << 8) | query j
  int* wq_qi;             // [kFilterWarps][kQueueCap] query index
  uint64_t* full_bar;     // [kMaxSlots]
  uint64_t* empty_bar;    // [kMaxSlots]
  uint64_t* itm_full;     // [kStateBufsDeep] item state published
  uint64_t* q_empty;      // [kStateBufsDeep] item state released by every filter warp
  uint64_t* qb_full;      // [kItemBufs] the item's queries have landed in their buffer (not WIDE)
  uint64_t* qb_empty;     // [kItemBufs] the item's MMAs have read them
  uint64_t* acc_full;     // [kMaxAccSets]
  uint64_t* acc_empty;    // [kMaxAccSets]
  uint64_t* meta_full;    // [kMetaBufs]
  uint64_t* meta_empty;   // [kMetaBufs]
  uint32_t* tmem_slot;
};

// Depth of the item-state ring in shared memory (sized for the deepest user of a group size).
__host__ __device__ constexpr int state_bufs(int nq) { return nq == kTcNQ ? kStateBufsDeep : kItemBufs; }

__host__ __device__ inline size_t smem_fixed_bytes(int kblocks, bool wide, int nq) {
  return (wide ? 0 : (size_t)kItemBufs * kblocks * nq * 128) + (size_t)kMetaBufs * kMetaBytes +
         4 * state_bufs(nq) * nq * sizeof(float) + state_bufs(nq) * sizeof(TcItem) +
         3 * kFilterWarps * kQueueCap * sizeof(int) +
         (2 * kMaxSlots + 2 * kStateBufsDeep + 2 * kItemBufs + 2 * kMaxAccSets + 2 * kMetaBufs) * sizeof(uint64_t) + 16;
}

__device__ __forceinline__ SmemLayout carve_smem(unsigned char* base, int n_slots, int kblocks, bool wide, int nq) {
  SmemLayout s;
  s.slots = base;
  s.qbuf = s.slots + (size_t)n_slots * (kSlotBytes + (wide ? nq * 128u : 0u));
  s.meta = s.qbuf + (wide ? 0 : (size_t)kItemBufs * kblocks * nq * 128);
  s.thr_s = reinterpret_cast<float*>(s.meta + (size_t)kMetaBufs * kMetaBytes);
  const int ns = state_bufs(nq);
  s.tau_s = s.thr_s + ns * nq;
  s.qn2_s = s.tau_s + ns * nq;
  s.qi_s = reinterpret_cast<int*>(s.qn2_s + ns * nq);
  s.itm = reinterpret_cast<TcItem*>(s.qi_s + ns * nq);
  s.wq_row = reinterpret_cast<int*>(s.itm + ns);
  s.wq_meta = s.wq_row + kFilterWarps * kQueueCap;
  s.wq_qi = s.wq_meta + kFilterWarps * kQueueCap;
  s.full_bar = reinterpret_cast<uint64_t*>(s.wq_qi + kFilterWarps * kQueueCap);
  s.empty_bar = s.full_bar + kMaxSlots;
  s.itm_full = s.empty_bar + kMaxSlots;
  s.q_empty = s.itm_full + kStateBufsDeep;
  s.qb_full = s.q_empty + kStateBufsDeep;
  s.qb_empty = s.qb_full + kItemBufs;
  s.acc_full = s.qb_empty + kItemBufs;
  s.acc_empty = s.acc_full + kMaxAccSets;
  s.meta_full = s.acc_empty + kMaxAccSets;
  s.meta_empty = s.meta_full + kMetaBufs;
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.meta_empty + kMetaBufs);
  return s;
}

// What the producer knows about an item it has not published yet (one pipeline stage per field group).
template <int CH>
struct Ahead {
  TcItem rec;     // warp-uniform
  int qi[CH];     // lane l, chunk c: query index of the item's pair 32 c + l (-1 past nq)
  float qn2[CH];  // |q|^2 of that query
};

// What the scorer needs of the kernel arguments, by value (the scorer is one out-of-line function per
// kernel: three call sites, and its unrolled distance loop is the largest piece of code in the kernel).
struct ScoreCtx {
  const float* xs;
  const int* ids;
  const float* qs;
  const int* pq;
  float* tau_g;
  const float* tau0;
  int* ladder;
  int* cand_n;
  TcCand* cand;
  const unsigned char* qbuf;  // shared memory: the item ring's query boxes (not QGLOBAL)
  const int* qi_s;            // shared memory: the item ring's query indices
  int cap, k, d, d_pad, kblocks;
};

// Scores the first n (<= 32) entries of a warp's survivor queue, one per lane, and appends the
// candidates within their query's bound to its buffer.  Returns nothing; the caller advances the queue.
// Queue entry: the row of x_sorted and, QGLOBAL, the pair index (query vector = qs[pair], query index =
// pq[pair]) or else (item slot << 8) | query j of the item whose queries are in shared memory.
// Inlined at its three call sites: out of line (one copy, less instruction-cache pressure) it measured 10 %
// faster on config 2 but 10 - 60 % slower on configs 3, 5 and on 8-GPU shards (profiles/r2_experiments).
#ifndef NLSH_SCORE_INLINE  // A/B: -DNLSH_SCORE_INLINE=__noinline__
#define NLSH_SCORE_INLINE __forceinline__
#endif
#ifndef NLSH_TC_QS_BATCH  // row float4s in flight per lane when the query is read from shared memory
#define NLSH_TC_QS_BATCH 8
#endif
template <int METRIC, bool QGLOBAL, bool V8>
__device__ NLSH_SCORE_INLINE void score_batch(const ScoreCtx a, const int* q_row, const int* q_meta, const int* q_qi,
                                         unsigned head, int n, int lane, unsigned& n_appended) {
  if (lane < n) {
    const unsigned e = (head + (unsigned)lane) & (kQueueCap - 1);
    const int row = q_row[e];
    const int meta = q_meta[e];
    const int qi = q_qi[e];
    // what depends only on (row, query) is requested before the row is read: one L2 round trip for a batch's
    // loads, one more for the append and the ladder counts of the lanes that found a candidate
    const int id = __ldg(a.ids + row);
    // the bound is read fresh: other items of the same query may have lowered it since this item was picked up
    const float ext = __ldcg(a.tau_g + qi);
    const float t0 = a.ladder != nullptr ? __ldg(a.tau0 + qi) : 0.f;
    float dist;
    if (QGLOBAL) {
      TcQueryGlobal q;
      q.q = a.qs + (size_t)meta * a.d_pad;
      dist = tc_thread_distance<METRIC, 16, V8>(a.xs + (size_t)row * a.d_pad, q, a.d);
    } else {
      const int islot = meta >> 8, j = meta & 255;
      TcQueryShared q;
      q.base = a.qbuf + (size_t)islot * a.kblocks * kQBoxBytes + j * 128;
      q.sw = j & 7;
      dist = tc_thread_distance<METRIC, NLSH_TC_QS_BATCH, V8>(a.xs + (size_t)row * a.d_pad, q, a.d);
    }
    if (dist <= ext) {
      const int pos = atomicAdd(a.cand_n + qi, 1);  // its round trip overlaps the ladder's
      if (a.ladder != nullptr && t0 > 0.f && t0 < pos_inf()) {
        // threshold ladder: count this candidate at the tightest level it satisfies, then look for the
        // tightest level whose cumulative count reaches k (read after the add: the lanes of a batch often
        // score candidates of the same query, and each should see the others' counts)
        int l = (int)((1.0f - dist / t0) * (float)kTcLadderDen);
        l = l < 0 ? 0 : (l > kTcLadder - 1 ? kTcLadder - 1 : l);
        while (l > 0 && !(dist <= t0 * (1.0f - (float)l * (1.0f / kTcLadderDen)))) --l;
        int* lad = a.ladder + (size_t)qi * kTcLadder;
        atomicAdd(lad + l, 1);
        if (l > 0) {
          int cnt[kTcLadder];
#pragma unroll
          for (int m4 = 0; m4 < kTcLadder; m4 += 4) {
            const int4 c4 = __ldcg(reinterpret_cast<const int4*>(lad + m4));
            cnt[m4] = c4.x;
            cnt[m4 + 1] = c4.y;
            cnt[m4 + 2] = c4.z;
            cnt[m4 + 3] = c4.w;
          }
          int cum = 0, best = 0;
#pragma unroll
          for (int m = kTcLadder - 1; m >= 1; --m) {
            cum += cnt[m];
            if (best == 0 && cum >= a.k) best = m;
          }
          if (best > 0)
            atomicMin(reinterpret_cast<int*>(a.tau_g + qi),
                      __float_as_int(t0 * (1.0f - (float)best * (1.0f / kTcLadderDen))));
        }
      }
      if (pos < a.cap) {
        TcCand c;
        c.d = dist;
        c.id = id;
        a.cand[(size_t)qi * a.cap + pos] = c;
      }
      ++n_appended;
    }
  }
  __syncwarp();
}

// NLSH_TC_LB: the thread count the scan kernel is compiled for (it always launches kThreads = 352) - a register cap
// in disguise for A/B builds (`make variant NAME=r128 DEFS=-DNLSH_TC_LB=512`: 128 registers, 12 - 112 bytes of
// spills).  The idea was to leave room on every SM for the small kernels of the next batch beside the persistent
// scan CTA; measured, the lanes of nlsh.parallel.PipelinedSearch overlap no better (config 4: 1.131 against 1.133
// ms per batch, shard-sized 0.370 against 0.367; profiles/r2_experiments/ab_scan_regcap.txt), so the default stays.
#ifndef NLSH_TC_LB
#define NLSH_TC_LB kThreads
#endif
template <int METRIC, int NQ, bool WIDE, bool QGLOBAL, bool V8>
__global__ void __launch_bounds__(NLSH_TC_LB, 1)
    scan_tc_kernel(const TcScanArgs a, const __grid_constant__ CUtensorMap map_x,
                   const __grid_constant__ CUtensorMap map_x32, const __grid_constant__ CUtensorMap map_q) {
  static_assert(!WIDE || QGLOBAL, "wide rows: the queries are not resident in shared memory");
  static_assert(NQ == kTcNQ || (NQ == kTcNQMax && WIDE), "128 queries per item: their K blocks travel with the tiles");
  constexpr int kCh = NQ / 32;                  // 32-column chunks of the accumulator
  constexpr int kSets = kTmemCols / NQ;         // TMEM accumulator ring depth
  static_assert(kSets % kGroups == 0 && kSets >= 2 && kSets <= kMaxAccSets, "accumulator ring");
  constexpr uint32_t kQBytes = NQ * 128u;       // one K block of the item's queries
  // Item-state ring depth.  With the scorer's queries in global memory an item's query buffer is dead once its
  // MMAs have run, so the state ring (thresholds, query indices: 0.5 KB per item) runs 16 items deep while
  // the 16 KB query buffers recycle behind the MMA issuer - on 8-GPU shards (2 - 3 tiles per item) four items
  // were not enough to keep tiles in flight while a filter warp scores a batch.  Scorer queries in shared
  // memory: a queue entry pins its item's buffer, one ring, four deep.
  constexpr int kSB = (QGLOBAL && NQ == kTcNQ) ? kStateBufsDeep : kItemBufs;
  extern __shared__ unsigned char stc_smem_raw[];
  unsigned char* base = stc_smem_raw + ((1024u - (smem_u32(stc_smem_raw) & 1023u)) & 1023u);
  const SmemLayout s = carve_smem(base, a.n_slots, a.kblocks, WIDE, NQ);
  constexpr uint32_t kSlotStride = kSlotBytes + (WIDE ? kQBytes : 0u);  // WIDE: rows + the queries' K block

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const unsigned n_slots = (unsigned)a.n_slots;
  const int kblocks = a.kblocks;
  // accumulator sets in use (a power of two <= kSets): how many tiles the MMAs may run ahead of the filter
  const unsigned sets_log2 = (unsigned)a.sets_log2 < (unsigned)(31 - __clz(kSets)) ? (unsigned)a.sets_log2 : (unsigned)(31 - __clz(kSets));
  const unsigned sets_mask = (1u << sets_log2) - 1u;

  if (tid == 0) {
    for (int i = 0; i < a.n_slots; ++i) {
      mbar_init(&s.full_bar[i], 1);
      mbar_init(&s.empty_bar[i], 1);
    }
    for (int i = 0; i < kMetaBufs; ++i) {
      mbar_init(&s.meta_full[i], 1);
      mbar_init(&s.meta_empty[i], 4);
    }
    for (int i = 0; i < kStateBufsDeep; ++i) {
      mbar_init(&s.itm_full[i], 1);
      mbar_init(&s.q_empty[i], kFilterWarps);
    }
    for (int i = 0; i < kItemBufs; ++i) {
      mbar_init(&s.qb_full[i], 1);
      mbar_init(&s.qb_empty[i], 1);
    }
    for (int i = 0; i < kSets; ++i) {
      mbar_init(&s.acc_full[i], 1);
      mbar_init(&s.acc_empty[i], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tc_alloc(s.tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s.tmem_slot;

  if (warp == 0) {
    // =================================== planner ==========================================
    ROLE_T0();
    int total = *a.n_items;
    if (total > a.max_items) total = a.max_items;
    // lane 0 draws the next item index; the value is broadcast one iteration later, so the atomic's round
    // trip is never waited for (inline PTX: the compiler turns atomicAdd on a uniform address into a
    // warp-aggregated atomic with a shuffle right behind it)
    auto claim = [&]() {
      int v = 0;
      if (lane == 0) asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(v) : "l"(a.item_counter) : "memory");
      return v;
    };
    auto load_rec = [&](int idx) {
      TcItem r;
      if (idx < total) {
        r = a.items[idx];
      } else {
        r.row0 = r.row1 = r.pair_base = r.nq = r.chunk = 0;
        r.pad[0] = r.pad[1] = r.pad[2] = 0;
      }
      return r;
    };
    auto load_pair = [&](Ahead<kCh>& h) {
#pragma unroll
      for (int c = 0; c < kCh; ++c) {
        h.qi[c] = -1;
        h.qn2[c] = 0.f;
        if (32 * c + lane < h.rec.nq) {
          h.qi[c] = __ldg(a.pq + h.rec.pair_base + 32 * c + lane);
          h.qn2[c] = __ldg(a.pqn2 + h.rec.pair_base + 32 * c + lane);
        }
      }
    };
    auto load_tau = [&](const Ahead<kCh>& h, float (&tau)[kCh]) {
#pragma unroll
      for (int c = 0; c < kCh; ++c) tau[c] = h.qi[c] >= 0 ? __ldcg(a.tau_g + h.qi[c]) : neg_inf();
    };
    // Threshold refresh: other items of the same queries (on any SM) lower tau_g while the published items
    // are streamed.  Whenever the planner has to wait for an item slot it re-reads tau_g for the queries of
    // every published slot and rewrites the thresholds that got lower (the filter may read the old or the
    // new value: both are upper bounds; the planner is the only writer of the item slots).
    auto refresh = [&](int n_live) {
#ifndef NLSH_NO_REFRESH  // A/B: -DNLSH_NO_REFRESH
      float t[kSB][kCh];
#pragma unroll
      for (int b = 0; b < kSB; ++b) {
#pragma unroll
        for (int c = 0; c < kCh; ++c) {
          const int qi = b < n_live ? s.qi_s[b * NQ + 32 * c + lane] : -1;
          t[b][c] = qi >= 0 ? __ldcg(a.tau_g + qi) : pos_inf();
        }
      }
#pragma unroll
      for (int b = 0; b < kSB; ++b) {
#pragma unroll
        for (int c = 0; c < kCh; ++c) {
          const int e = b * NQ + 32 * c + lane;
          if (t[b][c] < s.tau_s[e]) {
            s.tau_s[e] = t[b][c];
            s.thr_s[e] = make_thr<METRIC>(t[b][c], s.qn2_s[e], a.l2_slack);
          }
        }
      }
      __syncwarp();
#endif
    };
    // look-ahead pipeline: index of item i+3, record of i+2, pair state of i+1, tau_g of i
    int idx3 = __shfl_sync(NLSH_FULL_MASK, claim(), 0);
    Ahead<kCh> cur, nx1, nx2;
    cur.rec = load_rec(idx3);
    idx3 = __shfl_sync(NLSH_FULL_MASK, claim(), 0);
    nx1.rec = load_rec(idx3);
    idx3 = __shfl_sync(NLSH_FULL_MASK, claim(), 0);
    nx2.rec = load_rec(idx3);
    int pend = claim();
    load_pair(cur);
    load_pair(nx1);
    float tau_cur[kCh], tau_nx1[kCh];
    load_tau(cur, tau_cur);
    unsigned icount = 0;
    while (true) {
      // issue the look-ahead loads first: they complete while the planner waits for the item slot
      idx3 = __shfl_sync(NLSH_FULL_MASK, pend, 0);
      const TcItem rec3 = load_rec(idx3);                      // item i+3
      pend = claim();                                          // index of item i+4
      load_pair(nx2);                                          // item i+2 (its record arrived last iteration)
      load_tau(nx1, tau_nx1);                                  // item i+1

      const int islot = (int)(icount % kSB);         // item-state slot
      const int qslot = (int)(icount % kItemBufs);   // query-buffer slot
      {
        const unsigned par = ((icount / kSB) & 1u) ^ 1u;
        const unsigned qpar = ((icount / kItemBufs) & 1u) ^ 1u;
        ROLE_TIC();
        while (!mbar_try_wait_short(&s.q_empty[islot], par)) refresh(icount < kSB ? (int)icount : kSB);
        if (!WIDE)  // (WIDE: no query buffers, the queries travel with the row tiles)
          while (!mbar_try_wait_short(&s.qb_empty[qslot], qpar)) refresh(icount < kSB ? (int)icount : kSB);
        ROLE_TOC(0);
      }
      const TcItem rec = cur.rec;
      if (rec.nq == 0) {
        if (lane == 0) {
          s.itm[islot].nq = 0;  // end of work
          mbar_arrive(&s.itm_full[islot]);
        }
        break;
      }
#pragma unroll
      for (int c = 0; c < kCh; ++c) {  // lane l: state of the item's query 32 c + l
        float th = neg_inf(), tau = neg_inf();
        if (32 * c + lane < rec.nq) {
          tau = tau_cur[c];
          th = make_thr<METRIC>(tau, cur.qn2[c], a.l2_slack);
        }
        const int e = islot * NQ + 32 * c + lane;
        s.thr_s[e] = th;
        s.tau_s[e] = tau;
        s.qn2_s[e] = cur.qn2[c];
        s.qi_s[e] = cur.qi[c];
      }
      if (lane == 0) s.itm[islot] = rec;
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s.itm_full[islot]);  // the streamer starts on the row tiles, the filter sees the thresholds
        if (!WIDE) {
          mbar_arrive_expect_tx(&s.qb_full[qslot], (unsigned)kblocks * kQBytes);
          unsigned char* qdst = s.qbuf + (size_t)qslot * kblocks * kQBytes;
          for (int kb = 0; kb < kblocks; ++kb)
            tma_load_2d(qdst + kb * kQBytes, &map_q, kb * kTcBK, rec.pair_base, &s.qb_full[qslot]);
        }
      }
      __syncwarp();
      ++icount;
      cur = nx1;
#pragma unroll
      for (int c = 0; c < kCh; ++c) tau_cur[c] = tau_nx1[c];
      nx1 = nx2;
      nx2.rec = rec3;
    }
    ROLE_REPORT(6);  // wait q_empty + qb_empty (threshold refresh rounds included), -, -, -, total
  } else if (warp == kStreamWarp) {
    // =================================== tile streamer ====================================
    if (lane == 0) {
      ROLE_T0();
      unsigned sl = 0, sl_par = 0;  // slot ring position and the parity of its current round
      unsigned icount = 0, tcount = 0;
      const long long n_rows4 = a.n_rows & ~3ll;
      while (true) {
        const int islot = (int)(icount % kSB);
        ROLE_WAIT(0, &s.itm_full[islot], (icount / kSB) & 1u);
        const TcItem rec = s.itm[islot];
        if (rec.nq == 0) break;
        const int n_tiles = (rec.row1 - rec.row0 + kTile - 1) / kTile;
#pragma unroll 1
        for (int t = 0; t < n_tiles; ++t, ++tcount) {
          const int trow0 = rec.row0 + t * kTile;
          {
            // The tile's row norms: a plain bulk copy needs a 16-byte aligned source, so it starts
            // at the 4-row boundary below the tile and stops at the last whole group of 4 rows of
            // the array (the filter reads the <= 3 rows after that directly).
            const unsigned mb = tcount % kMetaBufs;
            ROLE_WAIT(1, &s.meta_empty[mb], ((tcount / kMetaBufs) & 1u) ^ 1u);
            const long long m0 = trow0 & ~3ll;
            long long avail = n_rows4 - m0;
            if (avail > kMetaFloats) avail = kMetaFloats;
            if (avail > 0) {
              mbar_arrive_expect_tx(&s.meta_full[mb], (unsigned)avail * 4u);
              bulk_g2s(s.meta + (size_t)mb * kMetaBytes, a.xnorm + m0, (unsigned)avail * 4u, &s.meta_full[mb]);
            } else {
              mbar_arrive(&s.meta_full[mb]);
            }
          }
          // the ragged last tile of a chunk is fetched in 32-row boxes: the rows behind it belong to the
          // next bucket and would be read for nothing (26 % of the traffic with 305-row buckets)
          const int rows_left = rec.row1 - trow0;
          const int sub = rows_left >= kTile ? 0 : (rows_left + 31) >> 5;
          const unsigned qbytes = WIDE ? kQBytes : 0u;
          const unsigned tx = (sub == 0 ? kSlotBytes : (unsigned)sub * kSubBoxBytes) + qbytes;
          for (int kb = 0; kb < kblocks; ++kb) {
            ROLE_WAIT(2, &s.empty_bar[sl], sl_par ^ 1u);
            unsigned char* dst = s.slots + (size_t)sl * kSlotStride;
            // a box is always written in full (rows / columns past the tensor are zero filled)
            mbar_arrive_expect_tx(&s.full_bar[sl], tx);
            if (sub == 0) {
              tma_load_2d(dst, &map_x, kb * kTcBK, trow0, &s.full_bar[sl]);
            } else {
              for (int b = 0; b < sub; ++b)
                tma_load_2d(dst + b * kSubBoxBytes, &map_x32, kb * kTcBK, trow0 + 32 * b, &s.full_bar[sl]);
            }
            if (WIDE) tma_load_2d(dst + kSlotBytes, &map_q, kb * kTcBK, rec.pair_base, &s.full_bar[sl]);
            if (++sl == n_slots) {
              sl = 0;
              sl_par ^= 1u;
            }
          }
        }
        ++icount;
      }
      ROLE_REPORT(21);  // wait itm_full, meta_empty, slot empty, -, total
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================== MMA issuer =======================================
    if (lane == 0) {
      ROLE_T0();
      const uint32_t idesc = make_tf32_idesc(NQ);
      unsigned sl = 0, sl_par = 0;  // slot ring position and the parity of its current round
      unsigned icount = 0, tcount = 0;
      while (true) {
        const int islot = (int)(icount % kSB);
        const int qslot = (int)(icount % kItemBufs);
        ROLE_WAIT(0, &s.itm_full[islot], (icount / kSB) & 1u);
        const int nq = s.itm[islot].nq;
        if (nq == 0) break;
        const int n_tiles = (s.itm[islot].row1 - s.itm[islot].row0 + kTile - 1) / kTile;
        if (!WIDE) ROLE_WAIT(0, &s.qb_full[qslot], (icount / kItemBufs) & 1u);
        const unsigned char* qsrc = s.qbuf + (size_t)qslot * kblocks * kQBytes;
        for (int t = 0; t < n_tiles; ++t, ++tcount) {
          const unsigned set = tcount & sets_mask;
          ROLE_WAIT(1, &s.acc_empty[set], ((tcount >> sets_log2) & 1u) ^ 1u);  // the filter drained this set
          tc_fence_after();
          const uint32_t acc = tmem_base + set * (uint32_t)NQ;
          for (int kb = 0; kb < kblocks; ++kb) {
            ROLE_WAIT(2, &s.full_bar[sl], sl_par);
            tc_fence_after();
            const unsigned char* slot = s.slots + (size_t)sl * kSlotStride;
            const uint64_t da = make_kmajor_sw128_desc(slot);
            const uint64_t db = make_kmajor_sw128_desc(WIDE ? slot + kSlotBytes : qsrc + kb * kQBytes);
#pragma unroll
            for (int k8 = 0; k8 < kTcBK / 8; ++k8)  // UMMA K = 8 tf32 = 32 bytes: +2 in (addr >> 4)
              tc_mma_tf32(acc, da + (uint64_t)(k8 * 2), db + (uint64_t)(k8 * 2), idesc,
                          (kb > 0 || k8 > 0) ? 1u : 0u);
            tc_commit(&s.empty_bar[sl]);  // the slot may be refilled once these MMAs have read it
            if (++sl == n_slots) {
              sl = 0;
              sl_par ^= 1u;
            }
          }
          tc_commit(&s.acc_full[set]);
        }
        if (!WIDE) tc_commit(&s.qb_empty[qslot]);  // the item's query buffer may be refilled
        ++icount;
      }
      ROLE_REPORT(11);  // wait itm_full + qb_full, acc_empty, slot full, -, total
    }
    __syncwarp();
  } else {
    // =================================== filter + score ===================================
    const int fw = warp - 2;            // 0 .. kFilterWarps - 1
    const int group = fw >> 2;          // tiles with tcount % kGroups == group are this warp's
    const int quarter = warp & 3;       // TMEM lanes this warp may read
    const int r_local = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int* q_row = s.wq_row + fw * kQueueCap;
    int* q_meta = s.wq_meta + fw * kQueueCap;
    int* q_qi = s.wq_qi + fw * kQueueCap;
    ScoreCtx sc;
    sc.xs = a.xs;
    sc.ids = a.ids;
    sc.qs = a.qs;
    sc.pq = a.pq;
    sc.tau_g = a.tau_g;
    sc.tau0 = a.tau0;
    sc.ladder = a.ladder;
    sc.cand_n = a.cand_n;
    sc.cand = a.cand;
    sc.qbuf = s.qbuf;
    sc.qi_s = s.qi_s;
    sc.cap = a.cap;
    sc.k = a.k;
    sc.d = a.d;
    sc.d_pad = a.d_pad;
    sc.kblocks = kblocks;
    unsigned head = 0;  // queue front (index mod kQueueCap)
    int count = 0;      // queued survivors (warp-uniform)
    unsigned n_surv = 0, n_batches = 0, n_flush = 0, n_appended = 0;
    unsigned icount = 0, tcount = 0;
    // !QGLOBAL: queue entries refer to the item slot whose queries they need, so a slot is released only when
    // the queue's front has passed the item's last entry - after the next full batch as a rule, and by
    // scoring a partial batch only when the warp would otherwise wait (nothing is published to filter)
    unsigned enq_total = 0, done_total = 0;  // entries queued / scored so far
    unsigned wm[kSB + 1];                    // enq_total at the end of each filtered, unreleased item (!QGLOBAL)
    int npend = 0;
    unsigned rel = 0;                        // item count of the oldest unreleased item
    auto release_ready = [&]() {
      while (npend > 0 && (int)(done_total - wm[0]) >= 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.q_empty[rel % kSB]);
        ++rel;
        --npend;
#pragma unroll
        for (int i = 0; i < kSB; ++i) wm[i] = wm[i + 1];
      }
    };
    ROLE_T0();
    while (true) {
      const int islot = (int)(icount % kSB);
      if (!QGLOBAL && npend > 0 && !mbar_try_wait_short(&s.itm_full[islot], (icount / kSB) & 1u)) {
        if (count > 0) {
          score_batch<METRIC, QGLOBAL, V8>(sc, q_row, q_meta, q_qi, head, count, lane, n_appended);
          head = (head + (unsigned)count) & (kQueueCap - 1);
          done_total += (unsigned)count;
          count = 0;
          ++n_flush;
        }
        release_ready();
      }
      ROLE_WAIT(0, &s.itm_full[islot], (icount / kSB) & 1u);
      const TcItem rec = s.itm[islot];
      if (rec.nq == 0) break;
      const float* th = s.thr_s + islot * NQ;
      const int n_tiles = (rec.row1 - rec.row0 + kTile - 1) / kTile;
      const int n_ch = (rec.nq + 31) >> 5;  // column chunks that hold queries
      for (int t = 0; t < n_tiles; ++t, ++tcount) {
        if ((int)(tcount % kGroups) != group) continue;
        const int row = rec.row0 + t * kTile + r_local;
        const bool valid = row < rec.row1;
        const unsigned mb = tcount % kMetaBufs;
        ROLE_WAIT(1, &s.meta_full[mb], (tcount / kMetaBufs) & 1u);
        float xn = reinterpret_cast<const float*>(s.meta + (size_t)mb * kMetaBytes)[((rec.row0 + t * kTile) & 3) + r_local];
        if (row >= (a.n_rows & ~3ll)) xn = valid ? a.xnorm[row] : 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.meta_empty[mb]);
        float ra, rb;  // bound = ra * dot + rb
        if (METRIC == NLSH_METRIC_L2) {
          ra = -2.0f;
          rb = (1.0f - kFilterC) * xn;
        } else {
          ra = -1.0f / fmaxf(sqrtf(xn), 1e-8f);
          rb = 0.f;
        }
        const unsigned set = tcount & sets_mask;
        ROLE_WAIT(2, &s.acc_full[set], (tcount >> sets_log2) & 1u);
        tc_fence_after();
        unsigned masks[kCh];
#pragma unroll
        for (int c = 0; c < kCh; ++c) {
          masks[c] = 0;
          if (c < n_ch) {  // warp-uniform
            uint32_t v0[16], v1[16];
            tc_ld16_nowait(lane_base + set * (uint32_t)NQ + 32u * c, v0);
            tc_ld16_nowait(lane_base + set * (uint32_t)NQ + 32u * c + 16u, v1);
            tc_wait_ld2(v0, v1);
            unsigned mask = 0;
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const float4 t0 = *reinterpret_cast<const float4*>(th + 32 * c + j4);
              const float4 t1 = *reinterpret_cast<const float4*>(th + 32 * c + 16 + j4);
              mask |= (fmaf(ra, __uint_as_float(v0[j4 + 0]), rb) <= t0.x ? 1u : 0u) << (j4 + 0);
              mask |= (fmaf(ra, __uint_as_float(v0[j4 + 1]), rb) <= t0.y ? 1u : 0u) << (j4 + 1);
              mask |= (fmaf(ra, __uint_as_float(v0[j4 + 2]), rb) <= t0.z ? 1u : 0u) << (j4 + 2);
              mask |= (fmaf(ra, __uint_as_float(v0[j4 + 3]), rb) <= t0.w ? 1u : 0u) << (j4 + 3);
              mask |= (fmaf(ra, __uint_as_float(v1[j4 + 0]), rb) <= t1.x ? 1u : 0u) << (16 + j4 + 0);
              mask |= (fmaf(ra, __uint_as_float(v1[j4 + 1]), rb) <= t1.y ? 1u : 0u) << (16 + j4 + 1);
              mask |= (fmaf(ra, __uint_as_float(v1[j4 + 2]), rb) <= t1.z ? 1u : 0u) << (16 + j4 + 2);
              mask |= (fmaf(ra, __uint_as_float(v1[j4 + 3]), rb) <= t1.w ? 1u : 0u) << (16 + j4 + 3);
            }
            masks[c] = valid ? mask : 0u;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.acc_empty[set]);  // TMEM set free for the tile kSets ahead
        // queue this warp's survivors, one per lane and round; a full batch is scored right away
#pragma unroll
        for (int c = 0; c < kCh; ++c) {
          unsigned mask = masks[c];
          while (true) {
            const bool has = mask != 0;
            const unsigned b = __ballot_sync(NLSH_FULL_MASK, has);
            if (b == 0) break;
            if (has) {
              const int j = 32 * c + __ffs(mask) - 1;
              mask &= mask - 1;
              const unsigned e = (head + (unsigned)count + __popc(b & ((1u << lane) - 1u))) & (kQueueCap - 1);
              q_row[e] = row;
              q_meta[e] = QGLOBAL ? rec.pair_base + j : ((islot << 8) | j);
              q_qi[e] = s.qi_s[islot * NQ + j];
            }
            const int added = __popc(b);
            count += added;
            enq_total += (unsigned)added;
            n_surv += (unsigned)added;
            __syncwarp();
            if (count >= 32) {
              ROLE_TIC();
              score_batch<METRIC, QGLOBAL, V8>(sc, q_row, q_meta, q_qi, head, 32, lane, n_appended);
              ROLE_TOC(3);
              head = (head + 32u) & (kQueueCap - 1);
              count -= 32;
              done_total += 32u;
              ++n_batches;
              if (!QGLOBAL) release_ready();
            }
          }
        }
      }
      if (QGLOBAL) {  // entries do not refer to the item slot and stay queued across items
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.q_empty[islot]);
      } else {
#pragma unroll
        for (int i = 0; i <= kSB; ++i)
          if (i == npend) wm[i] = enq_total;
        ++npend;
        release_ready();
      }
      ++icount;
    }
    if (count > 0) {
      score_batch<METRIC, QGLOBAL, V8>(sc, q_row, q_meta, q_qi, head, count, lane, n_appended);
      ++n_flush;
    }
    ROLE_REPORT(16);  // (sum over the 8 filter warps) wait itm_full, meta_full, acc_full, full-batch scoring, total
    if (a.stats != nullptr) {
      // debug counters: [0] survivors of the filter, [1] full batches, [2] end-of-item batches,
      // [3] candidates appended (n_appended is per lane), [4] tiles this warp filtered
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n_appended += __shfl_xor_sync(NLSH_FULL_MASK, n_appended, o);
      if (lane == 0) {
        atomicAdd(a.stats + 0, (unsigned long long)n_surv);
        atomicAdd(a.stats + 1, (unsigned long long)n_batches);
        atomicAdd(a.stats + 2, (unsigned long long)n_flush);
        atomicAdd(a.stats + 3, (unsigned long long)n_appended);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tc_dealloc(tmem_base, kTmemCols);
}

// qs[i] = qn[pairs[i] / p] (pair order = grouped by bucket), pq[i] = pairs[i] / p, pqn2[i] = |qs[i]|^2;
// one warp per pair.
__global__ void __launch_bounds__(256)
    gather_pair_queries_kernel(const float* __restrict__ qn, const int* __restrict__ pairs,
                               const int* __restrict__ n_valid, long long n_pairs, int p, int d_pad,
                               float* __restrict__ qs, int* __restrict__ pq, float* __restrict__ pqn2) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
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
    if (lane == 0) {
      pq[i] = (int)q;
      pqn2[i] = ss;
    }
  }
}

// Seed of tau_g: the exact k-th best distance of each query among the first `seed_rows` rows of its
// probed buckets, taken in probe order (the first bucket alone when it is large enough), one warp per
// query.  Any k candidates bound the final k-th distance from above, so this is a valid threshold from
// the very first tile; without it every row would survive the filter.  The query sits in shared memory
// (one float4 chunk per lane) and a ring of three 4-row steps is in flight per warp, 8 lanes per row.
// The bound is inflated by the worst-case difference between two fp32 summation orders of the same d
// terms (2 d 2^-24 relative for the positive terms of L2, 4 d 2^-24 absolute for the angular distance of
// a unit-norm query), because the scan sums them in another order.
constexpr int kMaxSeedRows = 4096;

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
__global__ void __launch_bounds__(128, 5)
    seed_tau_kernel(const float* __restrict__ qn, const int* __restrict__ probes,
                    const int* __restrict__ offsets, const float* __restrict__ xs, int n_buckets, int p,
                    int d, int d_pad, int k, int seed_rows, int seed_div, long long n_queries,
                    float* __restrict__ tau_g, float* __restrict__ tau0) {
  __shared__ float4 q_s[4][32];  // per warp: 16-byte chunk c of its query at [c]
  const int wb = threadIdx.x >> 5;
  const long long q = (long long)blockIdx.x * 4 + wb;
  if (q >= n_queries) return;
  const int lane = lane_id();
  const int l8 = lane & 7, g = lane >> 3;
  q_s[wb][lane] = lane < (d_pad >> 2) ? *reinterpret_cast<const float4*>(qn + q * d_pad + 4 * lane)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  const float4* qs = q_s[wb];
  WarpTopK<1, int> top;
  top.init(NLSH_ID_SENTINEL);
  int budget = -1;  // sample rows still to take; sized by the first non-empty bucket
  int taken = 0;
  for (int j = 0; j < p && budget != 0; ++j) {  // warp-uniform
    const int b = probes[q * p + j];
    if (b < 0 || b >= n_buckets) continue;
    bool dup = false;
    for (int e = 0; e < j; ++e) dup |= probes[q * p + e] == b;  // the reference probes a set
    if (dup) continue;
    const int r0 = offsets[b];
    const int size = offsets[b + 1] - r0;
    if (size <= 0) continue;
    if (budget < 0) {
      // a large bucket gets a larger sample: its rows below the bound all become candidates
      // (about k * size / sample of them), and those must fit the candidate buffer
      budget = seed_rows;
      const int want = size / seed_div;
      if (want > budget) budget = want;
      if (budget > kMaxSeedRows) budget = kMaxSeedRows;
    }
    const int n = size < budget ? size : budget;
    const float* xb = xs + (size_t)r0 * d_pad;
    float4 x0[4], x1[4], x2[4];
    seed_load_rows(x0, xb, g, n, d_pad, l8);
    seed_load_rows(x1, xb, 4 + g, n, d_pad, l8);
    seed_load_rows(x2, xb, 8 + g, n, d_pad, l8);
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
    budget -= n;
    taken += n;
  }
  if (lane == 0) {
    float t = pos_inf();
    if (taken >= k && top.tau < pos_inf()) {
      const float u = (float)d * 5.9604645e-8f;  // d * 2^-24
      t = METRIC == NLSH_METRIC_L2 ? top.tau * (1.0f + 2.5f * u) + 1e-30f : top.tau + 4.5f * u;
    }
    tau_g[q] = t;
    tau0[q] = t;
  }
}

// The same seed for wide rows (d_pad > 128) or k > 32: a block of 8 warps per query, one sample row per lane
// and step, every distance by one thread in the scorer's own arithmetic (tc_thread_distance), lists of up to
// 128 entries per warp merged through shared memory.  The rows of a step are 32 different rows, so a warp is
// latency bound: the 8 warps of a query take every eighth step.
constexpr int kSeedWarps = 8;

template <int METRIC, int KPL>
__global__ void __launch_bounds__(32 * kSeedWarps)
    seed_tau_generic_kernel(const float* __restrict__ qn, const int* __restrict__ probes,
                            const int* __restrict__ offsets, const float* __restrict__ xs, int n_buckets, int p,
                            int d, int d_pad, int k, int seed_rows, int seed_div, long long n_queries,
                            float* __restrict__ tau_g, float* __restrict__ tau0) {
  __shared__ float sh_d[kSeedWarps][32 * KPL];
  const long long q = blockIdx.x;
  if (q >= n_queries) return;
  const int warp = threadIdx.x >> 5;
  const int lane = lane_id();
  TcQueryGlobal qg;
  qg.q = qn + (size_t)q * d_pad;
  WarpTopK<KPL, int> top;
  top.init(NLSH_ID_SENTINEL);
  int budget = -1, taken = 0;
  for (int j = 0; j < p && budget != 0; ++j) {  // block-uniform
    const int b = probes[q * p + j];
    if (b < 0 || b >= n_buckets) continue;
    bool dup = false;
    for (int e = 0; e < j; ++e) dup |= probes[q * p + e] == b;
    if (dup) continue;
    const int r0 = offsets[b];
    const int size = offsets[b + 1] - r0;
    if (size <= 0) continue;
    if (budget < 0) {
      budget = seed_rows;
      const int want = size / seed_div;
      if (want > budget) budget = want;
      if (budget > kMaxSeedRows) budget = kMaxSeedRows;
    }
    const int n = size < budget ? size : budget;
    for (int base = 32 * warp; base < n; base += 32 * kSeedWarps) {
      const int r = base + lane;
      float dist = 0.f;
      if (r < n) dist = tc_thread_distance<METRIC, 8, false>(xs + (size_t)(r0 + r) * d_pad, qg, d);
      top.offer(dist, taken + r, r < n, k);
    }
    budget -= n;
    taken += n;
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) sh_d[warp][j * 32 + lane] = top.d[j];
  __syncthreads();
  if (warp != 0) return;
  for (int w = 1; w < kSeedWarps; ++w) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const float v = sh_d[w][j * 32 + lane];
      top.offer(v, (w << 16) | (j * 32 + lane), v < pos_inf(), k);  // only the k-th best VALUE matters here
    }
  }
  if (lane == 0) {
    float t = pos_inf();
    // same arithmetic as the scorer: a sample row's distance is bit-identical there, no inflation needed
    if (taken >= k && top.tau < pos_inf()) t = top.tau;
    tau_g[q] = t;
    tau0[q] = t;
  }
}

size_t scan_tc_smem(int kblocks, int n_slots, bool wide, int nq) {
  return (size_t)n_slots * (kSlotBytes + (wide ? nq * 128u : 0u)) + smem_fixed_bytes(kblocks, wide, nq) + 1024;
}

}  // namespace

bool nlsh_scan_tc_supported(int d, int k, int metric) {
  return d >= 1 && (d + 3) / 4 * 4 <= kMaxWideKBlocks * kTcBK && k >= 1 && k <= NLSH_MAX_K &&
         (metric == NLSH_METRIC_L2 || metric == NLSH_METRIC_ANGULAR);
}

namespace {
__global__ void copy_tau_kernel(const float* __restrict__ src, long long n, float* __restrict__ tau_g,
                                float* __restrict__ tau0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float t = src[i];
    tau_g[i] = t;
    tau0[i] = t;
  }
}
}  // namespace

int nlsh_scan_tc_prepare(const float* qn, const int* pairs, const int* n_valid, long long n_pairs,
                         int p, int d_pad, float* qs, int* pq, float* pqn2, float* tau_g, float* tau0,
                         const float* tau_seed, long long n_queries, const int* probes, const int* offsets,
                         const float* xs, long long n_rows, int n_buckets, int d, int k, int metric,
                         cudaStream_t st) {
  long long blocks = (n_pairs * 32 + 255) / 256;
  const long long cap = (long long)nlsh_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  gather_pair_queries_kernel<<<(unsigned)blocks, 256, 0, st>>>(qn, pairs, n_valid, n_pairs, p, d_pad, qs, pq,
                                                             pqn2);
  NLSH_CUDA_TRY(nlsh_post_launch());
  if (tau_seed != nullptr) {  // bounds computed elsewhere (another rank's slice of the seeding)
    copy_tau_kernel<<<(unsigned)((n_queries + 255) / 256), 256, 0, st>>>(tau_seed, n_queries, tau_g, tau0);
    return nlsh_check_cuda(nlsh_post_launch(), "copy_tau_kernel launch");
  }
  return nlsh_scan_tc_seed(qn, n_queries, probes, p, offsets, xs, n_rows, n_buckets, d, d_pad, k, metric, tau_g,
                           tau0, st);
}

int nlsh_scan_tc_seed(const float* qn, long long n_queries, const int* probes, int p, const int* offsets,
                      const float* xs, long long n_rows, int n_buckets, int d, int d_pad, int k, int metric,
                      float* tau_g, float* tau0, cudaStream_t st, int sample_rows) {
  // Sample rows per query: half the average bucket, 32 .. 128, and at least 1/seed_div of the query's own
  // first bucket (2441-row buckets: 152 rows), so that the rows of a large bucket that fall below the bound
  // stay a few dozen.  The scan's time hardly depends on the survivors (config 4: 1.08 M survivors 0.96 ms,
  // 2.47 M 1.03 ms) while the sample costs 0.55 us per row and batch: 152 rows 1.25 ms per call, 384 rows
  // 1.32 ms; on an 8-GPU shard 128 .. 192 rows are equal, 64 rows lose (profiles/r2_experiments).
  // NLSH_SCAN_SEED=<rows> overrides the base sample; the scan needs the seed (every tau_g is written here).
  const long long avg = n_buckets > 0 ? n_rows / n_buckets : 0;
  int seed_rows = (int)(avg / 2 / 32 * 32);
  if (seed_rows < 32) seed_rows = 32;
  if (seed_rows > 128) seed_rows = 128;
  // the caller's choice (a rank that seeds only its 1/N slice of the queries can afford a larger sample: on an
  // 8-GPU shard 192 rows against 128 take the scan from 0.176 to 0.154 ms, profiles/r2_experiments)
  if (sample_rows > 0) seed_rows = sample_rows;
  if (const char* env = getenv("NLSH_SCAN_SEED")) seed_rows = atoi(env);
  if (seed_rows < 0) seed_rows = 0;  // 0: no sample, tau_g = +inf (every row is scored; A/B and tests only)
  if (seed_rows > kMaxSeedRows) seed_rows = kMaxSeedRows;
  int seed_div = 16;
  if (const char* env = getenv("NLSH_SCAN_SEED_DIV")) seed_div = atoi(env);
  if (seed_div < 1) seed_div = 1;
  if (seed_rows == 0) seed_div = 1 << 30;
  // a sample of at least a few times k rows; 8 k for long lists (config 5, k = 100, wide rows: every survivor
  // costs the scorer 7.7 KB of reads - 400 sample rows 2.8 M survivors and a 6.4 ms scan, 800 rows 1.8 M and 4.7 ms)
  if (seed_rows > 0 && seed_rows < (k > 32 ? 8 : 4) * k) seed_rows = (k > 32 ? 8 : 4) * k;
  const unsigned sb = (unsigned)((n_queries + 3) / 4);
#define NLSH_SEED_ARGS qn, probes, offsets, xs, n_buckets, p, d, d_pad, k, seed_rows, seed_div, n_queries, tau_g, tau0
  if (d_pad <= kMaxKBlocks * kTcBK && k <= 32) {
    if (metric == NLSH_METRIC_L2)
      seed_tau_kernel<NLSH_METRIC_L2><<<sb, 128, 0, st>>>(NLSH_SEED_ARGS);
    else
      seed_tau_kernel<NLSH_METRIC_ANGULAR><<<sb, 128, 0, st>>>(NLSH_SEED_ARGS);
  } else if (metric == NLSH_METRIC_L2) {
    const unsigned gb = (unsigned)n_queries;
    if (k <= 32) seed_tau_generic_kernel<NLSH_METRIC_L2, 1><<<gb, 32 * kSeedWarps, 0, st>>>(NLSH_SEED_ARGS);
    else if (k <= 64) seed_tau_generic_kernel<NLSH_METRIC_L2, 2><<<gb, 32 * kSeedWarps, 0, st>>>(NLSH_SEED_ARGS);
    else seed_tau_generic_kernel<NLSH_METRIC_L2, 4><<<gb, 32 * kSeedWarps, 0, st>>>(NLSH_SEED_ARGS);
  } else {
    const unsigned gb = (unsigned)n_queries;
    if (k <= 32) seed_tau_generic_kernel<NLSH_METRIC_ANGULAR, 1><<<gb, 32 * kSeedWarps, 0, st>>>(NLSH_SEED_ARGS);
    else if (k <= 64) seed_tau_generic_kernel<NLSH_METRIC_ANGULAR, 2><<<gb, 32 * kSeedWarps, 0, st>>>(NLSH_SEED_ARGS);
    else seed_tau_generic_kernel<NLSH_METRIC_ANGULAR, 4><<<gb, 32 * kSeedWarps, 0, st>>>(NLSH_SEED_ARGS);
  }
#undef NLSH_SEED_ARGS
  return nlsh_check_cuda(nlsh_post_launch(), "seed_tau_kernel launch");
}

namespace {
template <int METRIC, int NQ, bool WIDE, bool QGLOBAL, bool V8>
int launch_variant(const TcScanArgs& a, const CUtensorMap& map_x, const CUtensorMap& map_x32,
                   const CUtensorMap& map_q, int grid, size_t smem, cudaStream_t st) {
  auto kern = scan_tc_kernel<METRIC, NQ, WIDE, QGLOBAL, V8>;
  NLSH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kThreads, smem, st>>>(a, map_x, map_x32, map_q);
  return nlsh_check_cuda(nlsh_post_launch(), "scan_tc_kernel launch");
}

template <int METRIC>
int launch_metric(const TcScanArgs& a, const CUtensorMap& map_x, const CUtensorMap& map_x32,
                  const CUtensorMap& map_q, bool wide, bool qglobal, bool v8, int grid, size_t smem, cudaStream_t st) {
#define NLSH_TC_LAUNCH(NQ, WIDE, QGLOBAL)                                                                  \
  return v8 ? launch_variant<METRIC, NQ, WIDE, QGLOBAL, true>(a, map_x, map_x32, map_q, grid, smem, st)    \
            : launch_variant<METRIC, NQ, WIDE, QGLOBAL, false>(a, map_x, map_x32, map_q, grid, smem, st)
  if (a.nq_group == kTcNQMax) { NLSH_TC_LAUNCH(kTcNQMax, true, true); }
  if (wide) { NLSH_TC_LAUNCH(kTcNQ, true, true); }
  if (qglobal) { NLSH_TC_LAUNCH(kTcNQ, false, true); }
  NLSH_TC_LAUNCH(kTcNQ, false, false);
#undef NLSH_TC_LAUNCH
}
}  // namespace

int nlsh_scan_tc_launch(int metric, TcScanArgs a, cudaStream_t st) {
  a.kblocks = (a.d_pad + kTcBK - 1) / kTcBK;
  a.l2_slack = 2.1e-6f * sqrtf((float)a.d);
  if (a.nq_group != kTcNQMax) a.nq_group = kTcNQ;
  // wide rows, or 128 queries per item: the queries' K blocks travel with the row tiles
  const bool wide = a.kblocks > kMaxKBlocks || a.nq_group == kTcNQMax;
  // Scorer query source.  The item's shared-memory copy halves the scorer's L2 requests (a survivor then costs
  // its 512-byte row alone) but ties queue entries to item slots: fine when an item yields a batch or more per
  // warp (buckets of >= 16 tiles; config 3's 9-tile buckets measured 0.22 ms global against 0.26 shared, config 4's
  // 19-tile buckets 0.89 against 0.87), while on 8-GPU shards (2 - 3 tiles per bucket) warps would sit on several
  // unreleased items, so short items take the pair-ordered global copy (QGLOBAL) - as wide rows must.
  // NLSH_TC_QGLOBAL=0/1 overrides (A/B runs).
  bool qglobal = wide || a.avg_item_rows < 16 * kTile;
  if (const char* env = getenv("NLSH_TC_QGLOBAL")) qglobal = wide || atoi(env) != 0;
  int n_slots = kMaxSlots;
  // each K block's slot is freed by its own tcgen05.commit, so any ring depth >= 2 makes progress.  With the
  // queries read from global memory the scorer lives on L1 hits, and shared memory above 196 KB leaves L1 only
  // 28 KB: those variants stop at the 196 KB carve-out (measured: config 3 0.33 -> 0.23 ms, shard 0.213 -> 0.184).
  size_t smem_limit = (qglobal ? 195 : 224) * 1024;  // a block's share of a carve-out is 1 KB less
  if (const char* env = getenv("NLSH_TC_SMEM_KB")) {  // A/B runs: up to the 227 KB a block may opt in to
    const int v = atoi(env);
    if (v >= 64 && v <= 227) smem_limit = (size_t)v * 1024;
  }
  while (n_slots > 3 && scan_tc_smem(a.kblocks, n_slots, wide, a.nq_group) > smem_limit) --n_slots;
  if (const char* env = getenv("NLSH_TC_SLOTS")) {  // A/B runs
    const int v = atoi(env);
    if (v >= 3 && v <= kMaxSlots && scan_tc_smem(a.kblocks, v, wide, a.nq_group) <= 227 * 1024) n_slots = v;
  }
  a.n_slots = n_slots;
  // Tiles between the TMA and the filter: slot ring + accumulator ring.  The scorer re-reads surviving rows from
  // L2, and 148 SMs x 16 tiles x 64 KB in flight (150 MB) is more than the L2 holds, so the accumulator ring
  // is used 8 deep by default (NLSH_TC_SETS=2/4/8/16: A/B runs).
  a.sets_log2 = 3;
  if (const char* env = getenv("NLSH_TC_SETS")) {
    const int v = atoi(env);
    a.sets_log2 = v >= 16 ? 4 : (v >= 8 ? 3 : (v >= 4 ? 2 : 1));
  }
  const size_t smem = scan_tc_smem(a.kblocks, n_slots, wide, a.nq_group);
  CUtensorMap map_x, map_x32, map_q;
  int rc;
  if ((rc = tc_make_map(&map_x, a.xs, a.n_rows, a.d_pad, kTile)) != NLSH_OK) return rc;
  if ((rc = tc_make_map(&map_x32, a.xs, a.n_rows, a.d_pad, 32)) != NLSH_OK) return rc;
  if ((rc = tc_make_map(&map_q, a.qs, a.n_pairs, a.d_pad, a.nq_group)) != NLSH_OK) return rc;
  // One persistent CTA per SM; `sm_reserve` SMs are left to other streams (the launch-bound front part of
  // the next batch in nlsh.parallel.PipelinedSearch).  NLSH_TC_GRID=<CTAs> overrides (A/B runs).
  int grid = nlsh_num_sms() - a.sm_reserve;
  if (const char* env = getenv("NLSH_TC_GRID")) grid = atoi(env);
  if (grid < 1) grid = 1;
  if (grid > nlsh_num_sms()) grid = nlsh_num_sms();
  // 32-byte row loads in the scorer when the query comes from shared memory and every row of xs starts on a
  // 32-byte boundary (with both operands from global memory they measured slower: shard 0.194 against 0.176 ms,
  // config 5 3.83 against 3.72); NLSH_TC_V8=0/1 overrides (A/B runs)
  bool v8 = !qglobal && a.d_pad % 8 == 0 && (reinterpret_cast<uintptr_t>(a.xs) & 31) == 0 &&
            (reinterpret_cast<uintptr_t>(a.qs) & 31) == 0;
  if (const char* env = getenv("NLSH_TC_V8"))
    v8 = atoi(env) != 0 && a.d_pad % 8 == 0 && (reinterpret_cast<uintptr_t>(a.xs) & 31) == 0 &&
         (reinterpret_cast<uintptr_t>(a.qs) & 31) == 0;
  if (metric == NLSH_METRIC_L2)
    return launch_metric<NLSH_METRIC_L2>(a, map_x, map_x32, map_q, wide, qglobal, v8, grid, smem, st);
  return launch_metric<NLSH_METRIC_ANGULAR>(a, map_x, map_x32, map_q, wide, qglobal, v8, grid, smem, st);
}
