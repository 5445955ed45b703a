/*
 * nlsh_b200.h — C ABI of libnlsh_b200.so, the B200 (sm_100a) implementation of the
 * Neural-LSH index-build + query hot path.
 *
 * The reference (stegben/neural-locality-sensitive-hashing) has no C ABI for this path:
 * it is Python + one Cython helper.  Each entry point below names the reference
 * interface (file:line under /root/reference) whose arithmetic it replaces.  The Python
 * package `nlsh` in this repo binds these with ctypes and re-exposes the reference's own
 * API (nlsh.utils.hash_codes, nlsh.hashings.*.hash, nlsh.indexer.build_index / Indexer,
 * precompute.self_get_knn_pt); INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every pointer marked "device" is a CUDA device pointer owned by the caller;
 *    the library never allocates persistent device memory: scratch is passed in and
 *    sized with the matching *_workspace_bytes() call (same arguments);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    all work is enqueued on it, no entry point synchronises the device;
 *  - return value: 0 on success, negative NLSH_ERR_* otherwise, text via
 *    nlsh_last_error() (thread-local);
 *  - all matrices are row-major, fp32 unless stated.
 */
#ifndef NLSH_B200_H
#define NLSH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLSH_B200_VERSION 310 /* 0.3.1: + nlsh_sample_probes, nlsh_query_seed_tau_rows; scan pipeline with a tile streamer */

#define NLSH_OK 0
#define NLSH_ERR_INVALID (-1)   /* bad argument (-> ValueError in the Python layer) */
#define NLSH_ERR_CUDA (-2)      /* CUDA runtime error (-> RuntimeError) */
#define NLSH_ERR_WORKSPACE (-3) /* workspace too small (-> RuntimeError) */

/* hidden-layer activations (encoders.py:24-55 uses ReLU; the SIREN trunk of
 * encoders.py:58-79 uses sin(w0 * .), exposed but validated only against a restatement) */
#define NLSH_ACT_IDENTITY 0
#define NLSH_ACT_RELU 1
#define NLSH_ACT_SIN 2

/* output heads: hashings.py:24-26 (sigmoid / tanh) and hashings.py:104-106 (softmax) */
#define NLSH_HEAD_SIGMOID 0 /* bit_i = sigmoid(l_i) > 0.5          (fp32 semantics, see below) */
#define NLSH_HEAD_TANH 1    /* bit_i = tanh(l_i)/2 + 0.5 > 0.5                                */
#define NLSH_HEAD_SOFTMAX 2 /* code  = argmax_i l_i  (Categorical.hash, hashings.py:131-133)  */

/* candidate-scan / kNN metrics */
#define NLSH_METRIC_L2 0      /* nlsh/data.py:192-201: ||q - x + 1e-6||_2 (F.pairwise_distance) */
#define NLSH_METRIC_ANGULAR 1 /* nlsh/data.py:100-109: 1 - cos(q, x), norms clamped at 1e-8     */
#define NLSH_METRIC_L2SQ 2    /* precompute.py:37-54 (_l2): squared L2, no eps, no sqrt         */
#define NLSH_METRIC_COSINE 3  /* precompute.py:22-34 (_cosine_distance): 1 - <q/|q|, x/|x|>     */

#define NLSH_MAX_K 128        /* top-k capacity of the warp-resident lists */
#define NLSH_MAX_HASH_BITS 15 /* utils.pyx:6-15 returns int16: codes are < 2^15 */
#define NLSH_MAX_LAYERS 16

/* One Linear(+activation) layer: y = act(x W^T + b), W is [out_dim, in_dim] row-major
 * exactly as torch.nn.Linear stores it (encoders.py:41-48, hashings.py:19). */
typedef struct nlsh_layer {
  const float* weight; /* device, [out_dim, in_dim] */
  const float* bias;   /* device, [out_dim] or NULL  */
  int32_t in_dim;
  int32_t out_dim;
  int32_t act;     /* NLSH_ACT_* applied after the bias */
  float act_scale; /* w0 of NLSH_ACT_SIN, ignored otherwise */
} nlsh_layer_t;

int nlsh_version(void);
const char* nlsh_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Host helper: MSB-first bit pack of 0/1 int32 bits into int16 codes.
 * Replaces nlsh/utils.pyx:6-15 (binarr_to_int) as looped by utils.pyx:18-32 (hash_codes);
 * the accumulator is int32 and the result is truncated to int16 exactly as the Cython
 * return type does.  bits is a HOST array [n, s, hs] with element strides (in elements).
 * out is a HOST array [n, s] (contiguous).
 * ------------------------------------------------------------------------------------- */
int nlsh_pack_codes_host(const int32_t* bits, int64_t n, int64_t s, int64_t hs,
                         int64_t stride_n, int64_t stride_s, int64_t stride_b, int16_t* out);

/* ---------------------------------------------------------------------------------------
 * Learned-hasher forward + bucket-code epilogue.
 * Replaces MultiLayerRelu.forward (encoders.py:24-55) + _Hasher.forward
 * (nlsh/hashings.py:13-27, 97-107) + the `probs > 0.5` / bit-pack of hashings.py:66-76 and
 * utils.pyx:6-15 (n == 1 case), or argmax for the softmax head (hashings.py:131-133).
 *
 * layers[0..n_layers-1]: the trunk followed by the output layer (act = IDENTITY); the
 * last layer's out_dim is hash_size.  logits_out (device, [n, hash_size], may be NULL)
 * receives the pre-sigmoid outputs of the last layer; codes_out (device int32 [n], may be
 * NULL) receives the bucket code.  Thresholds reproduce torch's fp32 sigmoid / tanh
 * comparison bit-for-bit: sigmoid(l) > 0.5  <=>  l > 1.5*2^-24;  tanh(l)/2+0.5 > 0.5  <=>
 * l > 2^-24 (verified against torch 2.11 CPU, tests/test_gpu_hash.py::test_threshold_dead_band_bit_exact and tests/test_oracle_golden.py).
 * ------------------------------------------------------------------------------------- */
size_t nlsh_mlp_workspace_bytes(int64_t n, const nlsh_layer_t* layers, int32_t n_layers);
int nlsh_mlp_hash_f32(const float* x, int64_t n, int32_t d, const nlsh_layer_t* layers,
                      int32_t n_layers, int32_t head, float* logits_out, int32_t* codes_out,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Bucket codes from given logits only (the epilogue above as a stand-alone call; used to
 * prove "codes bit-exact from identical fp32 logits"). */
int nlsh_codes_from_logits(const float* logits, int64_t n, int32_t hash_size, int32_t head,
                           int32_t* codes_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Deterministic multi-probe: the p most probable codes under independent bits
 * (sigmoid / tanh heads: flip-sets in increasing sum of |logit|; softmax head: top-p
 * logits).  Replaces the Bernoulli sampling of hashings.py:77-81: probe 0 is
 * always the hard code of nlsh_mlp_hash_f32.  probes_out: device int32 [n, p], row i holds
 * distinct codes in increasing cost order, padded with -1 when 2^hash_size < p.
 * ------------------------------------------------------------------------------------- */
int nlsh_topp_probes(const float* logits, int64_t n, int32_t hash_size, int32_t head,
                     int32_t p, int32_t* probes_out, void* stream);

/* Sampled multi-probe, the reference's own semantics (hashings.py:77-81: n - 1 draws of
 * torch.distributions.Bernoulli(probs) next to the hard code): probes_out[i, 0] is the hard code,
 * probes_out[i, j >= 1] the packed code of an independent draw, bit b set with probability
 * sigmoid(logit_b) (tanh head: tanh(logit_b) / 2 + 0.5).  Uniforms come from Philox-4x32-10 keyed by
 * `seed` with counter (row, j, b / 4): the same (seed, logits) always gives the same probes; rows may
 * repeat a code (the reference collects them into a set, utils.pyx:18-32; the scan ignores repeats).
 * The stream of torch's own sampler is not reproduced (it depends on torch's launch geometry): parity
 * is in distribution, and reference-sampled sets can be replayed through `probes` of the scan. */
int nlsh_sample_probes(const float* logits, int64_t n, int32_t hash_size, int32_t head,
                       int32_t p, uint64_t seed, int32_t* probes_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Index build: codes -> CSR bucket layout + bucket-contiguous copy of the vectors.
 * Replaces build_index (nlsh/indexer.py:6-24) for single-code rows (hash_times = 1, the
 * only way Indexer._build_index calls it, indexer.py:36-38).
 *   offsets_out int32 [n_buckets + 1]  bucket c owns ids_out[offsets[c] : offsets[c+1]]
 *   ids_out     int32 [n]              row ids, ASCENDING inside each bucket (pinned by
 *                                      nlsh/tests/test_indexer.py:14-26), buckets in
 *                                      ascending code order
 *   x_sorted_out fp32 [n, d_pad]       row i is x[ids_out[i]], d_pad = d rounded up to a multiple
 *                                      of 4 (zero filled) so every row starts 16-byte aligned for
 *                                      the bulk-async copies (may be NULL, then x may be NULL)
 *   x_sqnorm_out fp32 [n]              |x_sorted row|^2 (may be NULL; needs x_sorted_out).  The
 *                                      query path's tensor-core filter bounds distances with it.
 * codes must lie in [0, n_buckets).
 * ------------------------------------------------------------------------------------- */
size_t nlsh_build_workspace_bytes(int64_t n, int32_t n_buckets);
int nlsh_build_csr(const int32_t* codes, int64_t n, int32_t n_buckets, const float* x, int32_t d,
                   int32_t* offsets_out, int32_t* ids_out, float* x_sorted_out, float* x_sqnorm_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Query: multi-probe candidate scan + exact distance + top-k, batched over all queries.
 * Replaces the per-query loop of Indexer.query (nlsh/indexer.py:62-95): index_select
 * gather (77-82), distance_func (84-87 -> nlsh/data.py:109 / :201), topk (90-91).
 *   xq        device fp32 [n_queries, d]
 *   probes    device int32 [n_queries, p]; -1 = unused slot; duplicates inside a row are
 *             scanned once (the reference probes a Python set)
 *   offsets / ids / x_sorted: the CSR produced by nlsh_build_csr (x_sorted row stride d_pad)
 *   x_sqnorm  device fp32 [n_rows] from nlsh_build_csr, or NULL.  With it (and d <= 4096)
 *             the scan runs a tcgen05 tf32 GEMM of each row tile against the bucket's queries as
 *             a FILTER: pairs whose distance lower bound exceeds an upper bound of the query's final
 *             k-th best distance are dropped, the rest are scored exactly as below and the k best
 *             of them selected.  Results are the same either way.
 *   max_bucket_rows: max over buckets of offsets[c+1]-offsets[c] (host knows it from build)
 *   ids_out   device int64 [n_queries, k]  row ids (+ id_offset) by ascending distance,
 *             ties broken by smaller id; -1 past the number of candidates
 *   dists_out device fp32 [n_queries, k]   +inf past the number of candidates
 *   ncand_out device int32 [n_queries]     candidates scanned (indexer.py:71,94)
 *   flags: bit 0 = use the synchronous-staging kernel instead of the bulk-async (TMA) ring;
 *          bit 1 = do not use the tensor-core filter even when x_sqnorm is given (debug / A-B only);
 *          bit 2 = tensor-core scan with candidate buffers of only k entries per query, so that
 *                  (almost) every query takes the exact re-scan of the overflow path (tests only);
 *          bit 3 = L2: leave the SQUARED distances in dists_out (a caller that merges the lists of several
 *                  shards takes the root after the merge: two candidates whose roots round to the same
 *                  float then keep the order of their squared distances, as in an unsharded search);
 *          bits 8-15 = number of SMs the tensor-core scan's persistent grid leaves free for kernels of
 *                  other streams (batches in flight on several streams overlap the launch-bound front
 *                  part of one batch with the scan of another; 0 = use every SM)
 * ------------------------------------------------------------------------------------- */
size_t nlsh_query_workspace_bytes(int64_t n_queries, int32_t p, int32_t k, int32_t d,
                                  int32_t n_buckets, int64_t n_rows, int64_t max_bucket_rows);
int nlsh_query_scan_topk(const float* xq, int64_t n_queries, int32_t d, const int32_t* probes,
                         int32_t p, const int32_t* offsets, int32_t n_buckets, const int32_t* ids,
                         const float* x_sorted, const float* x_sqnorm, int64_t n_rows,
                         int64_t max_bucket_rows,
                         int32_t metric, int32_t k, int64_t id_offset, int64_t* ids_out,
                         float* dists_out, int32_t* ncand_out, void* workspace,
                         size_t workspace_bytes, uint32_t flags, void* stream);

/* The same with the queries' distance bounds supplied: tau_seed device fp32 [n_queries] (or NULL = compute them
 * here), each an upper bound of the query's final k-th best distance in the scan's units (squared for L2), e.g.
 * from nlsh_query_seed_tau on ANY shard of the database: k rows within the bound exist somewhere, so no shard
 * needs candidates beyond it.  Row-sharded search (nlsh/parallel.py) lets each rank seed 1/N of the queries and
 * all-gathers the bounds with the probe matrix.  Only the tensor-core scan uses them; results do not change. */
int nlsh_query_scan_topk_seeded(const float* xq, int64_t n_queries, int32_t d, const int32_t* probes,
                                int32_t p, const int32_t* offsets, int32_t n_buckets, const int32_t* ids,
                                const float* x_sorted, const float* x_sqnorm, int64_t n_rows,
                                int64_t max_bucket_rows, int32_t metric, int32_t k, int64_t id_offset,
                                const float* tau_seed, int64_t* ids_out, float* dists_out,
                                int32_t* ncand_out, void* workspace, size_t workspace_bytes, uint32_t flags,
                                void* stream);

/* Distance bounds of a batch of queries from a sample of the rows of their probed buckets (the seed of the
 * tensor-core scan's filter, scan_tc.cu::seed_tau_kernel): tau_out[q] = the exact k-th best distance of query q
 * among the first rows of its probed buckets (in the scan's units, inflated by the rounding bound of another
 * summation order), +inf when they hold fewer than k rows.  No reference counterpart (the reference scores
 * every candidate, nlsh/indexer.py:84-91). */
size_t nlsh_query_seed_workspace_bytes(int64_t n_queries, int32_t d);
int nlsh_query_seed_tau(const float* xq, int64_t n_queries, int32_t d, const int32_t* probes, int32_t p,
                        const int32_t* offsets, int32_t n_buckets, const float* x_sorted, int64_t n_rows,
                        int32_t metric, int32_t k, float* tau_out, void* workspace, size_t workspace_bytes,
                        void* stream);
/* The same with the base sample size chosen by the caller (sample_rows rows of the query's first probed
 * bucket(s), 0 = the library's rule): a rank that seeds only its 1/N slice of the queries affords more rows. */
int nlsh_query_seed_tau_rows(const float* xq, int64_t n_queries, int32_t d, const int32_t* probes, int32_t p,
                             const int32_t* offsets, int32_t n_buckets, const float* x_sorted, int64_t n_rows,
                             int32_t metric, int32_t k, int32_t sample_rows, float* tau_out, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Which kernel nlsh_query_scan_topk runs for this shape with / without x_sqnorm and default flags:
 * 1 = tensor-core filtered scan (scan_tc.cu: d <= 4096, k <= 128 and at least ~4 (query, probe) pairs
 * per bucket, i.e. bucket tiles are shared between queries), 0 = fp32 SIMT scan (scan.cu). */
int nlsh_query_scan_impl(int32_t d, int32_t k, int32_t metric, int32_t has_sqnorm, int64_t n_queries,
                         int32_t p, int32_t n_buckets);

/* ---------------------------------------------------------------------------------------
 * Brute-force kNN (ground truth / training labels).
 * Replaces self_get_knn_pt (precompute.py:57-67) with distance_func = _l2
 * (precompute.py:37-54, NLSH_METRIC_L2SQ) or _cosine_distance (precompute.py:22-34,
 * NLSH_METRIC_COSINE); NLSH_METRIC_L2 / ANGULAR give the scan metrics for held-out
 * queries.  exclude_self != 0: query i (global index self_offset + i) never returns db
 * row self_offset + i (the reference drops the first hit instead, precompute.py:66).
 * ------------------------------------------------------------------------------------- */
size_t nlsh_knn_workspace_bytes(int64_t n_queries, int64_t n_rows, int32_t d, int32_t k);
int nlsh_knn_bruteforce(const float* xq, int64_t n_queries, const float* xdb, int64_t n_rows,
                        int32_t d, int32_t metric, int32_t k, int32_t exclude_self,
                        int64_t self_offset, int64_t id_offset, int64_t* ids_out, float* dists_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Merge of G per-shard top-k lists (after the NCCL all-gather of the sharded search; the
 * reference has no multi-GPU path).  List l lives at dists + l * dist_stride / ids +
 * l * id_stride (strides in elements; 0 = contiguous [G, n_queries, k]), [n_queries, k] each,
 * ascending, ids < 0 = empty slot - so the lists can be read in place from the packed
 * all-gather buffer.  ncand (optional, with ncand_out): per-shard candidate counts
 * [G][n_queries] at stride ncand_stride, summed into ncand_out.  Output as nlsh_query_scan_topk.
 * ------------------------------------------------------------------------------------- */
int nlsh_merge_topk(const float* dists, const int64_t* ids, int64_t dist_stride, int64_t id_stride,
                    const int32_t* ncand, int64_t ncand_stride, int32_t n_lists, int64_t n_queries,
                    int32_t k, int64_t* ids_out, float* dists_out, int32_t* ncand_out, void* stream);

/* recall@k on device: mean over queries of |gt[q,:k_gt] ∩ pred[q,:k_pred]| / k_gt, the
 * definition of nlsh/metrics.py:4-25 (negative pred ids never match). hits_out: device
 * int32 [n_queries]. */
int nlsh_recall_hits(const int64_t* gt, int32_t k_gt, const int64_t* pred, int32_t k_pred,
                     int64_t n_queries, int32_t* hits_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): number of kernels this library has launched in the
 * process, and a ring of CUDA-event pairs recorded on the launching stream around the
 * candidate-scan kernel of every nlsh_query_scan_topk call while enabled (per thread, up to
 * 256 calls).  nlsh_profile_read waits for the events, writes the kernel durations in ms,
 * returns how many it wrote (or a negative error) and rewinds the ring.
 * ------------------------------------------------------------------------------------- */
long long nlsh_kernel_launch_count(void);
int nlsh_profile_enable(int on);
int nlsh_profile_read(float* ms_out, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* NLSH_B200_H */
