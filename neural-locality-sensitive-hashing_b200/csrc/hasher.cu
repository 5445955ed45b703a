// Learned-hasher forward (encoders.py:24-55 + nlsh/hashings.py:13-27), the bucket-code
// epilogue (hashings.py:66-76 + utils.pyx:6-15) and the deterministic multi-probe
// enumeration that replaces the Bernoulli sampling of hashings.py:77-81.
//
// Round-1 kernel for the dense contraction: an fp32 register-tiled SIMT GEMM with the
// bias + activation fused into its epilogue.  It is exact fp32 (sequential-K FMA), which
// is what the 1e-5 logit tolerance needs; the tcgen05 3xTF32 version replaces it next.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

// tensor-core path (tc_linear.cu)
bool nlsh_tc_layer_supported(int in_dim, int out_dim);
int nlsh_tc_split(const float* x, size_t n, float* hi, float* lo, cudaStream_t st);
int nlsh_tc_linear(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo,
                   const float* bias, int M, int N, int K, int act, float act_scale, float* out_hi,
                   float* out_lo, float* out_full, int ld_out, int* codes_out, int head,
                   cudaStream_t st, const float* head_w = nullptr, const float* head_b = nullptr,
                   int head_n = 0);

namespace {

constexpr int kMlpChunkRows = 32768;  // activations of one chunk stay L2 resident (<= 32 MB each)

__device__ __forceinline__ float apply_act(float v, int act, float scale) {
  if (act == NLSH_ACT_RELU) return fmaxf(v, 0.f);
  if (act == NLSH_ACT_SIN) return sinf(scale * v);
  return v;
}

// C[M, N] = act(A[M, K] * W[N, K]^T + bias[N]); A, W, C row-major.
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256)
    linear_act_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W,
                      const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N,
                      int K, int act, float act_scale, int vec_ok) {
  constexpr int BK = 16;
  constexpr int NTX = BN / TN;
  constexpr int NT = (BM / TM) * NTX;
  static_assert(NT == 256, "tile shape must give 256 threads");
  constexpr int A_VECS = BM * BK / 4;
  constexpr int W_VECS = BN * BK / 4;
  constexpr int A_PER_T = (A_VECS + NT - 1) / NT;
  constexpr int W_PER_T = (W_VECS + NT - 1) / NT;

  __shared__ float As[2][BK][BM + 4];
  __shared__ float Ws[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % NTX;
  const int ty = tid / NTX;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_PER_T], rw[W_PER_T];

  auto fetch = [&](const float* __restrict__ base, int ld, int row0, int rows_total, int k0,
                   int idx) -> float4 {
    const int r = idx / (BK / 4);
    const int kq = idx % (BK / 4);
    const int gr = row0 + r;
    const int gk = k0 + kq * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gr < rows_total) {
      const float* p = base + (size_t)gr * ld + gk;
      if (vec_ok && gk + 3 < K) {
        v = *reinterpret_cast<const float4*>(p);
      } else {
        if (gk + 0 < K) v.x = p[0];
        if (gk + 1 < K) v.y = p[1];
        if (gk + 2 < K) v.z = p[2];
        if (gk + 3 < K) v.w = p[3];
      }
    }
    return v;
  };
  auto load_regs = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER_T; ++i) {
      const int idx = tid + i * NT;
      if (idx < A_VECS) ra[i] = fetch(A, lda, m0, M, k0, idx);
    }
#pragma unroll
    for (int i = 0; i < W_PER_T; ++i) {
      const int idx = tid + i * NT;
      if (idx < W_VECS) rw[i] = fetch(W, K, n0, N, k0, idx);
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER_T; ++i) {
      const int idx = tid + i * NT;
      if (idx < A_VECS) {
        const int r = idx / (BK / 4), kq = idx % (BK / 4);
        As[buf][kq * 4 + 0][r] = ra[i].x;
        As[buf][kq * 4 + 1][r] = ra[i].y;
        As[buf][kq * 4 + 2][r] = ra[i].z;
        As[buf][kq * 4 + 3][r] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < W_PER_T; ++i) {
      const int idx = tid + i * NT;
      if (idx < W_VECS) {
        const int r = idx / (BK / 4), kq = idx % (BK / 4);
        Ws[buf][kq * 4 + 0][r] = rw[i].x;
        Ws[buf][kq * 4 + 1][r] = rw[i].y;
        Ws[buf][kq * 4 + 2][r] = rw[i].z;
        Ws[buf][kq * 4 + 3][r] = rw[i].w;
      }
    }
  };

  const int n_ktiles = (K + BK - 1) / BK;
  load_regs(0);
  store_smem(0);
  __syncthreads();

  for (int kt = 0; kt < n_ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_ktiles) load_regs((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[buf][k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Ws[buf][k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < n_ktiles) {
      store_smem(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn < N) {
        float v = acc[i][j] + (bias ? bias[gn] : 0.f);
        C[(size_t)gm * ldc + gn] = apply_act(v, act, act_scale);
      }
    }
  }
}

// Bucket code from fp32 logits.  torch evaluates sigmoid(l) > 0.5 in fp32, which is true
// exactly for l > 1.5 * 2^-24 (and tanh(l)/2 + 0.5 > 0.5 exactly for l > 2^-24): SURVEY Q2.
__device__ __forceinline__ float head_threshold(int head) {
  return head == NLSH_HEAD_TANH ? 5.9604644775390625e-08f : 8.940696716308594e-08f;
}

__global__ void codes_kernel(const float* __restrict__ logits, long long n, int hs, int head,
                             int* __restrict__ codes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* l = logits + i * hs;
  int code = 0;
  if (head == NLSH_HEAD_SOFTMAX) {
    float best = l[0];
    for (int j = 1; j < hs; ++j) {
      const float v = l[j];
      if (v > best) {  // first maximum wins, as torch.argmax
        best = v;
        code = j;
      }
    }
  } else {
    const float thr = head_threshold(head);
    for (int j = 0; j < hs; ++j) code = (code << 1) | (l[j] > thr ? 1 : 0);  // MSB first
  }
  codes[i] = code;
}

// One warp per row: the p cheapest flip-masks (cost = sum of |logit| over flipped bits,
// added in logit-index order; ties by smaller mask) XOR the hard code.
// Only the min(p, hs) cheapest bits (ties: lower bit position first) can occur in one of the p
// cheapest masks (each of those single-bit masks, and the empty mask, precedes any mask that
// contains a costlier bit), so 2^min(p, hs) masks are enumerated instead of 2^hs; the result
// is identical to the full enumeration of oracle.topp_probes.
template <int KPL>
__global__ void __launch_bounds__(128)
    probes_topp_kernel(const float* __restrict__ logits, long long n, int hs, int head, int p,
                            int* __restrict__ probes) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = lane_id();
  const float thr = head_threshold(head);
  float a[NLSH_MAX_HASH_BITS];  // a[i] = |logit i|; logit i is bit position hs - 1 - i
  int base = 0;
#pragma unroll
  for (int i = 0; i < NLSH_MAX_HASH_BITS; ++i) {
    a[i] = 0.f;
    if (i < hs) {
      const float v = logits[row * hs + i];
      a[i] = fabsf(v);
      base = (base << 1) | (v > thr ? 1 : 0);
    }
  }
  // lane = bit position: rank of this bit among all bits by (cost, position)
  float my_cost = 0.f;
#pragma unroll
  for (int i = 0; i < NLSH_MAX_HASH_BITS; ++i)
    if (i < hs && lane == hs - 1 - i) my_cost = a[i];
  int rank = 0;
#pragma unroll
  for (int i = 0; i < NLSH_MAX_HASH_BITS; ++i) {
    if (i < hs) {
      const int pos = hs - 1 - i;
      rank += (a[i] < my_cost || (a[i] == my_cost && pos < lane)) ? 1 : 0;
    }
  }
  const int n_sel = p < hs ? p : hs;
  const unsigned sel = __ballot_sync(NLSH_FULL_MASK, lane < hs && rank < n_sel);  // selected positions

  WarpTopK<KPL, int> top;
  top.init(NLSH_ID_SENTINEL);
  const int n_masks = 1 << n_sel;
  for (int m0 = 0; m0 < n_masks; m0 += 32) {
    const int mc = m0 + lane;  // compact mask over the selected bits (ascending position)
    float cost = 0.f;
    int full = 0;
#pragma unroll
    for (int i = 0; i < NLSH_MAX_HASH_BITS; ++i) {
      if (i < hs) {
        const int pos = hs - 1 - i;
        if ((sel >> pos) & 1u) {
          const int j = __popc(sel & ((1u << pos) - 1u));
          if ((mc >> j) & 1) {
            cost = __fadd_rn(cost, a[i]);
            full |= 1 << pos;
          }
        }
      }
    }
    top.offer(cost, full, mc < n_masks, p);
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < p) {
      const int m = top.id[j];
      probes[row * p + pos] = (m == NLSH_ID_SENTINEL) ? -1 : (base ^ m);
    }
  }
}

// Softmax head: the p largest logits (ties by smaller index).
template <int KPL>
__global__ void __launch_bounds__(128)
    probes_softmax_kernel(const float* __restrict__ logits, long long n, int hs, int p,
                          int* __restrict__ probes) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = lane_id();
  WarpTopK<KPL, int> top;
  top.init(NLSH_ID_SENTINEL);
  for (int c0 = 0; c0 < hs; c0 += 32) {
    const int c = c0 + lane;
    const float v = (c < hs) ? -logits[row * hs + c] : 0.f;
    top.offer(v, c, c < hs, p);
  }
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int pos = j * 32 + lane;
    if (pos < p) {
      const int c = top.id[j];
      probes[row * p + pos] = (c == NLSH_ID_SENTINEL) ? -1 : c;
    }
  }
}

int validate_layers(int32_t d, const nlsh_layer_t* layers, int32_t n_layers, int32_t head) {
  NLSH_REQUIRE(layers != nullptr, "mlp: layers is NULL");
  NLSH_REQUIRE(n_layers >= 1 && n_layers <= NLSH_MAX_LAYERS, "mlp: n_layers=%d outside [1, %d]",
               n_layers, NLSH_MAX_LAYERS);
  NLSH_REQUIRE(head == NLSH_HEAD_SIGMOID || head == NLSH_HEAD_TANH || head == NLSH_HEAD_SOFTMAX,
               "mlp: unknown head %d", head);
  int prev = d;
  for (int l = 0; l < n_layers; ++l) {
    NLSH_REQUIRE(layers[l].weight != nullptr, "mlp: layer %d has no weight", l);
    NLSH_REQUIRE(layers[l].in_dim == prev, "mlp: layer %d in_dim=%d but previous width is %d", l,
                 layers[l].in_dim, prev);
    NLSH_REQUIRE(layers[l].out_dim >= 1, "mlp: layer %d out_dim=%d", l, layers[l].out_dim);
    NLSH_REQUIRE(layers[l].act >= NLSH_ACT_IDENTITY && layers[l].act <= NLSH_ACT_SIN,
                 "mlp: layer %d unknown activation %d", l, layers[l].act);
    prev = layers[l].out_dim;
  }
  if (head != NLSH_HEAD_SOFTMAX)
    NLSH_REQUIRE(prev <= NLSH_MAX_HASH_BITS,
                 "mlp: hash_size=%d exceeds %d bits (codes are int16 in the reference)", prev,
                 NLSH_MAX_HASH_BITS);
  return NLSH_OK;
}

struct MlpPlan {
  int64_t chunk;
  int width;
  size_t buf_floats;
};

MlpPlan mlp_plan(int64_t n, const nlsh_layer_t* layers, int32_t n_layers) {
  MlpPlan p;
  p.chunk = n < kMlpChunkRows ? (n > 0 ? n : 1) : kMlpChunkRows;
  p.width = 1;
  for (int l = 0; l < n_layers; ++l)
    if (layers[l].out_dim > p.width) p.width = layers[l].out_dim;
  p.buf_floats = (size_t)p.chunk * p.width;
  return p;
}

int launch_linear(const float* A, int lda, const nlsh_layer_t& L, float* C, int ldc, int M,
                  cudaStream_t st) {
  const int N = L.out_dim, K = L.in_dim;
  const int vec_ok = (K % 4 == 0) && (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(L.weight) & 15) == 0);
  if (N <= 16) {
    dim3 grid((M + 127) / 128, (N + 15) / 16);
    linear_act_kernel<128, 16, 8, 1><<<grid, 256, 0, st>>>(A, lda, L.weight, L.bias, C, ldc, M, N, K,
                                                          L.act, L.act_scale, vec_ok);
  } else if (N <= 64) {
    dim3 grid((M + 127) / 128, (N + 63) / 64);
    linear_act_kernel<128, 64, 8, 4><<<grid, 256, 0, st>>>(A, lda, L.weight, L.bias, C, ldc, M, N, K,
                                                          L.act, L.act_scale, vec_ok);
  } else {
    dim3 grid((M + 127) / 128, (N + 127) / 128);
    linear_act_kernel<128, 128, 8, 8><<<grid, 256, 0, st>>>(A, lda, L.weight, L.bias, C, ldc, M, N,
                                                           K, L.act, L.act_scale, vec_ok);
  }
  return nlsh_check_cuda(nlsh_post_launch(), "linear_act_kernel launch");
}

// ---- tensor-core path (tcgen05, 3xTF32 split) ------------------------------------------------
// NLSH_MLP_IMPL=simt forces the fp32 SIMT kernels, NLSH_MLP_IMPL=tc the tensor cores (A/B runs).  The
// default is the tensor cores whenever every layer has in_dim % 4 == 0 and out_dim <= 256 and no layer
// ends in a sine: sin(w0 x) with w0 = 30 (the first SIREN layer, encoders.py:58-79) amplifies the
// 2^-21-relative error of the 3xTF32 split thirty-fold, measured on a B200 (profiles/r2_siren.jsonl) as
// 0.9e-5 .. 1.7e-5 of the row scale against the fp32 oracle where the fp32 SIMT kernels hold 1e-6 .. 6e-6,
// so sine trunks take the SIMT kernels and keep the 1e-5 logit bar.
bool mlp_use_tc(const nlsh_layer_t* layers, int32_t n_layers) {
  const char* env = getenv("NLSH_MLP_IMPL");
  if (env != nullptr && strcmp(env, "simt") == 0) return false;
  const bool forced = env != nullptr && strcmp(env, "tc") == 0;
  for (int l = 0; l < n_layers; ++l) {
    if (!nlsh_tc_layer_supported(layers[l].in_dim, layers[l].out_dim)) return false;
    if (layers[l].act == NLSH_ACT_SIN && !forced) return false;
  }
  return true;
}

struct TcPlan {
  int64_t chunk;
  size_t act_floats;     // per activation buffer: chunk * widest layer input
  size_t weight_floats;  // all layers
  int hs;
};

TcPlan tc_plan(int64_t n, const nlsh_layer_t* layers, int32_t n_layers) {
  TcPlan p;
  int width = 4;
  p.weight_floats = 0;
  for (int l = 0; l < n_layers; ++l) {
    if (layers[l].in_dim > width) width = layers[l].in_dim;
    p.weight_floats += (size_t)layers[l].in_dim * layers[l].out_dim;
  }
  // Rows per chunk of the tensor-core path: whole waves of 128-row CTAs - two per SM-count (37 888 rows on 148 SMs:
  // a 32 768-row chunk is 1.73 waves and pays for two) - while a chunk's split activations stay about L2 sized;
  // one wave for very wide inputs.  NLSH_MLP_CHUNK=<rows> overrides (A/B runs).
  const int64_t wave = (int64_t)nlsh_num_sms() * 128;
  int64_t chunk = width > 512 ? wave : 2 * wave;
  if (const char* env = getenv("NLSH_MLP_CHUNK")) {
    const long long v = atoll(env);
    if (v >= 128) chunk = v;
  }
  if (n < chunk) chunk = n > 0 ? n : 1;
  p.chunk = chunk;
  p.act_floats = (size_t)chunk * width;
  p.hs = layers[n_layers - 1].out_dim;
  return p;
}

size_t tc_workspace_bytes(int64_t n, const nlsh_layer_t* layers, int32_t n_layers) {
  const TcPlan p = tc_plan(n, layers, n_layers);
  WorkspaceCarver ws(nullptr);
  ws.take<float>(p.weight_floats);
  ws.take<float>(p.weight_floats);
  for (int i = 0; i < 4; ++i) ws.take<float>(p.act_floats);
  ws.take<float>((size_t)p.chunk * p.hs);
  return ws.total();
}

int mlp_hash_tc(const float* x, int64_t n, int32_t d, const nlsh_layer_t* layers, int32_t n_layers,
                int32_t head, float* logits_out, int32_t* codes_out, void* workspace,
                cudaStream_t st) {
  const TcPlan p = tc_plan(n, layers, n_layers);
  WorkspaceCarver ws(workspace);
  float* w_hi = ws.take<float>(p.weight_floats);
  float* w_lo = ws.take<float>(p.weight_floats);
  float* act[4];
  for (int i = 0; i < 4; ++i) act[i] = ws.take<float>(p.act_floats);
  float* logits_tmp = ws.take<float>((size_t)p.chunk * p.hs);

  // The output layer (out_dim = hash_size <= 16, no activation) rides in the epilogue of the layer before it
  // (tc_linear.cu, "fused head") instead of being a tcgen05 layer of its own; NLSH_MLP_FUSE_HEAD=0: A/B runs.
  bool fuse_head = n_layers >= 2 && layers[n_layers - 1].out_dim <= 16 &&
                   layers[n_layers - 1].act == NLSH_ACT_IDENTITY &&
                   layers[n_layers - 1].in_dim == layers[n_layers - 2].out_dim;
  if (const char* env = getenv("NLSH_MLP_FUSE_HEAD")) fuse_head = fuse_head && atoi(env) != 0;
  const int n_tc_layers = fuse_head ? n_layers - 1 : n_layers;
  // weights change between index builds while training: split them at every call (tiny)
  size_t w_off[NLSH_MAX_LAYERS];
  size_t off = 0;
  for (int l = 0; l < n_tc_layers; ++l) {
    w_off[l] = off;
    const size_t cnt = (size_t)layers[l].in_dim * layers[l].out_dim;
    int rc = nlsh_tc_split(layers[l].weight, cnt, w_hi + off, w_lo + off, st);
    if (rc != NLSH_OK) return rc;
    off += cnt;
  }
  for (int64_t r0 = 0; r0 < n; r0 += p.chunk) {
    const int rows = (int)((n - r0) < p.chunk ? (n - r0) : p.chunk);
    int rc = nlsh_tc_split(x + (size_t)r0 * d, (size_t)rows * d, act[0], act[1], st);
    if (rc != NLSH_OK) return rc;
    int cur = 0;  // act[cur], act[cur + 1] hold the split input of the next layer
    float* logits_chunk = logits_out ? logits_out + (size_t)r0 * p.hs : logits_tmp;
    for (int l = 0; l < n_tc_layers; ++l) {
      const bool last = (l == n_tc_layers - 1);
      const bool fused = last && fuse_head;
      const nlsh_layer_t& out_l = layers[n_layers - 1];
      const int nxt = cur ^ 2;
      rc = nlsh_tc_linear(act[cur], act[cur + 1], w_hi + w_off[l], w_lo + w_off[l], layers[l].bias,
                          rows, layers[l].out_dim, layers[l].in_dim, layers[l].act,
                          layers[l].act_scale, last ? nullptr : act[nxt], last ? nullptr : act[nxt + 1],
                          last ? logits_chunk : nullptr, fused ? out_l.out_dim : layers[l].out_dim,
                          last && codes_out ? codes_out + r0 : nullptr, head, st,
                          fused ? out_l.weight : nullptr, fused ? out_l.bias : nullptr, fused ? out_l.out_dim : 0);
      if (rc != NLSH_OK) return rc;
      cur = nxt;
    }
  }
  return NLSH_OK;
}

}  // namespace

extern "C" size_t nlsh_mlp_workspace_bytes(int64_t n, const nlsh_layer_t* layers, int32_t n_layers) {
  if (layers == nullptr || n_layers < 1 || n_layers > NLSH_MAX_LAYERS) return 0;
  const MlpPlan p = mlp_plan(n, layers, n_layers);
  WorkspaceCarver ws(nullptr);
  ws.take<float>(p.buf_floats);
  ws.take<float>(p.buf_floats);
  ws.take<float>((size_t)p.chunk * layers[n_layers - 1].out_dim);
  size_t need = ws.total();
  if (mlp_use_tc(layers, n_layers)) {
    const size_t tc = tc_workspace_bytes(n, layers, n_layers);
    if (tc > need) need = tc;
  }
  return need;
}

extern "C" int nlsh_codes_from_logits(const float* logits, int64_t n, int32_t hash_size,
                                      int32_t head, int32_t* codes_out, void* stream) {
  NLSH_REQUIRE(n >= 0 && hash_size >= 1, "codes: bad shape n=%lld hash_size=%d", (long long)n,
               hash_size);
  NLSH_REQUIRE(head >= NLSH_HEAD_SIGMOID && head <= NLSH_HEAD_SOFTMAX, "codes: unknown head %d", head);
  if (head != NLSH_HEAD_SOFTMAX)
    NLSH_REQUIRE(hash_size <= NLSH_MAX_HASH_BITS, "codes: hash_size=%d exceeds %d", hash_size,
                 NLSH_MAX_HASH_BITS);
  if (n == 0) return NLSH_OK;
  NLSH_REQUIRE(logits && codes_out, "codes: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  const long long blocks = (n + threads - 1) / threads;
  codes_kernel<<<(unsigned)blocks, threads, 0, st>>>(logits, n, hash_size, head, codes_out);
  return nlsh_check_cuda(nlsh_post_launch(), "codes_kernel launch");
}

extern "C" int nlsh_mlp_hash_f32(const float* x, int64_t n, int32_t d, const nlsh_layer_t* layers,
                                 int32_t n_layers, int32_t head, float* logits_out,
                                 int32_t* codes_out, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  NLSH_REQUIRE(n >= 0 && d >= 1, "mlp: bad shape n=%lld d=%d", (long long)n, d);
  int rc = validate_layers(d, layers, n_layers, head);
  if (rc != NLSH_OK) return rc;
  if (n == 0) return NLSH_OK;
  NLSH_REQUIRE(x != nullptr, "mlp: x is NULL");
  const size_t need = nlsh_mlp_workspace_bytes(n, layers, n_layers);
  if (workspace == nullptr || workspace_bytes < need) {
    nlsh_set_error("mlp: workspace %zu bytes < required %zu", workspace_bytes, need);
    return NLSH_ERR_WORKSPACE;
  }
  if (mlp_use_tc(layers, n_layers) && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    return mlp_hash_tc(x, n, d, layers, n_layers, head, logits_out, codes_out, workspace,
                       static_cast<cudaStream_t>(stream));
  const MlpPlan plan = mlp_plan(n, layers, n_layers);
  const int hs = layers[n_layers - 1].out_dim;
  WorkspaceCarver ws(workspace);
  float* buf[2];
  buf[0] = ws.take<float>(plan.buf_floats);
  buf[1] = ws.take<float>(plan.buf_floats);
  float* logits_tmp = ws.take<float>((size_t)plan.chunk * hs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  for (int64_t r0 = 0; r0 < n; r0 += plan.chunk) {
    const int rows = (int)((n - r0) < plan.chunk ? (n - r0) : plan.chunk);
    const float* in = x + (size_t)r0 * d;
    int ld_in = d;
    float* logits_chunk = logits_out ? logits_out + (size_t)r0 * hs : logits_tmp;
    for (int l = 0; l < n_layers; ++l) {
      const bool last = (l == n_layers - 1);
      float* out = last ? logits_chunk : buf[l & 1];
      const int ld_out = layers[l].out_dim;
      rc = launch_linear(in, ld_in, layers[l], out, ld_out, rows, st);
      if (rc != NLSH_OK) return rc;
      in = out;
      ld_in = ld_out;
    }
    if (codes_out) {
      rc = nlsh_codes_from_logits(logits_chunk, rows, hs, head, codes_out + r0, stream);
      if (rc != NLSH_OK) return rc;
    }
  }
  return NLSH_OK;
}

// ---- sampled multi-probe (hashings.py:77-81) ------------------------------------------------
// Philox-4x32-10 (Salmon et al. 2011), the counter-based generator torch's CUDA sampler is built on: one
// call = four 32-bit words from (counter, key); no state, so a (row, probe, bit) triple always gets the
// same uniform for a given seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// probes[i, 0] = hard code (`dist.probs > 0.5`), probes[i, j] for j >= 1 = packed code of one draw of
// Bernoulli(probs) per bit, bit b set iff u(i, j, b) < prob_b with u = (word >> 8) * 2^-24 in [0, 1).
// One thread per (row, probe).
__global__ void __launch_bounds__(256)
    probes_sample_kernel(const float* __restrict__ logits, long long n, int hs, int head, int p,
                         unsigned long long seed, int* __restrict__ probes) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * p) return;
  const long long row = t / p;
  const int j = (int)(t - row * p);
  const float thr = head == NLSH_HEAD_TANH ? 5.9604644775390625e-08f : 8.940696716308594e-08f;
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  int code = 0;
  for (int b0 = 0; b0 < hs; b0 += 4) {
    const uint4 w = philox4x32_10(make_uint4((unsigned)row, (unsigned)(row >> 32), (unsigned)j, (unsigned)(b0 >> 2)), key);
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (b0 + e < hs) {
        const float l = logits[row * hs + b0 + e];
        int bit;
        if (j == 0) {
          bit = l > thr ? 1 : 0;  // the reference's fp32 `probs > 0.5`, as in codes_kernel
        } else {
          const float prob = head == NLSH_HEAD_TANH ? tanhf(l) * 0.5f + 0.5f : 1.0f / (1.0f + expf(-l));
          const float u = (float)(ws[e] >> 8) * 5.9604644775390625e-08f;
          bit = u < prob ? 1 : 0;
        }
        code = (code << 1) | bit;  // MSB first (utils.pyx:12-14)
      }
    }
  }
  probes[t] = code;
}

extern "C" int nlsh_sample_probes(const float* logits, int64_t n, int32_t hash_size, int32_t head,
                                  int32_t p, uint64_t seed, int32_t* probes_out, void* stream) {
  NLSH_REQUIRE(n >= 0 && hash_size >= 1 && hash_size <= NLSH_MAX_HASH_BITS,
               "sample probes: bad shape n=%lld hash_size=%d", (long long)n, hash_size);
  NLSH_REQUIRE(p >= 1 && p <= 1024, "sample probes: p=%d outside [1, 1024]", p);
  NLSH_REQUIRE(head == NLSH_HEAD_SIGMOID || head == NLSH_HEAD_TANH,
               "sample probes: head %d has no Bernoulli sampling (hashings.py:66-92 is the bit hasher)", head);
  if (n == 0) return NLSH_OK;
  NLSH_REQUIRE(logits && probes_out, "sample probes: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = (long long)n * p;
  probes_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(logits, n, hash_size, head, p, seed,
                                                                       probes_out);
  return nlsh_check_cuda(nlsh_post_launch(), "probes_sample_kernel launch");
}

extern "C" int nlsh_topp_probes(const float* logits, int64_t n, int32_t hash_size, int32_t head,
                                int32_t p, int32_t* probes_out, void* stream) {
  NLSH_REQUIRE(n >= 0 && hash_size >= 1, "probes: bad shape n=%lld hash_size=%d", (long long)n,
               hash_size);
  NLSH_REQUIRE(p >= 1 && p <= NLSH_MAX_K, "probes: p=%d outside [1, %d]", p, NLSH_MAX_K);
  NLSH_REQUIRE(head >= NLSH_HEAD_SIGMOID && head <= NLSH_HEAD_SOFTMAX, "probes: unknown head %d",
               head);
  if (head != NLSH_HEAD_SOFTMAX)
    NLSH_REQUIRE(hash_size <= NLSH_MAX_HASH_BITS, "probes: hash_size=%d exceeds %d", hash_size,
                 NLSH_MAX_HASH_BITS);
  if (n == 0) return NLSH_OK;
  NLSH_REQUIRE(logits && probes_out, "probes: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps_per_block = 4;
  const unsigned blocks = (unsigned)((n + warps_per_block - 1) / warps_per_block);
  const int kpl = p <= 32 ? 1 : (p <= 64 ? 2 : 4);
  if (head == NLSH_HEAD_SOFTMAX) {
    if (kpl == 1)
      probes_softmax_kernel<1><<<blocks, 128, 0, st>>>(logits, n, hash_size, p, probes_out);
    else if (kpl == 2)
      probes_softmax_kernel<2><<<blocks, 128, 0, st>>>(logits, n, hash_size, p, probes_out);
    else
      probes_softmax_kernel<4><<<blocks, 128, 0, st>>>(logits, n, hash_size, p, probes_out);
  } else {
    if (kpl == 1)
      probes_topp_kernel<1><<<blocks, 128, 0, st>>>(logits, n, hash_size, head, p, probes_out);
    else if (kpl == 2)
      probes_topp_kernel<2><<<blocks, 128, 0, st>>>(logits, n, hash_size, head, p, probes_out);
    else
      probes_topp_kernel<4><<<blocks, 128, 0, st>>>(logits, n, hash_size, head, p, probes_out);
  }
  return nlsh_check_cuda(nlsh_post_launch(), "probes kernel launch");
}
