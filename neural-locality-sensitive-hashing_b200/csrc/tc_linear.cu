// Tensor-core path of the hasher's dense contraction (encoders.py:41-55 nn.Linear chains,
// nlsh/hashings.py:19-22 output layer): C[M, N] = act(A[M, K] W[N, K]^T + b) on the 5th-gen
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) at fp32-equivalent accuracy.
//
// Accuracy: a tf32 operand keeps 11 significant bits, far from the 1e-5 logit bar, so every
// fp32 operand x is split exactly into hi = x with the 13 low mantissa bits cleared and
// lo = x - hi, and each product is issued three times into the same fp32 accumulator:
// hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-22 relative).  The tensor core truncates
// when it adds into the fp32 accumulator, a one-sided error that grows with the number of
// accumulation steps, so the large hi*hi term is spread round-robin over up to four TMEM
// accumulators and the two small cross terms go to their own accumulator; the epilogue adds
// them with ordinary round-to-nearest fp32 adds.  The split operands are
// materialised in global memory (weights once per call; activations by the previous layer's
// epilogue, so a chunk's hi/lo activations stay L2 resident) and streamed by TMA.
//
// One CTA = one 128-row tile of A against all N <= 256 columns.  Warp roles: warp 4 lane 0
// issues the TMA loads (SWIZZLE_128B, K-major, 32 fp32 = 128 bytes per row per K block) into
// a ring of stages; warp 5 lane 0 issues the tcgen05.mma's and commits stage releases /
// accumulator completion to mbarriers; warps 0-3 are the epilogue: tcgen05.ld the 128 x N
// accumulator (one row per thread), add bias, apply the activation, and either write the
// next layer's hi/lo operands or the logits + bucket codes.
#include <string.h>

#include "tc_common.cuh"

namespace {

constexpr int kTcThreads = 192;

struct TcArgs {
  const float* bias;  // [N] or nullptr
  float* out_hi;      // [M, ld_out] split outputs for the next layer (nullptr on the last)
  float* out_lo;
  float* out_full;    // [M, ld_out] plain fp32 output (logits), may be nullptr
  int* codes_out;     // [M] bucket codes (last layer), may be nullptr
  int M, N, K;
  int n_pad;          // UMMA N: N rounded up to a multiple of 16
  int ld_out;
  int act;
  float act_scale;
  int head;
  int stages;
  int n_main;         // TMEM accumulators for hi*hi (k-block kb uses kb % n_main); one more for the cross terms
  int tmem_cols;      // power of two >= max(32, (n_main + 1) * n_pad)
  // Fused head: the NEXT (last) layer of the hasher, out_dim = head_n <= 16 (nlsh/hashings.py:19-22 output_layer),
  // computed by the epilogue on the CUDA cores from this layer's activations while they are in registers:
  // logits[row, j] = head_b[j] + sum_c act(...)[row, c] * head_w[j, c] (fp32 FMA chain in ascending c).  As a
  // tcgen05 layer it is 96 MMAs of N = 16 - the issue cost of a 256-wide layer for 1/16 of its work - plus a
  // launch and a round trip of the split activations.
  const float* head_w;  // [head_n, N] fp32 or nullptr
  const float* head_b;  // [head_n] or nullptr
  int head_n;
};

constexpr int kHeadMax = 16;

template <bool FUSED>  // FUSED: the hasher's output layer rides in this layer's epilogue (TcArgs::head_w)
__global__ void __launch_bounds__(kTcThreads, 1)
    tc_linear_kernel(const TcArgs a, const __grid_constant__ CUtensorMap map_a_hi,
                     const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_w_hi,
                     const __grid_constant__ CUtensorMap map_w_lo) {
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  const size_t a_bytes = (size_t)kTcBM * kTcBK * sizeof(float);    // 16 KB
  const size_t w_bytes = (size_t)a.n_pad * kTcBK * sizeof(float);  // n_pad * 128 B (multiple of 2 KB)
  const size_t stage_bytes = 2 * a_bytes + 2 * w_bytes;
  unsigned char* tail = base + (size_t)a.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);  // [stages <= 4]
  uint64_t* empty_bar = full_bar + 4;                      // [stages <= 4]
  uint64_t* acc_bar = empty_bar + 4;                       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  float* head_s = reinterpret_cast<float*>(tail + 256);  // [n_pad][kHeadMax] head weights, column-major groups

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int m0 = blockIdx.x * kTcBM;
  const int n_kblocks = (a.K + kTcBK - 1) / kTcBK;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 5) {  // one warp allocates the accumulator columns and owns the deallocation
    tc_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------- TMA producer ---------------------------------------
    if (lane == 0) {
      int s = 0;
      unsigned par = 0;  // running stage / round parity instead of a run-time modulo per K block
      for (int kb = 0; kb < n_kblocks; ++kb) {
        mbar_wait_poll(&empty_bar[s], par ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], (unsigned)stage_bytes);
        unsigned char* st = base + (size_t)s * stage_bytes;
        tma_load_2d(st, &map_a_hi, kb * kTcBK, m0, &full_bar[s]);
        tma_load_2d(st + a_bytes, &map_a_lo, kb * kTcBK, m0, &full_bar[s]);
        tma_load_2d(st + 2 * a_bytes, &map_w_hi, kb * kTcBK, 0, &full_bar[s]);
        tma_load_2d(st + 2 * a_bytes + w_bytes, &map_w_lo, kb * kTcBK, 0, &full_bar[s]);
        if (++s == a.stages) {
          s = 0;
          par ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------- MMA issuer -----------------------------------------
    if (lane == 0) {
      const uint32_t idesc = make_tf32_idesc(a.n_pad);
      int s = 0;
      unsigned par = 0;
      for (int kb = 0; kb < n_kblocks; ++kb) {
        mbar_wait_poll(&full_bar[s], par);
        tc_fence_after();
        unsigned char* st = base + (size_t)s * stage_bytes;
        const uint64_t da_hi = make_kmajor_sw128_desc(st);
        const uint64_t da_lo = make_kmajor_sw128_desc(st + a_bytes);
        const uint64_t dw_hi = make_kmajor_sw128_desc(st + 2 * a_bytes);
        const uint64_t dw_lo = make_kmajor_sw128_desc(st + 2 * a_bytes + w_bytes);
        const uint32_t acc_main = tmem_base + (uint32_t)((kb % a.n_main) * a.n_pad);
        const uint32_t acc_cross = tmem_base + (uint32_t)(a.n_main * a.n_pad);
#pragma unroll
        for (int k = 0; k < kTcBK / 8; ++k) {  // UMMA K = 8 tf32 = 32 bytes: +2 in (addr >> 4)
          const uint64_t adv = (uint64_t)(k * 2);
          tc_mma_tf32(acc_main, da_hi + adv, dw_hi + adv, idesc, (kb >= a.n_main || k > 0) ? 1u : 0u);
          tc_mma_tf32(acc_cross, da_hi + adv, dw_lo + adv, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_mma_tf32(acc_cross, da_lo + adv, dw_hi + adv, idesc, 1u);
        }
        tc_commit(&empty_bar[s]);  // the stage may be refilled once these MMAs have read it
        if (++s == a.stages) {
          s = 0;
          par ^= 1u;
        }
      }
      tc_commit(acc_bar);  // accumulator complete
    }
    __syncwarp();
  } else {
    // ------------------------------- epilogue (warps 0-3) -------------------------------
    constexpr bool fused = FUSED;
    if (fused) {  // stage the head weights while the MMAs run: head_s[c][j] = head_w[j][c], zero padded
#pragma unroll 4
      for (int idx = tid; idx < a.n_pad * kHeadMax; idx += 128) {  // coalesced reads along c
        const int j = idx / a.n_pad, c = idx - j * a.n_pad;
        head_s[c * kHeadMax + j] = (j < a.head_n && c < a.N) ? a.head_w[(size_t)j * a.N + c] : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
    }
    float logit[FUSED ? kHeadMax : 1];
#pragma unroll
    for (int j = 0; j < (FUSED ? kHeadMax : 1); ++j) logit[j] = 0.f;
    mbar_wait_poll(acc_bar, 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;  // TMEM lane = accumulator row
    const bool row_ok = row < a.M;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool last = a.out_hi == nullptr;
    const float thr = a.head == NLSH_HEAD_TANH ? 5.9604644775390625e-08f : 8.940696716308594e-08f;
    int code = 0;
    float best = 0.f;
    const int n_main_used = a.n_main < n_kblocks ? a.n_main : n_kblocks;
    for (int c0 = 0; c0 < a.n_pad; c0 += 16) {
      float v[16], t[16];
      {  // first hi*hi accumulator and the cross-term accumulator in flight together (warp-collective loads)
        uint32_t vr[16], tr[16];
        tc_ld16_nowait(lane_addr + (uint32_t)c0, vr);
        tc_ld16_nowait(lane_addr + (uint32_t)(a.n_main * a.n_pad + c0), tr);
        tc_wait_ld2(vr, tr);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] = __uint_as_float(vr[i]);
          t[i] = __uint_as_float(tr[i]);
        }
      }
      float cross[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) cross[i] = t[i];
      for (int j = 1; j < n_main_used; ++j) {
        tc_ld16(lane_addr + (uint32_t)(j * a.n_pad + c0), t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += cross[i];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = c0 + i;
        float x = v[i] + ((a.bias != nullptr && col < a.N) ? a.bias[col] : 0.f);
        v[i] = tc_act(x, a.act, a.act_scale);
      }
      if constexpr (FUSED) {  // 16 columns x 16 head outputs (columns / outputs past the layer hold zero weights)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4* w = reinterpret_cast<const float4*>(head_s + (size_t)(c0 + i) * kHeadMax);
#pragma unroll
          for (int j4 = 0; j4 < kHeadMax / 4; ++j4) {
            const float4 ww = w[j4];  // the same address in every lane: a broadcast
            logit[4 * j4 + 0] = fmaf(v[i], ww.x, logit[4 * j4 + 0]);
            logit[4 * j4 + 1] = fmaf(v[i], ww.y, logit[4 * j4 + 1]);
            logit[4 * j4 + 2] = fmaf(v[i], ww.z, logit[4 * j4 + 2]);
            logit[4 * j4 + 3] = fmaf(v[i], ww.w, logit[4 * j4 + 3]);
          }
        }
        continue;
      }
      if (!row_ok) continue;
      if (!last) {
        float* ph = a.out_hi + (size_t)row * a.ld_out + c0;
        float* pl = a.out_lo + (size_t)row * a.ld_out + c0;
        if (c0 + 16 <= a.N && (a.ld_out & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 h, l;
            h.x = tf32_hi(v[i + 0]); l.x = v[i + 0] - h.x;
            h.y = tf32_hi(v[i + 1]); l.y = v[i + 1] - h.y;
            h.z = tf32_hi(v[i + 2]); l.z = v[i + 2] - h.z;
            h.w = tf32_hi(v[i + 3]); l.w = v[i + 3] - h.w;
            *reinterpret_cast<float4*>(ph + i) = h;
            *reinterpret_cast<float4*>(pl + i) = l;
          }
        } else {
          for (int i = 0; i < 16 && c0 + i < a.N; ++i) {
            const float h = tf32_hi(v[i]);
            ph[i] = h;
            pl[i] = v[i] - h;
          }
        }
      } else {
        for (int i = 0; i < 16 && c0 + i < a.N; ++i) {
          const int col = c0 + i;
          if (a.out_full) a.out_full[(size_t)row * a.ld_out + col] = v[i];
          if (a.head == NLSH_HEAD_SOFTMAX) {
            if (col == 0 || v[i] > best) {
              best = v[i];
              code = col;
            }
          } else {
            code = (code << 1) | (v[i] > thr ? 1 : 0);  // MSB first (utils.pyx:12-14)
          }
        }
      }
    }
    if constexpr (FUSED) {
      if (row_ok) {
#pragma unroll
      for (int j = 0; j < kHeadMax; ++j) {
        if (j < a.head_n) {
          const float l = logit[j] + (a.head_b != nullptr ? a.head_b[j] : 0.f);
          if (a.out_full) a.out_full[(size_t)row * a.ld_out + j] = l;
          if (a.head == NLSH_HEAD_SOFTMAX) {
            if (j == 0 || l > best) {
              best = l;
              code = j;
            }
          } else {
            code = (code << 1) | (l > thr ? 1 : 0);  // MSB first (utils.pyx:12-14)
          }
        }
      }
      }
    }
    if ((last || fused) && row_ok && a.codes_out) a.codes_out[row] = code;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

// x -> (hi, lo): hi = x with the 13 low mantissa bits cleared (a tf32 value), lo = x - hi (exact).
__global__ void split_tf32_kernel(const float* __restrict__ x, size_t n, float* __restrict__ hi,
                                  float* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t j = i; j < n; j += stride) {
    const float v = x[j];
    const float h = tf32_hi(v);
    hi[j] = h;
    lo[j] = v - h;
  }
}

}  // namespace

typedef CUresult (*TcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tc_make_map(CUtensorMap* map, const float* base, long long rows, int cols, int box_rows) {
  static TcEncodeTiledFn encode = nullptr;
  if (encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NLSH_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      nlsh_set_error("cuTensorMapEncodeTiled is not available from this driver");
      return NLSH_ERR_CUDA;
    }
    encode = reinterpret_cast<TcEncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims,
                            strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    nlsh_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%d box_rows=%d)",
                   (int)r, rows, cols, box_rows);
    return NLSH_ERR_CUDA;
  }
  return NLSH_OK;
}


// ---- interface used by hasher.cu ---------------------------------------------------------
bool nlsh_tc_layer_supported(int in_dim, int out_dim) {
  return in_dim % 4 == 0 && in_dim >= 4 && out_dim >= 1 && out_dim <= 256;
}

int nlsh_tc_split(const float* x, size_t n, float* hi, float* lo, cudaStream_t st) {
  if (n == 0) return NLSH_OK;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  split_tf32_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, hi, lo);
  return nlsh_check_cuda(nlsh_post_launch(), "split_tf32_kernel launch");
}

// One layer on the tensor cores.  a_hi / a_lo: [M, K] split input; w_hi / w_lo: [N, K] split
// weights; outputs as in TcArgs.
int nlsh_tc_linear(const float* a_hi, const float* a_lo, const float* w_hi, const float* w_lo,
                   const float* bias, int M, int N, int K, int act, float act_scale, float* out_hi,
                   float* out_lo, float* out_full, int ld_out, int* codes_out, int head,
                   cudaStream_t st, const float* head_w, const float* head_b, int head_n) {
  TcArgs a{};
  a.head_w = head_w;
  a.head_b = head_b;
  a.head_n = head_w != nullptr ? head_n : 0;
  if (head_w != nullptr && (head_n < 1 || head_n > kHeadMax)) {
    nlsh_set_error("tc_linear: fused head of %d outputs (1..%d)", head_n, kHeadMax);
    return NLSH_ERR_INVALID;
  }
  a.bias = bias;
  a.out_hi = out_hi;
  a.out_lo = out_lo;
  a.out_full = out_full;
  a.codes_out = codes_out;
  a.M = M;
  a.N = N;
  a.K = K;
  a.n_pad = (N + 15) / 16 * 16;
  a.ld_out = ld_out;
  a.act = act;
  a.act_scale = act_scale;
  a.head = head;
  a.n_main = 512 / a.n_pad - 1;
  if (a.n_main > 4) a.n_main = 4;
  if (a.n_main < 1) a.n_main = 1;
  a.tmem_cols = 32;
  while (a.tmem_cols < (a.n_main + 1) * a.n_pad) a.tmem_cols *= 2;
  const size_t stage_bytes =
      2 * (size_t)kTcBM * kTcBK * sizeof(float) + 2 * (size_t)a.n_pad * kTcBK * sizeof(float);
  const size_t head_bytes = head_w != nullptr ? (size_t)a.n_pad * kHeadMax * sizeof(float) : 0;
  a.stages = 4;
  while (a.stages > 1 && a.stages * stage_bytes + 2048 + head_bytes > 220 * 1024) --a.stages;
  const size_t smem = a.stages * stage_bytes + 2048 + head_bytes;

  CUtensorMap m_a_hi, m_a_lo, m_w_hi, m_w_lo;
  int rc;
  if ((rc = tc_make_map(&m_a_hi, a_hi, M, K, kTcBM)) != NLSH_OK) return rc;
  if ((rc = tc_make_map(&m_a_lo, a_lo, M, K, kTcBM)) != NLSH_OK) return rc;
  if ((rc = tc_make_map(&m_w_hi, w_hi, N, K, a.n_pad)) != NLSH_OK) return rc;
  if ((rc = tc_make_map(&m_w_lo, w_lo, N, K, a.n_pad)) != NLSH_OK) return rc;
  auto kern = head_w != nullptr ? tc_linear_kernel<true> : tc_linear_kernel<false>;
  NLSH_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((M + kTcBM - 1) / kTcBM);
  kern<<<grid, kTcThreads, smem, st>>>(a, m_a_hi, m_a_lo, m_w_hi, m_w_lo);
  return nlsh_check_cuda(nlsh_post_launch(), "tc_linear_kernel launch");
}
