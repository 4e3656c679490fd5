// Persistent warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   D[M,N] = op(A) * op(B)  (+ bias[N]) (GELU) (* gelu'(aux)) (+ residual) (+= D)
//
// * operands are bf16 (or fp16), accumulation is fp32 in TMEM, output bf16/fp16/fp32;
// * A and B tiles are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a kStages-deep
//   shared-memory ring guarded by full/empty mbarriers;
// * one thread issues tcgen05.mma (128 x BN x 16 per instruction, cta_group::1) and releases ring
//   slots with tcgen05.commit;
// * two TMEM accumulator stages (2 x BN columns) let the four epilogue warps drain tile i
//   (tcgen05.ld -> registers -> fused epilogue -> 16-byte global stores) while tile i+1 is being
//   multiplied;
// * persistent: grid = min(tiles, 148); tiles are visited in grouped order so concurrently
//   running CTAs share A/B panels in the 126 MB L2.
//
// Layouts: each operand is either "K-major" (the contraction index is contiguous in memory:
// A stored [M,K], B stored [N,K] — the forward x @ W^T case) or "MN-major" (A stored [K,M],
// B stored [K,N] — what the backward GEMMs dX = dY @ W and dW = dY^T @ X need), selected per
// operand through the UMMA instruction descriptor + shared-memory descriptor.
//
// The reference has no GEMM of its own (all math is stock TF/cuBLAS); this kernel
// is the compute half of the fused all-gather->GEMM / GEMM->reduce-scatter kernels (tp_fused.cu).
#include "epl_common.cuh"
#include <algorithm>
#include <cstdio>

namespace epl {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 bytes = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;    // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2..5: epilogue
constexpr int kGroupM = 8;

enum Epilogue : int { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_GELU = 2, EPI_DGELU = 3, EPI_BIAS_RESIDUAL = 4 };

struct GemmParams {
  int M, N, K;
  int ldd;                 // row stride of D / aux / pre / residual (elements)
  void* D;
  const void* bias;        // [N] (same dtype as D when 16-bit, else bf16)
  void* pre;               // EPI_BIAS_GELU: optional pre-activation output
  const void* aux;         // EPI_DGELU: pre-activation input; EPI_BIAS_RESIDUAL: residual input
  int epilogue;
  int accumulate;          // D += result
  int out_dtype;           // EPL_F32 / EPL_BF16 / EPL_F16
  int a_mn_major, b_mn_major;
  int ab_format;           // 1 = bf16, 0 = fp16
  float alpha;
};

EPL_DEVICE float gelu_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
}
EPL_DEVICE float gelu_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float t = tanhf(k0 * (x + k1 * x * x * x));
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x * x);
}

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BN * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 128) ? 6 : (BN <= 160 ? 5 : 4);
  static constexpr int kBarrierOffset = kStages * kStageBytes;
  static constexpr int kTotalBytes = kBarrierOffset + 256 + 1024;   // barriers + alignment slack
};

EPL_DEVICE void tile_coords(int tile, int m_blocks, int n_blocks, int& mb, int& nb) {
  const int per_group = kGroupM * n_blocks;
  const int group = tile / per_group;
  const int first_m = group * kGroupM;
  const int rows = min(m_blocks - first_m, kGroupM);
  const int in_group = tile - group * per_group;
  mb = first_m + in_group % rows;
  nb = in_group / rows;
}

template <typename OutT>
EPL_DEVICE void store_chunk(const GemmParams& p, int row, int col0, const float (&v)[32]) {
  OutT* drow = reinterpret_cast<OutT*>(p.D) + (size_t)row * p.ldd;
  constexpr int E = 16 / sizeof(OutT);
  const bool vec_ok = (p.ldd % E == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0);
#pragma unroll
  for (int g = 0; g < 32 / E; ++g) {
    const int c = col0 + g * E;
    if (vec_ok && c + E <= p.N) {
      Vec<OutT, E> o;
      if (p.accumulate) {
        Vec<OutT, E> old = ld_vec<OutT, E>(drow + c);
#pragma unroll
        for (int e = 0; e < E; ++e) o.v[e] = from_f32<OutT>(v[g * E + e] + to_f32<OutT>(old.v[e]));
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) o.v[e] = from_f32<OutT>(v[g * E + e]);
      }
      st_vec<OutT, E>(drow + c, o);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (c + e < p.N) {
          float r = v[g * E + e];
          if (p.accumulate) r += to_f32<OutT>(drow[c + e]);
          drow[c + e] = from_f32<OutT>(r);
        }
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  constexpr int kTmemCols = 512;                       // 2 accumulator stages of BN (<= 256) fp32 columns
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarrierOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int num_tiles = m_blocks * n_blocks;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mb, nb;
        tile_coords(tile, m_blocks, n_blocks, mb, nb);
        const int m0 = mb * BLOCK_M, n0 = nb * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * L::kStageBytes;
          unsigned char* sb = sa + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          const int k0 = kb * BLOCK_K;
          if (!p.a_mn_major) {
            tma_load_2d(sa, &map_a, &full_bar[stage], k0, m0);                    // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / 64; ++a)                                  // box {64 m, 64 k} per atom
              tma_load_2d(sa + a * (BLOCK_K * 128), &map_a, &full_bar[stage], m0 + a * 64, k0);
          }
          if (!p.b_mn_major) {
            tma_load_2d(sb, &map_b, &full_bar[stage], k0, n0);                    // box {64 k, BN n}
          } else {
#pragma unroll
            for (int a = 0; a < BN / 64; ++a)
              tma_load_2d(sb + a * (BLOCK_K * 128), &map_b, &full_bar[stage], n0 + a * 64, k0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(BLOCK_M, BN, p.ab_format, p.a_mn_major, p.b_mn_major);
      // K-major: atoms of 8 rows x 128 B, SBO = 1024 B, +32 B per UMMA_K.  MN-major: 64-wide atoms of
      // BLOCK_K rows x 128 B, LBO = BLOCK_K*128 B between atoms, SBO = 1024 B per 8 k-rows, +2048 B per UMMA_K.
      const uint32_t a_lbo = p.a_mn_major ? BLOCK_K * 128 : 16, b_lbo = p.b_mn_major ? BLOCK_K * 128 : 16;
      const uint32_t a_kstep = p.a_mn_major ? UMMA_K * 128 : UMMA_K * 2, b_kstep = p.b_mn_major ? UMMA_K * 128 : UMMA_K * 2;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
            umma_f16(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);                 // slot is free once these MMAs have read it
          if (kb == k_blocks - 1) umma_commit(&tmem_full[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue (warps 2..5) =======================
    const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mb, nb;
      tile_coords(tile, m_blocks, n_blocks, mb, nb);
      const int row = mb * BLOCK_M + quarter * 32 + lane;
      const int n0 = nb * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + c * 32, r);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        if (row < p.M && col0 < p.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          if (p.epilogue == EPI_BIAS || p.epilogue == EPI_BIAS_GELU || p.epilogue == EPI_BIAS_RESIDUAL) {
            if (p.bias != nullptr) {
              const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(p.bias) + col0;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __bfloat162float(b[j]);
            }
          }
          if (p.epilogue == EPI_BIAS_GELU) {
            if (p.pre != nullptr) {
              __nv_bfloat16* prow = reinterpret_cast<__nv_bfloat16*>(p.pre) + (size_t)row * p.ldd + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (col0 + j + 8 <= p.N && (p.ldd & 7) == 0) {
                  Vec<__nv_bfloat16, 8> o;
#pragma unroll
                  for (int e = 0; e < 8; ++e) { o.v[e] = __float2bfloat16_rn(v[j + e]); v[j + e] = __bfloat162float(o.v[e]); }
                  st_vec<__nv_bfloat16, 8>(prow + j, o);
                } else {
                  for (int e = 0; e < 8; ++e) if (col0 + j + e < p.N) {
                    __nv_bfloat16 h = __float2bfloat16_rn(v[j + e]); prow[j + e] = h; v[j + e] = __bfloat162float(h);
                  }
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
          } else if (p.epilogue == EPI_DGELU || p.epilogue == EPI_BIAS_RESIDUAL) {
            const __nv_bfloat16* arow = reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)row * p.ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float a8[8];
              if (col0 + j + 8 <= p.N && (p.ldd & 7) == 0) {
                Vec<__nv_bfloat16, 8> a = ld_vec<__nv_bfloat16, 8>(arow + j);
#pragma unroll
                for (int e = 0; e < 8; ++e) a8[e] = __bfloat162float(a.v[e]);
              } else {
                for (int e = 0; e < 8; ++e) a8[e] = (col0 + j + e < p.N) ? __bfloat162float(arow[j + e]) : 0.f;
              }
#pragma unroll
              for (int e = 0; e < 8; ++e)
                v[j + e] = (p.epilogue == EPI_DGELU) ? v[j + e] * gelu_grad_f(a8[e]) : v[j + e] + a8[e];
            }
          }
          if (p.out_dtype == EPL_BF16) store_chunk<__nv_bfloat16>(p, row, col0, v);
          else if (p.out_dtype == EPL_F32) store_chunk<float>(p, row, col0, v);
          else store_chunk<__half>(p, row, col0, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || sym == nullptr) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] (cols contiguous, row stride ld elements), box {box_cols, box_rows}, 128B swizzle
static int make_map_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                       uint32_t box_rows, int is_fp16) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -10;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, is_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

template <int BN>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, int num_sms, cudaStream_t st) {
  using L = SmemLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotalBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M, n_blocks = (p.N + BN - 1) / BN;
  const int grid = std::min(m_blocks * n_blocks, num_sms);
  gemm_tcgen05_kernel<BN><<<grid, kGemmThreads, L::kTotalBytes, st>>>(ma, mb, p);
  return EPL_CHECK_LAUNCH();
}

static int pick_bn(int N, int b_mn_major, int forced) {
  if (forced == 128 || forced == 256 || (forced == 160 && !b_mn_major)) return forced;
  // measured on B200 (tools/gemm_bench.py): the 128x256 tile wins whenever N spans more than one narrow tile,
  // even with 12% padding (N = 1600): fewer A re-reads and the lowest shared-memory bytes per MMA cycle.
  if (N <= 128) return 128;
  if (N <= 160 && !b_mn_major) return 160;
  return 256;
}

}  // namespace epl
using namespace epl;

// A: K-major  -> memory [M, K] (lda = row stride);  MN-major -> memory [K, M].
// B: K-major  -> memory [N, K] (ldb);               MN-major -> memory [K, N].
// Alignment: base pointers 16 B aligned, lda/ldb multiples of 8 elements.
extern "C" int epl_gemm(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                        int a_mn_major, int b_mn_major, const void* bias, void* pre, const void* aux, int epilogue,
                        int accumulate, int out_dtype, float alpha, int is_fp16, int force_bn, int num_sms, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int bn = pick_bn(N, b_mn_major, force_bn);
  CUtensorMap ma, mb;
  int rc;
  if (!a_mn_major) rc = make_map_2d(&ma, A, M, K, lda, BLOCK_K, BLOCK_M, is_fp16);
  else rc = make_map_2d(&ma, A, K, M, lda, 64, BLOCK_K, is_fp16);
  if (rc) return rc;
  if (!b_mn_major) rc = make_map_2d(&mb, B, N, K, ldb, BLOCK_K, bn, is_fp16);
  else rc = make_map_2d(&mb, B, K, N, ldb, 64, BLOCK_K, is_fp16);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.D = D; p.bias = bias; p.pre = pre; p.aux = aux; p.epilogue = epilogue;
  p.accumulate = accumulate; p.out_dtype = out_dtype; p.a_mn_major = a_mn_major; p.b_mn_major = b_mn_major; p.alpha = alpha; p.ab_format = is_fp16 ? 0 : 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_sms <= 0) num_sms = kNumSMs;
  if (bn == 256) return launch_gemm<256>(ma, mb, p, num_sms, st);
  if (bn == 160) return launch_gemm<160>(ma, mb, p, num_sms, st);
  return launch_gemm<128>(ma, mb, p, num_sms, st);
}
