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
// Two kernels share this file:
//   gemm_tcgen05_kernel<BN, comm>  1-CTA tiles (described above); also the compute half of the fused collective modes
//                                  COMM_AG (all-gather -> GEMM), COMM_RS (GEMM -> reduce-scatter), COMM_AGB (weight gather);
//   gemm2_tcgen05_kernel           the default for M, N >= 256: CTA pairs (cta_group::2), 256 x {256,192,128} tiles,
//                                  6-stage ring, 8 epilogue warps, aux-operand prefetch, red.add accumulation, dynamic
//                                  (atomic-counter) tile scheduler, narrow MMA on the ragged last column tile.
//
// The reference has no GEMM of its own (all math is stock TF/cuBLAS).
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
  int bn2;                 // 2-CTA kernel: tile width (128 / 192 / 256), chosen per shape by pick_bn2()
  uint32_t* sched_counter; // 2-CTA kernel: global tile counter of the dynamic scheduler (one per stream, self-resetting)
  int fp8;                 // 2-CTA kernel: A and B are e4m3 (1 byte), K-major; tcgen05.mma kind::f8f6f4, 128 K-elements per k-block
  const float* scale_a;    // fp8: device scalars, the de-quantisation factors of A and B (epilogue multiplies the accumulator)
  const float* scale_b;
};

// Fused collective (tensor-parallel) state.  mode 1: all-gather -> GEMM, mode 2: GEMM -> reduce-scatter.
constexpr int kMaxPeers = 8;
enum CommMode : int { COMM_NONE = 0, COMM_AG = 1, COMM_RS = 2, COMM_AGB = 3 };   // AGB: the gathered operand is B (ZeRO-3 weights)
struct CommParams {
  int rank, world;
  uint32_t epoch;              // strictly increasing per launch on this workspace; 0 = read it from local_sync[31] (+1): the launch
                               // then carries no per-call host value and can be replayed from a CUDA graph (a one-thread
                               // kernel launched right after bumps the stored epoch, see epl_gemm_fused)
  int copy_ctas;               // mode 1: trailing CTAs of the grid that run the NVLink copy role
  int rows_per_rank;           // rows of the gathered operand owned by each rank (M / world, or N / world for COMM_AGB)
  uint32_t* flags[kMaxPeers];  // symmetric signal pad of every rank: [0,8) start slots, [8,16) end slots
  uint32_t* local_sync;        // device-local words: [0] start-go, [1] arrival counter, [2] end-go, [8+r] chunk-r arrivals
  const void* ag_src[kMaxPeers];   // mode 1: every rank's shard [rows_per_rank, K] (contiguous rows)
  void* ag_dst;                    // mode 1: local gathered A [M, K] (the buffer map_a describes)
  void* rs_stage[kMaxPeers];       // mode 2: every rank's staging [world, rows_per_rank, ldd]; this rank writes slot `rank`
  void* rs_out;                    // mode 2: local reduced output [rows_per_rank, ldd]
};

// tanh-approximation GELU on the MUFU tanh unit (abs error ~2^-11, below bf16 resolution): 6 / 9 FMA-pipe instructions
// per element.  tanhf() costs ~25 and made the fused GELU epilogues slower than the main loop they hide behind.
EPL_DEVICE float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
EPL_DEVICE float gelu_f(float x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, k01, k0));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
EPL_DEVICE float gelu_grad_f(float x) {
  const float k0 = 0.7978845608028654f, k01 = 0.7978845608028654f * 0.044715f;
  const float x2 = x * x;
  const float t = tanh_approx(x * fmaf(x2, k01, k0));
  const float du = fmaf(x2, 3.f * k01, k0);
  const float sech2 = fmaf(-t, t, 1.f);
  return fmaf(0.5f * x * sech2, du, fmaf(0.5f, t, 0.5f));
}

template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BN * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 128) ? 6 : (BN <= 160 ? 5 : 4);
  static constexpr int kBarrierOffset = kStages * kStageBytes;
  static constexpr int kStagingOffset = kBarrierOffset + 256;       // 4 epilogue warps x 4 KB (reduce-scatter epilogue transposes here)
  static constexpr int kTotalBytes = kStagingOffset + 16384 + 1024; // + alignment slack
};

// weight-gather order: all m-blocks of one n-block before the next n-block, starting with the n-blocks whose weights
// are local; the A panel (activations) stays L2-resident across n-blocks
EPL_DEVICE void tile_coords_n_outer(int tile, int m_blocks, int n_blocks, int rank, int world, int& mb, int& nb) {
  const int per = max(n_blocks / world, 1);
  nb = (tile / m_blocks + rank * per) % n_blocks;
  mb = tile % m_blocks;
}

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
EPL_DEVICE void store_chunk(const GemmParams& p, void* row_ptr, int col0, const float (&v)[32]) {
  OutT* drow = reinterpret_cast<OutT*>(row_ptr);
  constexpr int E = 16 / sizeof(OutT);
  const bool vec_ok = (p.ldd % E == 0) && ((reinterpret_cast<uintptr_t>(row_ptr) & 15) == 0);
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

EPL_DEVICE uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
EPL_DEVICE void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
EPL_DEVICE void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
EPL_DEVICE void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// threads [0, world) of the calling CTA: tell every peer "slot[rank] = epoch", wait until every peer told us
EPL_DEVICE void cross_gpu_signal_wait(const CommParams& c, int slot_base) {
  if ((int)threadIdx.x < c.world) {
    __threadfence_system();
    st_release_sys(c.flags[threadIdx.x] + slot_base + c.rank, c.epoch);
    while (ld_acquire_sys(c.flags[c.rank] + slot_base + threadIdx.x) < c.epoch) {}
  }
}

// m-block visiting order: all-gather starts with the local rows (already here), reduce-scatter ends with them
template <int kComm>
EPL_DEVICE int rotate_mb(int mb, int m_blocks, const CommParams& c) {
  if constexpr (kComm == COMM_NONE || kComm == COMM_AGB) return mb;
  const int per = max(m_blocks / c.world, 1);
  const int shift = (kComm == COMM_AG ? c.rank : c.rank + 1) * per;
  return (mb + shift) % m_blocks;
}

// ---- copy role of the all-gather -> GEMM kernel: pull every rank's shard into the local gathered buffer ----------
EPL_DEVICE void ag_copy_role(const GemmParams& p, const CommParams& c, int gemm_ctas, unsigned char* copy_smem) {
  const int cid = blockIdx.x - gemm_ctas;
  if (cid == 0) {
    cross_gpu_signal_wait(c, 0);                       // every rank's shard is ready to be read
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); st_release_gpu(c.local_sync, c.epoch); }
  } else if (threadIdx.x == 0) {
    while (ld_acquire_gpu(c.local_sync) < c.epoch) {}
  }
  __syncthreads();
  // Bulk-DMA ring: one thread streams [peer global -> shared -> local global] with cp.async.bulk; kCopySlots x 32 KB are in
  // flight per copy CTA, so a handful of CTAs keeps ~2.5 MB outstanding — enough to cover NVLink latency at full bandwidth
  // without spending registers or issue slots on the copy (the 20 copy CTAs would otherwise top out near 150 GB/s).
  constexpr int kCopySlots = 5;                           // 4 loads in flight + 1 slot draining (160 KB fits every tile config)
  constexpr uint32_t kPiece = 32768;
  uint64_t* cbar = reinterpret_cast<uint64_t*>(copy_smem + kCopySlots * kPiece);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kCopySlots; ++i) mbar_init(&cbar[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t chunk_bytes = (size_t)c.rows_per_rank * p.K * 2;
    size_t per_cta = (chunk_bytes / c.copy_ctas + 15) & ~(size_t)15;
    const size_t b0 = min((size_t)cid * per_cta, chunk_bytes), b1 = min(b0 + per_cta, chunk_bytes);
    const int pieces = (int)((b1 - b0 + kPiece - 1) / kPiece);
    uint32_t it = 0;                                      // global piece counter -> ring slot + phase
    for (int step = 0; step < c.world; ++step) {
      const int src = (c.rank + step) % c.world;          // local chunk first, then ring order
      const unsigned char* from = reinterpret_cast<const unsigned char*>(c.ag_src[src]) + b0;
      unsigned char* to = reinterpret_cast<unsigned char*>(c.ag_dst) + (size_t)src * chunk_bytes + b0;
      for (int q = 0; q < pieces + kCopySlots - 2; ++q) {
        if (q < pieces) {                                 // issue load q
          const uint32_t slot = (it + q) % kCopySlots;
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the slot's previous store (two groups back) has drained
          const uint32_t bytes = (uint32_t)min((size_t)kPiece, (b1 - b0) - (size_t)q * kPiece);
          mbar_expect_tx(&cbar[slot], bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       :: "r"(smem_u32(copy_smem + slot * kPiece)), "l"(from + (size_t)q * kPiece), "r"(bytes), "r"(smem_u32(&cbar[slot])) : "memory");
        }
        const int d = q - (kCopySlots - 2);               // store piece d once its load has landed
        if (d >= 0) {
          const uint32_t slot = (it + d) % kCopySlots, ph = ((it + d) / kCopySlots) & 1;
          mbar_wait(&cbar[slot], ph);
          const uint32_t bytes = (uint32_t)min((size_t)kPiece, (b1 - b0) - (size_t)d * kPiece);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       :: "l"(to + (size_t)d * kPiece), "r"(smem_u32(copy_smem + slot * kPiece)), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      it += pieces;
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // this source's rows are in local memory
      fence_proxy_async_all();
      __threadfence();
      red_release_gpu_add(c.local_sync + 8 + src, 1u);
    }
  }
  __syncthreads();
  // end barrier: nobody may overwrite its shard before every peer has finished reading it
  __shared__ int last_copy;
  if (threadIdx.x == 0) last_copy = (atomicAdd(c.local_sync + 1, 1u) == (uint32_t)c.copy_ctas * c.epoch - 1u);
  __syncthreads();
  if (last_copy) cross_gpu_signal_wait(c, 8);
}

template <int BN, int kComm>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p, const CommParams c_in) {
  CommParams c = c_in;
  if constexpr (kComm != COMM_NONE) {
    if (c.epoch == 0) c.epoch = *reinterpret_cast<volatile uint32_t*>(c.local_sync + 31) + 1u;
  }
  using L = SmemLayout<BN>;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int gemm_ctas = (kComm == COMM_AG || kComm == COMM_AGB) ? (int)gridDim.x - c.copy_ctas : (int)gridDim.x;
  if constexpr (kComm == COMM_AG || kComm == COMM_AGB) {
    if ((int)blockIdx.x >= gemm_ctas) { ag_copy_role(p, c, gemm_ctas, smem); return; }
  }
  constexpr int kStages = L::kStages;
  constexpr int kTmemCols = 512;                       // 2 accumulator stages of BN (<= 256) fp32 columns
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

  if constexpr (kComm == COMM_RS) {
    // every rank's staging buffer is free again (its previous reduction has completed): open the gate for stores
    if (blockIdx.x == 0 && warp == 0) {
      cross_gpu_signal_wait(c, 0);
      __syncwarp();
      if (lane == 0) { __threadfence(); st_release_gpu(c.local_sync, c.epoch); }
    }
  }
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gemm_ctas) {
        int mb, nb;
        if constexpr (kComm == COMM_AGB) tile_coords_n_outer(tile, m_blocks, n_blocks, c.rank, c.world, mb, nb);
        else tile_coords(tile, m_blocks, n_blocks, mb, nb);
        mb = rotate_mb<kComm>(mb, m_blocks, c);
        const int m0 = mb * BLOCK_M, n0 = nb * BN;
        if constexpr (kComm == COMM_AGB) {             // the gathered weight rows of this tile must have landed
          const int s_lo = n0 / c.rows_per_rank, s_hi = min(n0 + BN - 1, p.N - 1) / c.rows_per_rank;
          for (int sr = s_lo; sr <= s_hi; ++sr)
            while (ld_acquire_gpu(c.local_sync + 8 + sr) < (uint32_t)c.copy_ctas * c.epoch) {}
          fence_proxy_async_all();
        }
        if constexpr (kComm == COMM_AG) {              // the gathered rows of this tile must have landed
          const int s_lo = m0 / c.rows_per_rank, s_hi = min(m0 + BLOCK_M - 1, p.M - 1) / c.rows_per_rank;
          for (int sr = s_lo; sr <= s_hi; ++sr)
            while (ld_acquire_gpu(c.local_sync + 8 + sr) < (uint32_t)c.copy_ctas * c.epoch) {}
          fence_proxy_async_all();
        }
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gemm_ctas) {
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
    const CommParams& cm = c;                             // (the chunk loop below reuses the name `c`)
    (void)cm;
    int acc = 0; uint32_t acc_phase = 0;
    bool rs_go = false;
    (void)rs_go;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gemm_ctas) {
      int mb, nb;
      if constexpr (kComm == COMM_AGB) tile_coords_n_outer(tile, m_blocks, n_blocks, c.rank, c.world, mb, nb);
      else tile_coords(tile, m_blocks, n_blocks, mb, nb);
      mb = rotate_mb<kComm>(mb, m_blocks, c);
      const int row = mb * BLOCK_M + quarter * 32 + lane;
      const int n0 = nb * BN;
      unsigned char* drow = reinterpret_cast<unsigned char*>(p.D);
      const size_t out_es = p.out_dtype == EPL_F32 ? 4 : 2;
      if constexpr (kComm == COMM_RS) {                // the tile goes to the rank that owns these rows
        if (!rs_go) {
          if (lane == 0) while (ld_acquire_gpu(c.local_sync) < c.epoch) {}
          __syncwarp();
          rs_go = true;
        }
        const int owner = min(row / c.rows_per_rank, c.world - 1);
        drow = reinterpret_cast<unsigned char*>(c.rs_stage[owner]) +
               ((size_t)c.rank * c.rows_per_rank + (row - owner * c.rows_per_rank)) * p.ldd * out_es;
      } else {
        drow += (size_t)row * p.ldd * out_es;
      }
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
          if constexpr (kComm != COMM_RS) {
            if (p.out_dtype == EPL_BF16) store_chunk<__nv_bfloat16>(p, drow, col0, v);
            else if (p.out_dtype == EPL_F32) store_chunk<float>(p, drow, col0, v);
            else store_chunk<__half>(p, drow, col0, v);
          } else {
            // stage this row's 32 values (64 B) in the warp's shared-memory tile [32 rows x 128 B], 16-byte chunks XOR-swizzled
            unsigned char* wrow = smem + L::kStagingOffset + (warp - 2) * 4096 + lane * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 w;
              w.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); w.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
              w.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); w.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
              *reinterpret_cast<uint4*>(wrow + ((((c & 1) * 4 + g) ^ (lane & 7)) << 4)) = w;
            }
          }
        }
        if constexpr (kComm == COMM_RS) {
          // every second chunk (or the last one): the warp writes its 32 x 128 B tile with full 128-byte lines per row,
          // 4 rows per instruction — NVLink sees whole cache lines instead of 32 scattered 16-byte packets
          if ((c & 1) || c == BN / 32 - 1) {
            __syncwarp();
            const int cbase = n0 + (c & ~1) * 32;
            const int row0 = mb * BLOCK_M + quarter * 32;
            const unsigned char* wtile = smem + L::kStagingOffset + (warp - 2) * 4096;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int ri = it * 4 + (lane >> 3), ch = lane & 7;
              const int grow = row0 + ri, col = cbase + ch * 8;
              const bool have = (ch < 4) || (c & 1);                       // second half exists only after an odd chunk
              if (have && grow < p.M && col + 8 <= p.N) {
                const uint4 w = *reinterpret_cast<const uint4*>(wtile + ri * 128 + ((ch ^ (ri & 7)) << 4));
                const int owner = min(grow / cm.rows_per_rank, cm.world - 1);
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(cm.rs_stage[owner]) +
                                     ((size_t)cm.rank * cm.rows_per_rank + (grow - owner * cm.rows_per_rank)) * p.ldd + col;
                *reinterpret_cast<uint4*>(dst) = w;
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (kComm == COMM_RS) {
      // ---- all of this CTA's tiles are stored; the last CTA of the grid runs the cross-GPU barrier ------------------
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2) {
        int is_last = 0;
        if (lane == 0) {
          __threadfence_system();
          is_last = (atomicAdd(c.local_sync + 1, 1u) == (uint32_t)gemm_ctas * c.epoch - 1u);
        }
        is_last = __shfl_sync(0xffffffffu, is_last, 0);
        if (is_last) {
          if (lane < c.world) {
            __threadfence_system();
            st_release_sys(c.flags[lane] + 8 + c.rank, c.epoch);          // "my partial tiles are in your staging buffer"
            while (ld_acquire_sys(c.flags[c.rank] + 8 + lane) < c.epoch) {}
          }
          __syncwarp();
          if (lane == 0) { __threadfence(); st_release_gpu(c.local_sync + 2, c.epoch); }
        }
        if (lane == 0) while (ld_acquire_gpu(c.local_sync + 2) < c.epoch) {}
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- local reduction over the `world` staging slots -> rs_out (bf16), 8 columns per thread ----------------------
      const int et = (warp - 2) * 32 + lane;                              // 0..127
      const int vec_per_row = p.N / 8;
      const __nv_bfloat16* stage = reinterpret_cast<const __nv_bfloat16*>(c.rs_stage[c.rank]);
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(c.rs_out);
      const long total = (long)c.rows_per_rank * vec_per_row;
      for (long i = (long)blockIdx.x * 128 + et; i < total; i += (long)gemm_ctas * 128) {
        const int r = (int)(i / vec_per_row), cv = (int)(i % vec_per_row);
        float accv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < c.world; ++w) {
          Vec<__nv_bfloat16, 8> part = ld_vec<__nv_bfloat16, 8>(stage + ((size_t)w * c.rows_per_rank + r) * p.ldd + cv * 8);
#pragma unroll
          for (int e = 0; e < 8; ++e) accv[e] += __bfloat162float(part.v[e]);
        }
        Vec<__nv_bfloat16, 8> o;
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = __float2bfloat16_rn(accv[e]);
        st_vec<__nv_bfloat16, 8>(out + (size_t)r * p.ldd + cv * 8, o);
      }
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
  CUresult r;
  EPL_ENCODE_RETRY(r, ptr, enc(map, is_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  if (r != CUDA_SUCCESS)
    fprintf(stderr, "[epl] cuTensorMapEncodeTiled failed (CUresult %d): ptr %p rows %llu cols %llu ld %llu box %u x %u\n", (int)r, ptr,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
  return r == CUDA_SUCCESS ? 0 : -11;
}

template <int BN, int kComm>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& p, const CommParams& c, int num_sms,
                       cudaStream_t st) {
  using L = SmemLayout<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, kComm>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotalBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int m_blocks = (p.M + BLOCK_M - 1) / BLOCK_M, n_blocks = (p.N + BN - 1) / BN;
  int grid;
  if (kComm == COMM_NONE) grid = std::min(m_blocks * n_blocks, num_sms);
  else grid = num_sms;       // fused kernels: the whole grid must be co-resident (CTAs spin on flags), counters assume a fixed size
  gemm_tcgen05_kernel<BN, kComm><<<grid, kGemmThreads, L::kTotalBytes, st>>>(ma, mb, p, c);
  return EPL_CHECK_LAUNCH();
}


// =================================================================================================================
// 2-CTA variant: a CTA pair (cluster of 2, same TPC) computes one 256 x 256 tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows), so shared-memory fill traffic
// per SM drops from 48 KB to 32 KB per k-block and B is read by the tensor cores from both SMs' shared memory.
// The leader CTA (rank 0) issues the MMAs; completion is multicast to the barriers of both CTAs; both CTAs run their own
// TMA producer and their own epilogue (each owns 128 accumulator rows in its own TMEM).
// =================================================================================================================
constexpr int kStages2 = 6;                         // 6 x (16 KB A + 16 KB B-half) = 192 KB
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address with the CTA-rank bit cleared -> leader CTA

EPL_DEVICE uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
EPL_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int kCols> EPL_DEVICE void tmem_alloc2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_result)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols> EPL_DEVICE void tmem_dealloc2(uint32_t tmem_addr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_addr), "n"(kCols) : "memory");
}
EPL_DEVICE void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// same tile, e4m3 / e5m2 operands (32 K-elements = 32 bytes per instruction: twice the math per shared-memory byte of bf16)
EPL_DEVICE void umma_f8_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (when all previously issued MMAs are complete) on the barrier at this smem offset in BOTH CTAs of the pair
EPL_DEVICE void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// TMA load into this CTA's shared memory, completion bytes reported to the LEADER CTA's barrier at the same offset
EPL_DEVICE void tma_load_2d_2cta(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
EPL_DEVICE void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

constexpr int kGemm2Threads = 320;    // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue (two per TMEM lane quarter, 128 columns each)

EPL_DEVICE uint4 ld_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
EPL_DEVICE void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 32 accumulator columns of one output row -> epilogue math -> global.  `ax` = the matching 32 bf16 of the aux operand
// (pre-activation for EPI_DGELU, residual for EPI_BIAS_RESIDUAL), prefetched by the caller BEFORE it waited for the
// accumulator so the load latency hides behind the main loop; kFast = the whole chunk is in range and rows are 16-byte
// aligned (everything except the ragged last N tile).
template <bool kFast>
EPL_DEVICE void epilogue_chunk32(const GemmParams& p, int row, int col0, const uint32_t (&r)[32], const uint4 (&ax)[4]) {
  float v[32];
  float alpha = p.alpha;
  if (p.scale_a != nullptr) alpha *= __ldg(p.scale_a) * __ldg(p.scale_b);     // fp8: per-tensor de-quantisation factors
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
  if (p.bias != nullptr && (p.epilogue == EPI_BIAS || p.epilogue == EPI_BIAS_GELU || p.epilogue == EPI_BIAS_RESIDUAL)) {
    const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(p.bias) + col0;
    if (kFast) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 bv = *reinterpret_cast<const uint4*>(b + g * 8);          // warp-uniform address: one broadcast load
        const uint32_t w[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16x2(w[e]); v[g * 8 + 2 * e] += f.x; v[g * 8 + 2 * e + 1] += f.y; }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __bfloat162float(b[j]);
    }
  }
  if (p.epilogue == EPI_BIAS_GELU) {
    if (p.pre != nullptr) {
      __nv_bfloat16* prow = reinterpret_cast<__nv_bfloat16*>(p.pre) + (size_t)row * p.ldd + col0;
      if (kFast) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
          o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + g * 8) = o;
          const uint32_t w[4] = {o.x, o.y, o.z, o.w};                        // GELU sees exactly the stored (rounded) pre-activation
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16x2(w[e]); v[g * 8 + 2 * e] = f.x; v[g * 8 + 2 * e + 1] = f.y; }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) {
          const __nv_bfloat16 h = __float2bfloat16_rn(v[j]); prow[j] = h; v[j] = __bfloat162float(h);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_f(v[j]);
  } else if (p.epilogue == EPI_DGELU || p.epilogue == EPI_BIAS_RESIDUAL) {
    float a[32];
    if (kFast) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t w[4] = {ax[g].x, ax[g].y, ax[g].z, ax[g].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = unpack_bf16x2(w[e]); a[g * 8 + 2 * e] = f.x; a[g * 8 + 2 * e + 1] = f.y; }
      }
    } else {
      const __nv_bfloat16* arow = reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)row * p.ldd + col0;
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = (col0 + j < p.N) ? __bfloat162float(arow[j]) : 0.f;
    }
    if (p.epilogue == EPI_DGELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= gelu_grad_f(a[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += a[j];
    }
  }
  if (kFast && p.out_dtype == EPL_BF16 && !p.accumulate) {
    __nv_bfloat16* drow = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 o;
      o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
      o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
      *reinterpret_cast<uint4*>(drow + g * 8) = o;
    }
  } else if (kFast && p.out_dtype == EPL_F32 && p.accumulate && (p.ldd & 3) == 0) {
    // weight-gradient accumulation into the fp32 main-grad buffer: fire-and-forget vector reductions, no read-modify-write
    // round trip in the epilogue (each element receives exactly one contribution per GEMM, so the result is deterministic)
    float* drow = reinterpret_cast<float*>(p.D) + (size_t)row * p.ldd + col0;
#pragma unroll
    for (int g = 0; g < 8; ++g) red_add_v4_f32(drow + g * 4, v[g * 4 + 0], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
  } else {
    const size_t out_es = p.out_dtype == EPL_F32 ? 4 : 2;
    unsigned char* drow = reinterpret_cast<unsigned char*>(p.D) + (size_t)row * p.ldd * out_es;
    if (p.out_dtype == EPL_BF16) store_chunk<__nv_bfloat16>(p, drow, col0, v);
    else if (p.out_dtype == EPL_F32) store_chunk<float>(p, drow, col0, v);
    else store_chunk<__half>(p, drow, col0, v);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Dynamic tile scheduler of the 2-CTA kernel.  Tiles are not pre-assigned (tile = cluster_id + i * num_clusters): the
// leader CTA's producer thread draws the next tile index from a global counter (atomicAdd) and publishes it through a
// 4-deep shared-memory ring to every role of BOTH CTAs of the pair (its own MMA / epilogue warps through local shared
// memory, the peer CTA's producer / epilogue warps through distributed shared memory + a remote mbarrier arrive).
// Why: (1) a concurrently running collective kernel (the fused gradient reduce-scatter + AdamW + all-gather of
// csrc/symm.cu, launched per bucket while backward is still running) holds a few SMs; with static assignment the CTA
// pairs that cannot be scheduled start late and then still walk their whole tile list (measured: 122.8 vs 117.7 ms/step),
// with a counter the running pairs simply absorb the work; (2) ragged last-column tiles are cheaper (narrow MMA, see
// below) and a counter balances uneven tiles for free.
// The counter resets itself: a launch performs exactly num_tiles + num_clusters draws (every pair ends on one terminal
// draw), so the pair that receives the value num_tiles + num_clusters - 1 knows it is the last and stores 0.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kSched = 4;

EPL_DEVICE uint32_t mapa_u32(uint32_t cta_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(cta_rank));
  return r;
}
EPL_DEVICE void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(cluster_addr), "r"(v) : "memory");
}
EPL_DEVICE void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
EPL_DEVICE void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

struct TileRing {
  uint64_t* full;        // [kSched] per CTA: the slot holds a tile index
  uint64_t* empty;       // [kSched] leader CTA only: every consumer of both CTAs has read the slot
  volatile int* tile;    // [kSched] per CTA
};

// consumer side (every role except the leader's producer): next tile index, slot handed back to the leader
EPL_DEVICE int ring_next(const TileRing& r, uint32_t& it, bool whole_warp, int lane) {
  const uint32_t slot = it & (kSched - 1), par = (it / kSched) & 1;
  mbar_wait_cluster(&r.full[slot], par);
  const int t = r.tile[slot];
  if (whole_warp) __syncwarp();
  if (!whole_warp || lane == 0) mbar_arrive_cluster_release(mapa_u32(smem_u32(&r.empty[slot]), 0));
  ++it;
  return t;
}

// Tile order of the 2-CTA kernel: all full-width tiles first (grouped for L2 reuse), the cheap ragged last-column tiles at the
// very end — with the atomic-counter scheduler this is longest-processing-time-first: the last round of a launch is filled with
// quarter-cost tiles instead of ending on a full one (N = 1600: 192 full + 32 ragged tiles on 74 CTA pairs).
EPL_DEVICE void tile_coords2(int tile, int m_blocks, int n_blocks, bool ragged, int& mb, int& nb) {
  const int n_full = ragged ? n_blocks - 1 : n_blocks;
  const int full_tiles = m_blocks * n_full;
  if (tile < full_tiles) { tile_coords(tile, m_blocks, n_full, mb, nb); return; }
  mb = tile - full_tiles;
  nb = n_full;
}

// width of the MMA for the n-block at column n0: the ragged last column tile multiplies only the columns that exist
// (rounded up to 32 so that each CTA of the pair holds a multiple of 16) instead of a full BN-wide tile
EPL_DEVICE int tile_n_eff(int N, int n0, int BN) { return min(BN, (N - n0 + 31) & ~31); }

// Epilogue of the cta_group::2 kernel: each CTA drains its 128 rows x BN columns with 8 warps (two per TMEM lane quarter).
EPL_DEVICE void gemm2_epilogue(const GemmParams& p, uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty, const TileRing& ring,
                               int cta, int warp, int lane, int m_blocks, int n_blocks) {
  constexpr int BM2 = 256, kMaxBN = 256;
  const int BN = p.bn2;
  const int num_tiles = m_blocks * n_blocks;
  const int quarter = warp & 3, half = (warp - 2) >> 2;
  const bool need_aux = p.epilogue == EPI_DGELU || p.epilogue == EPI_BIAS_RESIDUAL;
  const bool rows_aligned = (p.ldd & 7) == 0;
  const int half_cols = BN / 2;                                // 64 / 96 / 128 columns per epilogue warp
  const int nchunks = half_cols / 32;
  int acc = 0; uint32_t acc_phase = 0;
  uint32_t it = 0;
  for (;;) {
    const int tile = ring_next(ring, it, true, lane);
    if (tile >= num_tiles) break;
    int mb, nb;
    tile_coords2(tile, m_blocks, n_blocks, (p.N % BN) != 0 && n_blocks > 1, mb, nb);
    const int row = mb * BM2 + cta * BLOCK_M + quarter * 32 + lane;
    const int n0 = nb * BN + half * half_cols;                 // this warp's columns
    const bool row_ok = row < p.M;
    const bool pf = need_aux && rows_aligned && row_ok;
    const __nv_bfloat16* arow = reinterpret_cast<const __nv_bfloat16*>(p.aux) + (size_t)row * p.ldd + n0;
    // aux for the first 64 columns is requested before the accumulator wait: its latency hides behind the main loop
    uint4 ax0[8], ax1[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) ax0[g] = (pf && n0 + g * 8 + 8 <= p.N) ? ld_nc_v4(arow + g * 8) : make_uint4(0, 0, 0, 0);
    mbar_wait(&tmem_full[acc], acc_phase);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kMaxBN + half * half_cols;
    uint32_t r0[32], r1[32];
    // ---- columns [0, 64) ----
    tmem_ld_32x32(taddr, r0);
    tmem_ld_32x32(taddr + 32, r1);
#pragma unroll
    for (int g = 0; g < 8; ++g)
      ax1[g] = (pf && 64 + g * 8 < half_cols && n0 + 64 + g * 8 + 8 <= p.N) ? ld_nc_v4(arow + 64 + g * 8) : make_uint4(0, 0, 0, 0);
    tmem_ld_wait();
    if (nchunks == 2) {                                        // accumulator stage drained into registers: hand it back early
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
    }
    if (row_ok) {
      const uint4 (&a0)[4] = *reinterpret_cast<const uint4 (*)[4]>(&ax0[0]);
      const uint4 (&a1)[4] = *reinterpret_cast<const uint4 (*)[4]>(&ax0[4]);
      if (rows_aligned && n0 + 32 <= p.N) epilogue_chunk32<true>(p, row, n0, r0, a0);
      else if (n0 < p.N) epilogue_chunk32<false>(p, row, n0, r0, a0);
      if (rows_aligned && n0 + 64 <= p.N) epilogue_chunk32<true>(p, row, n0 + 32, r1, a1);
      else if (n0 + 32 < p.N) epilogue_chunk32<false>(p, row, n0 + 32, r1, a1);
    }
    if (nchunks > 2) {
      // ---- columns [64, 96) or [64, 128) ----
      tmem_ld_32x32(taddr + 64, r0);
      if (nchunks > 3) tmem_ld_32x32(taddr + 96, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (row_ok) {
        const uint4 (&a0)[4] = *reinterpret_cast<const uint4 (*)[4]>(&ax1[0]);
        const uint4 (&a1)[4] = *reinterpret_cast<const uint4 (*)[4]>(&ax1[4]);
        if (rows_aligned && n0 + 96 <= p.N) epilogue_chunk32<true>(p, row, n0 + 64, r0, a0);
        else if (n0 + 64 < p.N) epilogue_chunk32<false>(p, row, n0 + 64, r0, a0);
        if (nchunks > 3) {
          if (rows_aligned && n0 + 128 <= p.N) epilogue_chunk32<true>(p, row, n0 + 96, r1, a1);
          else if (n0 + 96 < p.N) epilogue_chunk32<false>(p, row, n0 + 96, r1, a1);
        }
      }
    }
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const GemmParams p) {
  // The tile is 256 rows (128 per CTA) x BN columns; BN is a runtime value (128 / 192 / 256) so one kernel serves every
  // width pick_bn2() selects.  Shared-memory stages and TMEM accumulator stages keep the 256-wide stride.
  constexpr int BM2 = 256, kMaxBN = 256;
  constexpr int kABytes = BLOCK_M * BLOCK_K * 2, kBBytes = (kMaxBN / 2) * BLOCK_K * 2, kStageBytes = kABytes + kBBytes;
  const int BN = p.bn2;
  // bytes one CTA's producer puts on the stage barrier: A half + B half (MN-major B is fetched in whole 64-column atoms)
  const uint32_t b_boxes = (uint32_t)((BN / 2 + 63) / 64);
  const uint32_t tx_bytes = kABytes + (p.b_mn_major ? b_boxes * (BLOCK_K * 128) : (uint32_t)(BN / 2) * BLOCK_K * 2);
  constexpr int kTmemCols = 512;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tmem_full = empty_bar + kStages2;
  uint64_t* tmem_empty = tmem_full + 2;
  TileRing ring;
  ring.full = tmem_empty + 2;
  ring.empty = ring.full + kSched;
  ring.tile = reinterpret_cast<volatile int*>(ring.empty + kSched);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(ring.tile) + kSched);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int m_blocks = (p.M + BM2 - 1) / BM2;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int num_tiles = m_blocks * n_blocks;
  const int block_k = p.fp8 ? 2 * BLOCK_K : BLOCK_K;        // elements per 128-byte k-block
  const int k_blocks = (p.K + block_k - 1) / block_k;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < kStages2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 16); }   // 8 epilogue warps x 2 CTAs
    // ring: 1 publisher; consumers = peer producer + MMA thread + 16 epilogue warps
    for (int s = 0; s < kSched; ++s) { mbar_init(&ring.full[s], 1); mbar_init(&ring.empty[s], 18); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2<kTmemCols>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs); the leader's is also the tile scheduler ==============
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      int tile = leader ? (int)atomicAdd(p.sched_counter, 1u) : 0;
      for (;;) {
        int next = 0;
        if (leader) {
          const uint32_t slot = it & (kSched - 1), par = (it / kSched) & 1;
          mbar_wait_cluster(&ring.empty[slot], par ^ 1);
          ring.tile[slot] = tile;
          st_cluster_u32(mapa_u32(smem_u32(const_cast<int*>(&ring.tile[slot])), 1), (uint32_t)tile);
          mbar_arrive(&ring.full[slot]);
          mbar_arrive_cluster_release(mapa_u32(smem_u32(&ring.full[slot]), 1));
          ++it;
          if (tile >= num_tiles) {
            if (tile == num_tiles + num_clusters - 1) *reinterpret_cast<volatile uint32_t*>(p.sched_counter) = 0u;   // last draw of the launch
            break;
          }
          next = (int)atomicAdd(p.sched_counter, 1u);          // in flight while this tile's loads are issued
        } else {
          tile = ring_next(ring, it, false, 0);
          if (tile >= num_tiles) break;
        }
        int mb, nb;
        tile_coords2(tile, m_blocks, n_blocks, (p.N % BN) != 0 && n_blocks > 1, mb, nb);
        const int n_eff = tile_n_eff(p.N, nb * BN, BN);
        const int m0 = mb * BM2 + (int)cta * BLOCK_M, n0 = nb * BN + (int)cta * (n_eff / 2);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * kStageBytes;
          unsigned char* sb = sa + kABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * tx_bytes);             // bytes of both CTAs land on the leader's barrier
          const int k0 = kb * block_k;
          if (!p.a_mn_major) {
            tma_load_2d_2cta(sa, &map_a, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / 64; ++a) tma_load_2d_2cta(sa + a * (BLOCK_K * 128), &map_a, &full_bar[stage], m0 + a * 64, k0);
          }
          if (!p.b_mn_major) {
            tma_load_2d_2cta(sb, &map_b, &full_bar[stage], k0, n0);
          } else {
            for (uint32_t a = 0; a < b_boxes; ++a) tma_load_2d_2cta(sb + a * (BLOCK_K * 128), &map_b, &full_bar[stage], n0 + (int)a * 64, k0);
          }
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        tile = next;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (leader && lane == 0) {
      const uint32_t a_lbo = p.a_mn_major ? BLOCK_K * 128 : 16, b_lbo = p.b_mn_major ? BLOCK_K * 128 : 16;
      const uint32_t a_kstep = p.a_mn_major ? UMMA_K * 128 : UMMA_K * 2, b_kstep = p.b_mn_major ? UMMA_K * 128 : UMMA_K * 2;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      uint32_t it = 0;
      for (;;) {
        const int tile = ring_next(ring, it, false, 0);
        if (tile >= num_tiles) break;
        int mb, nb;
        tile_coords2(tile, m_blocks, n_blocks, (p.N % BN) != 0 && n_blocks > 1, mb, nb);
        const uint32_t idesc = make_idesc_f16(BM2, tile_n_eff(p.N, nb * BN, BN), p.ab_format, p.a_mn_major, p.b_mn_major);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kMaxBN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(sa + k * a_kstep, a_lbo, 1024);
            const uint64_t db = make_smem_desc_sw128(sb + k * b_kstep, b_lbo, 1024);
            if (p.fp8) umma_f8_2cta(tmem_d, da, db, idesc, (kb | k) != 0);
            else umma_f16_2cta(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          umma_commit_2cta(&empty_bar[stage]);
          if (kb == k_blocks - 1) umma_commit_2cta(&tmem_full[acc]);
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    gemm2_epilogue(p, tmem_base, tmem_full, tmem_empty, ring, (int)cta, warp, lane, m_blocks, n_blocks);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<kTmemCols>(tmem_base);
  }
}

// One draw counter per stream (launches on one stream are serialised, and the counter is back at zero when a launch ends).
constexpr int kSchedSlots = 64;
static uint32_t* sched_pool() {
  static uint32_t* pool = nullptr;
  if (pool == nullptr) {
    if (cudaMalloc(&pool, kSchedSlots * 128) != cudaSuccess) { pool = nullptr; return nullptr; }
    cudaMemset(pool, 0, kSchedSlots * 128);
  }
  return pool;
}
static uint32_t* sched_counter_for(cudaStream_t st) {
  static cudaStream_t owner[kSchedSlots];
  static int used = 0;
  uint32_t* pool = sched_pool();
  if (pool == nullptr) return nullptr;
  for (int i = 0; i < used; ++i) if (owner[i] == st) return pool + i * 32;
  if (used == kSchedSlots) used = 1;                 // recycle (slot 0 stays with the first stream seen)
  owner[used] = st;
  return pool + (used++) * 32;
}

// after an aborted launch (sticky error, killed kernel) the counters may be non-zero: tests / error paths call this
extern "C" int epl_gemm_reset_scheduler() {
  uint32_t* pool = sched_pool();
  if (pool == nullptr) return -12;
  return (int)cudaMemset(pool, 0, kSchedSlots * 128);
}

static int launch_gemm2(const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, int num_sms, cudaStream_t st) {
  constexpr int kSmem = kStages2 * (BLOCK_M * BLOCK_K * 2 + 128 * BLOCK_K * 2) + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  p.sched_counter = sched_counter_for(st);
  if (p.sched_counter == nullptr) return -12;
  const int tiles = ((p.M + 255) / 256) * ((p.N + p.bn2 - 1) / p.bn2);
  int clusters = std::min(tiles, num_sms / 2);
  gemm2_tcgen05_kernel<<<clusters * 2, kGemm2Threads, kSmem, st>>>(ma, mb, p);
  return EPL_CHECK_LAUNCH();
}

// Tile width of the 2-CTA kernel.  Measured on B200 (profiles/r1_gemm_bench_v3_widths.txt): the time of one tile is almost
// independent of its width (192-wide: 0.95x, 128-wide: 0.87x of the 256-wide tile) because the main loop is bound by
// bytes in flight from L2 (6 stages x 32 KB per CTA), not by MMA issue — so a narrower tile never pays for the waves it
// saves, not even on N = 1600 where 256 leaves 3.03 waves.  The widths stay selectable (force_bn 448 / 384) for shapes
// with fewer tiles than CTA pairs, where they do win.
static int pick_bn2(int M, int N, int num_sms) {
  const long pairs = std::max(num_sms / 2, 1), mt = (M + 255) / 256;
  const int widths[3] = {256, 192, 128};
  const double tile_time[3] = {1.0, 0.95, 0.87};
  int best = 256;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const long tiles = mt * ((N + widths[i] - 1) / widths[i]);
    const long waves = (tiles + pairs - 1) / pairs;
    // a narrower width must promise >= 10 % (short-K shapes fall short of the calibrated tile times)
    const double cost = (double)waves * tile_time[i] * (i == 0 ? 1.0 : 1.10);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = widths[i]; }
  }
  return best;
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
  // The 2-CTA kernel (cta_group::2, 256 x {128,192,256} tiles) is the default whenever the problem spans at least one
  // such tile: measured +8-12 % over the 1-CTA 128x256 kernel on every GPT-2-XL shape
  // (profiles/r1_gemm_bench_v2_with_2cta.txt).  force_bn: 0 = pick the width per shape, 512 / 448 / 384 = force 256 / 192 / 128.
  const bool two_cta_forced = force_bn == 512 || force_bn == 448 || force_bn == 384;
  if ((two_cta_forced || force_bn == 0) && M >= 256 && N >= 256) {
    const int sms = num_sms > 0 ? num_sms : kNumSMs;
    const int bn2 = two_cta_forced ? force_bn - 256 : pick_bn2(M, N, sms);
    CUtensorMap ma2, mb2;
    int rc2 = !a_mn_major ? make_map_2d(&ma2, A, M, K, lda, BLOCK_K, BLOCK_M, is_fp16) : make_map_2d(&ma2, A, K, M, lda, 64, BLOCK_K, is_fp16);
    if (rc2) return rc2;
    rc2 = !b_mn_major ? make_map_2d(&mb2, B, N, K, ldb, BLOCK_K, bn2 / 2, is_fp16) : make_map_2d(&mb2, B, K, N, ldb, 64, BLOCK_K, is_fp16);
    if (rc2) return rc2;
    GemmParams p2;
    p2.M = M; p2.N = N; p2.K = K; p2.ldd = ldd; p2.D = D; p2.bias = bias; p2.pre = pre; p2.aux = aux; p2.epilogue = epilogue;
    p2.accumulate = accumulate; p2.out_dtype = out_dtype; p2.a_mn_major = a_mn_major; p2.b_mn_major = b_mn_major; p2.alpha = alpha;
    p2.ab_format = is_fp16 ? 0 : 1; p2.bn2 = bn2; p2.fp8 = 0; p2.scale_a = p2.scale_b = nullptr;
    if (accumulate && out_dtype == EPL_BF16 && epilogue == EPI_NONE && (ldd & 7) == 0) {
      // D += result for a bf16 D is "residual add with the residual = D": the old values are prefetched before the accumulator
      // wait (aux path) instead of being loaded, added and stored chunk by chunk with the load latency exposed in the epilogue.
      // Weight-gradient GEMMs of micro-batches 2..M (pipelines, gradient accumulation) and of tied weights take this path.
      p2.epilogue = EPI_BIAS_RESIDUAL; p2.bias = nullptr; p2.aux = D; p2.accumulate = 0;
    }
    return launch_gemm2(ma2, mb2, p2, sms, (cudaStream_t)stream);
  }
  const int bn = pick_bn(N, b_mn_major, two_cta_forced ? 0 : force_bn);
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
  p.fp8 = 0; p.scale_a = p.scale_b = nullptr; p.sched_counter = nullptr; p.bn2 = 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_sms <= 0) num_sms = kNumSMs;
  CommParams c{};
  if (bn == 256) return launch_gemm<256, COMM_NONE>(ma, mb, p, c, num_sms, st);
  if (bn == 160) return launch_gemm<160, COMM_NONE>(ma, mb, p, c, num_sms, st);
  return launch_gemm<128, COMM_NONE>(ma, mb, p, c, num_sms, st);
}

// 1-byte tensor map: [rows, cols] of e4m3, cols contiguous, row stride ld bytes, box {128 bytes, box_rows}, 128B swizzle
static int make_map_2d_u8(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -10;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld};
  cuuint32_t box[2] = {128, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r;
  EPL_ENCODE_RETRY(r, ptr, enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  return r == CUDA_SUCCESS ? 0 : -11;
}

// fp8 (e4m3 x e4m3 -> fp32 accumulate) forward GEMM on the 2-CTA kernel: D[M,N] = (A_q[M,K] @ B_q[N,K]^T) * scale_a * scale_b
// (+ the usual epilogues).  A_q / B_q come from epl_quantize_e4m3 (csrc/act.cu) together with their de-quantisation factors.
extern "C" int epl_gemm_fp8(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                            const void* bias, void* pre, const void* aux, int epilogue, int out_dtype, float alpha,
                            const void* scale_a, const void* scale_b, int num_sms, void* stream) {
  if (M < 256 || N < 256 || (K & 15) || (lda & 15) || (ldb & 15)) return -21;
  const int sms = num_sms > 0 ? num_sms : kNumSMs;
  const int bn2 = pick_bn2(M, N, sms);
  CUtensorMap ma, mb;
  int rc = make_map_2d_u8(&ma, A, M, K, lda, BLOCK_M);
  if (rc) return rc;
  rc = make_map_2d_u8(&mb, B, N, K, ldb, bn2 / 2);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.D = D; p.bias = bias; p.pre = pre; p.aux = aux; p.epilogue = epilogue;
  p.accumulate = 0; p.out_dtype = out_dtype; p.a_mn_major = 0; p.b_mn_major = 0; p.alpha = alpha;
  p.ab_format = 0;                               // kind::f8f6f4 format code 0 = E4M3
  p.bn2 = bn2; p.fp8 = 1; p.scale_a = (const float*)scale_a; p.scale_b = (const float*)scale_b;
  return launch_gemm2(ma, mb, p, sms, (cudaStream_t)stream);
}

__global__ void bump_epoch_kernel(uint32_t* word) { *word = *word + 1u; }

// Fused tensor-parallel GEMMs.  mode 1 (all-gather -> GEMM): A is the local gathered buffer `ag_dst` [M, K] which the
// kernel's copy CTAs fill from `ag_src[r]` (each [M/world, K], contiguous).  mode 2 (GEMM -> reduce-scatter): the tiles are
// written to `rs_stage[owner]` slot `rank`, then reduced into `rs_out` [M/world, ldd] (bf16 only).
extern "C" int epl_gemm_fused(int mode, const void* A, const void* B, int M, int N, int K, int lda, int ldb, int ldd,
                              int b_mn_major, const void* bias, void* pre, int epilogue, void* D,
                              int rank, int world, unsigned epoch, int copy_ctas, void* const* flag_ptrs, void* local_sync,
                              void* const* ag_src, void* const* rs_stage, void* rs_out, int is_fp16, void* stream) {
  if (world > kMaxPeers || (mode == COMM_AGB ? N % world : M % world)) return -20;
  const int bn = pick_bn(N, b_mn_major, 0);
  CUtensorMap ma, mb;
  int rc = make_map_2d(&ma, A, M, K, lda, BLOCK_K, BLOCK_M, is_fp16);
  if (rc) return rc;
  if (!b_mn_major) rc = make_map_2d(&mb, B, N, K, ldb, BLOCK_K, bn, is_fp16);
  else rc = make_map_2d(&mb, B, K, N, ldb, 64, BLOCK_K, is_fp16);
  if (rc) return rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.D = D; p.bias = bias; p.pre = pre; p.aux = nullptr; p.epilogue = epilogue;
  p.accumulate = 0; p.out_dtype = is_fp16 ? EPL_F16 : EPL_BF16; p.a_mn_major = 0; p.b_mn_major = b_mn_major; p.alpha = 1.f;
  p.ab_format = is_fp16 ? 0 : 1;
  p.fp8 = 0; p.scale_a = p.scale_b = nullptr; p.sched_counter = nullptr; p.bn2 = 0;
  CommParams c{};
  c.rank = rank; c.world = world; c.epoch = epoch; c.copy_ctas = copy_ctas;
  c.rows_per_rank = (mode == COMM_AGB ? N : M) / world;
  c.local_sync = (uint32_t*)local_sync; c.ag_dst = const_cast<void*>(mode == COMM_AGB ? B : A); c.rs_out = rs_out;
  for (int i = 0; i < world; ++i) {
    c.flags[i] = (uint32_t*)flag_ptrs[i];
    c.ag_src[i] = ag_src ? ag_src[i] : nullptr;
    c.rs_stage[i] = rs_stage ? rs_stage[i] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc2 = -22;
  if (mode == COMM_AG) {
    if (lda != K || (K % 8)) return -21;
    if (bn == 256) rc2 = launch_gemm<256, COMM_AG>(ma, mb, p, c, kNumSMs, st);
    else if (bn == 160) rc2 = launch_gemm<160, COMM_AG>(ma, mb, p, c, kNumSMs, st);
    else rc2 = launch_gemm<128, COMM_AG>(ma, mb, p, c, kNumSMs, st);
  } else if (mode == COMM_AGB) {
    if (b_mn_major || ldb != K || (K % 8)) return -21;
    if (bn == 256) rc2 = launch_gemm<256, COMM_AGB>(ma, mb, p, c, kNumSMs, st);
    else if (bn == 160) rc2 = launch_gemm<160, COMM_AGB>(ma, mb, p, c, kNumSMs, st);
    else rc2 = launch_gemm<128, COMM_AGB>(ma, mb, p, c, kNumSMs, st);
  } else if (mode == COMM_RS) {
    if (is_fp16 || (N % 8) || (ldd % 8)) return -21;
    if (bn == 256) rc2 = launch_gemm<256, COMM_RS>(ma, mb, p, c, kNumSMs, st);
    else if (bn == 160) rc2 = launch_gemm<160, COMM_RS>(ma, mb, p, c, kNumSMs, st);
    else rc2 = launch_gemm<128, COMM_RS>(ma, mb, p, c, kNumSMs, st);
  }
  if (rc2 == 0 && epoch == 0) {                    // device-side epoch: advance it once the whole grid has finished
    bump_epoch_kernel<<<1, 1, 0, st>>>(c.local_sync + 31);
    rc2 = EPL_CHECK_LAUNCH();
  }
  return rc2;
}
