// Symmetric memory over NVLink peer mappings + the fused data-parallel kernel (K1).
//
// Symmetric memory: every rank cudaMalloc's an identically sized buffer, exports it with
// cudaIpcGetMemHandle, the 64-byte handles travel through the control plane (torch.distributed
// store) and every rank maps every peer's buffer (cudaIpcOpenMemHandle, lazy peer access).
// Kernels then receive a table of peer pointers and use plain 16-byte ld/st.global over
// NVLink-5 / NVSwitch (measured ~770 GB/s per direction per GPU).
//
// K1 — epl_fused_rs_adam_ag: ONE kernel per gradient bucket that replaces the reference's
//   all-reduce (ncclAllReduce per fused buffer, graph_editor.py:670-725) + divide + unfused Adam,
//   or its ZeRO-v1 reduce -> apply -> broadcast chain (runtime/zero.py:88-175):
//     1. cross-GPU barrier (release/acquire flags in peer memory): every rank's gradients are ready;
//     2. each rank owns 1/W of the bucket; for its shard it loads the W partial gradients straight
//        from the peers' bucket buffers (reduce-scatter by P2P loads), un-scales, applies AdamW on
//        its fp32 master/m/v shard, and stores the new bf16 weights into EVERY rank's parameter
//        buffer (all-gather by P2P stores);
//     3. cross-GPU barrier: all parameter shards have landed.
//   Gradients cross NVLink once (W-1)/W * bucket bytes in, weights once out — the collective's
//   minimum — and never touch HBM in between: no reduced-gradient buffer, no separate Adam pass.
#include "epl_common.cuh"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <type_traits>

namespace epl {

constexpr int kMaxPeers = 8;

struct PeerTable {
  void* ptr[kMaxPeers];
};

struct FusedDpArgs {
  PeerTable grads;          // bucket gradient buffer of every rank (element type G)
  PeerTable params;         // bucket parameter buffer of every rank (element type O)
  PeerTable flags;          // uint32 flags[2][kMaxPeers] of every rank (this bucket's slot)
  uint32_t* local_sync;     // [0] = go flag, [1] = arrival counter (device-local)
  float* master; float* m; float* v; const float* mask;   // this rank's fp32 shard state
  int64_t shard_start;      // element offset of this rank's shard inside the bucket
  int64_t shard_n;          // elements in the shard (multiple of 8)
  int rank, world;
  uint32_t epoch;           // strictly increasing per launch on this bucket
  float lr, beta1, beta2, eps, weight_decay, grad_scale, inv_c1, inv_c2;
};

// ---- cross-GPU barrier executed by one CTA; the rest of the grid waits on a device-local flag -------------
EPL_DEVICE void barrier_start(const FusedDpArgs& a) {
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      uint32_t* peer = reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]);
      st_release_sys(peer + a.rank, a.epoch);                           // slot 0: "my gradients are ready"
      const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]);
      while (ld_acquire_sys(mine + threadIdx.x) < a.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.local_sync), "r"(a.epoch) : "memory");
    }
  } else {
    if (threadIdx.x == 0) {
      uint32_t v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.local_sync) : "memory"); } while (v < a.epoch);
    }
  }
  __syncthreads();
}

EPL_DEVICE void barrier_end(const FusedDpArgs& a) {
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();                                              // my P2P stores are visible system-wide
    uint32_t prev = atomicAdd(a.local_sync + 1, 1u);
    is_last = (prev == gridDim.x * a.epoch - 1);                         // counter is never reset: epoch * grid arrivals
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      uint32_t* peer = reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]);
      st_release_sys(peer + kMaxPeers + a.rank, a.epoch);                // slot 1: "my shard is written everywhere"
      const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]);
      while (ld_acquire_sys(mine + kMaxPeers + threadIdx.x) < a.epoch) {}
    }
  }
}

template <typename T> EPL_DEVICE void accumulate8(float (&acc)[8], const int4& raw);
template <> EPL_DEVICE void accumulate8<__nv_bfloat16>(float (&acc)[8], const int4& raw) {
  const uint32_t w[4] = {(uint32_t)raw.x, (uint32_t)raw.y, (uint32_t)raw.z, (uint32_t)raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = unpack_bf16x2(w[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
}
template <> EPL_DEVICE void accumulate8<__half>(float (&acc)[8], const int4& raw) {
  const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); acc[2 * i] += f.x; acc[2 * i + 1] += f.y; }
}

EPL_DEVICE int4 ld_peer(const void* p) {          // peer memory: bypass L1 allocation, data is read once
  int4 r;
  asm volatile("ld.global.relaxed.sys.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
EPL_DEVICE void st_peer(void* p, const int4& v) {
  asm volatile("st.global.relaxed.sys.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// G = gradient / weight element type (bf16 or fp16), 8 elements (16 bytes) per thread per iteration
template <typename G, bool kHasMask>
__global__ void __launch_bounds__(512, 1) fused_rs_adam_ag_kernel(const FusedDpArgs a) {
  barrier_start(a);
  const int64_t nvec = a.shard_n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const int64_t e = a.shard_start + (i << 3);            // element offset inside the bucket
    // ---- reduce-scatter: W peer loads in flight, starting with a different peer on every rank -----------------
    int4 raw[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < a.world) {
        const int src = (a.rank + p) % a.world;
        raw[p] = ld_peer(reinterpret_cast<const G*>(a.grads.ptr[src]) + e);
      }
    }
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) if (p < a.world) accumulate8<G>(g, raw[p]);
    // ---- AdamW on the fp32 shard ------------------------------------------------------------------------------
    float4 p0 = reinterpret_cast<const float4*>(a.master)[2 * i], p1 = reinterpret_cast<const float4*>(a.master)[2 * i + 1];
    float4 m0 = reinterpret_cast<const float4*>(a.m)[2 * i], m1 = reinterpret_cast<const float4*>(a.m)[2 * i + 1];
    float4 v0 = reinterpret_cast<const float4*>(a.v)[2 * i], v1 = reinterpret_cast<const float4*>(a.v)[2 * i + 1];
    float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float kk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if constexpr (kHasMask) {
      float4 k0 = reinterpret_cast<const float4*>(a.mask)[2 * i], k1 = reinterpret_cast<const float4*>(a.mask)[2 * i + 1];
      kk[0] = k0.x; kk[1] = k0.y; kk[2] = k0.z; kk[3] = k0.w; kk[4] = k1.x; kk[5] = k1.y; kk[6] = k1.z; kk[7] = k1.w;
    }
    uint32_t packed[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gr = g[j] * a.grad_scale;
      mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * gr;
      vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * gr * gr;
      pp[j] -= a.lr * ((mm[j] * a.inv_c1) / (sqrtf(vv[j] * a.inv_c2) + a.eps) + a.weight_decay * kk[j] * pp[j]);
    }
    reinterpret_cast<float4*>(a.master)[2 * i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(a.master)[2 * i + 1] = make_float4(pp[4], pp[5], pp[6], pp[7]);
    reinterpret_cast<float4*>(a.m)[2 * i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(a.m)[2 * i + 1] = make_float4(mm[4], mm[5], mm[6], mm[7]);
    reinterpret_cast<float4*>(a.v)[2 * i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    reinterpret_cast<float4*>(a.v)[2 * i + 1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
    if constexpr (sizeof(G) == 2 && std::is_same<G, __nv_bfloat16>::value) {
#pragma unroll
      for (int j = 0; j < 4; ++j) packed[j] = pack_bf16x2(pp[2 * j], pp[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) { __half2 h = __floats2half2_rn(pp[2 * j], pp[2 * j + 1]); packed[j] = *reinterpret_cast<uint32_t*>(&h); }
    }
    const int4 w = make_int4((int)packed[0], (int)packed[1], (int)packed[2], (int)packed[3]);
    // ---- all-gather: push the new weights into every rank's parameter buffer -------------------------------------
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < a.world) {
        const int dst = (a.rank + p) % a.world;
        st_peer(reinterpret_cast<G*>(a.params.ptr[dst]) + e, w);
      }
    }
  }
  barrier_end(a);
}

// ---------------------------------------------------------------------------------------------------------
// K1 v2 — the same reduce-scatter + AdamW + all-gather, restructured so that it can run on a handful of SMs next
// to the backward GEMMs (and at full NVLink rate on the whole GPU for the last bucket):
//   * peer gradients do not travel through registers: one thread per CTA streams 4 KB pieces of every peer's
//     gradient shard into a shared-memory ring with cp.async.bulk (completion on an mbarrier), so the bytes in
//     flight per CTA (kStages-1 stages x W x 4 KB) are independent of the thread count — 8 CTAs keep ~0.7 MB
//     outstanding, the v1 kernel needed all 148 SMs x 512 threads for the same;
//   * the new weights are staged in shared memory (3 x 4 KB) and leave as W bulk stores per piece
//     (cp.async.bulk.global.shared::cta) — in- and outbound NVLink traffic overlap inside every CTA;
//   * the reduction order is fixed (rank 0 .. W-1) so the result is bit-reproducible against an fp32 reference;
//   * hyper-parameters that change every step (lr, bias corrections, gradient scale) and the barrier epoch are
//     read from device memory (`dyn`), so the launch can be replayed from a CUDA graph.
// ---------------------------------------------------------------------------------------------------------
constexpr int kK1Piece = 2048;                    // elements per piece (4 KB of bf16 per peer)
constexpr int kK1Threads = kK1Piece / 8;          // one 16-byte vector per thread per piece
constexpr int kK1OutBufs = 8;                     // weight staging buffers: kK1OutBufs - 1 store groups in flight per CTA (a peer
                                                  // store keeps its source buffer for an NVLink round trip: with 3 buffers
                                                  // a CTA moved one 4 KB piece per ~2 us regardless of everything else)

struct FusedDpDyn {                               // device-resident, updated by a tiny kernel / memcpy before each step
  float lr, inv_c1, inv_c2, grad_scale;
};

EPL_DEVICE void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
EPL_DEVICE void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}

template <typename G, bool kHasMask>
__global__ void __launch_bounds__(kK1Threads, 1) fused_rs_adam_ag_v2_kernel(const FusedDpArgs a, const FusedDpDyn* __restrict__ dyn,
                                                                         int stages) {
  extern __shared__ __align__(128) unsigned char k1_smem[];
  const int W = a.world;
  const uint32_t stage_bytes = (uint32_t)W * kK1Piece * 2;
  unsigned char* ring = k1_smem;                                            // [stages][W][4 KB]
  unsigned char* obuf = ring + (size_t)stages * stage_bytes;               // [kK1OutBufs][4 KB]
  uint64_t* full = reinterpret_cast<uint64_t*>(obuf + kK1OutBufs * kK1Piece * 2);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  // epoch 0 = take it from device memory (local_sync[2] holds the epoch of the previous launch on this bucket): the launch
  // carries no per-step host value and can be replayed from a CUDA graph
  FusedDpArgs aa = a;
  if (aa.epoch == 0) aa.epoch = *reinterpret_cast<volatile uint32_t*>(a.local_sync + 2) + 1u;
  barrier_start(aa);                                                        // (contains __syncthreads)
  const float lr = dyn->lr, inv_c1 = dyn->inv_c1, inv_c2 = dyn->inv_c2, gscale = dyn->grad_scale;
  const int64_t npieces = (a.shard_n + kK1Piece - 1) / kK1Piece;
  const int64_t first = blockIdx.x, step = gridDim.x;
  const int64_t mine = first < npieces ? (npieces - first + step - 1) / step : 0;   // pieces of this CTA

  auto issue = [&](int64_t jl) {                                            // thread 0: loads of local piece index jl
    const int64_t piece = first + jl * step;
    const int64_t e0 = piece * kK1Piece;
    const uint32_t bytes = (uint32_t)(min((int64_t)kK1Piece, a.shard_n - e0) * 2);
    const int slot = (int)(jl % stages);
    mbar_expect_tx(&full[slot], bytes * (uint32_t)W);
    for (int p = 0; p < W; ++p) {
      const int src = (a.rank + p) % W;                                     // issue order rotated: all links busy at once
      bulk_g2s(ring + (size_t)slot * stage_bytes + (size_t)src * (kK1Piece * 2),
               reinterpret_cast<const G*>(a.grads.ptr[src]) + a.shard_start + e0, bytes, &full[slot]);
    }
  };
  if (tid == 0) for (int64_t jl = 0; jl < min(mine, (int64_t)stages - 1); ++jl) issue(jl);

  // optimizer state of the NEXT piece is requested one iteration ahead (register double buffer): its HBM latency overlaps the
  // current piece's reduction + AdamW instead of being exposed once per piece (measured before: 2 us per 4 KB piece per CTA)
  float4 nx[8];
  auto load_state = [&](int64_t jl_) {
    const int64_t e0_ = (first + jl_ * step) * kK1Piece;
    const int nvec_ = (int)(min((int64_t)kK1Piece, a.shard_n - e0_) >> 3);
    if (tid < nvec_) {
      const int64_t i_ = (e0_ >> 3) + tid;
      nx[0] = reinterpret_cast<const float4*>(a.master)[2 * i_]; nx[1] = reinterpret_cast<const float4*>(a.master)[2 * i_ + 1];
      nx[2] = reinterpret_cast<const float4*>(a.m)[2 * i_]; nx[3] = reinterpret_cast<const float4*>(a.m)[2 * i_ + 1];
      nx[4] = reinterpret_cast<const float4*>(a.v)[2 * i_]; nx[5] = reinterpret_cast<const float4*>(a.v)[2 * i_ + 1];
      if constexpr (kHasMask) { nx[6] = reinterpret_cast<const float4*>(a.mask)[2 * i_]; nx[7] = reinterpret_cast<const float4*>(a.mask)[2 * i_ + 1]; }
    }
  };
  if (mine > 0) load_state(0);
  for (int64_t jl = 0; jl < mine; ++jl) {
    const int64_t piece = first + jl * step;
    const int64_t e0 = piece * kK1Piece;
    const int nvec = (int)(min((int64_t)kK1Piece, a.shard_n - e0) >> 3);
    const int slot = (int)(jl % stages);
    const uint32_t phase = (uint32_t)((jl / stages) & 1);
    const bool active = tid < nvec;
    const int64_t i = (e0 >> 3) + tid;                                      // 8-element vector index inside the shard
    const float4 p0 = nx[0], p1 = nx[1], m0 = nx[2], m1 = nx[3], v0 = nx[4], v1 = nx[5], k0 = nx[6], k1 = nx[7];
    if (jl + 1 < mine) load_state(jl + 1);
    if (tid == 0 && jl + stages - 1 < mine) issue(jl + stages - 1);        // refill the slot consumed in the previous iteration
    mbar_wait(&full[slot], phase);
    unsigned char* ob = obuf + (size_t)(jl % kK1OutBufs) * (kK1Piece * 2);
    if (active) {
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const unsigned char* st = ring + (size_t)slot * stage_bytes + (size_t)tid * 16;
      for (int src = 0; src < W; ++src) accumulate8<G>(g, *reinterpret_cast<const int4*>(st + (size_t)src * (kK1Piece * 2)));
      float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      float kk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if constexpr (kHasMask) { kk[0] = k0.x; kk[1] = k0.y; kk[2] = k0.z; kk[3] = k0.w; kk[4] = k1.x; kk[5] = k1.y; kk[6] = k1.z; kk[7] = k1.w; }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gr = g[j] * gscale;
        mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * gr;
        vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * gr * gr;
        pp[j] -= lr * ((mm[j] * inv_c1) / (sqrtf(vv[j] * inv_c2) + a.eps) + a.weight_decay * kk[j] * pp[j]);
      }
      reinterpret_cast<float4*>(a.master)[2 * i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
      reinterpret_cast<float4*>(a.master)[2 * i + 1] = make_float4(pp[4], pp[5], pp[6], pp[7]);
      reinterpret_cast<float4*>(a.m)[2 * i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      reinterpret_cast<float4*>(a.m)[2 * i + 1] = make_float4(mm[4], mm[5], mm[6], mm[7]);
      reinterpret_cast<float4*>(a.v)[2 * i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      reinterpret_cast<float4*>(a.v)[2 * i + 1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
      uint32_t packed[4];
      if constexpr (std::is_same<G, __nv_bfloat16>::value) {
#pragma unroll
        for (int j = 0; j < 4; ++j) packed[j] = pack_bf16x2(pp[2 * j], pp[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { __half2 h = __floats2half2_rn(pp[2 * j], pp[2 * j + 1]); packed[j] = *reinterpret_cast<uint32_t*>(&h); }
      }
      *reinterpret_cast<int4*>(ob + (size_t)tid * 16) = make_int4((int)packed[0], (int)packed[1], (int)packed[2], (int)packed[3]);
    }
    fence_proxy_async();                                                    // staged weights -> visible to the bulk-copy engine
    __syncthreads();                                                        // ring slot consumed by everyone, staging complete
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)nvec * 16;
      for (int p = 0; p < W; ++p) {
        const int dst = (a.rank + p) % W;
        bulk_s2g(reinterpret_cast<G*>(a.params.ptr[dst]) + a.shard_start + e0, ob, bytes);
      }
      tma_store_commit();
      tma_store_wait_read<kK1OutBufs - 2>();                                // the staging buffer the NEXT piece will use is free again
    }
  }
  if (tid == 0) {
    tma_store_wait<0>();                                                    // every weight store of this CTA has been performed
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  // end barrier; unlike v1 the arrival counter is reset by the last CTA, so the grid size may change between launches
  __syncthreads();
  __shared__ int is_last;
  if (tid == 0) {
    __threadfence_system();
    is_last = (atomicAdd(a.local_sync + 1, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (tid < W) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[tid]) + kMaxPeers + a.rank, aa.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + kMaxPeers + tid) < aa.epoch) {}
    }
    __syncthreads();
    if (tid == 0) { a.local_sync[1] = 0u; a.local_sync[2] = aa.epoch; __threadfence(); }
  }
}

// Stand-alone device barrier over the same flag protocol (used by tests and by the Python engine between phases)
__global__ void symm_barrier_kernel(PeerTable flags, int rank, int world, uint32_t epoch) {
  if (threadIdx.x < world) {
    __threadfence_system();
    uint32_t* peer = reinterpret_cast<uint32_t*>(flags.ptr[threadIdx.x]);
    st_release_sys(peer + rank, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.ptr[rank]);
    while (ld_acquire_sys(mine + threadIdx.x) < epoch) {}
  }
}

// peer-copy bandwidth probe: dst (local) <- src (peer), 16-byte vectors
__global__ void __launch_bounds__(512) peer_copy_kernel(const int4* __restrict__ src, int4* __restrict__ dst, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) dst[i] = ld_peer(src + i);
}

}  // namespace epl
using namespace epl;

extern "C" {

int epl_symm_alloc(int64_t bytes, void** out) {
  cudaError_t e = cudaMalloc(out, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaMemset(*out, 0, (size_t)bytes);
}
int epl_symm_free(void* p) { return (int)cudaFree(p); }
int epl_symm_export(void* p, void* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
int epl_symm_import(const void* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
}
int epl_symm_unimport(void* p) { return (int)cudaIpcCloseMemHandle(p); }

int epl_symm_barrier(void* const* flag_ptrs, int rank, int world, unsigned epoch, void* stream) {
  PeerTable t;
  for (int i = 0; i < kMaxPeers; ++i) t.ptr[i] = i < world ? flag_ptrs[i] : nullptr;
  symm_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t, rank, world, epoch);
  return EPL_CHECK_LAUNCH();
}

int epl_peer_copy(const void* src, void* dst, int64_t bytes, int blocks, void* stream) {
  peer_copy_kernel<<<blocks, 512, 0, (cudaStream_t)stream>>>((const int4*)src, (int4*)dst, bytes / 16);
  return EPL_CHECK_LAUNCH();
}

// grad/param element dtype: EPL_BF16 or EPL_F16.  shard_n must be a multiple of 8, shard_start too.
int epl_fused_rs_adam_ag(void* const* grad_ptrs, void* const* param_ptrs, void* const* flag_ptrs, void* local_sync,
                         void* master, void* m, void* v, const void* mask, int64_t shard_start, int64_t shard_n,
                         int rank, int world, unsigned epoch, int dtype, float lr, float beta1, float beta2, float eps,
                         float weight_decay, float grad_scale, float inv_c1, float inv_c2, int blocks, void* stream) {
  if (world > kMaxPeers || (shard_n & 7) || (shard_start & 7)) return -20;
  FusedDpArgs a;
  for (int i = 0; i < kMaxPeers; ++i) {
    a.grads.ptr[i] = i < world ? grad_ptrs[i] : nullptr;
    a.params.ptr[i] = i < world ? param_ptrs[i] : nullptr;
    a.flags.ptr[i] = i < world ? flag_ptrs[i] : nullptr;
  }
  a.local_sync = (uint32_t*)local_sync;
  a.master = (float*)master; a.m = (float*)m; a.v = (float*)v; a.mask = (const float*)mask;
  a.shard_start = shard_start; a.shard_n = shard_n; a.rank = rank; a.world = world; a.epoch = epoch;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_scale = grad_scale;
  a.inv_c1 = inv_c1; a.inv_c2 = inv_c2;
  cudaStream_t st = (cudaStream_t)stream;
  if (blocks <= 0) blocks = kNumSMs;
  blocks = std::min(blocks, kNumSMs);            // every CTA must be co-resident: the grid spins on a flag
  if (dtype == EPL_BF16) {
    if (mask) fused_rs_adam_ag_kernel<__nv_bfloat16, true><<<blocks, 512, 0, st>>>(a);
    else fused_rs_adam_ag_kernel<__nv_bfloat16, false><<<blocks, 512, 0, st>>>(a);
  } else if (dtype == EPL_F16) {
    if (mask) fused_rs_adam_ag_kernel<__half, true><<<blocks, 512, 0, st>>>(a);
    else fused_rs_adam_ag_kernel<__half, false><<<blocks, 512, 0, st>>>(a);
  } else {
    return -1;
  }
  return EPL_CHECK_LAUNCH();
}

// v2: bulk-copy ring; `dyn` = device pointer to {lr, inv_c1, inv_c2, grad_scale}; epoch is still a launch argument
// (a graph-captured launch bumps it with epl_symm_epoch_add on the flags instead, see below).
int epl_fused_rs_adam_ag_v2(void* const* grad_ptrs, void* const* param_ptrs, void* const* flag_ptrs, void* local_sync,
                            void* master, void* m, void* v, const void* mask, int64_t shard_start, int64_t shard_n,
                            int rank, int world, unsigned epoch, int dtype, const void* dyn, float beta1, float beta2, float eps,
                            float weight_decay, int blocks, void* stream) {
  if (world > kMaxPeers || (shard_n & 7) || (shard_start & 7)) return -20;
  FusedDpArgs a;
  for (int i = 0; i < kMaxPeers; ++i) {
    a.grads.ptr[i] = i < world ? grad_ptrs[i] : nullptr;
    a.params.ptr[i] = i < world ? param_ptrs[i] : nullptr;
    a.flags.ptr[i] = i < world ? flag_ptrs[i] : nullptr;
  }
  a.local_sync = (uint32_t*)local_sync;
  a.master = (float*)master; a.m = (float*)m; a.v = (float*)v; a.mask = (const float*)mask;
  a.shard_start = shard_start; a.shard_n = shard_n; a.rank = rank; a.world = world; a.epoch = epoch;
  a.lr = 0.f; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_scale = 1.f;
  a.inv_c1 = a.inv_c2 = 1.f;
  const int stage_bytes = world * kK1Piece * 2;
  int stages = std::max(2, std::min(8, (128 * 1024) / stage_bytes));
  const int smem = stages * stage_bytes + kK1OutBufs * kK1Piece * 2 + 128;
  const int64_t npieces = (shard_n + kK1Piece - 1) / kK1Piece;
  if (blocks <= 0) blocks = kNumSMs;
  blocks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(blocks, kNumSMs), npieces));
  // CTAs are launched as pairs (cluster of 2 = one TPC) so that a small grid running next to the backward GEMMs takes whole
  // TPCs and leaves the GEMM's CTA pairs (cta_group::2 needs both SMs of a TPC) a clean set of SMs
  cudaStream_t st = (cudaStream_t)stream;
#define EPL_K1V2(G, MASK)                                                                                              \
  do {                                                                                                                 \
    static bool configured = false;                                                                                    \
    if (!configured) {                                                                                                 \
      cudaError_t e = cudaFuncSetAttribute(fused_rs_adam_ag_v2_kernel<G, MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           128 * 1024 + kK1OutBufs * kK1Piece * 2 + 128);                               \
      if (e != cudaSuccess) return (int)e;                                                                             \
      configured = true;                                                                                               \
    }                                                                                                                  \
    cudaLaunchConfig_t cfg = {};                                                                                       \
    cfg.gridDim = dim3(blocks, 1, 1); cfg.blockDim = dim3(kK1Threads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st; \
    cudaLaunchAttribute attr[1];                                                                                       \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                                  \
    attr[0].val.clusterDim.x = (blocks % 2 == 0) ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;   \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                                                \
    cudaError_t le = cudaLaunchKernelEx(&cfg, fused_rs_adam_ag_v2_kernel<G, MASK>, a, (const FusedDpDyn*)dyn, stages);  \
    if (le != cudaSuccess) return (int)le;                                                                             \
  } while (0)
  if (dtype == EPL_BF16) { if (mask) EPL_K1V2(__nv_bfloat16, true); else EPL_K1V2(__nv_bfloat16, false); }
  else if (dtype == EPL_F16) { if (mask) EPL_K1V2(__half, true); else EPL_K1V2(__half, false); }
  else return -1;
#undef EPL_K1V2
  return EPL_CHECK_LAUNCH();
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// K5 — all-to-all over NVLink peer memory (MoE dispatch / combine).
//   send: local [world, seg] buffer; segment j is stored straight into rank j's symmetric receive buffer at
//   slot `rank` (P2P 16-byte stores), bracketed by release/acquire flag barriers in peer memory.  One kernel,
//   no NCCL; the reference issues 2*world ncclSend/ncclRecv per call (tensorflow_nccl.h:186-206).
// ---------------------------------------------------------------------------------------------------------
namespace epl {
struct A2AArgs {
  PeerTable recv;           // every rank's receive buffer [world, seg_bytes]
  PeerTable flags;          // signal pads: [0,8) start, [8,16) end
  const int4* send;         // local [world, seg_bytes]
  uint32_t* local_sync;     // [0] go, [1] arrivals
  int64_t seg_vecs;         // 16-byte vectors per segment
  int rank, world;
  uint32_t epoch;
};

__global__ void a2a_bump_epoch_kernel(uint32_t* word) { *word = *word + 1u; }

__global__ void __launch_bounds__(512) alltoall_p2p_kernel(const A2AArgs a_in) {
  A2AArgs a = a_in;            // epoch 0: kept on the device (local_sync[3] = epoch of the previous call), graph-replayable
  if (a.epoch == 0) a.epoch = *reinterpret_cast<volatile uint32_t*>(a.local_sync + 3) + 1u;
  // start: every peer has finished consuming its receive buffer from the previous call
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]) + a.rank, a.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + threadIdx.x) < a.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.local_sync), "r"(a.epoch) : "memory"); }
  } else if (threadIdx.x == 0) {
    uint32_t v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.local_sync) : "memory"); } while (v < a.epoch);
  }
  __syncthreads();
  const int64_t total = a.seg_vecs * a.world;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int p = (int)(i / a.seg_vecs);
    const int dst = (a.rank + p) % a.world;                     // rotate so all links carry traffic at once
    const int64_t off = i - (int64_t)p * a.seg_vecs;
    int4 v = a.send[(int64_t)dst * a.seg_vecs + off];
    st_peer(reinterpret_cast<int4*>(a.recv.ptr[dst]) + (int64_t)a.rank * a.seg_vecs + off, v);
  }
  // end: my segments have landed everywhere and everyone's segments have landed here
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    is_last = (atomicAdd(a.local_sync + 1, 1u) == gridDim.x - 1);          // the last CTA resets the counter (grid may vary per call)
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]) + kMaxPeers + a.rank, a.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + kMaxPeers + threadIdx.x) < a.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) { a.local_sync[1] = 0u; __threadfence(); }
  }
}
}  // namespace epl

extern "C" int epl_alltoall_p2p(const void* send, void* const* recv_ptrs, void* const* flag_ptrs, void* local_sync,
                                int64_t seg_bytes, int rank, int world, unsigned epoch, int blocks, void* stream) {
  if (world > epl::kMaxPeers || (seg_bytes & 15)) return -20;
  epl::A2AArgs a;
  for (int i = 0; i < epl::kMaxPeers; ++i) {
    a.recv.ptr[i] = i < world ? recv_ptrs[i] : nullptr;
    a.flags.ptr[i] = i < world ? flag_ptrs[i] : nullptr;
  }
  a.send = (const int4*)send; a.local_sync = (uint32_t*)local_sync; a.seg_vecs = seg_bytes / 16;
  a.rank = rank; a.world = world; a.epoch = epoch;
  if (blocks <= 0) blocks = 64;
  epl::alltoall_p2p_kernel<<<std::min(blocks, epl::kNumSMs), 512, 0, (cudaStream_t)stream>>>(a);
  if (epoch == 0) epl::a2a_bump_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((uint32_t*)local_sync + 3);
  return EPL_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------------------
// K5b — MoE dispatch fused into the all-to-all: the routed [E, C, M] tensor is never materialised.  `index[e*C + c]` is the
// row of `x` ([tokens, M] bf16/fp16) routed to slot c of expert e, or -1 for an empty slot.  Each slot is read from `x`
// (or zero-filled) and stored straight into the destination rank's symmetric receive buffer, laid out
// [local expert][src rank][C][M] — what the expert GEMMs consume (T = world * C rows per local expert).  One warp per slot, 16-byte vectors.
// (Reference: einsum "gsec,gsm->egcm" followed by 2*world ncclSend/Recv, parallel/hooks.py:758-794.)
// ---------------------------------------------------------------------------------------------------------
namespace epl {
struct A2AGatherArgs {
  PeerTable recv;           // every rank's receive buffer [e_local][world][C][M]
  PeerTable flags;
  const int4* x;            // local tokens [tokens, M]
  const int* index;         // [E * C]
  uint32_t* local_sync;
  int64_t row_vecs;         // 16-byte vectors per row (M * 2 / 16)
  int E, C, e_local;
  int rank, world;
  uint32_t epoch;
};

__global__ void __launch_bounds__(512) alltoall_gather_p2p_kernel(const A2AGatherArgs a_in) {
  A2AGatherArgs a = a_in;
  if (a.epoch == 0) a.epoch = *reinterpret_cast<volatile uint32_t*>(a.local_sync + 3) + 1u;
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]) + a.rank, a.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + threadIdx.x) < a.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.local_sync), "r"(a.epoch) : "memory"); }
  } else if (threadIdx.x == 0) {
    uint32_t v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.local_sync) : "memory"); } while (v < a.epoch);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int64_t slots = (int64_t)a.E * a.C;
  for (int64_t s0 = (int64_t)blockIdx.x * warps + warp; s0 < slots; s0 += (int64_t)gridDim.x * warps) {
    // rotate the expert order by rank so that all ranks do not hammer the same destination at the same time
    const int e = (int)((s0 / a.C + (int64_t)a.rank * a.e_local) % a.E), c = (int)(s0 % a.C);
    const int row = a.index[(int64_t)e * a.C + c];
    const int dst = e / a.e_local, el = e - dst * a.e_local;
    int4* out = reinterpret_cast<int4*>(a.recv.ptr[dst]) + (((int64_t)el * a.world + a.rank) * a.C + c) * a.row_vecs;   // [e_local][src][C][M]
    const int4* in = a.x + (int64_t)max(row, 0) * a.row_vecs;
    for (int64_t v = lane; v < a.row_vecs; v += 32) {
      int4 val = row >= 0 ? in[v] : make_int4(0, 0, 0, 0);
      st_peer(out + v, val);
    }
  }
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    is_last = (atomicAdd(a.local_sync + 1, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]) + kMaxPeers + a.rank, a.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + kMaxPeers + threadIdx.x) < a.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) { a.local_sync[1] = 0u; __threadfence(); }
  }
}
}  // namespace epl

extern "C" int epl_alltoall_gather_p2p(const void* x, const void* index, void* const* recv_ptrs, void* const* flag_ptrs, void* local_sync,
                                       int64_t row_bytes, int E, int C, int rank, int world, unsigned epoch, int blocks, void* stream) {
  if (world > epl::kMaxPeers || (row_bytes & 15) || E % world) return -20;
  epl::A2AGatherArgs a;
  for (int i = 0; i < epl::kMaxPeers; ++i) {
    a.recv.ptr[i] = i < world ? recv_ptrs[i] : nullptr;
    a.flags.ptr[i] = i < world ? flag_ptrs[i] : nullptr;
  }
  a.x = (const int4*)x; a.index = (const int*)index; a.local_sync = (uint32_t*)local_sync; a.row_vecs = row_bytes / 16;
  a.E = E; a.C = C; a.e_local = E / world; a.rank = rank; a.world = world; a.epoch = epoch;
  if (blocks <= 0) blocks = 64;
  epl::alltoall_gather_p2p_kernel<<<std::min(blocks, epl::kNumSMs), 512, 0, (cudaStream_t)stream>>>(a);
  if (epoch == 0) epl::a2a_bump_epoch_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((uint32_t*)local_sync + 3);
  return EPL_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------------------
// NVLS — in-switch reduction through a multicast mapping (cuMulticastCreate / cuMulticastBindMem; the mapping is set up by
// runtime/nvls.py).  One-shot all-reduce: after a flag barrier every rank owns 1/W of the buffer; for its slice it issues
// multimem.ld_reduce (the NVSwitch adds the W copies and returns ONE value: (W-1)/W fewer bytes into the GPU than W peer
// loads) and multimem.st (one store, the switch broadcasts it to all W copies).  bf16 inputs accumulate in fp32 inside the
// switch (.acc::f32).
// ---------------------------------------------------------------------------------------------------------
namespace epl {
struct NvlsArgs {
  void* mc;                 // multicast address of the symmetric buffer
  PeerTable flags;
  uint32_t* local_sync;
  int64_t nvec;             // 16-byte vectors in the buffer
  int rank, world;
  uint32_t epoch;
};

template <int kDtype>   // EPL_BF16 or EPL_F32
__global__ void __launch_bounds__(512) nvls_allreduce_kernel(const NvlsArgs a_in) {
  NvlsArgs a = a_in;
  if (a.epoch == 0) a.epoch = *reinterpret_cast<volatile uint32_t*>(a.local_sync + 3) + 1u;
  if (blockIdx.x == 0) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]) + a.rank, a.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + threadIdx.x) < a.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a.local_sync), "r"(a.epoch) : "memory"); }
  } else if (threadIdx.x == 0) {
    uint32_t v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.local_sync) : "memory"); } while (v < a.epoch);
  }
  __syncthreads();
  const int64_t per = (a.nvec + a.world - 1) / a.world;
  const int64_t lo = per * a.rank, hi = min(lo + per, a.nvec);
  int4* base = reinterpret_cast<int4*>(a.mc);
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r0, r1, r2, r3;
    if constexpr (kDtype == EPL_BF16) {
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                   : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(base + i) : "memory");
      asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" :: "l"(base + i), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    } else {
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "l"(base + i) : "memory");
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(base + i), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
    }
  }
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    is_last = (atomicAdd(a.local_sync + 1, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]) + kMaxPeers + a.rank, a.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + kMaxPeers + threadIdx.x) < a.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) { a.local_sync[1] = 0u; __threadfence(); }
  }
}
}  // namespace epl

extern "C" int epl_nvls_allreduce(void* mc_ptr, void* const* flag_ptrs, void* local_sync, int64_t nbytes, int dtype, int rank, int world,
                                  unsigned epoch, int blocks, void* stream) {
  if (world > epl::kMaxPeers || (nbytes & 15)) return -20;
  epl::NvlsArgs a;
  for (int i = 0; i < epl::kMaxPeers; ++i) a.flags.ptr[i] = i < world ? flag_ptrs[i] : nullptr;
  a.mc = mc_ptr; a.local_sync = (uint32_t*)local_sync; a.nvec = nbytes / 16; a.rank = rank; a.world = world; a.epoch = epoch;
  if (blocks <= 0) blocks = 32;
  blocks = std::min(blocks, epl::kNumSMs);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EPL_BF16) epl::nvls_allreduce_kernel<EPL_BF16><<<blocks, 512, 0, st>>>(a);
  else if (dtype == EPL_F32) epl::nvls_allreduce_kernel<EPL_F32><<<blocks, 512, 0, st>>>(a);
  else return -1;
  if (epoch == 0) epl::a2a_bump_epoch_kernel<<<1, 1, 0, st>>>((uint32_t*)local_sync + 3);
  return EPL_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------------------
// K1 over NVLS: the reduce-scatter is ONE multimem.ld_reduce per 16 bytes of this rank's shard (the NVSwitch adds the W
// replicas of the gradient bucket in fp32 and returns bf16: only 1/W of the bucket enters the GPU instead of (W-1)/W), the
// all-gather is ONE multimem.st per 16 bytes (the switch replicates the new weights into every rank's parameter bucket).
// Per direction and GPU the phase moves (1 + 1/W) x bucket bytes instead of 2 x (W-1)/W x bucket bytes of the peer-pointer
// kernel (W = 8: 3.5 vs 5.45 GB for GPT-2-XL).  Four vectors per thread are requested before the first is consumed.
// ---------------------------------------------------------------------------------------------------------
namespace epl {
struct FusedNvlsArgs {
  const void* mc_grads;     // multicast address of this bucket's gradients (element 0 of the bucket)
  void* mc_params;          // multicast address of this bucket's parameters
  PeerTable flags;
  uint32_t* local_sync;
  float* master; float* m; float* v; const float* mask;
  int64_t shard_start, shard_n;
  int rank, world;
  uint32_t epoch;
  float beta1, beta2, eps, weight_decay;
};

template <bool kHasMask>
__global__ void __launch_bounds__(256) fused_nvls_adam_kernel(const FusedNvlsArgs a, const FusedDpDyn* __restrict__ dyn) {
  FusedDpArgs b;                                                  // reuse the flag-barrier helpers of the peer-pointer kernel
  b.flags = a.flags; b.local_sync = a.local_sync; b.rank = a.rank; b.world = a.world;
  b.epoch = a.epoch ? a.epoch : *reinterpret_cast<volatile uint32_t*>(a.local_sync + 2) + 1u;
  barrier_start(b);
  const float lr = dyn->lr, inv_c1 = dyn->inv_c1, inv_c2 = dyn->inv_c2, gscale = dyn->grad_scale;
  const int64_t nvec = a.shard_n >> 3;
  const int4* gsrc = reinterpret_cast<const int4*>(reinterpret_cast<const __nv_bfloat16*>(a.mc_grads) + a.shard_start);
  int4* pdst = reinterpret_cast<int4*>(reinterpret_cast<__nv_bfloat16*>(a.mc_params) + a.shard_start);
  constexpr int kU = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * kU) {
    uint32_t r[kU][4];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + (int64_t)u * stride;
      if (i < nvec)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                     : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "l"(gsrc + i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + (int64_t)u * stride;
      if (i >= nvec) continue;
      float g[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16x2(r[u][j]); g[2 * j] = f.x; g[2 * j + 1] = f.y; }
      float4 p0 = reinterpret_cast<const float4*>(a.master)[2 * i], p1 = reinterpret_cast<const float4*>(a.master)[2 * i + 1];
      float4 m0 = reinterpret_cast<const float4*>(a.m)[2 * i], m1 = reinterpret_cast<const float4*>(a.m)[2 * i + 1];
      float4 v0 = reinterpret_cast<const float4*>(a.v)[2 * i], v1 = reinterpret_cast<const float4*>(a.v)[2 * i + 1];
      float pp[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      float kk[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if constexpr (kHasMask) {
        float4 k0 = reinterpret_cast<const float4*>(a.mask)[2 * i], k1 = reinterpret_cast<const float4*>(a.mask)[2 * i + 1];
        kk[0] = k0.x; kk[1] = k0.y; kk[2] = k0.z; kk[3] = k0.w; kk[4] = k1.x; kk[5] = k1.y; kk[6] = k1.z; kk[7] = k1.w;
      }
      uint32_t packed[4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gr = g[j] * gscale;
        mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * gr;
        vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * gr * gr;
        pp[j] -= lr * ((mm[j] * inv_c1) / (sqrtf(vv[j] * inv_c2) + a.eps) + a.weight_decay * kk[j] * pp[j]);
      }
      reinterpret_cast<float4*>(a.master)[2 * i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
      reinterpret_cast<float4*>(a.master)[2 * i + 1] = make_float4(pp[4], pp[5], pp[6], pp[7]);
      reinterpret_cast<float4*>(a.m)[2 * i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
      reinterpret_cast<float4*>(a.m)[2 * i + 1] = make_float4(mm[4], mm[5], mm[6], mm[7]);
      reinterpret_cast<float4*>(a.v)[2 * i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      reinterpret_cast<float4*>(a.v)[2 * i + 1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
#pragma unroll
      for (int j = 0; j < 4; ++j) packed[j] = pack_bf16x2(pp[2 * j], pp[2 * j + 1]);
      asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};"
                   :: "l"(pdst + i), "r"(packed[0]), "r"(packed[1]), "r"(packed[2]), "r"(packed[3]) : "memory");
    }
  }
  // end barrier (self-resetting arrival counter, device-side epoch), as in the v2 kernel
  __syncthreads();
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    is_last = (atomicAdd(a.local_sync + 1, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if ((int)threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(a.flags.ptr[threadIdx.x]) + kMaxPeers + a.rank, b.epoch);
      while (ld_acquire_sys(reinterpret_cast<const uint32_t*>(a.flags.ptr[a.rank]) + kMaxPeers + threadIdx.x) < b.epoch) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) { a.local_sync[1] = 0u; a.local_sync[2] = b.epoch; __threadfence(); }
  }
}
}  // namespace epl

extern "C" int epl_fused_nvls_adam(const void* mc_grads, void* mc_params, void* const* flag_ptrs, void* local_sync, void* master, void* m,
                                   void* v, const void* mask, int64_t shard_start, int64_t shard_n, int rank, int world, unsigned epoch,
                                   const void* dyn, float beta1, float beta2, float eps, float weight_decay, int blocks, void* stream) {
  using namespace epl;
  if (world > kMaxPeers || (shard_n & 7) || (shard_start & 7)) return -20;
  FusedNvlsArgs a;
  a.mc_grads = mc_grads; a.mc_params = mc_params;
  for (int i = 0; i < kMaxPeers; ++i) a.flags.ptr[i] = i < world ? flag_ptrs[i] : nullptr;
  a.local_sync = (uint32_t*)local_sync; a.master = (float*)master; a.m = (float*)m; a.v = (float*)v; a.mask = (const float*)mask;
  a.shard_start = shard_start; a.shard_n = shard_n; a.rank = rank; a.world = world; a.epoch = epoch;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  if (blocks <= 0) blocks = kNumSMs * 2;
  blocks = std::min(blocks, kNumSMs * 4);        // all CTAs must become resident eventually; none waits on another except at the start flag
  cudaStream_t st = (cudaStream_t)stream;
  if (mask) fused_nvls_adam_kernel<true><<<blocks, 256, 0, st>>>(a, (const FusedDpDyn*)dyn);
  else fused_nvls_adam_kernel<false><<<blocks, 256, 0, st>>>(a, (const FusedDpDyn*)dyn);
  return EPL_CHECK_LAUNCH();
}
