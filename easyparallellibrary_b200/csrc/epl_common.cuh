// Shared device helpers for the EPL-B200 kernels (sm_100a only).
//
// Everything here is thin inline PTX: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences), cluster helpers, vector
// loads/stores with cache hints, and acquire/release system-scope accesses for
// NVLink peer memory.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define EPL_HOST_DEVICE __host__ __device__ __forceinline__
#define EPL_DEVICE __device__ __forceinline__

// dtype codes shared with python (ops/_lib.py)
enum EplDtype : int { EPL_F32 = 0, EPL_BF16 = 1, EPL_F16 = 2 };

#define EPL_CHECK_LAUNCH() (int)cudaGetLastError()

// Autograd's per-device worker threads bind their CUDA context lazily: a thread whose first CUDA call is a DRIVER call of ours
// (cuTensorMapEncodeTiled) has no current context and gets CUDA_ERROR_INVALID_CONTEXT (seen on B200: stage 0 of a pipeline,
// whose backward starts with a GEMM).  Bind the primary context of the device that owns `ptr` and evaluate the call again.
static inline void epl_bind_context_for(const void* ptr) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) == cudaSuccess && a.type == cudaMemoryTypeDevice) cudaSetDevice(a.device);
  cudaFree(0);
}
#define EPL_ENCODE_RETRY(r, ptr, call)                 \
  do {                                                 \
    r = (call);                                        \
    if (r == CUDA_ERROR_INVALID_CONTEXT) {             \
      epl_bind_context_for(ptr);                       \
      r = (call);                                      \
    }                                                  \
  } while (0)

namespace epl {

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------------------------------------
// scalar conversion
// ---------------------------------------------------------------------------------------------
template <typename T> EPL_DEVICE float to_f32(T v);
template <> EPL_DEVICE float to_f32<float>(float v) { return v; }
template <> EPL_DEVICE float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> EPL_DEVICE float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> EPL_DEVICE T from_f32(float v);
template <> EPL_DEVICE float from_f32<float>(float v) { return v; }
template <> EPL_DEVICE __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> EPL_DEVICE __half from_f32<__half>(float v) { return __float2half_rn(v); }

EPL_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
EPL_DEVICE float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 8 x 16-bit (or 4 x fp32) vector = 16 bytes
template <typename T, int N> struct alignas(sizeof(T) * N) Vec { T v[N]; };

template <typename T, int N>
EPL_DEVICE Vec<T, N> ld_vec(const T* p) { return *reinterpret_cast<const Vec<T, N>*>(p); }
template <typename T, int N>
EPL_DEVICE void st_vec(T* p, const Vec<T, N>& v) { *reinterpret_cast<Vec<T, N>*>(p) = v; }

// streaming (read-once) 16 B load / store
EPL_DEVICE int4 ld_stream(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
EPL_DEVICE void st_stream(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
EPL_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
EPL_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// system-scope acquire/release (flags in peer memory over NVLink)
// ---------------------------------------------------------------------------------------------
EPL_DEVICE void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
EPL_DEVICE uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
EPL_DEVICE uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
EPL_DEVICE void red_release_sys_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.sys.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
EPL_DEVICE void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
EPL_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

EPL_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
EPL_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
EPL_DEVICE void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
EPL_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
EPL_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
EPL_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
EPL_DEVICE void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(desc) : "memory");
}
// 2-D tile load: coordinates are (inner, outer) element offsets
EPL_DEVICE void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(desc), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
EPL_DEVICE void tma_store_2d(const void* desc, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :: "l"(desc), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
EPL_DEVICE void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> EPL_DEVICE void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> EPL_DEVICE void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
EPL_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05
// ---------------------------------------------------------------------------------------------
EPL_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
EPL_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCols>
EPL_DEVICE void tmem_alloc(uint32_t* smem_result) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(smem_result)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
EPL_DEVICE void tmem_dealloc(uint32_t tmem_addr) {    // the same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_addr), "n"(kCols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16/fp16 inputs, fp32 accumulate; one thread issues
EPL_DEVICE void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when every previously issued tcgen05.mma of this thread has completed
EPL_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp receives row (lane base + t), 32 consecutive columns
EPL_DEVICE void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
EPL_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 format: version field = 1).
//   start address >> 4 in bits [0,14); leading byte offset >> 4 in [16,30); stride byte offset >> 4 in [32,46);
//   version in [46,48); layout type in [61,64): 2 = 128-byte swizzle.
EPL_DEVICE uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16: D = fp32, A/B = bf16 (1) or fp16 (0); major: 0 = K-major, 1 = MN-major.
EPL_HOST_DEVICE uint32_t make_idesc_f16(int m, int n, int ab_format, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                                // c_format = F32
  d |= (uint32_t)(ab_format & 7) << 7;         // a_format
  d |= (uint32_t)(ab_format & 7) << 10;        // b_format
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)((n >> 3) & 0x3F) << 17;
  d |= (uint32_t)((m >> 4) & 0x1F) << 24;
  return d;
}

EPL_DEVICE bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace epl
