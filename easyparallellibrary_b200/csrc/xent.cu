// Fused softmax cross-entropy, forward + gradient in one pass over the logits (sm_100a).
//
// One CTA per row.  The row (up to ~110 K bf16 logits) is staged in shared memory with 16-byte loads, so HBM
// sees exactly one read of the logits and one write of d(logits):
//   loss[r]    = logsumexp(z) - z[label]
//   dlogits[r] = (softmax(z) - onehot(label)) * grad_scale          (written in place when dlogits == logits)
// Rows whose label == ignore_index produce loss 0 and zero gradient.
//
// With `vocab_offset/valid` a rank computes its shard of a vocabulary-parallel loss: it then returns the local
// max and sum-exp instead (see tp_xent below) and the Python side combines them with one tiny all-reduce
// (the reference needs four collectives, epl/ops/distributed_losses.py:58-151).
#include "epl_common.cuh"
#include <algorithm>

namespace epl {

constexpr int kXentThreads = 512;

EPL_DEVICE float block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  for (int w = 0; w < kXentThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

// mode 0: full softmax-xent (+grad).  mode 1: statistics only (local max, local sum-exp, local target logit).
// mode 2: gradient from global statistics (gmax, gsum) for vocab-parallel.
template <typename T>
__global__ void __launch_bounds__(kXentThreads)
xent_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, float* __restrict__ loss,
            T* __restrict__ dlogits, float* __restrict__ stats, const float* __restrict__ gstats, int V, int ld,
            float grad_scale, int64_t ignore_index, int vocab_start, int mode) {
  constexpr int E = 16 / sizeof(T);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* row = reinterpret_cast<T*>(smem_raw);
  __shared__ float red[kXentThreads / 32];
  const int r = blockIdx.x;
  const T* src = logits + (size_t)r * ld;
  const int nvec = V / E;               // V is padded to a multiple of E by the caller (pad logits = -inf)
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nvec; i += kXentThreads) {
    Vec<T, E> v = ld_vec<T, E>(src + i * E);
    st_vec<T, E>(row + i * E, v);
#pragma unroll
    for (int e = 0; e < E; ++e) mx = fmaxf(mx, to_f32<T>(v.v[e]));
  }
  for (int i = nvec * E + threadIdx.x; i < V; i += kXentThreads) { row[i] = src[i]; mx = fmaxf(mx, to_f32<T>(src[i])); }
  mx = block_reduce(mx, red, true);
  if (mode == 2) mx = gstats[2 * r];
  float sum = 0.f;
  const float kLog2e = 1.4426950408889634f;
  for (int i = threadIdx.x; i < V; i += kXentThreads) sum += exp2f((to_f32<T>(row[i]) - mx) * kLog2e);
  sum = block_reduce(sum, red, false);
  const int64_t label = labels[r];
  const int local = (int)(label - vocab_start);
  const bool has = label != ignore_index && local >= 0 && local < V;
  if (mode == 1) {
    if (threadIdx.x == 0) {
      stats[3 * r] = mx;
      stats[3 * r + 1] = sum;
      stats[3 * r + 2] = has ? to_f32<T>(row[local]) : 0.f;
    }
    return;
  }
  if (mode == 2) sum = gstats[2 * r + 1];
  const bool ignored = label == ignore_index;
  if (mode == 0 && threadIdx.x == 0) loss[r] = ignored ? 0.f : (logf(sum) + mx - to_f32<T>(row[local]));
  if (dlogits == nullptr) return;
  const float inv = ignored ? 0.f : grad_scale / sum;
  T* dst = dlogits + (size_t)r * ld;
  for (int i = threadIdx.x; i < nvec; i += kXentThreads) {
    Vec<T, E> v = ld_vec<T, E>(row + i * E), o;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      int c = i * E + e;
      float p = exp2f((to_f32<T>(v.v[e]) - mx) * kLog2e) * inv;
      if (has && c == local) p -= (ignored ? 0.f : grad_scale);
      o.v[e] = from_f32<T>(p);
    }
    st_vec<T, E>(dst + i * E, o);
  }
  for (int i = nvec * E + threadIdx.x; i < V; i += kXentThreads) {
    float p = exp2f((to_f32<T>(row[i]) - mx) * kLog2e) * inv;
    if (has && i == local) p -= (ignored ? 0.f : grad_scale);
    dst[i] = from_f32<T>(p);
  }
}

// x *= *scalar (device scalar), used to fold an upstream gradient into a pre-computed dlogits
template <typename T>
__global__ void __launch_bounds__(256) scale_by_device_scalar_kernel(T* __restrict__ x, const float* __restrict__ s, int64_t nvec) {
  constexpr int E = 16 / sizeof(T);
  const float k = *s;
  if (k == 1.f) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Vec<T, E> v = ld_vec<T, E>(x + i * E);
#pragma unroll
    for (int e = 0; e < E; ++e) v.v[e] = from_f32<T>(to_f32<T>(v.v[e]) * k);
    st_vec<T, E>(x + i * E, v);
  }
}

}  // namespace epl
using namespace epl;

template <typename T>
static int launch_xent(const void* logits, const void* labels, void* loss, void* dlogits, void* stats, const void* gstats,
                       int rows, int V, int ld, float grad_scale, int64_t ignore_index, int vocab_start, int mode,
                       cudaStream_t st) {
  size_t smem = ((size_t)V * sizeof(T) + 15) / 16 * 16;
  if (smem > 200 * 1024) return -3;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(xent_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    configured = true;
  }
  xent_kernel<T><<<rows, kXentThreads, smem, st>>>((const T*)logits, (const int64_t*)labels, (float*)loss, (T*)dlogits,
                                                   (float*)stats, (const float*)gstats, V, ld, grad_scale, ignore_index,
                                                   vocab_start, mode);
  return EPL_CHECK_LAUNCH();
}

// logits: [rows, ld] with V valid columns (ld >= V, ld % (16/sizeof(T)) == 0).
extern "C" int epl_xent(const void* logits, const void* labels, void* loss, void* dlogits, void* stats, const void* gstats,
                        int rows, int V, int ld, float grad_scale, int64_t ignore_index, int vocab_start, int mode,
                        int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) return 0;
  if (dtype == EPL_F32) return launch_xent<float>(logits, labels, loss, dlogits, stats, gstats, rows, V, ld, grad_scale, ignore_index, vocab_start, mode, st);
  if (dtype == EPL_BF16) return launch_xent<__nv_bfloat16>(logits, labels, loss, dlogits, stats, gstats, rows, V, ld, grad_scale, ignore_index, vocab_start, mode, st);
  return launch_xent<__half>(logits, labels, loss, dlogits, stats, gstats, rows, V, ld, grad_scale, ignore_index, vocab_start, mode, st);
}

extern "C" int epl_scale_by_device_scalar(void* x, const void* scalar, int64_t n, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EPL_F32) { int64_t nv = n / 4; scale_by_device_scalar_kernel<float><<<(int)std::min<int64_t>((nv + 255) / 256, kNumSMs * 16), 256, 0, st>>>((float*)x, (const float*)scalar, nv); }
  else if (dtype == EPL_BF16) { int64_t nv = n / 8; scale_by_device_scalar_kernel<__nv_bfloat16><<<(int)std::min<int64_t>((nv + 255) / 256, kNumSMs * 16), 256, 0, st>>>((__nv_bfloat16*)x, (const float*)scalar, nv); }
  else { int64_t nv = n / 8; scale_by_device_scalar_kernel<__half><<<(int)std::min<int64_t>((nv + 255) / 256, kNumSMs * 16), 256, 0, st>>>((__half*)x, (const float*)scalar, nv); }
  return EPL_CHECK_LAUNCH();
}
