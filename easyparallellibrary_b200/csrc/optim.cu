// Fused optimizer kernels over flat shards (sm_100a).
//
// One pass over HBM per parameter: read fp32 master, m, v and the (bf16/fp16/fp32)
// gradient; un-scale the gradient (loss scale, 1/replicas, clip coefficient);
// AdamW update; write master, m, v and the low-precision model weight.
// 28 B/param with bf16 grads+weights -> the roofline is HBM copy bandwidth.
//
// Replaces the ~12 unfused TF ops per variable of the reference optimizer
// (epl/ops/adam_weight_decay_optimizer.py:117-153).
#include "epl_common.cuh"
#include <algorithm>

namespace epl {

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, grad_scale, inv_c1, inv_c2;
  int vec_ok;
  const float* dyn;      // optional device {lr, inv_c1, inv_c2, grad_scale}: overrides the launch values (CUDA-graph replays)
};

template <typename G, typename O, bool kHasOut, bool kHasMask>
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ master, const G* __restrict__ grad,
                                                     float* __restrict__ m, float* __restrict__ v,
                                                     O* __restrict__ out, const float* __restrict__ mask,
                                                     int64_t n, AdamArgs a) {
  if (a.dyn != nullptr) { a.lr = a.dyn[0]; a.inv_c1 = a.dyn[1]; a.inv_c2 = a.dyn[2]; a.grad_scale = a.dyn[3]; }
  const int64_t nvec = a.vec_ok ? (n >> 2) : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 p4 = reinterpret_cast<const float4*>(master)[i];
    float4 m4 = reinterpret_cast<const float4*>(m)[i];
    float4 v4 = reinterpret_cast<const float4*>(v)[i];
    Vec<G, 4> g4 = ld_vec<G, 4>(grad + 4 * i);
    float4 k4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if constexpr (kHasMask) k4 = reinterpret_cast<const float4*>(mask)[i];
    float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float kk[4] = {k4.x, k4.y, k4.z, k4.w};
    Vec<O, 4> o4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float g = to_f32<G>(g4.v[j]) * a.grad_scale;
      mm[j] = a.beta1 * mm[j] + (1.f - a.beta1) * g;
      vv[j] = a.beta2 * vv[j] + (1.f - a.beta2) * g * g;
      float upd = (mm[j] * a.inv_c1) / (sqrtf(vv[j] * a.inv_c2) + a.eps) + a.weight_decay * kk[j] * pp[j];
      pp[j] -= a.lr * upd;
      if constexpr (kHasOut) o4.v[j] = from_f32<O>(pp[j]);
    }
    reinterpret_cast<float4*>(master)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if constexpr (kHasOut) st_vec<O, 4>(out + 4 * i, o4);
  }
  // tail (n % 4)
  const int64_t tail0 = nvec << 2;
  for (int64_t i = tail0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float g = to_f32<G>(grad[i]) * a.grad_scale;
    float mm = a.beta1 * m[i] + (1.f - a.beta1) * g;
    float vv = a.beta2 * v[i] + (1.f - a.beta2) * g * g;
    float k = 1.f;
    if constexpr (kHasMask) k = mask[i];
    float p = master[i];
    p -= a.lr * ((mm * a.inv_c1) / (sqrtf(vv * a.inv_c2) + a.eps) + a.weight_decay * k * p);
    master[i] = p; m[i] = mm; v[i] = vv;
    if constexpr (kHasOut) out[i] = from_f32<O>(p);
  }
}

template <typename G, typename O, bool kHasOut, bool kHasMom>
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ master, const G* __restrict__ grad,
                                                   float* __restrict__ mom, O* __restrict__ out, int64_t n,
                                                   float lr, float momentum, float weight_decay, float grad_scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float p = master[i];
    float g = to_f32<G>(grad[i]) * grad_scale + weight_decay * p;
    if constexpr (kHasMom) {
      float b = momentum * mom[i] + g;
      mom[i] = b;
      g = b;
    }
    p -= lr * g;
    master[i] = p;
    if constexpr (kHasOut) out[i] = from_f32<O>(p);
  }
}

// sum of squares of a flat buffer -> out[0] (atomicAdd; caller zeroes), and non-finite flag -> out[1]
template <typename T>
__global__ void __launch_bounds__(256) sumsq_kernel(const T* __restrict__ x, int64_t n, float* __restrict__ out) {
  float acc = 0.f;
  bool bad = false;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float f = to_f32<T>(x[i]);
    acc += f * f;
    bad |= !isfinite(f);
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  __shared__ int sbad;
  if (threadIdx.x == 0) sbad = 0;
  __syncthreads();
  if (bad) sbad = 1;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      atomicAdd(out, t);
      if (sbad) atomicExch(reinterpret_cast<int*>(out + 1), 0x3f800000);
    }
  }
}

template <typename G, typename O>
static int launch_adamw(float* master, const void* grad, float* m, float* v, void* out, const float* mask, int64_t n,
                        AdamArgs a, cudaStream_t st) {
  int64_t work = (n + 3) / 4;
  int blocks = (int)std::min<int64_t>((work + 255) / 256, (int64_t)kNumSMs * 8);
  if (blocks < 1) blocks = 1;
  const G* g = static_cast<const G*>(grad);
  O* o = static_cast<O*>(out);
  if (out && mask) adamw_kernel<G, O, true, true><<<blocks, 256, 0, st>>>(master, g, m, v, o, mask, n, a);
  else if (out) adamw_kernel<G, O, true, false><<<blocks, 256, 0, st>>>(master, g, m, v, o, mask, n, a);
  else if (mask) adamw_kernel<G, O, false, true><<<blocks, 256, 0, st>>>(master, g, m, v, o, mask, n, a);
  else adamw_kernel<G, O, false, false><<<blocks, 256, 0, st>>>(master, g, m, v, o, mask, n, a);
  return EPL_CHECK_LAUNCH();
}

}  // namespace epl

using namespace epl;

#define DISPATCH2(GD, OD, FN, ...)                                                            \
  switch ((GD) * 3 + (OD)) {                                                                 \
    case 0: return FN<float, float>(__VA_ARGS__);                                            \
    case 1: return FN<float, __nv_bfloat16>(__VA_ARGS__);                                    \
    case 2: return FN<float, __half>(__VA_ARGS__);                                           \
    case 3: return FN<__nv_bfloat16, float>(__VA_ARGS__);                                    \
    case 4: return FN<__nv_bfloat16, __nv_bfloat16>(__VA_ARGS__);                            \
    case 5: return FN<__nv_bfloat16, __half>(__VA_ARGS__);                                   \
    case 6: return FN<__half, float>(__VA_ARGS__);                                           \
    case 7: return FN<__half, __nv_bfloat16>(__VA_ARGS__);                                   \
    case 8: return FN<__half, __half>(__VA_ARGS__);                                          \
    default: return -1;                                                                      \
  }

extern "C" int epl_adamw(void* master, const void* grad, int grad_dtype, void* m, void* v, void* out, int out_dtype,
                         const void* mask, int64_t n, float lr, float beta1, float beta2, float eps,
                         float weight_decay, float grad_scale, float inv_c1, float inv_c2, void* stream) {
  auto al = [](const void* p, uintptr_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
  int vec_ok = al(master, 16) && al(m, 16) && al(v, 16) && al(mask, 16) && al(grad, grad_dtype == EPL_F32 ? 16 : 8) &&
               al(out, out_dtype == EPL_F32 ? 16 : 8);
  AdamArgs a{lr, beta1, beta2, eps, weight_decay, grad_scale, inv_c1, inv_c2, vec_ok, nullptr};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DISPATCH2(grad_dtype, out_dtype, launch_adamw, (float*)master, grad, (float*)m, (float*)v, out, (const float*)mask,
            n, a, st);
}

// same, with the per-step values {lr, inv_c1, inv_c2, grad_scale} read from device memory: the launch is step-invariant
extern "C" int epl_adamw_dyn(void* master, const void* grad, int grad_dtype, void* m, void* v, void* out, int out_dtype,
                             const void* mask, int64_t n, const void* dyn, float beta1, float beta2, float eps,
                             float weight_decay, void* stream) {
  auto al = [](const void* p, uintptr_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
  int vec_ok = al(master, 16) && al(m, 16) && al(v, 16) && al(mask, 16) && al(grad, grad_dtype == EPL_F32 ? 16 : 8) &&
               al(out, out_dtype == EPL_F32 ? 16 : 8);
  AdamArgs a{0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, 1.f, vec_ok, (const float*)dyn};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DISPATCH2(grad_dtype, out_dtype, launch_adamw, (float*)master, grad, (float*)m, (float*)v, out, (const float*)mask,
            n, a, st);
}

template <typename G, typename O>
static int launch_sgd(float* master, const void* grad, float* mom, void* out, int64_t n, float lr, float momentum,
                      float wd, float gs, cudaStream_t st) {
  int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8);
  if (blocks < 1) blocks = 1;
  const G* g = static_cast<const G*>(grad);
  O* o = static_cast<O*>(out);
  if (out && mom) sgd_kernel<G, O, true, true><<<blocks, 256, 0, st>>>(master, g, mom, o, n, lr, momentum, wd, gs);
  else if (out) sgd_kernel<G, O, true, false><<<blocks, 256, 0, st>>>(master, g, mom, o, n, lr, momentum, wd, gs);
  else if (mom) sgd_kernel<G, O, false, true><<<blocks, 256, 0, st>>>(master, g, mom, o, n, lr, momentum, wd, gs);
  else sgd_kernel<G, O, false, false><<<blocks, 256, 0, st>>>(master, g, mom, o, n, lr, momentum, wd, gs);
  return EPL_CHECK_LAUNCH();
}

extern "C" int epl_sgd(void* master, const void* grad, int grad_dtype, void* mom, void* out, int out_dtype, int64_t n,
                       float lr, float momentum, float weight_decay, float grad_scale, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DISPATCH2(grad_dtype, out_dtype, launch_sgd, (float*)master, grad, (float*)mom, out, n, lr, momentum, weight_decay,
            grad_scale, st);
}

extern "C" int epl_sumsq(const void* x, int dtype, int64_t n, void* out2, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 8);
  if (blocks < 1) blocks = 1;
  if (dtype == EPL_F32) sumsq_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, n, (float*)out2);
  else if (dtype == EPL_BF16) sumsq_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)x, n, (float*)out2);
  else sumsq_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)x, n, (float*)out2);
  return EPL_CHECK_LAUNCH();
}
