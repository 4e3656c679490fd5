// Rotary position embedding applied in place to the Q and K slices of a packed QKV tensor [B, S, 3, H, D] (sm_100a).
// One thread rotates 8 (x_i, x_{i+D/2}) pairs (two 16-byte loads, two 16-byte stores); cos/sin come from sincospif-free
// fast paths (__sincosf on a per-thread frequency).  The backward pass is the same kernel with the angle negated.
// Memory-bound: 2 x (2/3 of the tensor) bytes per call.
#include "epl_common.cuh"
#include <algorithm>

namespace epl {

template <typename T>
__global__ void __launch_bounds__(256) rope_kernel(T* __restrict__ qkv, int B, int S, int H, int D, float base, float sign, int pos_offset) {
  constexpr int E = 16 / sizeof(T);
  const int half = D / 2;
  const int vec_per_head = half / E;                       // threads per (b, s, which, h)
  const int64_t total = (int64_t)B * S * 2 * H * vec_per_head;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float log2_base = log2f(base);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    int64_t t = i;
    const int v = (int)(t % vec_per_head); t /= vec_per_head;
    const int h = (int)(t % H); t /= H;
    const int which = (int)(t % 2); t /= 2;                // 0 = Q, 1 = K
    const int s = (int)(t % S);
    const int64_t b = t / S;
    T* p = qkv + ((((b * S + s) * 3 + which) * H + h) * (int64_t)D) + v * E;
    Vec<T, E> lo = ld_vec<T, E>(p), hi = ld_vec<T, E>(p + half);
    const float pos = (float)(s + pos_offset) * sign;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float freq = exp2f(-log2_base * (float)(2 * (v * E + e)) / (float)D);
      float sn, cs;
      __sincosf(pos * freq, &sn, &cs);
      const float a = to_f32<T>(lo.v[e]), c = to_f32<T>(hi.v[e]);
      lo.v[e] = from_f32<T>(a * cs - c * sn);
      hi.v[e] = from_f32<T>(c * cs + a * sn);
    }
    st_vec<T, E>(p, lo);
    st_vec<T, E>(p + half, hi);
  }
}

}  // namespace epl
using namespace epl;

// qkv: [B, S, 3, H, D] contiguous; D/2 must be a multiple of 16/sizeof(T).  sign = +1 forward, -1 backward.
extern "C" int epl_rope(void* qkv, int B, int S, int H, int D, float base, float sign, int pos_offset, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t work = (int64_t)B * S * 2 * H * (D / 2) / (dtype == EPL_F32 ? 4 : 8);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)kNumSMs * 16));
  if (dtype == EPL_F32) rope_kernel<float><<<blocks, 256, 0, st>>>((float*)qkv, B, S, H, D, base, sign, pos_offset);
  else if (dtype == EPL_BF16) rope_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((__nv_bfloat16*)qkv, B, S, H, D, base, sign, pos_offset);
  else rope_kernel<__half><<<blocks, 256, 0, st>>>((__half*)qkv, B, S, H, D, base, sign, pos_offset);
  return EPL_CHECK_LAUNCH();
}
