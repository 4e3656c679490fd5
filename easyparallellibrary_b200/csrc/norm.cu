// LayerNorm / RMSNorm forward and backward (sm_100a).
//
// One warp per row, the row cached in registers (16-byte vector loads), fp32
// statistics.  Forward reads x once and writes y (+ mean/rstd); backward reads
// x, dy once, writes dx, and accumulates dgamma/dbeta per CTA in shared memory,
// then one partial row per CTA which a second tiny kernel reduces.  These are
// HBM-bound: roofline = 2 (fwd) / 3 (bwd) x row bytes over copy bandwidth.
//
// The reference leaves normalisation to unfused framework ops.
#include "epl_common.cuh"
#include <algorithm>

namespace epl {

constexpr int kWarpsPerCta = 4;

// T = io dtype, VPL = 16-byte vectors per lane (row length <= VPL * 32 * (16/sizeof(T)))
template <typename T, int VPL, bool kRms>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
norm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ y,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D, float eps) {
  constexpr int E = 16 / sizeof(T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D / E;
  for (int row = blockIdx.x * kWarpsPerCta + warp; row < rows; row += gridDim.x * kWarpsPerCta) {
    const T* xr = x + (size_t)row * D;
    float vals[VPL][E];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      int vi = lane + i * 32;
      if (vi < nvec) {
        Vec<T, E> v = ld_vec<T, E>(xr + vi * E);
#pragma unroll
        for (int e = 0; e < E; ++e) { vals[i][e] = to_f32<T>(v.v[e]); sum += vals[i][e]; }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) vals[i][e] = 0.f;
      }
    }
    float mean = 0.f;
    if constexpr (!kRms) mean = warp_sum(sum) / D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      int vi = lane + i * 32;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < E; ++e) { float d = vals[i][e] - mean; sq += d * d; }
      }
    }
    float rstd = rsqrtf(warp_sum(sq) / D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    T* yr = y + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      int vi = lane + i * 32;
      if (vi < nvec) {
        Vec<T, E> g = ld_vec<T, E>(gamma + vi * E);
        Vec<T, E> o;
        if (beta != nullptr) {
          Vec<T, E> b = ld_vec<T, E>(beta + vi * E);
#pragma unroll
          for (int e = 0; e < E; ++e)
            o.v[e] = from_f32<T>((vals[i][e] - mean) * rstd * to_f32<T>(g.v[e]) + to_f32<T>(b.v[e]));
        } else {
#pragma unroll
          for (int e = 0; e < E; ++e) o.v[e] = from_f32<T>((vals[i][e] - mean) * rstd * to_f32<T>(g.v[e]));
        }
        st_vec<T, E>(yr + vi * E, o);
      }
    }
  }
}

// Backward: one CTA (128 threads) per row at a time; thread t owns vectors t, t+128, ... of every row, so the
// parameter-gradient accumulators are private registers (no atomics) and only the two row statistics need a
// CTA reduction (one __syncthreads per row, double-buffered scratch).
constexpr int kBwdThreads = 128;
template <typename T, int VPT, bool kRms>
__global__ void __launch_bounds__(kBwdThreads)
norm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ gamma,
                const float* __restrict__ mean_in, const float* __restrict__ rstd_in, T* __restrict__ dx,
                float* __restrict__ part_dgamma, float* __restrict__ part_dbeta, int rows, int D,
                const T* __restrict__ dres) {
  constexpr int E = 16 / sizeof(T);
  __shared__ float red[2][2][kBwdThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D / E;
  float accg[VPT][E], accb[VPT][E], gam[VPT][E];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int vi = threadIdx.x + i * kBwdThreads;
    Vec<T, E> g;
    if (vi < nvec) g = ld_vec<T, E>(gamma + vi * E);
#pragma unroll
    for (int e = 0; e < E; ++e) { accg[i][e] = 0.f; accb[i][e] = 0.f; gam[i][e] = vi < nvec ? to_f32<T>(g.v[e]) : 0.f; }
  }
  int buf = 0;
  // software pipeline: the next row's x / dy vectors are requested before this row's reduction + barrier
  Vec<T, E> nx[VPT], nd[VPT];
#define EPL_NORM_FETCH(ROW)                                                        \
  _Pragma("unroll") for (int i = 0; i < VPT; ++i) {                                \
    int vi = threadIdx.x + i * kBwdThreads;                                        \
    if (vi < nvec) {                                                               \
      nx[i] = ld_vec<T, E>(x + (size_t)(ROW) * D + vi * E);                        \
      nd[i] = ld_vec<T, E>(dy + (size_t)(ROW) * D + vi * E);                       \
    }                                                                              \
  }
  if ((int)blockIdx.x < rows) { EPL_NORM_FETCH(blockIdx.x) }
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float mean = kRms ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    Vec<T, E> cx[VPT], cd[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) { cx[i] = nx[i]; cd[i] = nd[i]; }
    if (row + (int)gridDim.x < rows) { EPL_NORM_FETCH(row + gridDim.x) }
    float xh[VPT][E], gd[VPT][E];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      int vi = threadIdx.x + i * kBwdThreads;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          float d = to_f32<T>(cd[i].v[e]);
          xh[i][e] = (to_f32<T>(cx[i].v[e]) - mean) * rstd;
          gd[i][e] = d * gam[i][e];
          s1 += gd[i][e];
          s2 += gd[i][e] * xh[i][e];
          accg[i][e] += d * xh[i][e];
          accb[i][e] += d;
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { red[buf][0][warp] = s1; red[buf][1][warp] = s2; }
    __syncthreads();
    s1 = 0.f; s2 = 0.f;
#pragma unroll
    for (int w = 0; w < kBwdThreads / 32; ++w) { s1 += red[buf][0][w]; s2 += red[buf][1][w]; }
    buf ^= 1;
    s1 = kRms ? 0.f : s1 / D;
    s2 = s2 / D;
    T* dxr = dx + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      int vi = threadIdx.x + i * kBwdThreads;
      if (vi < nvec) {
        Vec<T, E> o;
        if (dres != nullptr) {               // fused residual-branch gradient: dx = LN'(dy) + d(skip)
          Vec<T, E> rv = ld_vec<T, E>(dres + (size_t)row * D + vi * E);
#pragma unroll
          for (int e = 0; e < E; ++e) o.v[e] = from_f32<T>(rstd * (gd[i][e] - s1 - xh[i][e] * s2) + to_f32<T>(rv.v[e]));
        } else {
#pragma unroll
          for (int e = 0; e < E; ++e) o.v[e] = from_f32<T>(rstd * (gd[i][e] - s1 - xh[i][e] * s2));
        }
        st_vec<T, E>(dxr + vi * E, o);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    int vi = threadIdx.x + i * kBwdThreads;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        part_dgamma[(size_t)blockIdx.x * D + vi * E + e] = accg[i][e];
        if (part_dbeta) part_dbeta[(size_t)blockIdx.x * D + vi * E + e] = accb[i][e];
      }
    }
  }
}

// out[d] = sum_p part[p][d]  (accumulate into existing grad when acc != 0).  Block = 32 columns x 16 part-lanes.
template <typename T>
__global__ void __launch_bounds__(512) norm_param_reduce_kernel(const float* __restrict__ part, int parts, int D,
                                                                 T* __restrict__ out, int accumulate) {
  __shared__ float s[16][33];
  const int d = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (d < D)
    for (int p = threadIdx.y; p < parts; p += 16) acc += part[(size_t)p * D + d];
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && d < D) {
#pragma unroll
    for (int k = 1; k < 16; ++k) acc += s[k][threadIdx.x];
    if (accumulate) acc += to_f32<T>(out[d]);
    out[d] = from_f32<T>(acc);
  }
}

template <typename T, bool kRms>
static int launch_fwd(const void* x, const void* g, const void* b, void* y, float* mean, float* rstd, int rows, int D,
                      float eps, cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  int nvec = D / E;
  int vpl = (nvec + 31) / 32;
  int grid = std::min((rows + kWarpsPerCta - 1) / kWarpsPerCta, kNumSMs * 8);
  const T *xp = (const T*)x, *gp = (const T*)g, *bp = (const T*)b;
  T* yp = (T*)y;
#define LAUNCH(V)                                                                                                  \
  norm_fwd_kernel<T, V, kRms><<<grid, kWarpsPerCta * 32, 0, st>>>(xp, gp, bp, yp, mean, rstd, rows, D, eps)
  if (vpl <= 1) LAUNCH(1); else if (vpl <= 2) LAUNCH(2); else if (vpl <= 4) LAUNCH(4); else if (vpl <= 8) LAUNCH(8);
  else if (vpl <= 16) LAUNCH(16); else if (vpl <= 32) LAUNCH(32); else return -2;
#undef LAUNCH
  return EPL_CHECK_LAUNCH();
}

template <typename T, bool kRms>
static int launch_bwd(const void* x, const void* dy, const void* g, const float* mean, const float* rstd, void* dx,
                      float* pg, float* pb, int grid, int rows, int D, const void* dres, cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  int nvec = D / E;
  int vpt = (nvec + kBwdThreads - 1) / kBwdThreads;
  const T *xp = (const T*)x, *dp = (const T*)dy, *gp = (const T*)g;
  T* dxp = (T*)dx;
#define LAUNCH(V) norm_bwd_kernel<T, V, kRms><<<grid, kBwdThreads, 0, st>>>(xp, dp, gp, mean, rstd, dxp, pg, pb, rows, D, (const T*)dres)
  if (vpt <= 1) LAUNCH(1); else if (vpt <= 2) LAUNCH(2); else if (vpt <= 4) LAUNCH(4); else if (vpt <= 8) LAUNCH(8); else return -2;
#undef LAUNCH
  return EPL_CHECK_LAUNCH();
}

}  // namespace epl

using namespace epl;

// dtype: EPL_F32 / EPL_BF16 / EPL_F16.  rms != 0 -> RMSNorm (beta, mean ignored).  D must be a multiple of 16/sizeof(T).
extern "C" int epl_norm_fwd(const void* x, const void* gamma, const void* beta, void* y, void* mean, void* rstd,
                            int rows, int D, float eps, int dtype, int rms, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) return 0;
#define GO(T) (rms ? launch_fwd<T, true>(x, gamma, nullptr, y, nullptr, (float*)rstd, rows, D, eps, st)             \
                   : launch_fwd<T, false>(x, gamma, beta, y, (float*)mean, (float*)rstd, rows, D, eps, st))
  if (dtype == EPL_F32) return GO(float);
  if (dtype == EPL_BF16) return GO(__nv_bfloat16);
  return GO(__half);
#undef GO
}

extern "C" int epl_norm_bwd_grid(int rows) { return std::min(rows, kNumSMs * 4); }

// workspace: 2 * grid * D floats (grid from epl_norm_bwd_grid)
extern "C" int epl_norm_bwd(const void* x, const void* dy, const void* gamma, const void* mean, const void* rstd,
                            void* dx, void* dgamma, void* dbeta, void* workspace, int rows, int D, int dtype, int rms,
                            int accumulate, const void* dres, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows <= 0) return 0;
  int grid = epl_norm_bwd_grid(rows);
  float* pg = (float*)workspace;
  float* pb = rms ? nullptr : pg + (size_t)grid * D;
  int rc;
#define GO(T) (rms ? launch_bwd<T, true>(x, dy, gamma, nullptr, (const float*)rstd, dx, pg, pb, grid, rows, D, dres, st)  \
                   : launch_bwd<T, false>(x, dy, gamma, (const float*)mean, (const float*)rstd, dx, pg, pb, grid, rows, D, dres, st))
  if (dtype == EPL_F32) rc = GO(float); else if (dtype == EPL_BF16) rc = GO(__nv_bfloat16); else rc = GO(__half);
#undef GO
  if (rc) return rc;
  int rb = (D + 31) / 32;
  dim3 rblock(32, 16);
#define RED(T)                                                                                                     \
  do {                                                                                                             \
    norm_param_reduce_kernel<T><<<rb, rblock, 0, st>>>(pg, grid, D, (T*)dgamma, accumulate);                          \
    if (pb) norm_param_reduce_kernel<T><<<rb, rblock, 0, st>>>(pb, grid, D, (T*)dbeta, accumulate);                   \
  } while (0)
  if (dtype == EPL_F32) RED(float); else if (dtype == EPL_BF16) RED(__nv_bfloat16); else RED(__half);
#undef RED
  return EPL_CHECK_LAUNCH();
}
