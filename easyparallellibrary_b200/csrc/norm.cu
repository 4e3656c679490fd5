// LayerNorm / RMSNorm forward and backward (sm_100a).
//
// These kernels are pure HBM streams (forward: read x, write y; backward: read x, dy, d(skip), write dx), so the design
// question is how many bytes each SM keeps in flight.  A register-resident row (one warp per row, 16-byte loads) topped out
// at 40-80 KB per SM and 1.4-2.8 TB/s because registers bound both the occupancy and the prefetch depth.  Here every warp
// owns a private ring of row buffers in shared memory filled by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx,
// issued by lane 0): bytes in flight = warps x stages x row bytes (~100-200 KB per SM), independent of registers, with no
// CTA-wide barrier anywhere.  The row is then read from shared memory (conflict-free 16-byte vectors) once per pass, so the
// kernels need no per-row-length template.
//
// The reference leaves normalisation to unfused framework ops.
#include "epl_common.cuh"
#include <algorithm>

namespace epl {

EPL_DEVICE void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct NormGeom {          // launch-time geometry shared by host and device
  int warps;               // warps per CTA
  int stages;              // ring depth per warp
  int row_bytes;           // D * sizeof(T), multiple of 16
  int operands;            // row buffers per stage (1 forward, 2-3 backward)
  int param_bytes;         // fp32 copies of gamma (and beta) at the start of shared memory
};

// The kernels are issue-bound once the memory side is fixed (the first TMA version spent 1200 warp instructions per
// 1600-element row re-reading and re-converting the row in three passes), so for rows of up to 8 vectors per lane
// (VPL > 0) the row is converted to fp32 registers ONCE and gamma / beta are staged as fp32 in shared memory once per
// CTA: ~500 instructions per row, fully unrolled.  VPL == 0 is the generic multi-pass path for longer rows.
template <typename T, int VPL>
EPL_DEVICE void norm_load_row(const T* buf, int lane, int nvec, float (&vals)[VPL > 0 ? VPL : 1][16 / sizeof(T)]) {
  constexpr int E = 16 / sizeof(T);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const Vec<T, E> v = ld_vec<T, E>(buf + vi * E);
#pragma unroll
      for (int e = 0; e < E; ++e) vals[i][e] = to_f32<T>(v.v[e]);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) vals[i][e] = 0.f;
    }
  }
}

// Forward.  smem: [gamma fp32][beta fp32][warps][stages][row_bytes] then barriers [warps][stages].
template <typename T, bool kRms, int VPL>
__global__ void __launch_bounds__(256)
norm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ y,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D, float eps, NormGeom geo) {
  constexpr int E = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D / E;
  float* sg = reinterpret_cast<float*>(smem);
  float* sb = sg + D;
  unsigned char* ring0 = smem + geo.param_bytes;
  unsigned char* ring = ring0 + (size_t)warp * geo.stages * geo.row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring0 + (size_t)geo.warps * geo.stages * geo.row_bytes) + warp * geo.stages;
  const int first = blockIdx.x * geo.warps + warp, stride = gridDim.x * geo.warps;
  if (lane == 0) {
    for (int s = 0; s < geo.stages; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
    for (int s = 0; s < geo.stages; ++s) {
      const int row = first + s * stride;
      if (row < rows) {
        mbar_expect_tx(&bars[s], geo.row_bytes);
        bulk_g2s(ring + (size_t)s * geo.row_bytes, x + (size_t)row * D, geo.row_bytes, &bars[s]);
      }
    }
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    sg[d] = to_f32<T>(gamma[d]);
    sb[d] = beta != nullptr ? to_f32<T>(beta[d]) : 0.f;
  }
  __syncthreads();
  int s = 0; uint32_t phase = 0;
  for (int row = first; row < rows; row += stride) {
    mbar_wait(&bars[s], phase);
    const T* buf = reinterpret_cast<const T*>(ring + (size_t)s * geo.row_bytes);
    T* yr = y + (size_t)row * D;
    if constexpr (VPL > 0) {
      float vals[VPL][E];
      norm_load_row<T, VPL>(buf, lane, nvec, vals);
      __syncwarp();                                         // slot consumed: refill it before doing the math
      if (lane == 0) {
        const int nrow = row + geo.stages * stride;
        if (nrow < rows) {
          mbar_expect_tx(&bars[s], geo.row_bytes);
          bulk_g2s(ring + (size_t)s * geo.row_bytes, x + (size_t)nrow * D, geo.row_bytes, &bars[s]);
        }
      }
      float sum = 0.f;
      if constexpr (!kRms) {
#pragma unroll
        for (int i = 0; i < VPL; ++i)
#pragma unroll
          for (int e = 0; e < E; ++e) sum += vals[i][e];
      }
      const float mean = kRms ? 0.f : warp_sum(sum) / D;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        if (lane + i * 32 < nvec) {
#pragma unroll
          for (int e = 0; e < E; ++e) { const float d = vals[i][e] - mean; sq = fmaf(d, d, sq); }
        }
      }
      const float rstd = rsqrtf(warp_sum(sq) / D + eps);
      if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        rstd_out[row] = rstd;
      }
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          Vec<T, E> o;
#pragma unroll
          for (int e = 0; e < E; e += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(sg + vi * E + e), b4 = *reinterpret_cast<const float4*>(sb + vi * E + e);
            o.v[e + 0] = from_f32<T>(fmaf((vals[i][e + 0] - mean) * rstd, g4.x, b4.x));
            o.v[e + 1] = from_f32<T>(fmaf((vals[i][e + 1] - mean) * rstd, g4.y, b4.y));
            o.v[e + 2] = from_f32<T>(fmaf((vals[i][e + 2] - mean) * rstd, g4.z, b4.z));
            o.v[e + 3] = from_f32<T>(fmaf((vals[i][e + 3] - mean) * rstd, g4.w, b4.w));
          }
          st_vec<T, E>(yr + vi * E, o);
        }
      }
    } else {
      float sum = 0.f;
      if constexpr (!kRms) {
        for (int vi = lane; vi < nvec; vi += 32) {
          const Vec<T, E> v = ld_vec<T, E>(buf + vi * E);
#pragma unroll
          for (int e = 0; e < E; ++e) sum += to_f32<T>(v.v[e]);
        }
      }
      const float mean = kRms ? 0.f : warp_sum(sum) / D;
      float sq = 0.f;
      for (int vi = lane; vi < nvec; vi += 32) {
        const Vec<T, E> v = ld_vec<T, E>(buf + vi * E);
#pragma unroll
        for (int e = 0; e < E; ++e) { const float d = to_f32<T>(v.v[e]) - mean; sq = fmaf(d, d, sq); }
      }
      const float rstd = rsqrtf(warp_sum(sq) / D + eps);
      if (lane == 0) {
        if (mean_out) mean_out[row] = mean;
        rstd_out[row] = rstd;
      }
      for (int vi = lane; vi < nvec; vi += 32) {
        const Vec<T, E> v = ld_vec<T, E>(buf + vi * E);
        Vec<T, E> o;
#pragma unroll
        for (int e = 0; e < E; ++e) o.v[e] = from_f32<T>(fmaf((to_f32<T>(v.v[e]) - mean) * rstd, sg[vi * E + e], sb[vi * E + e]));
        st_vec<T, E>(yr + vi * E, o);
      }
      __syncwarp();                                         // every lane is done reading this slot
      if (lane == 0) {
        const int nrow = row + geo.stages * stride;
        if (nrow < rows) {
          mbar_expect_tx(&bars[s], geo.row_bytes);
          bulk_g2s(ring + (size_t)s * geo.row_bytes, x + (size_t)nrow * D, geo.row_bytes, &bars[s]);
        }
      }
    }
    if (++s == geo.stages) { s = 0; phase ^= 1; }
  }
}

// Backward, input gradient.  smem: [gamma fp32][warps][stages][operands][row_bytes] then barriers; operands = x, dy, (d skip).
template <typename T, bool kRms, int VPL>
__global__ void __launch_bounds__(256)
norm_bwd_dx_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ gamma,
                   const float* __restrict__ mean_in, const float* __restrict__ rstd_in, T* __restrict__ dx,
                   int rows, int D, const T* __restrict__ dres, NormGeom geo) {
  constexpr int E = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D / E;
  const bool has_res = dres != nullptr;
  const size_t slot = (size_t)geo.operands * geo.row_bytes;
  float* sg = reinterpret_cast<float*>(smem);
  unsigned char* ring0 = smem + geo.param_bytes;
  unsigned char* ring = ring0 + (size_t)warp * geo.stages * slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring0 + (size_t)geo.warps * geo.stages * slot) + warp * geo.stages;
  const int first = blockIdx.x * geo.warps + warp, stride = gridDim.x * geo.warps;
  auto fill = [&](int st, int row) {
    unsigned char* dst = ring + (size_t)st * slot;
    mbar_expect_tx(&bars[st], geo.operands * geo.row_bytes);
    bulk_g2s(dst, x + (size_t)row * D, geo.row_bytes, &bars[st]);
    bulk_g2s(dst + geo.row_bytes, dy + (size_t)row * D, geo.row_bytes, &bars[st]);
    if (has_res) bulk_g2s(dst + 2 * (size_t)geo.row_bytes, dres + (size_t)row * D, geo.row_bytes, &bars[st]);
  };
  if (lane == 0) {
    for (int st = 0; st < geo.stages; ++st) mbar_init(&bars[st], 1);
    mbar_fence_init();
    for (int st = 0; st < geo.stages; ++st) {
      const int row = first + st * stride;
      if (row < rows) fill(st, row);
    }
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) sg[d] = to_f32<T>(gamma[d]);
  __syncthreads();
  int s = 0; uint32_t phase = 0;
  for (int row = first; row < rows; row += stride) {
    const float mean = kRms ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    mbar_wait(&bars[s], phase);
    const T* bx = reinterpret_cast<const T*>(ring + (size_t)s * slot);
    const T* bd = reinterpret_cast<const T*>(ring + (size_t)s * slot + geo.row_bytes);
    const T* br = reinterpret_cast<const T*>(ring + (size_t)s * slot + 2 * (size_t)geo.row_bytes);
    T* dxr = dx + (size_t)row * D;
    if constexpr (VPL > 0) {
      // xh = normalised input, gd = gamma * dy; the residual gradient stays packed until the last moment
      float xh[VPL][E], gd[VPL][E];
      Vec<T, E> vr[VPL];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          const Vec<T, E> vx = ld_vec<T, E>(bx + vi * E), vd = ld_vec<T, E>(bd + vi * E);
          if (has_res) vr[i] = ld_vec<T, E>(br + vi * E);
#pragma unroll
          for (int e = 0; e < E; e += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(sg + vi * E + e);
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              xh[i][e + k] = (to_f32<T>(vx.v[e + k]) - mean) * rstd;
              gd[i][e + k] = to_f32<T>(vd.v[e + k]) * gg[k];
              s1 += gd[i][e + k];
              s2 = fmaf(gd[i][e + k], xh[i][e + k], s2);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        const int nrow = row + geo.stages * stride;
        if (nrow < rows) fill(s, nrow);
      }
      s1 = kRms ? 0.f : warp_sum(s1) / D;
      s2 = warp_sum(s2) / D;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int vi = lane + i * 32;
        if (vi < nvec) {
          Vec<T, E> o;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            float v = rstd * (fmaf(-xh[i][e], s2, gd[i][e]) - s1);
            if (has_res) v += to_f32<T>(vr[i].v[e]);               // fused residual-branch gradient: dx = LN'(dy) + d(skip)
            o.v[e] = from_f32<T>(v);
          }
          st_vec<T, E>(dxr + vi * E, o);
        }
      }
    } else {
      float s1 = 0.f, s2 = 0.f;
      for (int vi = lane; vi < nvec; vi += 32) {
        const Vec<T, E> vx = ld_vec<T, E>(bx + vi * E), vd = ld_vec<T, E>(bd + vi * E);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float g = to_f32<T>(vd.v[e]) * sg[vi * E + e];
          s1 += g;
          s2 = fmaf(g, (to_f32<T>(vx.v[e]) - mean) * rstd, s2);
        }
      }
      s1 = kRms ? 0.f : warp_sum(s1) / D;
      s2 = warp_sum(s2) / D;
      for (int vi = lane; vi < nvec; vi += 32) {
        const Vec<T, E> vx = ld_vec<T, E>(bx + vi * E), vd = ld_vec<T, E>(bd + vi * E);
        Vec<T, E> vr;
        if (has_res) vr = ld_vec<T, E>(br + vi * E);
        Vec<T, E> o;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float xh = (to_f32<T>(vx.v[e]) - mean) * rstd;
          float v = rstd * (to_f32<T>(vd.v[e]) * sg[vi * E + e] - s1 - xh * s2);
          if (has_res) v += to_f32<T>(vr.v[e]);
          o.v[e] = from_f32<T>(v);
        }
        st_vec<T, E>(dxr + vi * E, o);
      }
      __syncwarp();
      if (lane == 0) {
        const int nrow = row + geo.stages * stride;
        if (nrow < rows) fill(s, nrow);
      }
    }
    if (++s == geo.stages) { s = 0; phase ^= 1; }
  }
}

// partial dgamma / dbeta: grid = (column strips of 32 vectors, row chunks); block = 8 warps, warp w takes rows w, w+8, ...
// of the chunk.  part_*[chunk][D].
constexpr int kParamWarps = 8;
constexpr int kParamRowsPerCta = 128;
template <typename T, bool kRms>
__global__ void __launch_bounds__(kParamWarps * 32)
norm_bwd_param_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean_in,
                      const float* __restrict__ rstd_in, float* __restrict__ part_dgamma, float* __restrict__ part_dbeta,
                      int rows, int D) {
  constexpr int E = 16 / sizeof(T);
  __shared__ float red[kParamWarps][32][2 * E + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vi = blockIdx.x * 32 + lane;
  const bool col_ok = vi < D / E;
  const int row0 = blockIdx.y * kParamRowsPerCta;
  const int row_end = min(row0 + kParamRowsPerCta, rows);
  float ag[E], ab[E];
#pragma unroll
  for (int e = 0; e < E; ++e) { ag[e] = 0.f; ab[e] = 0.f; }
  constexpr int kUnroll = 4;
  for (int r = row0 + warp; r < row_end; r += kParamWarps * kUnroll) {
    Vec<T, E> vx[kUnroll], vd[kUnroll];
    float mu[kUnroll], rs[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int rr = r + u * kParamWarps;
      const bool ok = col_ok && rr < row_end;
      if (ok) {
        vx[u] = ld_vec<T, E>(x + (size_t)rr * D + vi * E);
        vd[u] = ld_vec<T, E>(dy + (size_t)rr * D + vi * E);
      }
      mu[u] = (kRms || rr >= row_end) ? 0.f : mean_in[rr];
      rs[u] = rr < row_end ? rstd_in[rr] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (col_ok && r + u * kParamWarps < row_end) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float d = to_f32<T>(vd[u].v[e]);
          ag[e] += d * ((to_f32<T>(vx[u].v[e]) - mu[u]) * rs[u]);
          ab[e] += d;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) { red[warp][lane][e] = ag[e]; red[warp][lane][E + e] = ab[e]; }
  __syncthreads();
  // 32 lanes x 2E values, summed over the 8 warps by the first 2E warps-worth of threads
  for (int idx = threadIdx.x; idx < 32 * 2 * E; idx += kParamWarps * 32) {
    const int l = idx / (2 * E), k = idx % (2 * E);
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kParamWarps; ++w) t += red[w][l][k];
    const int col = (blockIdx.x * 32 + l) * E + (k % E);
    if (col < D) {
      if (k < E) part_dgamma[(size_t)blockIdx.y * D + col] = t;
      else if (part_dbeta) part_dbeta[(size_t)blockIdx.y * D + col] = t;
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

// ring geometry: as many warps x stages as fit ~100 KB per CTA (two CTAs per SM), at least 2 stages
static NormGeom norm_geometry(int row_bytes, int operands, int param_bytes) {
  NormGeom g;
  g.row_bytes = row_bytes; g.operands = operands; g.param_bytes = (param_bytes + 127) / 128 * 128;
  const long budget = 110 * 1024 - g.param_bytes;
  g.warps = 8;
  while (g.warps > 1 && (long)g.warps * 2 * operands * row_bytes > budget) g.warps /= 2;
  g.stages = (int)std::min<long>(4, std::max<long>(2, budget / ((long)g.warps * operands * row_bytes)));
  return g;
}
static size_t norm_smem(const NormGeom& g) {
  return (size_t)g.param_bytes + (size_t)g.warps * g.stages * g.operands * g.row_bytes + (size_t)g.warps * g.stages * 8;
}

template <typename K>
static int norm_configure(K kernel, size_t bytes) {
  if (bytes > 220 * 1024) return -2;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

template <typename T, bool kRms>
static int launch_fwd(const void* x, const void* g, const void* b, void* y, float* mean, float* rstd, int rows, int D,
                      float eps, cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  const NormGeom geo = norm_geometry(D * (int)sizeof(T), 1, D * 8);
  const size_t bytes = norm_smem(geo);
  const int grid = std::min((rows + geo.warps - 1) / geo.warps, kNumSMs * 2);
  const int vpl = (D / E + 31) / 32;
  int rc;
#define LAUNCH(V)                                                                                                       \
  do {                                                                                                                  \
    rc = norm_configure(norm_fwd_kernel<T, kRms, V>, bytes);                                                            \
    if (rc) return rc;                                                                                                  \
    norm_fwd_kernel<T, kRms, V><<<grid, geo.warps * 32, bytes, st>>>((const T*)x, (const T*)g, (const T*)b, (T*)y, mean, rstd, rows, D, eps, geo); \
  } while (0)
  if (vpl <= 1) LAUNCH(1); else if (vpl <= 2) LAUNCH(2); else if (vpl <= 4) LAUNCH(4); else if (vpl <= 8) LAUNCH(8); else LAUNCH(0);
#undef LAUNCH
  return EPL_CHECK_LAUNCH();
}

template <typename T, bool kRms>
static int launch_bwd(const void* x, const void* dy, const void* g, const float* mean, const float* rstd, void* dx,
                      float* pg, float* pb, int chunks, int rows, int D, const void* dres, cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  const int nvec = D / E;
  const NormGeom geo = norm_geometry(D * (int)sizeof(T), dres ? 3 : 2, D * 4);
  const size_t bytes = norm_smem(geo);
  const int grid = std::min((rows + geo.warps - 1) / geo.warps, kNumSMs * 2);
  const int vpl = (nvec + 31) / 32;
  int rc;
#define LAUNCH(V)                                                                                                       \
  do {                                                                                                                  \
    rc = norm_configure(norm_bwd_dx_kernel<T, kRms, V>, bytes);                                                         \
    if (rc) return rc;                                                                                                  \
    norm_bwd_dx_kernel<T, kRms, V><<<grid, geo.warps * 32, bytes, st>>>((const T*)x, (const T*)dy, (const T*)g, mean, rstd, (T*)dx, rows, D, \
                                                                        (const T*)dres, geo);                           \
  } while (0)
  if (vpl <= 1) LAUNCH(1); else if (vpl <= 2) LAUNCH(2); else if (vpl <= 4) LAUNCH(4); else if (vpl <= 8) LAUNCH(8); else LAUNCH(0);
#undef LAUNCH
  dim3 pgrid((nvec + 31) / 32, chunks);
  norm_bwd_param_kernel<T, kRms><<<pgrid, kParamWarps * 32, 0, st>>>((const T*)x, (const T*)dy, mean, rstd, pg, pb, rows, D);
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

// number of partial rows the parameter-gradient pass writes (workspace = 2 * this * D floats)
extern "C" int epl_norm_bwd_grid(int rows) { return (rows + kParamRowsPerCta - 1) / kParamRowsPerCta; }

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
