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
};

// Forward.  smem: [warps][stages][row_bytes] then barriers [warps][stages].
template <typename T, bool kRms>
__global__ void __launch_bounds__(256)
norm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta, T* __restrict__ y,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D, float eps, NormGeom geo) {
  constexpr int E = 16 / sizeof(T);
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = D / E;
  unsigned char* ring = smem + (size_t)warp * geo.stages * geo.row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)geo.warps * geo.stages * geo.row_bytes) + warp * geo.stages;
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
  __syncwarp();
  int s = 0; uint32_t phase = 0;
  for (int row = first; row < rows; row += stride) {
    mbar_wait(&bars[s], phase);
    const T* buf = reinterpret_cast<const T*>(ring + (size_t)s * geo.row_bytes);
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
      for (int e = 0; e < E; ++e) { const float d = to_f32<T>(v.v[e]) - mean; sq += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(sq) / D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    T* yr = y + (size_t)row * D;
    for (int vi = lane; vi < nvec; vi += 32) {
      const Vec<T, E> v = ld_vec<T, E>(buf + vi * E);
      const Vec<T, E> g = ld_vec<T, E>(gamma + vi * E);
      Vec<T, E> o;
      if (beta != nullptr) {
        const Vec<T, E> bb = ld_vec<T, E>(beta + vi * E);
#pragma unroll
        for (int e = 0; e < E; ++e) o.v[e] = from_f32<T>((to_f32<T>(v.v[e]) - mean) * rstd * to_f32<T>(g.v[e]) + to_f32<T>(bb.v[e]));
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) o.v[e] = from_f32<T>((to_f32<T>(v.v[e]) - mean) * rstd * to_f32<T>(g.v[e]));
      }
      st_vec<T, E>(yr + vi * E, o);
    }
    __syncwarp();                                           // every lane is done reading this slot
    if (lane == 0) {
      const int nrow = row + geo.stages * stride;
      if (nrow < rows) {
        mbar_expect_tx(&bars[s], geo.row_bytes);
        bulk_g2s(ring + (size_t)s * geo.row_bytes, x + (size_t)nrow * D, geo.row_bytes, &bars[s]);
      }
    }
    if (++s == geo.stages) { s = 0; phase ^= 1; }
  }
}

// Backward, input gradient.  smem: [warps][stages][operands][row_bytes] then barriers; operands = x, dy, (d skip).
template <typename T, bool kRms>
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
  unsigned char* ring = smem + (size_t)warp * geo.stages * slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)geo.warps * geo.stages * slot) + warp * geo.stages;
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
  __syncwarp();
  int s = 0; uint32_t phase = 0;
  for (int row = first; row < rows; row += stride) {
    const float mean = kRms ? 0.f : mean_in[row];
    const float rstd = rstd_in[row];
    mbar_wait(&bars[s], phase);
    const T* bx = reinterpret_cast<const T*>(ring + (size_t)s * slot);
    const T* bd = reinterpret_cast<const T*>(ring + (size_t)s * slot + geo.row_bytes);
    const T* br = reinterpret_cast<const T*>(ring + (size_t)s * slot + 2 * (size_t)geo.row_bytes);
    float s1 = 0.f, s2 = 0.f;
    for (int vi = lane; vi < nvec; vi += 32) {
      const Vec<T, E> vx = ld_vec<T, E>(bx + vi * E), vd = ld_vec<T, E>(bd + vi * E), g = ld_vec<T, E>(gamma + vi * E);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float gd = to_f32<T>(vd.v[e]) * to_f32<T>(g.v[e]);
        s1 += gd;
        s2 += gd * ((to_f32<T>(vx.v[e]) - mean) * rstd);
      }
    }
    s1 = kRms ? 0.f : warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    T* dxr = dx + (size_t)row * D;
    for (int vi = lane; vi < nvec; vi += 32) {
      const Vec<T, E> vx = ld_vec<T, E>(bx + vi * E), vd = ld_vec<T, E>(bd + vi * E), g = ld_vec<T, E>(gamma + vi * E);
      Vec<T, E> vr;
      if (has_res) vr = ld_vec<T, E>(br + vi * E);
      Vec<T, E> o;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float xh = (to_f32<T>(vx.v[e]) - mean) * rstd;
        float v = rstd * (to_f32<T>(vd.v[e]) * to_f32<T>(g.v[e]) - s1 - xh * s2);
        if (has_res) v += to_f32<T>(vr.v[e]);                     // fused residual-branch gradient: dx = LN'(dy) + d(skip)
        o.v[e] = from_f32<T>(v);
      }
      st_vec<T, E>(dxr + vi * E, o);
    }
    __syncwarp();
    if (lane == 0) {
      const int nrow = row + geo.stages * stride;
      if (nrow < rows) fill(s, nrow);
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
static NormGeom norm_geometry(int row_bytes, int operands) {
  NormGeom g;
  g.row_bytes = row_bytes; g.operands = operands;
  const int budget = 100 * 1024;
  g.warps = 8;
  while (g.warps > 1 && (size_t)g.warps * 2 * operands * row_bytes > (size_t)budget) g.warps /= 2;
  g.stages = (int)std::min<size_t>(4, std::max<size_t>(2, (size_t)budget / ((size_t)g.warps * operands * row_bytes)));
  return g;
}
static size_t norm_smem(const NormGeom& g) { return (size_t)g.warps * g.stages * g.operands * g.row_bytes + (size_t)g.warps * g.stages * 8; }

template <typename K>
static int norm_configure(K kernel, size_t bytes) {
  if (bytes > 220 * 1024) return -2;
  static size_t configured = 0;                 // one instance per kernel type (K is a distinct function-pointer type per signature)
  static K last = nullptr;
  if (last == kernel && bytes <= configured) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  last = kernel; configured = bytes;
  return 0;
}

template <typename T, bool kRms>
static int launch_fwd(const void* x, const void* g, const void* b, void* y, float* mean, float* rstd, int rows, int D,
                      float eps, cudaStream_t st) {
  const NormGeom geo = norm_geometry(D * (int)sizeof(T), 1);
  const size_t bytes = norm_smem(geo);
  int rc = norm_configure(norm_fwd_kernel<T, kRms>, bytes);
  if (rc) return rc;
  const int grid = std::min((rows + geo.warps - 1) / geo.warps, kNumSMs * 2);
  norm_fwd_kernel<T, kRms><<<grid, geo.warps * 32, bytes, st>>>((const T*)x, (const T*)g, (const T*)b, (T*)y, mean, rstd, rows, D, eps, geo);
  return EPL_CHECK_LAUNCH();
}

template <typename T, bool kRms>
static int launch_bwd(const void* x, const void* dy, const void* g, const float* mean, const float* rstd, void* dx,
                      float* pg, float* pb, int chunks, int rows, int D, const void* dres, cudaStream_t st) {
  constexpr int E = 16 / sizeof(T);
  const int nvec = D / E;
  const NormGeom geo = norm_geometry(D * (int)sizeof(T), dres ? 3 : 2);
  const size_t bytes = norm_smem(geo);
  int rc = norm_configure(norm_bwd_dx_kernel<T, kRms>, bytes);
  if (rc) return rc;
  const int grid = std::min((rows + geo.warps - 1) / geo.warps, kNumSMs * 2);
  norm_bwd_dx_kernel<T, kRms><<<grid, geo.warps * 32, bytes, st>>>((const T*)x, (const T*)dy, (const T*)g, mean, rstd, (T*)dx, rows, D,
                                                                   (const T*)dres, geo);
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
