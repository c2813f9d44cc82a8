// b200clip: HBM-bound row kernels of the CLIP head -- L2 normalisation (a-L2), LayerNorm (tail of a-P1/a-P2),
// column sums for bias gradients, casts.  One warp owns one row: 128-bit coalesced loads, fp32 statistics,
// warp-shuffle reductions.  Roofline: HBM bandwidth (bytes listed per kernel).
#include "common.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

constexpr int ROW_THREADS = 256;          // 8 warps = 8 rows per block pass
constexpr int MAX_D = 1024;               // per-lane register tile is MAX_V float4; kernels are templated on MAX_V
#define B200_DISPATCH_V(D, ...)                                      \
  do {                                                              \
    if ((D) <= 512) { constexpr int MAX_V = 4; __VA_ARGS__; }       \
    else            { constexpr int MAX_V = 8; __VA_ARGS__; }       \
  } while (0)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 w = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// ------------------------------------------------------------------------------------------------
// a-L2 forward: y = x / max(||x||, eps)        (0426/train.py:191-192; F.normalize, eps 1e-12)
// bytes/row: D*sizeof(in) read + D*2 (bf16 out) [+ D*4 (f32 out)] + 4
// ------------------------------------------------------------------------------------------------
template <typename TIn, int MAX_V>
__global__ void __launch_bounds__(ROW_THREADS) l2norm_fwd_kernel(const TIn* __restrict__ x, long long ldx,
                                                                 __nv_bfloat16* __restrict__ y_bf16,
                                                                 float* __restrict__ y_f32, float* __restrict__ inv_norm,
                                                                 int rows, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int nv = D >> 7;                                    // float4 per lane
  for (long long row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * (ROW_THREADS / 32)) {
    const TIn* xr = x + row * ldx;
    float4 v[MAX_V];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        v[i] = ld4(xr + i * 128 + lane * 4);
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && inv_norm) inv_norm[row] = inv;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const float4 o = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
        if (y_bf16) st4(y_bf16 + row * D + i * 128 + lane * 4, o);
        if (y_f32) st4(y_f32 + row * D + i * 128 + lane * 4, o);
      }
  }
}

// a-L2 backward: dx (+)= inv * (dy - yhat * (yhat . dy)) [+ addend * *addend_scale],  yhat = x * inv.   If the norm was
// clamped (inv == 1/eps) the clamp has zero gradient and dx = dy * inv.  The addend carries gradients that reach x
// directly (the BCE heads' input gradient computed in the forward pass, scaled by the upstream dLoss).
template <typename TIn, int MAX_V>
__global__ void __launch_bounds__(ROW_THREADS) l2norm_bwd_kernel(const float* __restrict__ dy, const TIn* __restrict__ x,
                                                                 long long ldx, const float* __restrict__ inv_norm,
                                                                 float* __restrict__ dx, int accumulate, int rows, int D,
                                                                 float eps, const float* __restrict__ addend,
                                                                 const float* __restrict__ addend_scale, int dy_partials) {
  const int lane = threadIdx.x & 31;
  const int nv = D >> 7;
  const float ascale = (addend && addend_scale) ? *addend_scale : 1.0f;
  const long long pstride = static_cast<long long>(rows) * D;     // dy = sum of dy_partials partial sums, pstride apart
  for (long long row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * (ROW_THREADS / 32)) {
    const float inv = inv_norm[row];
    const bool clamped = inv >= 1.0f / eps;
    float4 g[MAX_V], yh[MAX_V];
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        g[i] = ld4(dy + row * D + i * 128 + lane * 4);
        for (int s = 1; s < dy_partials; ++s) {
          const float4 t = ld4(dy + s * pstride + row * D + i * 128 + lane * 4);
          g[i].x += t.x; g[i].y += t.y; g[i].z += t.z; g[i].w += t.w;
        }
        const float4 xv = ld4(x + row * ldx + i * 128 + lane * 4);
        yh[i] = make_float4(xv.x * inv, xv.y * inv, xv.z * inv, xv.w * inv);
        dot += g[i].x * yh[i].x + g[i].y * yh[i].y + g[i].z * yh[i].z + g[i].w * yh[i].w;
      }
    dot = clamped ? 0.f : warp_sum(dot);
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        float4 o = make_float4(inv * (g[i].x - yh[i].x * dot), inv * (g[i].y - yh[i].y * dot),
                               inv * (g[i].z - yh[i].z * dot), inv * (g[i].w - yh[i].w * dot));
        float* d = dx + row * D + i * 128 + lane * 4;
        if (accumulate) {
          const float4 old = ld4(d);
          o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
        }
        if (addend) {
          const float4 a = ld4(addend + row * D + i * 128 + lane * 4);
          o.x += a.x * ascale; o.y += a.y * ascale; o.z += a.z * ascale; o.w += a.w * ascale;
        }
        st4(d, o);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward (0426/train.py:95, eps 1e-5, biased variance) fused with the optional L2-normalised bf16 copy
// that feeds the InfoNCE kernels.  bytes/row: D*4 read + D*4 (y f32) + D*2 (yhat bf16) + 12.
// ------------------------------------------------------------------------------------------------
template <int MAX_V>
__global__ void __launch_bounds__(ROW_THREADS) layernorm_fwd_kernel(const float* __restrict__ z,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta,
                                                                    float* __restrict__ y_f32,
                                                                    __nv_bfloat16* __restrict__ yhat_bf16,
                                                                    float* __restrict__ mean_out,
                                                                    float* __restrict__ rstd_out,
                                                                    float* __restrict__ inv_norm_out, int rows, int D,
                                                                    float ln_eps, float l2_eps) {
  const int lane = threadIdx.x & 31;
  const int nv = D >> 7;
  for (long long row = blockIdx.x * (ROW_THREADS / 32) + (threadIdx.x >> 5); row < rows;
       row += static_cast<long long>(gridDim.x) * (ROW_THREADS / 32)) {
    float4 v[MAX_V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        v[i] = ld4(z + row * D + i * 128 + lane * 4);
        s += v[i].x + v[i].y + v[i].z + v[i].w;
      }
    const float mu = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
        q += a * a + b * b + c * c + d * d;
      }
    const float rstd = rsqrtf(warp_sum(q) / D + ln_eps);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const float4 g = ld4(gamma + i * 128 + lane * 4);
        const float4 b = ld4(beta + i * 128 + lane * 4);
        v[i] = make_float4((v[i].x - mu) * rstd * g.x + b.x, (v[i].y - mu) * rstd * g.y + b.y,
                           (v[i].z - mu) * rstd * g.z + b.z, (v[i].w - mu) * rstd * g.w + b.w);
        ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        if (y_f32) st4(y_f32 + row * D + i * 128 + lane * 4, v[i]);
      }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mu;
      if (rstd_out) rstd_out[row] = rstd;
    }
    if (yhat_bf16) {
      const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), l2_eps);
      if (lane == 0 && inv_norm_out) inv_norm_out[row] = inv;
#pragma unroll
      for (int i = 0; i < MAX_V; ++i)
        if (i < nv)
          st4(yhat_bf16 + row * D + i * 128 + lane * 4,
              make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
    }
  }
}

// LayerNorm backward: dz = rstd * (g*dy - mean(g*dy) - zhat * mean(g*dy*zhat)); per-block partial dgamma/dbeta and
// column sums of dz (the bias gradient of the Linear feeding the residual sum), reduced by reduce_partials_kernel.
// Outputs dz both as f32 (residual branch + bias grads) and bf16 (operand of the dW2 / dh GEMMs).
// L2F = true fuses the backward of the L2 normalisation in front: dy is not read but formed per row from the gradient
// w.r.t. the normalised features (given as `l2.parts` partial sums), the bf16 normalised features and 1/||y||:
//   dy = inv (g - yhat (yhat . g)) [+ addend * *addend_scale]         (saves writing and re-reading dy: 134 MB at B = 32768)
struct L2BwdArgs {
  const float* dyhat; int parts;      // [parts][rows][D] partial sums of d loss / d yhat
  const __nv_bfloat16* yhat;          // [rows][D]
  const float* inv_norm;              // [rows]
  float eps;                          // F.normalize eps: a clamped norm has zero gradient through the norm
  const float* addend;                // [rows][D] f32 gradient that reaches y directly, or null
  const float* addend_scale;          // device scalar or null
};

template <int MAX_V, bool L2F>
__global__ void __launch_bounds__(ROW_THREADS, MAX_V <= 4 ? 2 : 1) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd,
                                                                    const float* __restrict__ gamma,
                                                                    float* __restrict__ dz_f32,
                                                                    __nv_bfloat16* __restrict__ dz_bf16,
                                                                    float* __restrict__ partial /*[grid][3][D]*/, int rows,
                                                                    int D, float drop_p, unsigned int drop_seed_arg,
                                                                    const unsigned int* __restrict__ drop_seed_dev, const L2BwdArgs l2) {
  const unsigned int drop_seed = drop_seed_arg + ((drop_p > 0.f && drop_seed_dev) ? __ldg(drop_seed_dev) : 0u);
  // per-warp column accumulators (dgamma, dbeta, column sum of dz) live in shared memory, [warp][3][D]: each lane owns its
  // columns, so the read-modify-write needs no synchronisation, and 48 registers per thread are free for loads in flight
  extern __shared__ __align__(16) float ln_acc[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nv = D >> 7;
  float* my_acc = ln_acc + static_cast<size_t>(warp) * 3 * D;
  for (int i = lane * 4; i < 3 * D; i += 128) *reinterpret_cast<float4*>(my_acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  auto acc_add = [&](int which, int i, float4 v) {
    float4* q = reinterpret_cast<float4*>(my_acc + which * D + i * 128 + lane * 4);
    float4 t = *q;
    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    *q = t;
  };
  for (long long row = blockIdx.x * (ROW_THREADS / 32) + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * (ROW_THREADS / 32)) {
    const float mu = mean[row], rs = rstd[row];
    float4 g[MAX_V], zh[MAX_V];
    float s1 = 0.f, s2 = 0.f;
    float4 dyv[MAX_V], zrow[MAX_V];
    // every global load of the row is issued before the first reduction: one DRAM round trip per row instead of two (the z
    // loads used to sit behind the warp reduction of the L2 backward: 2.9 TB/s at B = 32768, long-scoreboard bound)
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) zrow[i] = ld4(z + row * D + i * 128 + lane * 4);
    if constexpr (L2F) {
      const float inv = l2.inv_norm[row];
      const bool clamped = inv >= 1.0f / l2.eps;
      const long long pstride = static_cast<long long>(rows) * D;
      const float ascale = (l2.addend && l2.addend_scale) ? *l2.addend_scale : 1.0f;
      float4 yh[MAX_V], ad[MAX_V];
      float dot = 0.f;
      if (l2.addend) {
#pragma unroll
        for (int i = 0; i < MAX_V; ++i)
          if (i < nv) ad[i] = ld4(l2.addend + row * D + i * 128 + lane * 4);
      }
#pragma unroll
      for (int i = 0; i < MAX_V; ++i)
        if (i < nv) {
          float4 t = ld4(l2.dyhat + row * D + i * 128 + lane * 4);
#pragma unroll 1
          for (int sp = 1; sp < l2.parts; ++sp) {
            const float4 u = ld4(l2.dyhat + sp * pstride + row * D + i * 128 + lane * 4);
            t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
          }
          dyv[i] = t;
          yh[i] = ld4(l2.yhat + row * D + i * 128 + lane * 4);
          dot += t.x * yh[i].x + t.y * yh[i].y + t.z * yh[i].z + t.w * yh[i].w;
        }
      dot = clamped ? 0.f : warp_sum(dot);
#pragma unroll
      for (int i = 0; i < MAX_V; ++i)
        if (i < nv) {
          float4 o = make_float4(inv * (dyv[i].x - yh[i].x * dot), inv * (dyv[i].y - yh[i].y * dot),
                                 inv * (dyv[i].z - yh[i].z * dot), inv * (dyv[i].w - yh[i].w * dot));
          if (l2.addend) {
            const float4 a = ad[i];
            o.x += a.x * ascale; o.y += a.y * ascale; o.z += a.z * ascale; o.w += a.w * ascale;
          }
          dyv[i] = o;
        }
    }
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        float4 d;
        if constexpr (L2F) d = dyv[i]; else d = ld4(dy + row * D + i * 128 + lane * 4);
        const float4 zz = zrow[i];
        const float4 gm = ld4(gamma + i * 128 + lane * 4);
        zh[i] = make_float4((zz.x - mu) * rs, (zz.y - mu) * rs, (zz.z - mu) * rs, (zz.w - mu) * rs);
        acc_add(0, i, make_float4(d.x * zh[i].x, d.y * zh[i].y, d.z * zh[i].z, d.w * zh[i].w));
        acc_add(1, i, d);
        g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * zh[i].x + g[i].y * zh[i].y + g[i].z * zh[i].z + g[i].w * zh[i].w;
      }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const float4 o = make_float4(rs * (g[i].x - s1 - zh[i].x * s2), rs * (g[i].y - s1 - zh[i].y * s2),
                                     rs * (g[i].z - s1 - zh[i].z * s2), rs * (g[i].w - s1 - zh[i].w * s2));
        if (dz_f32) st4(dz_f32 + row * D + i * 128 + lane * 4, o);             // residual branch: unmasked
        float4 m = o;                                                          // fc branch: through the dropout mask
        if (drop_p > 0.f) {
          const float sc = 1.0f / (1.0f - drop_p);
          const uint32_t c0 = i * 128 + lane * 4;
          m.x = dropout_keep(drop_seed, (uint32_t)row, c0, (uint32_t)D, drop_p) ? o.x * sc : 0.f;
          m.y = dropout_keep(drop_seed, (uint32_t)row, c0 + 1, (uint32_t)D, drop_p) ? o.y * sc : 0.f;
          m.z = dropout_keep(drop_seed, (uint32_t)row, c0 + 2, (uint32_t)D, drop_p) ? o.z * sc : 0.f;
          m.w = dropout_keep(drop_seed, (uint32_t)row, c0 + 3, (uint32_t)D, drop_p) ? o.w * sc : 0.f;
        }
        acc_add(2, i, m);
        if (dz_bf16) st4(dz_bf16 + row * D + i * 128 + lane * 4, m);
      }
  }
  // block reduction of the per-warp partials (fixed warp order: deterministic)
  __syncthreads();
  float* pg = partial + static_cast<long long>(blockIdx.x) * 3 * D;
  for (int c = threadIdx.x; c < 3 * D; c += ROW_THREADS) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < ROW_THREADS / 32; ++w) acc += ln_acc[static_cast<size_t>(w) * 3 * D + c];
    pg[c] = acc;
  }
}

// out_k[i] (+)= sum over `nparts` rows of partial[part][k*seg + i]   (final stage of the two-stage column reductions)
// Block = 32 columns x 8 row-slices: coalesced 128-B reads, the 8 slices are folded through shared memory.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, long long part_stride,
                                                              int nparts, int n, int seg, float* __restrict__ out0,
                                                              float* __restrict__ out1, float* __restrict__ out2,
                                                              int accumulate) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (col < n)
    for (int p = sl; p < nparts; p += 8) acc += partial[p * part_stride + col];
  red[sl][cx] = acc;
  __syncthreads();
  if (sl == 0 && col < n) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][cx];
    const int which = col / seg, i = col - which * seg;
    float* o = which == 0 ? out0 : (which == 1 ? out1 : out2);
    if (o) o[i] = accumulate ? o[i] + v : v;
  }
}

// column sums of a [rows, N] matrix (bias gradients), stage 1: warp per row, lanes across columns (128-bit loads),
// per-lane register accumulators, block fold through shared memory -> partial[block][N]
template <typename T, int MAX_V>
__global__ void __launch_bounds__(ROW_THREADS) colsum_rows_kernel(const T* __restrict__ a, long long lda, int rows, int N,
                                                                  float* __restrict__ partial) {
  __shared__ float red[ROW_THREADS / 32][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = N >> 7;
  float4 acc[MAX_V];
#pragma unroll
  for (int i = 0; i < MAX_V; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row = blockIdx.x * (ROW_THREADS / 32) + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * (ROW_THREADS / 32)) {
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        const float4 v = ld4(a + row * lda + i * 128 + lane * 4);
        acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
      }
  }
  float* pg = partial + static_cast<long long>(blockIdx.x) * N;
  for (int i = 0; i < nv; ++i) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < MAX_V; ++k)
      if (k == i) v = acc[k];
    __syncthreads();
    *reinterpret_cast<float4*>(&red[warp][lane * 4]) = v;
    __syncthreads();
    if (threadIdx.x < 128) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < ROW_THREADS / 32; ++w) t += red[w][threadIdx.x];
      pg[i * 128 + threadIdx.x] = t;
    }
  }
}

__global__ void dropout_mask_kernel(float* __restrict__ out, long long rows, int cols, float p, unsigned int seed) {
  const long long n = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = dropout_keep(seed, static_cast<uint32_t>(i / cols), static_cast<uint32_t>(i % cols), static_cast<uint32_t>(cols), p)
                 ? 1.0f / (1.0f - p) : 0.f;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    st4(out + i * 4, ld4(in + i * 4));
}

static inline int part_grid(long long rows) {      // kernels that emit per-block partials: keep the partial count small
  const long long want = (rows + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32);
  const long long cap = static_cast<long long>(num_sms()) * 3;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}
static inline int row_grid(long long rows) {
  const long long want = (rows + ROW_THREADS / 32 - 1) / (ROW_THREADS / 32);
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200clip_l2norm_fwd(const void* x, int x_is_bf16, long long ldx, void* y_bf16, float* y_f32,
                                   float* inv_norm, long long rows, int D, float eps, void* stream) {
  B200_REQUIRE(rows >= 0 && D > 0 && D % 128 == 0 && D <= MAX_D, "l2norm_fwd: D=%d must be a multiple of 128, <= %d", D, MAX_D);
  B200_REQUIRE(aligned16(x) && (y_bf16 == nullptr || aligned16(y_bf16)) && (y_f32 == nullptr || aligned16(y_f32)),
               "l2norm_fwd: pointers must be 16-byte aligned");
  B200_REQUIRE(ldx % 4 == 0, "l2norm_fwd: ldx must be a multiple of 4");
  if (rows == 0) return B200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_is_bf16)
    B200_DISPATCH_V(D, (l2norm_fwd_kernel<__nv_bfloat16, MAX_V><<<row_grid(rows), ROW_THREADS, 0, s>>>(
        static_cast<const __nv_bfloat16*>(x), ldx, static_cast<__nv_bfloat16*>(y_bf16), y_f32, inv_norm, (int)rows, D, eps)));
  else
    B200_DISPATCH_V(D, (l2norm_fwd_kernel<float, MAX_V><<<row_grid(rows), ROW_THREADS, 0, s>>>(
        static_cast<const float*>(x), ldx, static_cast<__nv_bfloat16*>(y_bf16), y_f32, inv_norm, (int)rows, D, eps)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_l2norm_bwd(const float* dy, int dy_partials, const void* x, int x_is_bf16, long long ldx,
                                   const float* inv_norm, float* dx, int accumulate, long long rows, int D, float eps,
                                   const float* addend, const float* addend_scale, void* stream) {
  B200_REQUIRE(dy_partials >= 1 && dy_partials <= 8, "l2norm_bwd: dy_partials=%d out of range", dy_partials);
  B200_REQUIRE(rows >= 0 && D > 0 && D % 128 == 0 && D <= MAX_D, "l2norm_bwd: D=%d must be a multiple of 128, <= %d", D, MAX_D);
  B200_REQUIRE(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(addend), "l2norm_bwd: pointers must be 16-byte aligned");
  if (rows == 0) return B200_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (x_is_bf16)
    B200_DISPATCH_V(D, (l2norm_bwd_kernel<__nv_bfloat16, MAX_V><<<row_grid(rows), ROW_THREADS, 0, s>>>(
        dy, static_cast<const __nv_bfloat16*>(x), ldx, inv_norm, dx, accumulate, (int)rows, D, eps, addend, addend_scale, dy_partials)));
  else
    B200_DISPATCH_V(D, (l2norm_bwd_kernel<float, MAX_V><<<row_grid(rows), ROW_THREADS, 0, s>>>(
        dy, static_cast<const float*>(x), ldx, inv_norm, dx, accumulate, (int)rows, D, eps, addend, addend_scale, dy_partials)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_layernorm_fwd(const float* z, const float* gamma, const float* beta, float* y_f32, void* yhat_bf16,
                                      float* mean, float* rstd, float* inv_norm, long long rows, int D, float ln_eps,
                                      float l2_eps, void* stream) {
  B200_REQUIRE(rows >= 0 && D > 0 && D % 128 == 0 && D <= MAX_D, "layernorm_fwd: D=%d must be a multiple of 128, <= %d", D, MAX_D);
  B200_REQUIRE(aligned16(z) && aligned16(gamma) && aligned16(beta), "layernorm_fwd: pointers must be 16-byte aligned");
  if (rows == 0) return B200_OK;
  B200_DISPATCH_V(D, (layernorm_fwd_kernel<MAX_V><<<row_grid(rows), ROW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      z, gamma, beta, y_f32, static_cast<__nv_bfloat16*>(yhat_bf16), mean, rstd, inv_norm, (int)rows, D, ln_eps, l2_eps)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" size_t b200clip_layernorm_bwd_workspace_bytes(long long rows, int D) {
  return static_cast<size_t>(part_grid(rows)) * 3 * D * sizeof(float);
}

extern "C" int b200clip_layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd,
                                      const float* gamma, float* dz_f32, void* dz_bf16, float* dgamma, float* dbeta,
                                      float* dz_colsum, int accumulate_params, long long rows, int D, float drop_p,
                                      unsigned int drop_seed, const unsigned int* drop_seed_dev, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  B200_REQUIRE(rows > 0 && D > 0 && D % 128 == 0 && D <= MAX_D, "layernorm_bwd: D=%d must be a multiple of 128, <= %d", D, MAX_D);
  const int grid = std::min(part_grid(rows), (D <= 512 ? 2 : 1) * num_sms());   // resident CTAs per SM (register-bound): one wave
  if (workspace_bytes < static_cast<size_t>(grid) * 3 * D * sizeof(float))
    return fail(B200_ERR_WORKSPACE, "layernorm_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  const size_t ln_smem = static_cast<size_t>(ROW_THREADS / 32) * 3 * D * sizeof(float);
  B200_DISPATCH_V(D, (cudaFuncSetAttribute(layernorm_bwd_kernel<MAX_V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ln_smem)));
  B200_DISPATCH_V(D, (layernorm_bwd_kernel<MAX_V, false><<<grid, ROW_THREADS, ln_smem, s>>>(
      dy, z, mean, rstd, gamma, dz_f32, static_cast<__nv_bfloat16*>(dz_bf16), partial, (int)rows, D, drop_p, drop_seed, drop_seed_dev, L2BwdArgs{})));
  B200_LAUNCH_CHECK();
  reduce_partials_kernel<<<(3 * D + 31) / 32, 256, 0, s>>>(partial, 3LL * D, grid, 3 * D, D, dgamma, dbeta, dz_colsum,
                                                         accumulate_params);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// LayerNorm backward with the backward of the L2 normalisation (F.normalize of the LayerNorm output) fused in front.
extern "C" int b200clip_layernorm_l2_bwd(const float* dyhat, int dyhat_partials, const void* yhat_bf16, const float* inv_norm,
                                         float l2_eps, const float* addend, const float* addend_scale, const float* z,
                                         const float* mean, const float* rstd, const float* gamma, float* dz_f32,
                                         void* dz_bf16, float* dgamma, float* dbeta, float* dz_colsum, int accumulate_params,
                                         long long rows, int D, float drop_p, unsigned int drop_seed,
                                         const unsigned int* drop_seed_dev, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  B200_REQUIRE(rows > 0 && D > 0 && D % 128 == 0 && D <= MAX_D, "layernorm_l2_bwd: D=%d must be a multiple of 128, <= %d", D, MAX_D);
  B200_REQUIRE(dyhat && yhat_bf16 && inv_norm && dyhat_partials >= 1 && dyhat_partials <= 8, "layernorm_l2_bwd: missing arguments");
  B200_REQUIRE(aligned16(dyhat) && aligned16(yhat_bf16) && aligned16(addend) && aligned16(z), "layernorm_l2_bwd: pointers must be 16-byte aligned");
  const int grid = std::min(part_grid(rows), (D <= 512 ? 2 : 1) * num_sms());   // resident CTAs per SM (register-bound): one wave
  if (workspace_bytes < static_cast<size_t>(grid) * 3 * D * sizeof(float))
    return fail(B200_ERR_WORKSPACE, "layernorm_l2_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  L2BwdArgs l2{dyhat, dyhat_partials, static_cast<const __nv_bfloat16*>(yhat_bf16), inv_norm, l2_eps, addend, addend_scale};
  const size_t ln_smem = static_cast<size_t>(ROW_THREADS / 32) * 3 * D * sizeof(float);
  B200_DISPATCH_V(D, (cudaFuncSetAttribute(layernorm_bwd_kernel<MAX_V, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ln_smem)));
  B200_DISPATCH_V(D, (layernorm_bwd_kernel<MAX_V, true><<<grid, ROW_THREADS, ln_smem, s>>>(
      nullptr, z, mean, rstd, gamma, dz_f32, static_cast<__nv_bfloat16*>(dz_bf16), partial, (int)rows, D, drop_p, drop_seed, drop_seed_dev, l2)));
  B200_LAUNCH_CHECK();
  reduce_partials_kernel<<<(3 * D + 31) / 32, 256, 0, s>>>(partial, 3LL * D, grid, 3 * D, D, dgamma, dbeta, dz_colsum,
                                                         accumulate_params);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" size_t b200clip_colsum_workspace_bytes(long long rows, int N) {
  return static_cast<size_t>(part_grid(rows)) * N * sizeof(float) + 256;
}

extern "C" int b200clip_colsum(const void* a, int a_is_bf16, long long lda, long long rows, int N, float* out,
                               int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(rows > 0 && N > 0 && N % 128 == 0 && N <= 2048, "colsum: N=%d must be a multiple of 128, <= 2048", N);
  B200_REQUIRE(aligned16(a) && lda % 8 == 0, "colsum: operand must be 16-byte aligned with lda %% 8 == 0");
  const int grid = part_grid(rows);
  if (workspace_bytes < static_cast<size_t>(grid) * N * sizeof(float)) return fail(B200_ERR_WORKSPACE, "colsum: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
#define B200_COLSUM_LAUNCH(T, V) colsum_rows_kernel<T, V><<<grid, ROW_THREADS, 0, s>>>(static_cast<const T*>(a), lda, (int)rows, N, partial)
  if (a_is_bf16) {
    if (N <= 512) B200_COLSUM_LAUNCH(__nv_bfloat16, 4); else if (N <= 1024) B200_COLSUM_LAUNCH(__nv_bfloat16, 8); else B200_COLSUM_LAUNCH(__nv_bfloat16, 16);
  } else {
    if (N <= 512) B200_COLSUM_LAUNCH(float, 4); else if (N <= 1024) B200_COLSUM_LAUNCH(float, 8); else B200_COLSUM_LAUNCH(float, 16);
  }
#undef B200_COLSUM_LAUNCH
  B200_LAUNCH_CHECK();
  reduce_partials_kernel<<<(N + 31) / 32, 256, 0, s>>>(partial, N, grid, N, N, out, nullptr, nullptr, accumulate);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

__global__ void dropout_seed_advance_kernel(unsigned int* seed) { *seed = *seed * 1664525u + 1013904223u; }

// advances a device-resident dropout seed word (Numerical Recipes LCG); one node of a captured step graph, so that every
// replay applies a fresh keep-mask although the kernels' launch arguments are frozen
extern "C" int b200clip_dropout_seed_advance(unsigned int* seed_dev, void* stream) {
  B200_REQUIRE(seed_dev != nullptr, "dropout_seed_advance: null pointer");
  dropout_seed_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(seed_dev);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// the scaled keep-mask the fused kernels apply for (seed, p): mask[r][c] = keep ? 1/(1-p) : 0   (tests / debugging)
extern "C" int b200clip_dropout_mask(float* out, long long rows, int cols, float p, unsigned int seed, void* stream) {
  B200_REQUIRE(rows > 0 && cols > 0 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
  const long long n = rows * cols;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, static_cast<long long>(num_sms()) * 16));
  dropout_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, rows, cols, p, seed);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// AdamW for the head parameters (the reference trains with torch.optim.AdamW(lr=1e-4, weight_decay=0.01)).
// torch.optim.AdamW semantics (decoupled decay, bias-corrected moments, amsgrad off).  The step count lives on the device
// (adamw_tick increments it once per optimizer step), so the update can sit inside a captured CUDA graph.
// ------------------------------------------------------------------------------------------------------------------
__global__ void adamw_tick_kernel(float* __restrict__ step) { *step += 1.0f; }

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                                    float weight_decay, const float* __restrict__ step) {
  __shared__ float s_bc[2];
  if (threadIdx.x == 0) {                                    // bias corrections in double (1 - 0.999^t cancels badly in fp32)
    const double t = static_cast<double>(*step);
    s_bc[0] = static_cast<float>(1.0 - pow(static_cast<double>(beta1), t));
    s_bc[1] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), t)));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const float gi = g[i];
    float pi = p[i] * (1.0f - lr * weight_decay);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    pi -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi;
  }
}

extern "C" int b200clip_adamw_tick(float* step, void* stream) {
  B200_REQUIRE(step != nullptr, "adamw_tick: missing step counter");
  adamw_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(step);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, const float* step, void* stream) {
  B200_REQUIRE(n >= 0 && param && grad && exp_avg && exp_avg_sq && step, "adamw_step: missing arguments");
  B200_REQUIRE(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adamw_step: bad hyper-parameters");
  if (n == 0) return B200_OK;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, static_cast<long long>(num_sms()) * 8));
  adamw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                    weight_decay, step);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// strided 2-D cast (concatenating two [rows, cols] f32 views side by side into one bf16 matrix = two calls)
__global__ void __launch_bounds__(256) cast2d_f32_bf16_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
                                                              long long ld_out, long long rows, int cols4) {
  const long long n = rows * cols4;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) {
    const long long r = i / cols4;
    const int c = static_cast<int>(i - r * cols4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + r * ld_in + c);
    *reinterpret_cast<uint2*>(out + r * ld_out + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

extern "C" int b200clip_cast_f32_bf16_2d(const float* in, long long ld_in, void* out, long long ld_out, long long rows, int cols,
                                         void* stream) {
  B200_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && aligned16(in) &&
               (reinterpret_cast<uintptr_t>(out) & 7u) == 0, "cast_2d: cols / leading dimensions must be multiples of 4, pointers aligned");
  if (rows == 0) return B200_OK;
  const long long n = rows * (cols / 4);
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, static_cast<long long>(num_sms()) * 16));
  cast2d_f32_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, ld_in, static_cast<__nv_bfloat16*>(out), ld_out, rows, cols / 4);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
  B200_REQUIRE(n >= 0 && n % 4 == 0 && aligned16(in) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0,
               "cast: n must be a multiple of 4 and pointers aligned");
  if (n == 0) return B200_OK;
  const long long n4 = n / 4;
  const int grid = static_cast<int>(std::min<long long>((n4 + 255) / 256, static_cast<long long>(num_sms()) * 16));
  cast_f32_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, static_cast<__nv_bfloat16*>(out), n4);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// contract check of the flash InfoNCE path: flag |= 1 when some row of x is not a unit vector (| ||x||^2 - 1 | > tol) or is
// non-finite.  One warp per row, fp32 input.
namespace b200 {
__global__ void __launch_bounds__(256) rows_unit_check_kernel(const float* __restrict__ x, long long rows, int D, float tol,
                                                              int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const long long w = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const long long nw = (static_cast<long long>(gridDim.x) * 256) >> 5;
  for (long long row = w; row < rows; row += nw) {
    const float* p = x + row * D;
    float ss = 0.f;
    for (int i = lane; i < D; i += 32) { const float v = p[i]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0 && !(fabsf(ss - 1.f) <= tol)) atomicOr(flag, 1);
  }
}
}  // namespace b200

extern "C" int b200clip_rows_unit_check(const float* x, long long rows, int D, float tol, int* flag, void* stream) {
  B200_REQUIRE(x && flag && rows > 0 && D > 0, "rows_unit_check: bad arguments");
  const int grid = static_cast<int>(std::min<long long>((rows + 7) / 8, 4ll * b200::num_sms()));
  b200::rows_unit_check_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, rows, D, tol, flag);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
