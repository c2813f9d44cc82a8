// b200clip: the "small-C" members of the CLIP head -- every op whose right-hand side is a handful (C <= 32) of
// class vectors: multi-label BCE on sigmoid(cos/tau) (a-B, 0426/train.py:178-230), the FC classification adapter
// with BCE-with-logits (a-A, NB02 c28:50-52 / c29:23-25) and in-loop prediction (a-M, 0426/train.py:869-886).
// All are HBM-bound over the [B, D] feature matrix: one warp per row, 128-bit coalesced loads, class vectors
// resident in shared memory, scores reduced with a 31-shuffle halving exchange so lane c owns class c.
#include "common.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

constexpr int SC_THREADS = 256;
constexpr int SC_MAXC = 32;

__device__ __forceinline__ float warp_colsum32_sc(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// SC_HEAD2: both BCE heads of the fused step in ONE pass over the features: classes [0, c1) are the multi-label
// class texts (a-B), classes [c1, C) the FC adapter rows (a-A).
enum ScMode : int { SC_MLBCE = 0, SC_FCBCE = 1, SC_PREDICT = 2, SC_HEAD2 = 3 };

struct ScParams {
  const float* x; long long ldx;      // [B, D] features (f32)
  const float* cls;                   // [C, D] class vectors (f32)   (HEAD2: the c1 class texts)
  const float* cls2; int c1;          // HEAD2: [C - c1, D] FC adapter weight; c1 = number of class texts
  double total_elems2;                // HEAD2: mean divisor of the FC BCE
  const float* bias;                  // [C] (FC) or null
  const float* labels; int label_cols; long long ld_labels;   // [B, label_cols] (missing classes read as 0)
  int B, C, D;
  float inv_tau;                      // 1/tau (MLBCE, PREDICT)
  float threshold;
  int normalize_x, normalize_cls;     // F.normalize the operands first
  const float* label_sum;             // MLBCE: device scalar sum(labels) over the GLOBAL batch
  double total_elems;                 // MLBCE: B_glob * C ; FCBCE: B_glob * C (mean divisor)
  const float* grad_scale;            // optional upstream scalar
  // outputs
  float* dx; int dx_accumulate;       // [B, D] or null
  float* coef;                        // [B, C] d loss / d score (for the class-vector / weight gradients) or null
  float* pred;                        // [B, C] {0,1} or null
  float* logits;                      // [B, C] raw scores (FC: z = xW^T+b) or null
  float* xinv;                        // [B] 1/||x|| (MLBCE with normalize_x) or null
  double* partial;                    // [grid][3]
  unsigned int* counter;
  double* sums;                       // [3]: MLBCE pos, neg numerators ; FC BCE sum (HEAD2) -- modes use what they need
  float* loss2;                       // HEAD2: FC loss
  float* loss;                        // [1] or null (single-rank finalisation)
  int* status;                        // non-finite / >1000 guard flag (0426/train.py:224) or null
};

template <int MODE, int MAX_V>
__global__ void __launch_bounds__(SC_THREADS) smallc_kernel(const ScParams p) {
  extern __shared__ float s_cls[];                      // [C][D]
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = p.D >> 7;
  const int C = p.C, D = p.D;

  // stage (and optionally L2-normalise, F.normalize eps 1e-12) the class vectors
  const int c1 = (MODE == SC_HEAD2) ? p.c1 : C;
  for (int c = warp; c < C; c += SC_THREADS / 32) {
    const float* src = (c < c1) ? p.cls + static_cast<long long>(c) * D : p.cls2 + static_cast<long long>(c - c1) * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float v = src[d];
      ss += v * v;
    }
    ss = warp_sum(ss);
    const float inv = (p.normalize_cls && c < c1) ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
    for (int d = lane; d < D; d += 32) s_cls[c * D + d] = src[d] * inv;
  }
  __syncthreads();

  float gscale = 1.0f;
  if (p.grad_scale) gscale = *p.grad_scale;
  float Psum = 0.f, Nsum = 0.f;
  if (MODE == SC_MLBCE || MODE == SC_HEAD2) {
    Psum = *p.label_sum;
    Nsum = static_cast<float>(p.total_elems - static_cast<double>(Psum));
  }
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;

  const long long row_stride = static_cast<long long>(gridDim.x) * (SC_THREADS / 32);
  long long row = blockIdx.x * (SC_THREADS / 32) + warp;
  float4 xn[MAX_V];                                       // software prefetch of the next row
#pragma unroll
  for (int i = 0; i < MAX_V; ++i)
    if (i < nv && row < p.B) xn[i] = *reinterpret_cast<const float4*>(p.x + row * p.ldx + i * 128 + lane * 4);
  for (; row < p.B; row += row_stride) {
    float4 xv[MAX_V];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V; ++i)
      if (i < nv) {
        xv[i] = xn[i];
        ss += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
      }
    if (row + row_stride < p.B) {
#pragma unroll
      for (int i = 0; i < MAX_V; ++i)
        if (i < nv) xn[i] = *reinterpret_cast<const float4*>(p.x + (row + row_stride) * p.ldx + i * 128 + lane * 4);
    }
    float inv = 1.0f;
    if (p.normalize_x) {
      ss = warp_sum(ss);
      inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      if (lane == 0 && p.xinv) p.xinv[row] = inv;
    }
    float part[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      part[c] = 0.f;
      if (c < C) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < MAX_V; ++i)
          if (i < nv) {
            const float4 t = *reinterpret_cast<const float4*>(&s_cls[c * D + i * 128 + lane * 4]);
            a += xv[i].x * t.x + xv[i].y * t.y + xv[i].z * t.z + xv[i].w * t.w;
          }
        part[c] = a;
      }
    }
    const float dot_raw = warp_colsum32_sc(part, lane);
    const float dot = (MODE == SC_HEAD2 && lane >= c1) ? dot_raw : dot_raw * inv;   // lane c: <x_hat, cls_c> (FC rows: raw x)
    const bool active = lane < C;
    float y = 0.f;
    if (MODE == SC_HEAD2) {
      const int lc = lane < c1 ? lane : lane - c1;
      if (active && lc < p.label_cols) y = p.labels[row * p.ld_labels + lc];
    } else if (active && p.labels && lane < p.label_cols) y = p.labels[row * p.ld_labels + lane];
    float coef = 0.f;                                            // d loss / d score_c (score = logit fed to sigmoid)
    if (MODE == SC_MLBCE || (MODE == SC_HEAD2 && lane < c1)) {
      const float s = dot * p.inv_tau;                           // :195
      const float sc = fminf(fmaxf(s, -50.f), 50.f);             // :213
      const float pp = 1.0f / (1.0f + expf(-sc));              // :214
      const float qq = 1.0f - pp;                                // :215
      if (active) {
        acc0 += static_cast<double>(logf(pp + 1e-8f) * y);       // :218 numerator
        acc1 += static_cast<double>(logf(qq + 1e-8f) * (1.0f - y));   // :219 numerator
        const float inside = (fabsf(s) <= 50.f) ? 1.f : 0.f;
        const float dpos = -y * pp * qq / ((pp + 1e-8f) * (Psum + 1e-8f));
        const float dneg = (1.0f - y) * pp * qq / ((qq + 1e-8f) * (Nsum + 1e-8f));
        coef = 0.5f * (dpos + dneg) * inside * gscale;
      }
    } else if (MODE == SC_FCBCE || MODE == SC_HEAD2) {
      const int fc_idx = (MODE == SC_HEAD2) ? lane - c1 : lane;
      const float z = dot + (active && p.bias ? p.bias[fc_idx] : 0.f);
      if (active) {
        // BCEWithLogits: max(z,0) - z*y + log1p(exp(-|z|))
        const double bce = static_cast<double>(fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z))));
        if (MODE == SC_HEAD2) acc2 += bce; else acc0 += bce;
        const float sg = 1.0f / (1.0f + expf(-z));
        coef = (sg - y) * gscale / static_cast<float>(MODE == SC_HEAD2 ? p.total_elems2 : p.total_elems);
        const int Cf = (MODE == SC_HEAD2) ? C - c1 : C;
        if (p.pred) p.pred[row * Cf + fc_idx] = sg > p.threshold ? 1.f : 0.f;
        if (p.logits) p.logits[row * Cf + fc_idx] = z;
      }
    } else {
      if (active) {
        const float s = dot * p.inv_tau;                         // :881
        const float pr = 1.0f / (1.0f + expf(-s));             // :883
        p.pred[row * C + lane] = pr > p.threshold ? 1.f : 0.f;   // :885
        if (p.labels) {                                          // :441-447 accuracy counters
          acc0 += ((pr > p.threshold ? 1.f : 0.f) == y) ? 1.0 : 0.0;
        }
      }
    }
    if (MODE != SC_PREDICT) {
      if (MODE == SC_HEAD2) {
        if (p.coef && active && lane >= c1) p.coef[row * (C - c1) + (lane - c1)] = coef;    // FC rows only (feeds dW, db)
      } else if (p.coef && active) p.coef[row * C + lane] = coef;
      if (p.dx) {
        // d x_hat = sum_c coef_c * cls_c * (1/tau) ; then through the normalisation
        float4 g[MAX_V], g2[MAX_V];                              // g: through the normalisation ; g2: acts on x itself (HEAD2 FC rows)
#pragma unroll
        for (int i = 0; i < MAX_V; ++i) g[i] = g2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        float sdot = 0.f;                                        // <x_hat, d x_hat>
        const float kt = (MODE == SC_MLBCE || MODE == SC_HEAD2) ? p.inv_tau : 1.0f;
        for (int c = 0; c < c1; ++c) {
          const float cc = __shfl_sync(0xffffffffu, coef, c) * kt;
          const float dc = __shfl_sync(0xffffffffu, dot, c);
          sdot += cc * dc;
#pragma unroll
          for (int i = 0; i < MAX_V; ++i)
            if (i < nv) {
              const float4 t = *reinterpret_cast<const float4*>(&s_cls[c * D + i * 128 + lane * 4]);
              g[i].x += cc * t.x; g[i].y += cc * t.y; g[i].z += cc * t.z; g[i].w += cc * t.w;
            }
        }
        if (MODE == SC_HEAD2) {
          for (int c = c1; c < C; ++c) {
            const float cc = __shfl_sync(0xffffffffu, coef, c);
#pragma unroll
            for (int i = 0; i < MAX_V; ++i)
              if (i < nv) {
                const float4 t = *reinterpret_cast<const float4*>(&s_cls[c * D + i * 128 + lane * 4]);
                g2[i].x += cc * t.x; g2[i].y += cc * t.y; g2[i].z += cc * t.z; g2[i].w += cc * t.w;
              }
          }
        }
#pragma unroll
        for (int i = 0; i < MAX_V; ++i)
          if (i < nv) {
            float4 o = g[i];
            if (p.normalize_x) {
              o.x = inv * (g[i].x - xv[i].x * inv * sdot); o.y = inv * (g[i].y - xv[i].y * inv * sdot);
              o.z = inv * (g[i].z - xv[i].z * inv * sdot); o.w = inv * (g[i].w - xv[i].w * inv * sdot);
            }
            if (MODE == SC_HEAD2) { o.x += g2[i].x; o.y += g2[i].y; o.z += g2[i].z; o.w += g2[i].w; }
            float* d = p.dx + row * D + i * 128 + lane * 4;
            if (p.dx_accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(d);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(d) = o;
          }
      }
    }
  }

  if (p.partial == nullptr) return;
  // deterministic two-level reduction of the loss numerators
  __shared__ double red3[3][SC_THREADS / 32];
  double v[3] = {acc0, acc1, acc2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red3[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < SC_THREADS / 32; ++w) t += red3[threadIdx.x][w];
    p.partial[blockIdx.x * 3 + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) { s0 += p.partial[b * 3]; s1 += p.partial[b * 3 + 1]; s2 += p.partial[b * 3 + 2]; }
    p.sums[0] = s0; p.sums[1] = s1; p.sums[2] = s2;
    if (p.loss) {
      float l;
      if (MODE == SC_MLBCE || MODE == SC_HEAD2) {
        const float pos = static_cast<float>(-s0) / (Psum + 1e-8f);                 // :218
        const float neg = static_cast<float>(-s1) / (Nsum + 1e-8f);                 // :219
        l = (pos + neg) * 0.5f;                                                     // :221
        if (p.status) *p.status = (isnan(l) || isinf(l) || l > 1000.f) ? 1 : 0;      // :224
        if (MODE == SC_HEAD2 && p.loss2) *p.loss2 = static_cast<float>(s2 / p.total_elems2);
      } else {
        l = static_cast<float>(s0 / p.total_elems);                                  // FC mean / mean accuracy
      }
      *p.loss = l;
    }
    *p.counter = 0;
  }
}

// out[c][d] (+)= sum_rows coef[row][c] * x[row][d] * (row_scale[row] if given) ; bias_out[c] (+)= sum_rows coef[row][c]
// Two-stage and deterministic: each block owns a slab of rows, then reduce_partials.
__global__ void __launch_bounds__(256) skinny_outer_partial_kernel(const float* __restrict__ coef, int C,
                                                                   const float* __restrict__ x, long long ldx,
                                                                   const float* __restrict__ row_scale, int rows, int D,
                                                                   int rows_per_block, float* __restrict__ partial /*[grid][C*D + C]*/) {
  extern __shared__ float s_coef[];                    // [rows_per_block][C]
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  const int nr = r1 - r0;
  for (int i = threadIdx.x; i < nr * C; i += blockDim.x) {
    const int r = i / C;
    s_coef[i] = coef[static_cast<long long>(r0) * C + i] * (row_scale ? row_scale[r0 + r] : 1.0f);
  }
  __syncthreads();
  float* out = partial + static_cast<long long>(blockIdx.x) * (static_cast<long long>(C) * D + C);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc[SC_MAXC];
#pragma unroll
    for (int c = 0; c < SC_MAXC; ++c) acc[c] = 0.f;
    int r = 0;
    for (; r + 8 <= nr; r += 8) {
      float xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) xv[u] = x[static_cast<long long>(r0 + r + u) * ldx + d];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int c = 0; c < SC_MAXC; ++c)
          if (c < C) acc[c] += s_coef[(r + u) * C + c] * xv[u];
    }
    for (; r < nr; ++r) {
      const float xv = x[static_cast<long long>(r0 + r) * ldx + d];
#pragma unroll
      for (int c = 0; c < SC_MAXC; ++c)
        if (c < C) acc[c] += s_coef[r * C + c] * xv;
    }
#pragma unroll
    for (int c = 0; c < SC_MAXC; ++c)
      if (c < C) out[static_cast<long long>(c) * D + d] = acc[c];
  }
  if (threadIdx.x < C) {                                // bias grads use the UNSCALED coefficients
    float a = 0.f;
    for (int r = 0; r < nr; ++r) a += coef[static_cast<long long>(r0 + r) * C + threadIdx.x];
    out[static_cast<long long>(C) * D + threadIdx.x] = a;
  }
}

__global__ void __launch_bounds__(256) reduce_partials2_kernel(const float* __restrict__ partial, long long part_stride,
                                                               int nparts, float* __restrict__ out, int n, int accumulate) {
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
    out[col] = accumulate ? out[col] + v : v;
  }
}

// sum of a float vector into one float (label counts); single block, deterministic, 4 x 128-bit loads in flight per thread
__global__ void __launch_bounds__(1024) sum_f32_kernel(const float* __restrict__ a, long long n, float* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  const long long n4 = ((reinterpret_cast<uintptr_t>(a) & 15u) == 0) ? n / 4 : 0;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  long long i = threadIdx.x;
  for (; i + 3 * 1024 < n4; i += 4 * 1024) {
    const float4 v0 = a4[i], v1 = a4[i + 1024], v2 = a4[i + 2048], v3 = a4[i + 3072];
    acc += static_cast<double>((v0.x + v0.y) + (v0.z + v0.w)) + static_cast<double>((v1.x + v1.y) + (v1.z + v1.w)) +
           static_cast<double>((v2.x + v2.y) + (v2.z + v2.w)) + static_cast<double>((v3.x + v3.y) + (v3.z + v3.w));
  }
  for (; i < n4; i += 1024) {
    const float4 v = a4[i];
    acc += static_cast<double>((v.x + v.y) + (v.z + v.w));
  }
  for (long long j = n4 * 4 + threadIdx.x; j < n; j += 1024) acc += static_cast<double>(a[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 32; ++w) s += red[w];
    *out = static_cast<float>(s);
  }
}

static int sc_grid(long long rows) {
  const long long want = (rows + SC_THREADS / 32 - 1) / (SC_THREADS / 32);
  const long long cap = static_cast<long long>(num_sms()) * 4;
  return static_cast<int>(std::max<long long>(1, std::min(want, cap)));
}

template <int MODE>
static int launch_smallc(const ScParams& p, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(p.C) * p.D * sizeof(float);
  const int grid = sc_grid(p.B);
  if (p.D <= 512) {
    auto k = smallc_kernel<MODE, 4>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SC_THREADS, smem, s>>>(p);
  } else {
    auto k = smallc_kernel<MODE, 8>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SC_THREADS, smem, s>>>(p);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static int check_sc(const char* who, const void* x, long long ldx, long long B, int C, int D) {
  B200_REQUIRE(B > 0 && C > 0 && C <= SC_MAXC, "%s: need B>0 and 0 < C <= %d (got B=%lld C=%d)", who, SC_MAXC, B, C);
  B200_REQUIRE(D > 0 && D % 128 == 0 && D <= 1024, "%s: D=%d must be a multiple of 128, <= 1024", who, D);
  B200_REQUIRE(aligned16(x) && ldx % 4 == 0, "%s: x must be 16-byte aligned with ldx %% 4 == 0", who);
  B200_REQUIRE(static_cast<size_t>(C) * D * 4 <= 200 * 1024, "%s: C*D too large for shared memory", who);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200clip_smallc_workspace_bytes(long long rows, int C, int D) {
  const size_t loss_part = static_cast<size_t>(sc_grid(rows)) * 3 * sizeof(double) + 256;
  const int rpb = 64;
  const size_t outer = static_cast<size_t>((rows + rpb - 1) / rpb) * (static_cast<size_t>(C) * D + C) * sizeof(float);
  return loss_part + outer + 256;
}

extern "C" int b200clip_sum_f32(const float* a, long long n, float* out, void* stream) {
  B200_REQUIRE(n >= 0 && out != nullptr, "sum_f32: bad arguments");
  sum_f32_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a, n, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_mlbce_fwd_bwd(const float* image_features, long long ldx, const float* text_features,
                                      const float* labels, int label_cols, long long ld_labels, long long B, int C, int D,
                                      float temperature, const float* label_sum, double total_elems,
                                      const float* grad_scale, float* d_image, int d_image_accumulate, float* coef,
                                      float* x_inv_norm, double* sums, float* loss, int* status, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  int rc = check_sc("mlbce", image_features, ldx, B, C, D);
  if (rc) return rc;
  B200_REQUIRE(temperature > 0.f && labels != nullptr && label_sum != nullptr && sums != nullptr, "mlbce: missing arguments");
  B200_REQUIRE(label_cols > 0, "mlbce: label_cols must be positive");
  if (workspace_bytes < b200clip_smallc_workspace_bytes(B, C, D)) return fail(B200_ERR_WORKSPACE, "mlbce: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ScParams p{};
  p.x = image_features; p.ldx = ldx; p.cls = text_features; p.labels = labels; p.label_cols = std::min(label_cols, C);
  p.ld_labels = ld_labels; p.B = (int)B; p.C = C; p.D = D; p.inv_tau = 1.0f / temperature; p.normalize_x = 1; p.normalize_cls = 1;
  p.label_sum = label_sum; p.total_elems = total_elems; p.grad_scale = grad_scale; p.dx = d_image; p.dx_accumulate = d_image_accumulate;
  p.coef = coef; p.xinv = x_inv_norm; p.partial = static_cast<double*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(sc_grid(B)) * 3 * sizeof(double));
  p.sums = sums; p.loss = loss; p.status = status;
  B200_CHECK_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s));
  return launch_smallc<SC_MLBCE>(p, s);
}

extern "C" int b200clip_fc_bce_fwd_bwd(const float* x, long long ldx, const float* weight, const float* bias,
                                       const float* labels, long long ld_labels, long long B, int C, int D,
                                       double total_elems, float threshold, const float* grad_scale, float* d_x,
                                       int d_x_accumulate, float* coef, float* pred, float* logits, double* sums,
                                       float* loss, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_sc("fc_bce", x, ldx, B, C, D);
  if (rc) return rc;
  B200_REQUIRE(sums != nullptr, "fc_bce: missing arguments");
  if (workspace_bytes < b200clip_smallc_workspace_bytes(B, C, D)) return fail(B200_ERR_WORKSPACE, "fc_bce: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ScParams p{};
  p.x = x; p.ldx = ldx; p.cls = weight; p.bias = bias; p.labels = labels; p.label_cols = C; p.ld_labels = ld_labels;
  p.B = (int)B; p.C = C; p.D = D; p.inv_tau = 1.0f; p.threshold = threshold; p.total_elems = total_elems; p.grad_scale = grad_scale;
  p.dx = d_x; p.dx_accumulate = d_x_accumulate; p.coef = coef; p.pred = pred; p.logits = logits; p.partial = static_cast<double*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(sc_grid(B)) * 3 * sizeof(double));
  p.sums = sums; p.loss = loss;
  B200_CHECK_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s));
  return launch_smallc<SC_FCBCE>(p, s);
}

extern "C" int b200clip_predict_multilabel(const float* image_features, long long ldx, const float* text_features,
                                           long long B, int C, int D, float temperature, float threshold, float* pred,
                                           void* stream) {
  int rc = check_sc("predict_multilabel", image_features, ldx, B, C, D);
  if (rc) return rc;
  B200_REQUIRE(pred != nullptr && temperature > 0.f, "predict_multilabel: missing arguments");
  ScParams p{};
  p.x = image_features; p.ldx = ldx; p.cls = text_features; p.B = (int)B; p.C = C; p.D = D; p.inv_tau = 1.0f / temperature;
  p.threshold = threshold; p.pred = pred;
  return launch_smallc<SC_PREDICT>(p, static_cast<cudaStream_t>(stream));
}

// out_w[C,D] (+)= coef^T (x * row_scale) ; out_b[C] (+)= column sums of coef
extern "C" int b200clip_skinny_outer(const float* coef, int C, const float* x, long long ldx, const float* row_scale,
                                     long long rows, int D, float* out_w, float* out_b, int accumulate, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  B200_REQUIRE(rows > 0 && C > 0 && C <= SC_MAXC && D > 0, "skinny_outer: bad shape");
  if (workspace_bytes < b200clip_smallc_workspace_bytes(rows, C, D)) return fail(B200_ERR_WORKSPACE, "skinny_outer: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rpb = 64;
  const int nblk = static_cast<int>((rows + rpb - 1) / rpb);
  float* partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(sc_grid(rows)) * 3 * sizeof(double) + 256);
  const long long stride = static_cast<long long>(C) * D + C;
  skinny_outer_partial_kernel<<<nblk, 256, rpb * C * sizeof(float), s>>>(coef, C, x, ldx, row_scale, (int)rows, D, rpb, partial);
  B200_LAUNCH_CHECK();
  const int nw = C * D;
  reduce_partials2_kernel<<<(nw + 31) / 32, 256, 0, s>>>(partial, stride, nblk, out_w, nw, accumulate);
  B200_LAUNCH_CHECK();
  if (out_b) {
    reduce_partials2_kernel<<<1, 256, 0, s>>>(partial + nw, stride, nblk, out_b, C, accumulate);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}

// Both BCE heads of the fused step in one pass over the image features (SC_HEAD2): classes [0,c1) = class texts of
// multilabel_contrastive_loss (0426/train.py:178-230), [c1, c1+c2) = rows of the FC adapter (NB02 c28:50-52).
// sums[3] = {MLBCE pos numerator, MLBCE neg numerator, FC BCE sum}; d_image accumulates BOTH heads' input gradients;
// fc_coef [B, c2] = d loss / d z for b200clip_skinny_outer (dW, db).
extern "C" int b200clip_bce_heads_fwd_bwd(const float* image_features, long long ldx, const float* text_features, int c1,
                                          const float* fc_weight, const float* fc_bias, int c2, const float* labels,
                                          int label_cols, long long ld_labels, long long B, int D, float temperature,
                                          const float* label_sum, double total_elems_text, double total_elems_fc,
                                          const float* grad_scale, float* d_image, int d_image_accumulate, float* fc_coef,
                                          double* sums, float* loss_text, float* loss_fc, int* status, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  int rc = check_sc("bce_heads", image_features, ldx, B, c1 + c2, D);
  if (rc) return rc;
  B200_REQUIRE(c1 > 0 && c2 > 0 && temperature > 0.f && labels && label_sum && sums && text_features && fc_weight, "bce_heads: missing arguments");
  if (workspace_bytes < b200clip_smallc_workspace_bytes(B, c1 + c2, D)) return fail(B200_ERR_WORKSPACE, "bce_heads: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ScParams p{};
  p.x = image_features; p.ldx = ldx; p.cls = text_features; p.cls2 = fc_weight; p.c1 = c1; p.bias = fc_bias; p.labels = labels;
  p.label_cols = std::min(label_cols, std::min(c1, c2)); p.ld_labels = ld_labels; p.B = (int)B; p.C = c1 + c2; p.D = D;
  p.inv_tau = 1.0f / temperature; p.normalize_x = 1; p.normalize_cls = 1; p.label_sum = label_sum; p.total_elems = total_elems_text;
  p.total_elems2 = total_elems_fc; p.grad_scale = grad_scale; p.dx = d_image; p.dx_accumulate = d_image_accumulate; p.coef = fc_coef;
  p.partial = static_cast<double*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(sc_grid(B)) * 3 * sizeof(double));
  p.sums = sums; p.loss = loss_text; p.loss2 = loss_fc; p.status = status; p.threshold = 0.5f;
  B200_CHECK_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s));
  return launch_smallc<SC_HEAD2>(p, s);
}
