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

// Register blocking: a warp owns SC_R = 4 rows at a time.  Every 128-bit read of a class vector from shared memory is
// used for 4 rows (16 FMAs per LDS.128 instead of 4): the one-row-per-warp version was shared-memory-bandwidth bound
// (234 us at B=32768, C=32; ncu profiles/r1_launches_summary_v4.txt).  Scores are reduced 32 values at a time
// (SC_R rows x 32/SC_R classes) with the halving exchange, so after chunk k lane L holds (row L / JC, class JC*k + L % JC);
// SC_R = 4 for D <= 512, 2 for D <= 1024 (register budget).
template <int MODE, int MAX_V, int SC_R>
__global__ void __launch_bounds__(SC_THREADS, 2) smallc_kernel(const ScParams p) {
  constexpr int JC = 32 / SC_R;                         // classes per reduction chunk
  constexpr int NCH = SC_MAXC / JC;                     // chunks
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
  const int my_r = lane / JC, my_j = lane % JC;
  const float kt = (MODE == SC_MLBCE || MODE == SC_HEAD2) ? p.inv_tau : 1.0f;

  const long long row_stride = static_cast<long long>(gridDim.x) * (SC_THREADS / 32) * SC_R;
  for (long long row0 = (static_cast<long long>(blockIdx.x) * (SC_THREADS / 32) + warp) * SC_R; row0 < p.B; row0 += row_stride) {
    float4 xv[SC_R][MAX_V];
    float inv[SC_R];
#pragma unroll
    for (int r = 0; r < SC_R; ++r) {
      const bool ok = row0 + r < p.B;
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < MAX_V; ++i) {
        xv[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nv && ok) xv[r][i] = *reinterpret_cast<const float4*>(p.x + (row0 + r) * p.ldx + i * 128 + lane * 4);
        ss += xv[r][i].x * xv[r][i].x + xv[r][i].y * xv[r][i].y + xv[r][i].z * xv[r][i].z + xv[r][i].w * xv[r][i].w;
      }
      inv[r] = 1.0f;
      if (p.normalize_x) {
        ss = warp_sum(ss);
        inv[r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        if (lane == 0 && p.xinv && ok) p.xinv[row0 + r] = inv[r];
      }
    }
    float my_inv = inv[0];
#pragma unroll
    for (int r = 1; r < SC_R; ++r) my_inv = (my_r == r) ? inv[r] : my_inv;
    const long long row = row0 + my_r;                    // the row this lane post-processes
    const bool row_ok = row < p.B;

    float dotk[NCH], coefk[NCH];                            // chunk k: (row my_r, class JC*k + my_j)
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      dotk[k] = 0.f;
      coefk[k] = 0.f;
      if (JC * k >= C) continue;                            // uniform
      float part[32];
#pragma unroll
      for (int j = 0; j < JC; ++j) {
        const int c = JC * k + j;
#pragma unroll
        for (int r = 0; r < SC_R; ++r) part[r * JC + j] = 0.f;
        if (c < C) {
#pragma unroll
          for (int i = 0; i < MAX_V; ++i)
            if (i < nv) {
              const float4 t = *reinterpret_cast<const float4*>(&s_cls[c * D + i * 128 + lane * 4]);
#pragma unroll
              for (int r = 0; r < SC_R; ++r)
                part[r * JC + j] += xv[r][i].x * t.x + xv[r][i].y * t.y + xv[r][i].z * t.z + xv[r][i].w * t.w;
            }
        }
      }
      const float dot_raw = warp_colsum32_sc(part, lane);
      const int cls_idx = JC * k + my_j;
      const float dot = (MODE == SC_HEAD2 && cls_idx >= c1) ? dot_raw : dot_raw * my_inv;   // <x_hat, cls_c> (FC rows: raw x)
      dotk[k] = dot;
      const bool active = cls_idx < C && row_ok;
      float y = 0.f;
      if (MODE == SC_HEAD2) {
        const int lc = cls_idx < c1 ? cls_idx : cls_idx - c1;
        if (active && lc < p.label_cols) y = p.labels[row * p.ld_labels + lc];
      } else if (active && p.labels && cls_idx < p.label_cols) y = p.labels[row * p.ld_labels + cls_idx];
      float coef = 0.f;                                            // d loss / d score_c (score = logit fed to sigmoid)
      if (MODE == SC_MLBCE || (MODE == SC_HEAD2 && cls_idx < c1)) {
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
        const int fc_idx = (MODE == SC_HEAD2) ? cls_idx - c1 : cls_idx;
        const float z = dot + ((cls_idx < C) && p.bias ? p.bias[fc_idx] : 0.f);
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
          p.pred[row * C + cls_idx] = pr > p.threshold ? 1.f : 0.f;   // :885
          if (p.labels) {                                          // :441-447 accuracy counters
            acc0 += ((pr > p.threshold ? 1.f : 0.f) == y) ? 1.0 : 0.0;
          }
        }
      }
      coefk[k] = coef;
      if (MODE != SC_PREDICT && p.coef && active) {
        if (MODE == SC_HEAD2) {
          if (cls_idx >= c1) p.coef[row * (C - c1) + (cls_idx - c1)] = coef;    // FC rows only (feeds dW, db)
        } else p.coef[row * C + cls_idx] = coef;
      }
    }

    if (MODE != SC_PREDICT && p.dx) {
      // d x_hat = sum_c coef_c * cls_c * (1/tau), then through the normalisation;  FC rows (HEAD2) act on x itself.
      // <x_hat, d x_hat> per row: the JC lanes of one group hold the row's classes
      float sd = 0.f;
#pragma unroll
      for (int k = 0; k < NCH; ++k)
        if (JC * k + my_j < c1) sd += coefk[k] * kt * dotk[k];
#pragma unroll
      for (int o = 1; o < JC; o <<= 1) sd += __shfl_xor_sync(0xffffffffu, sd, o);
      float sdot[SC_R];
#pragma unroll
      for (int r = 0; r < SC_R; ++r) sdot[r] = __shfl_sync(0xffffffffu, sd, r * JC);
#pragma unroll
      for (int i0 = 0; i0 < MAX_V; i0 += 2) {               // two float4 column groups at a time (register budget)
        if (i0 >= nv) continue;
        float4 g[SC_R][2];
#pragma unroll
        for (int r = 0; r < SC_R; ++r) g[r][0] = g[r][1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {              // pass 0: classes [0,c1) (through the normalisation); 1: [c1,C)
          if (pass == 1 && MODE != SC_HEAD2) continue;
#pragma unroll
          for (int k = 0; k < NCH; ++k) {
#pragma unroll
            for (int j = 0; j < JC; ++j) {
              const int c = JC * k + j;
              const bool take = pass == 0 ? (c < c1) : (c >= c1 && c < C);   // uniform
              if (!take) continue;
              float cc[SC_R];
#pragma unroll
              for (int r = 0; r < SC_R; ++r) cc[r] = __shfl_sync(0xffffffffu, coefk[k], r * JC + j) * (pass == 0 ? kt : 1.0f);
#pragma unroll
              for (int ii = 0; ii < 2; ++ii) {
                if (i0 + ii >= nv) continue;
                const float4 t = *reinterpret_cast<const float4*>(&s_cls[c * D + (i0 + ii) * 128 + lane * 4]);
#pragma unroll
                for (int r = 0; r < SC_R; ++r) {
                  g[r][ii].x += cc[r] * t.x; g[r][ii].y += cc[r] * t.y; g[r][ii].z += cc[r] * t.z; g[r][ii].w += cc[r] * t.w;
                }
              }
            }
          }
          if (pass == 0 && p.normalize_x) {
#pragma unroll
            for (int r = 0; r < SC_R; ++r)
#pragma unroll
              for (int ii = 0; ii < 2; ++ii) {
                if (i0 + ii >= MAX_V) continue;
                const float4 xr = xv[r][i0 + ii];
                const float a = inv[r], b = inv[r] * sdot[r];
                g[r][ii].x = a * (g[r][ii].x - xr.x * b); g[r][ii].y = a * (g[r][ii].y - xr.y * b);
                g[r][ii].z = a * (g[r][ii].z - xr.z * b); g[r][ii].w = a * (g[r][ii].w - xr.w * b);
              }
          }
        }
#pragma unroll
        for (int r = 0; r < SC_R; ++r) {
          if (row0 + r >= p.B) continue;
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            if (i0 + ii >= nv) continue;
            float4 o = g[r][ii];
            float* d = p.dx + (row0 + r) * D + (i0 + ii) * 128 + lane * 4;
            if (p.dx_accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(d);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(d) = o;
          }
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
// Two-stage and deterministic: each block owns a slab of rows (coefficients staged in shared memory, padded to 32
// classes so one broadcast LDS.128 serves 4 classes), each thread owns 2 columns d, d + 256; then reduce_partials2.
constexpr int SO_ROWS = 256;                            // rows per block
__global__ void __launch_bounds__(256) skinny_outer_partial_kernel(const float* __restrict__ coef, int C,
                                                                   const float* __restrict__ x, long long ldx,
                                                                   const float* __restrict__ row_scale, int rows, int D,
                                                                   float* __restrict__ partial /*[grid][C*D + C]*/) {
  __shared__ __align__(16) float s_coef[SO_ROWS][SC_MAXC];
  const int r0 = blockIdx.x * SO_ROWS;
  const int nr = min(rows, r0 + SO_ROWS) - r0;
  for (int i = threadIdx.x; i < SO_ROWS * SC_MAXC; i += blockDim.x) {
    const int r = i / SC_MAXC, c = i % SC_MAXC;
    s_coef[r][c] = (r < nr && c < C) ? coef[static_cast<long long>(r0 + r) * C + c] * (row_scale ? row_scale[r0 + r] : 1.0f) : 0.f;
  }
  __syncthreads();
  float* out = partial + static_cast<long long>(blockIdx.x) * (static_cast<long long>(C) * D + C);
  const int C4 = (C + 3) >> 2;
  for (int d = threadIdx.x; d < D; d += 2 * blockDim.x) {
    const int d2 = d + blockDim.x;
    const bool has2 = d2 < D;
    float acc0[SC_MAXC], acc1[SC_MAXC];
#pragma unroll
    for (int c = 0; c < SC_MAXC; ++c) acc0[c] = acc1[c] = 0.f;
#pragma unroll 2
    for (int r = 0; r < nr; ++r) {
      const float* xr = x + static_cast<long long>(r0 + r) * ldx;
      const float x0 = xr[d], x1 = has2 ? xr[d2] : 0.f;
#pragma unroll
      for (int c4 = 0; c4 < SC_MAXC / 4; ++c4) {
        if (c4 < C4) {
          const float4 k = *reinterpret_cast<const float4*>(&s_coef[r][c4 * 4]);
          acc0[c4 * 4] += k.x * x0; acc0[c4 * 4 + 1] += k.y * x0; acc0[c4 * 4 + 2] += k.z * x0; acc0[c4 * 4 + 3] += k.w * x0;
          acc1[c4 * 4] += k.x * x1; acc1[c4 * 4 + 1] += k.y * x1; acc1[c4 * 4 + 2] += k.z * x1; acc1[c4 * 4 + 3] += k.w * x1;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < SC_MAXC; ++c)
      if (c < C) {
        out[static_cast<long long>(c) * D + d] = acc0[c];
        if (has2) out[static_cast<long long>(c) * D + d2] = acc1[c];
      }
  }
  if (threadIdx.x < C) {                                // bias grads use the UNSCALED coefficients
    float a = 0.f;
    for (int r = 0; r < nr; ++r) a += coef[static_cast<long long>(r0 + r) * C + threadIdx.x];
    out[static_cast<long long>(C) * D + threadIdx.x] = a;
  }
}

__global__ void __launch_bounds__(256) reduce_partials2_kernel(const float* __restrict__ partial, long long part_stride,
                                                               int nparts, float* __restrict__ out, int n, int accumulate,
                                                               const float* __restrict__ scale) {
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
    if (scale) v *= *scale;
    out[col] = accumulate ? out[col] + v : v;
  }
}

// sum of a float vector into one float (label counts).  One thread-block CLUSTER of 8 CTAs (a single CTA pulls ~80 GB/s:
// 25 us for the 2 MB label matrix at B = 32768); every CTA reduces a contiguous slice, CTA 0 adds the 8 partials in rank
// order through distributed shared memory: deterministic, no workspace.
constexpr int SUM_CLUSTER = 8;
__global__ void __cluster_dims__(SUM_CLUSTER, 1, 1) __launch_bounds__(1024) sum_f32_kernel(const float* __restrict__ a, long long n,
                                                                                        float* __restrict__ out) {
  __shared__ double red[32];
  __shared__ double cta_sum;
  const uint32_t rank = cluster_ctarank();
  // slice boundaries in units of 4 floats so that the vector path stays aligned
  const long long n4_all = ((reinterpret_cast<uintptr_t>(a) & 15u) == 0) ? n / 4 : 0;
  const long long lo4 = n4_all * rank / SUM_CLUSTER, hi4 = n4_all * (rank + 1) / SUM_CLUSTER;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  double acc = 0.0;
  long long i = lo4 + threadIdx.x;
  for (; i + 3 * 1024 < hi4; i += 4 * 1024) {
    const float4 v0 = a4[i], v1 = a4[i + 1024], v2 = a4[i + 2048], v3 = a4[i + 3072];
    acc += static_cast<double>((v0.x + v0.y) + (v0.z + v0.w)) + static_cast<double>((v1.x + v1.y) + (v1.z + v1.w)) +
           static_cast<double>((v2.x + v2.y) + (v2.z + v2.w)) + static_cast<double>((v3.x + v3.y) + (v3.z + v3.w));
  }
  for (; i < hi4; i += 1024) {
    const float4 v = a4[i];
    acc += static_cast<double>((v.x + v.y) + (v.z + v.w));
  }
  if (rank == SUM_CLUSTER - 1)                               // scalar tail (and everything, if the base is unaligned)
    for (long long j = n4_all * 4 + threadIdx.x; j < n; j += 1024) acc += static_cast<double>(a[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 32; ++w) s += red[w];
    cta_sum = s;
  }
  cluster_sync_all();                                        // release/acquire at cluster scope: every cta_sum is visible
  if (rank == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (uint32_t r = 0; r < SUM_CLUSTER; ++r) {
      const uint32_t addr = mapa_u32(smem_u32(&cta_sum), r);
      double v;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr));
      s += v;
    }
    *out = static_cast<float>(s);
  }
  cluster_sync_all();                                        // keep every CTA's shared memory alive until CTA 0 has read it
}

// loss of the fused head from the six (all-reduced) numerators: sums6 = {sum log r, sum log c, sum S_ii, text BCE pos
// numerator, text BCE neg numerator, FC BCE sum}.  One thread; replaces a dozen scalar tensor ops per step.
__global__ void head_loss_finalize_kernel(const double* __restrict__ sums6, const float* __restrict__ label_sum,
                                          double inv_tau_nce, double b_glob, double total_text, double total_fc,
                                          float* __restrict__ loss, float* __restrict__ parts, int* __restrict__ status) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double l_nce = inv_tau_nce + (sums6[0] + sums6[1]) / (2.0 * b_glob) - sums6[2] / b_glob;   // fixed shift m (host.cuh nce_shift; 1/tau for tau >= 0.036)
  const float Psum = *label_sum;
  const float Nsum = static_cast<float>(total_text - static_cast<double>(Psum));
  const float pos = static_cast<float>(-sums6[3]) / (Psum + 1e-8f);                 // 0426/train.py:218
  const float neg = static_cast<float>(-sums6[4]) / (Nsum + 1e-8f);                 // :219
  const float l_text = (pos + neg) * 0.5f;                                          // :221
  const float l_fc = static_cast<float>(sums6[5] / total_fc);
  if (status) *status = (isnan(l_text) || isinf(l_text) || l_text > 1000.f) ? 1 : 0;   // :224
  parts[0] = static_cast<float>(l_nce); parts[1] = l_text; parts[2] = l_fc;
  *loss = static_cast<float>(l_nce) + l_text + l_fc;
}

// multilabel_asymmetric_loss (ASL), multimodal_attention/train.py:233-268 -- elementwise on the [B, C] logits, fused with its
// derivative.  Line references are to that file.  torch.clamp passes the gradient where the input is inside the closed range.
struct AslParams {
  const float* logits; const float* targets; long long n;
  float gamma_pos, gamma_neg, clip, eps;
  float inv_count;                    // 1/n for 'mean', 1 otherwise (scales the gradient)
  const float* grad_scale;            // upstream scalar gradient (mean / sum) or null
  const float* grad_elem;             // upstream elementwise gradient (reduction 'none') or null
  float* loss_elem;                   // [n] or null
  float* d_logits;                    // [n] or null
  double* partial; unsigned int* counter; double* sum; float* loss;
};

__global__ void __launch_bounds__(256) asl_kernel(const AslParams p) {
  __shared__ double red[8];
  __shared__ bool is_last;
  const float gs = p.grad_scale ? *p.grad_scale : 1.0f;
  double acc = 0.0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < p.n; i += static_cast<long long>(gridDim.x) * 256) {
    const float z = p.logits[i], t = p.targets[i];
    const float pr = 1.0f / (1.0f + expf(-z));                        // :247
    const float dpr = pr * (1.0f - pr);
    float qc = 1.0f - pr;                                             // :249
    float dqc = -dpr;
    if (p.clip > 0.f) {                                               // :251-252
      qc += p.clip;
      if (qc > 1.0f) { qc = 1.0f; dqc = 0.f; }
    }
    const float a = fmaxf(pr, p.eps), b = fmaxf(qc, p.eps);           // :254-255 clamp(min=eps)
    const float da = pr >= p.eps ? dpr : 0.f, db = qc >= p.eps ? dqc : 0.f;
    float pos = t * logf(a), dpos = t * da / a;
    float neg = (1.0f - t) * logf(b), dneg = (1.0f - t) * db / b;
    if (p.gamma_pos > 0.f) {                                          // :257-258
      const float w = powf(1.0f - pr, p.gamma_pos), dw = -p.gamma_pos * powf(1.0f - pr, p.gamma_pos - 1.0f) * dpr;
      dpos = dpos * w + pos * dw;
      pos *= w;
    }
    if (p.gamma_neg > 0.f) {                                          // :259-260 (focusing term uses the UNCLIPPED probability)
      const float w = powf(pr, p.gamma_neg), dw = p.gamma_neg * powf(pr, p.gamma_neg - 1.0f) * dpr;
      dneg = dneg * w + neg * dw;
      neg *= w;
    }
    const float l = -(pos + neg);                                     // :262
    acc += static_cast<double>(l);
    if (p.loss_elem) p.loss_elem[i] = l;
    if (p.d_logits) p.d_logits[i] = -(dpos + dneg) * p.inv_count * gs * (p.grad_elem ? p.grad_elem[i] : 1.0f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    p.partial[blockIdx.x] = s;
    __threadfence();
    is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += p.partial[b];      // fixed order: deterministic
    *p.sum = s;
    if (p.loss) *p.loss = static_cast<float>(s * static_cast<double>(p.inv_count));
    *p.counter = 0;
  }
}

static int sc_grid(long long rows) {
  const long long per_block = (SC_THREADS / 32) * 4;
  const long long want = (rows + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 2;
  return static_cast<int>(std::max<long long>(1, std::min(want, cap)));
}

template <int MODE>
static int launch_smallc(const ScParams& p, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(p.C) * p.D * sizeof(float);
  const int grid = sc_grid(p.B);
  if (p.D <= 512) {
    auto k = smallc_kernel<MODE, 4, 4>;
    B200_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, SC_THREADS, smem, s>>>(p);
  } else {
    auto k = smallc_kernel<MODE, 8, 2>;
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
  const int rpb = SO_ROWS;
  const size_t outer = static_cast<size_t>((rows + rpb - 1) / rpb) * (static_cast<size_t>(C) * D + C) * sizeof(float);
  return loss_part + outer + 256;
}

constexpr int ASL_MAX_GRID = 256;
extern "C" size_t b200clip_asl_workspace_bytes(void) { return ASL_MAX_GRID * sizeof(double) + 256; }

// reduction: 0 'none' (loss_elem required), 1 'mean', 2 'sum'.  loss[0] = mean or sum; d_logits (optional) = d loss / d logits
// times the upstream gradient (grad_scale scalar for mean/sum, grad_elem [n] for 'none').
extern "C" int b200clip_asl_fwd_bwd(const float* logits, const float* targets, long long n, float gamma_pos, float gamma_neg,
                                    float clip, float eps, int reduction, const float* grad_scale, const float* grad_elem,
                                    float* loss_elem, float* d_logits, double* sum, float* loss, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  B200_REQUIRE(n > 0 && logits && targets && sum, "asl: missing arguments");
  B200_REQUIRE(reduction >= 0 && reduction <= 2 && (reduction != 0 || loss_elem || d_logits), "asl: bad reduction mode");
  B200_REQUIRE(gamma_pos >= 0.f && gamma_neg >= 0.f && eps > 0.f, "asl: gamma_pos, gamma_neg >= 0 and eps > 0 expected");
  if (workspace_bytes < b200clip_asl_workspace_bytes()) return fail(B200_ERR_WORKSPACE, "asl: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AslParams p{};
  p.logits = logits; p.targets = targets; p.n = n; p.gamma_pos = gamma_pos; p.gamma_neg = gamma_neg; p.clip = clip; p.eps = eps;
  p.inv_count = reduction == 1 ? 1.0f / static_cast<float>(n) : 1.0f;
  p.grad_scale = grad_scale; p.grad_elem = grad_elem; p.loss_elem = loss_elem; p.d_logits = d_logits;
  p.partial = static_cast<double*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + ASL_MAX_GRID * sizeof(double));
  p.sum = sum; p.loss = loss;
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, ASL_MAX_GRID));
  B200_CHECK_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s));
  asl_kernel<<<grid, 256, 0, s>>>(p);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_head_loss_finalize(const double* sums6, const float* label_sum, float temperature_nce, double b_glob,
                                           double total_elems_text, double total_elems_fc, float* loss, float* parts3,
                                           int* status, void* stream) {
  B200_REQUIRE(sums6 && label_sum && loss && parts3 && temperature_nce > 0.f && b_glob > 0, "head_loss_finalize: bad arguments");
  head_loss_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(sums6, label_sum, nce_shift(temperature_nce),
                                                                             b_glob, total_elems_text, total_elems_fc, loss, parts3, status);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_sum_f32(const float* a, long long n, float* out, void* stream) {
  B200_REQUIRE(n >= 0 && out != nullptr, "sum_f32: bad arguments");
  sum_f32_kernel<<<SUM_CLUSTER, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a, n, out);
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
                                     long long rows, int D, float* out_w, float* out_b, int accumulate,
                                     const float* out_scale, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(rows > 0 && C > 0 && C <= SC_MAXC && D > 0, "skinny_outer: bad shape");
  if (workspace_bytes < b200clip_smallc_workspace_bytes(rows, C, D)) return fail(B200_ERR_WORKSPACE, "skinny_outer: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int rpb = SO_ROWS;
  const int nblk = static_cast<int>((rows + rpb - 1) / rpb);
  float* partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(sc_grid(rows)) * 3 * sizeof(double) + 256);
  const long long stride = static_cast<long long>(C) * D + C;
  skinny_outer_partial_kernel<<<nblk, 256, 0, s>>>(coef, C, x, ldx, row_scale, (int)rows, D, partial);
  B200_LAUNCH_CHECK();
  const int nw = C * D;
  reduce_partials2_kernel<<<(nw + 31) / 32, 256, 0, s>>>(partial, stride, nblk, out_w, nw, accumulate, out_scale);
  B200_LAUNCH_CHECK();
  if (out_b) {
    reduce_partials2_kernel<<<1, 256, 0, s>>>(partial + nw, stride, nblk, out_b, C, accumulate, out_scale);
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
