// b200clip: the two BCE heads of the fused step on warp-level tensor cores.
//
// multilabel_contrastive_loss on the 16 class texts (a-B, 0426/train.py:178-230) and the FC classification adapter with
// BCEWithLogits (a-A, NB02 c28:50-52 / c29:23-25) read the same image features.  The first version (smallc.cu, one
// warp per row on fp32 CUDA cores) was shared-memory-bandwidth bound: 250 us at B = 32768 for 134 MB of HBM traffic
// (20 us at the HBM roofline).  Here both the score product [B,512] x [512,32] and the input-gradient product
// [B,32] x [32,512] run on mma.sync.m16n8k16 (bf16 -> fp32), so the kernel is HBM-bound:
//   * operand = the L2-normalised bf16 features y_hat that LayerNorm already wrote for InfoNCE (33 MB instead of the
//     67 MB fp32 copy), plus 1/||y|| per row:  cos = y_hat . c_hat,  FC logit z = (y_hat . W) * ||y|| + b;
//   * no shared-memory staging of the features: a dot product is invariant under a permutation of K, so each lane's
//     16 contiguous bytes of a row are a valid A fragment for two k16 steps once the class fragments use the same
//     permutation (same trick as zeroshot.cu);
//   * the score accumulators ARE the A fragment of the second product (coefficients of 16 classes = one k16 step);
//   * output  d_y = (1/||y||) (G_m - y_hat (y_hat . G_m)) + G_f,  G_m = coef_m C_hat / tau,  G_f = coef_f W, for upstream
//     gradient 1 (backward scales it), and the FC coefficients * ||y|| in bf16 for the weight gradient below.
// skinny_outer_mma: dW_fc[16,512] = sum_rows coefn[row,:]^T y_hat[row,:] with the rows as the MMA K dimension
// (A = y_hat^T fragments via ldmatrix.trans), per-CTA partials, deterministic final reduction.
// Algorithmic bytes per row: 1024 (y_hat) + 64 (labels) + 2048 (d_y) + 32 (coefn)  ->  104 MB at B = 32768.
#include <algorithm>

#include "common.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

constexpr int HM_MAXD = 768;                             // widths: 512 (0426/config.py:30) and 768 (BASELINE.json configs[4])
constexpr int HM_C = 16;                                 // classes per head (class texts | FC rows)
constexpr int HM_THREADS = 256;
constexpr int HM_WARPS = HM_THREADS / 32;
// shared memory: class fragments for the score product (32 / 48 KB) + transposed class matrices for the gradient product
// (32 / 48 KB): two CTAs per SM at either width
template <int DD> struct HmCfg {
  static_assert(DD == 512 || DD == 768, "heads_mma widths");
  static constexpr int KS = DD / 32;                     // 32-element steps per row (two k16 MMAs each)
  static constexpr int FRAG_BYTES = KS * 4 * 32 * 16;
  static constexpr int CT_BYTES = 2 * DD * HM_C * 2;
  static constexpr int SMEM = FRAG_BYTES + CT_BYTES;
};

__device__ __forceinline__ void hmma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                           uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct HeadsParams {
  const __nv_bfloat16* yhat;          // [B, 512] L2-normalised features
  const float* inv_norm;              // [B] 1 / ||y||
  const float* cls;                   // [16, 512] class texts (normalised here, F.normalize eps 1e-12)
  const float* w;                     // [16, 512] FC adapter weight
  const float* bias;                  // [16] or null
  const float* labels; int label_cols; long long ld_labels;
  int B;
  float inv_tau;
  const float* label_sum;             // device scalar: sum(labels) over the GLOBAL batch
  double total_text, total_fc;        // B_glob * 16 each (mean divisors)
  float* dy;                          // [B, 512] f32 or null
  __nv_bfloat16* coefn;               // [B, 16] bf16: d loss / d z * ||y||   (or null)
  double* partial;                    // [grid][3 + 16]
  unsigned int* counter;
  double* sums;                       // [3] text pos numerator, text neg numerator, FC BCE sum
  float* db;                          // [16] sum_rows d loss / d z (unscaled) or null
};

template <int HM_D>
__global__ void __launch_bounds__(HM_THREADS, 2) bce_heads_mma_kernel(const HeadsParams p) {
  constexpr int HM_FRAG_BYTES = HmCfg<HM_D>::FRAG_BYTES, KS = HmCfg<HM_D>::KS;
  extern __shared__ __align__(128) uint8_t hm_smem[];
  // frag[(s*4 + t)*32 + lane] = V[8t + lane/4][32s + 8(lane%4) .. +7]   (V = 16 normalised class texts, then 16 FC rows)
  uint4* s_frag = reinterpret_cast<uint4*>(hm_smem);
  // ct[h][col][class] bf16, h = 0: c_hat / tau, h = 1: W          (B operand of the gradient product, class-contiguous)
  __nv_bfloat16* s_ct = reinterpret_cast<__nv_bfloat16*>(hm_smem + HM_FRAG_BYTES);
  __shared__ float s_cinv[HM_C];
  __shared__ double s_red[HM_WARPS][3 + HM_C];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & 3, r = lane >> 2;

  for (int c = warp; c < HM_C; c += HM_WARPS) {           // 1 / ||class text||
    float ss = 0.f;
    for (int d = lane; d < HM_D; d += 32) {
      const float v = p.cls[c * HM_D + d];
      ss += v * v;
    }
    ss = warp_sum(ss);
    if (lane == 0) s_cinv[c] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KS * 4 * 32; i += HM_THREADS) {
    const int l = i & 31, t = (i >> 5) & 3, s = i >> 7;
    const int n = 8 * t + (l >> 2), k0 = 32 * s + 8 * (l & 3);
    const float* src = (n < HM_C) ? p.cls + n * HM_D + k0 : p.w + (n - HM_C) * HM_D + k0;
    const float sc = (n < HM_C) ? s_cinv[n] : 1.0f;
    const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
    s_frag[i] = make_uint4(pack_bf16x2(v0.x * sc, v0.y * sc), pack_bf16x2(v0.z * sc, v0.w * sc),
                           pack_bf16x2(v1.x * sc, v1.y * sc), pack_bf16x2(v1.z * sc, v1.w * sc));
  }
  // pairs of classes; consecutive threads take consecutive columns of one class pair so that the global reads coalesce (with
  // the class pair fastest every load was a separate 32-byte sector: ~20 us of this kernel's 35 us at B = 4096)
  for (int i = threadIdx.x; i < 2 * HM_D * HM_C / 2; i += HM_THREADS) {
    const int col = i % HM_D, cp = (i / HM_D) & 7, h = i / (8 * HM_D);
    const int c0 = 2 * cp;
    float a, b;
    if (h == 0) {
      a = p.cls[c0 * HM_D + col] * s_cinv[c0] * p.inv_tau;
      b = p.cls[(c0 + 1) * HM_D + col] * s_cinv[c0 + 1] * p.inv_tau;
    } else {
      a = p.w[c0 * HM_D + col];
      b = p.w[(c0 + 1) * HM_D + col];
    }
    reinterpret_cast<uint32_t*>(s_ct)[(h * HM_D + col) * (HM_C / 2) + cp] = pack_bf16x2(a, b);
  }
  __syncthreads();

  const float Psum = *p.label_sum;
  const float Nsum = static_cast<float>(p.total_text - static_cast<double>(Psum));
  const float inv_total_fc = static_cast<float>(1.0 / p.total_fc);
  float bias_v[2][2];                                       // FC bias of this lane's classes: tile tt -> 8tt + 2q, +1
#pragma unroll
  for (int tt = 0; tt < 2; ++tt)
#pragma unroll
    for (int e = 0; e < 2; ++e) bias_v[tt][e] = p.bias ? p.bias[8 * tt + 2 * q + e] : 0.f;
  double acc_pos = 0.0, acc_neg = 0.0, acc_fc = 0.0;
  float db_acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};

  const int nblk = (p.B + 15) >> 4;
  for (int blk = blockIdx.x * HM_WARPS + warp; blk < nblk; blk += gridDim.x * HM_WARPS) {
    const int row_a = blk * 16 + r, row_b = row_a + 8;
    const bool ok_a = row_a < p.B, ok_b = row_b < p.B;
    const uint4* pa = reinterpret_cast<const uint4*>(p.yhat + static_cast<long long>(ok_a ? row_a : 0) * HM_D) + q;
    const uint4* pb = reinterpret_cast<const uint4*>(p.yhat + static_cast<long long>(ok_b ? row_b : 0) * HM_D) + q;
    float acc[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
#pragma unroll 8
    for (int s = 0; s < KS; ++s) {
      const uint4 xa = __ldg(pa + s * 4), xb = __ldg(pb + s * 4);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint4 f = s_frag[(s * 4 + t) * 32 + lane];
        hmma_16816(acc[t], xa.x, xb.x, xa.y, xb.y, f.x, f.y);
        hmma_16816(acc[t], xa.z, xb.z, xa.w, xb.w, f.z, f.w);
      }
    }
    // acc[t] = {(row_a, 8t+2q), (row_a, 8t+2q+1), (row_b, 8t+2q), (row_b, 8t+2q+1)}; t < 2: cos with class texts, t >= 2: y_hat . W
    const float inv_a = ok_a ? p.inv_norm[row_a] : 1.f, inv_b = ok_b ? p.inv_norm[row_b] : 1.f;
    const float nrm_a = 1.0f / inv_a, nrm_b = 1.0f / inv_b;
    float cm[2][4], cf[2][4];                               // coefficients d loss / d score in the accumulator layout
    float sd_a = 0.f, sd_b = 0.f;                           // y_hat . G_m per row (partial over this lane's classes)
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool is_b = e >= 2;
        const int row = is_b ? row_b : row_a;
        const bool ok = is_b ? ok_b : ok_a;
        const int c = 8 * tt + 2 * q + (e & 1);
        const float y = (ok && c < p.label_cols) ? p.labels[static_cast<long long>(row) * p.ld_labels + c] : 0.f;
        // ---- multi-label BCE on sigmoid(cos / tau): 0426/train.py:195-221 ----
        const float dot = acc[tt][e];
        const float s = dot * p.inv_tau;                           // :195
        const float sc = fminf(fmaxf(s, -50.f), 50.f);             // :213
        const float pp = 1.0f / (1.0f + expf(-sc));              // :214
        const float qq = 1.0f - pp;                                // :215
        float coef = 0.f;
        if (ok) {
          acc_pos += static_cast<double>(logf(pp + 1e-8f) * y);       // :218 numerator
          acc_neg += static_cast<double>(logf(qq + 1e-8f) * (1.0f - y));   // :219 numerator
          const float inside = (fabsf(s) <= 50.f) ? 1.f : 0.f;
          const float dpos = -y * pp * qq / ((pp + 1e-8f) * (Psum + 1e-8f));
          const float dneg = (1.0f - y) * pp * qq / ((qq + 1e-8f) * (Nsum + 1e-8f));
          coef = 0.5f * (dpos + dneg) * inside;
        }
        cm[tt][e] = coef;
        if (is_b) sd_b += coef * s; else sd_a += coef * s;         // coef * (1/tau) * cos
        // ---- FC adapter + BCEWithLogits ----
        const float z = acc[2 + tt][e] * (is_b ? nrm_b : nrm_a) + bias_v[tt][e & 1];
        float cfv = 0.f;
        if (ok) {
          acc_fc += static_cast<double>(fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z))));
          cfv = (1.0f / (1.0f + expf(-z)) - y) * inv_total_fc;
          db_acc[tt][e & 1] += cfv;
        }
        cf[tt][e] = cfv;
      }
    }
    sd_a += __shfl_xor_sync(0xffffffffu, sd_a, 1); sd_a += __shfl_xor_sync(0xffffffffu, sd_a, 2);
    sd_b += __shfl_xor_sync(0xffffffffu, sd_b, 1); sd_b += __shfl_xor_sync(0xffffffffu, sd_b, 2);
    if (p.coefn) {                                           // [row][16] bf16, classes 2q,2q+1 (tile 0) and 8+2q,8+2q+1 (tile 1)
      uint32_t* ca = reinterpret_cast<uint32_t*>(p.coefn + static_cast<long long>(row_a) * HM_C);
      uint32_t* cb = reinterpret_cast<uint32_t*>(p.coefn + static_cast<long long>(row_b) * HM_C);
      if (ok_a) { ca[q] = pack_bf16x2(cf[0][0] * nrm_a, cf[0][1] * nrm_a); ca[4 + q] = pack_bf16x2(cf[1][0] * nrm_a, cf[1][1] * nrm_a); }
      if (ok_b) { cb[q] = pack_bf16x2(cf[0][2] * nrm_b, cf[0][3] * nrm_b); cb[4 + q] = pack_bf16x2(cf[1][2] * nrm_b, cf[1][3] * nrm_b); }
    }
    if (p.dy) {
      // A fragments of the gradient product: rows x 16 classes (k = class): a0 (row_a, k 2q..), a1 (row_b, k 2q..),
      // a2 (row_a, k 8+2q..), a3 (row_b, k 8+2q..)
      const uint32_t am0 = pack_bf16x2(cm[0][0], cm[0][1]), am1 = pack_bf16x2(cm[0][2], cm[0][3]);
      const uint32_t am2 = pack_bf16x2(cm[1][0], cm[1][1]), am3 = pack_bf16x2(cm[1][2], cm[1][3]);
      const uint32_t af0 = pack_bf16x2(cf[0][0], cf[0][1]), af1 = pack_bf16x2(cf[0][2], cf[0][3]);
      const uint32_t af2 = pack_bf16x2(cf[1][0], cf[1][1]), af3 = pack_bf16x2(cf[1][2], cf[1][3]);
      const uint32_t* ctm = reinterpret_cast<const uint32_t*>(s_ct);
      const uint32_t* ctf = ctm + HM_D * (HM_C / 2);
      const uint32_t* ya = reinterpret_cast<const uint32_t*>(p.yhat + static_cast<long long>(ok_a ? row_a : 0) * HM_D) + q;
      const uint32_t* yb = reinterpret_cast<const uint32_t*>(p.yhat + static_cast<long long>(ok_b ? row_b : 0) * HM_D) + q;
      float* oa = p.dy + static_cast<long long>(row_a) * HM_D + 2 * q;
      float* ob = p.dy + static_cast<long long>(row_b) * HM_D + 2 * q;
      const float ka = inv_a * sd_a, kb = inv_b * sd_b;
#pragma unroll 4
      for (int j = 0; j < HM_D / 8; ++j) {
        const int col = 8 * j + r;                          // B fragment: (k = 2q, 2q+1 ; n = r) and (k = 8+2q.. ; n = r)
        float gm[4] = {0.f, 0.f, 0.f, 0.f}, gf[4] = {0.f, 0.f, 0.f, 0.f};
        hmma_16816(gm, am0, am1, am2, am3, ctm[col * 8 + q], ctm[col * 8 + 4 + q]);
        hmma_16816(gf, af0, af1, af2, af3, ctf[col * 8 + q], ctf[col * 8 + 4 + q]);
        // gm/gf = {(row_a, 8j+2q), (row_a, 8j+2q+1), (row_b, ..), (row_b, ..)}
        const uint32_t wa = __ldg(ya + j * 4), wb = __ldg(yb + j * 4);
        if (ok_a)
          *reinterpret_cast<float2*>(oa + j * 8) =
              make_float2(inv_a * gm[0] - bf16_lo(wa) * ka + gf[0], inv_a * gm[1] - bf16_hi(wa) * ka + gf[1]);
        if (ok_b)
          *reinterpret_cast<float2*>(ob + j * 8) =
              make_float2(inv_b * gm[2] - bf16_lo(wb) * kb + gf[2], inv_b * gm[3] - bf16_hi(wb) * kb + gf[3]);
      }
    }
  }

  // deterministic two-level reduction: loss numerators and the FC bias gradient
  double v3[3] = {acc_pos, acc_neg, acc_fc};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v3[k] += __shfl_xor_sync(0xffffffffu, v3[k], o);
    if (lane == 0) s_red[warp][k] = v3[k];
  }
#pragma unroll
  for (int tt = 0; tt < 2; ++tt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = db_acc[tt][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (r == 0) s_red[warp][3 + 8 * tt + 2 * q + e] = static_cast<double>(v);
    }
  __syncthreads();
  if (threadIdx.x < 3 + HM_C) {
    double t = 0.0;
    for (int w = 0; w < HM_WARPS; ++w) t += s_red[w][threadIdx.x];
    p.partial[blockIdx.x * (3 + HM_C) + threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    // fold with 8 threads per value (blocks j, j+8, ... each), then add the 8 in fixed order: deterministic
    __threadfence();
    __shared__ double fold[3 + HM_C][8];
    if (threadIdx.x < (3 + HM_C) * 8) {
      const int k = threadIdx.x >> 3, j = threadIdx.x & 7;
      double t = 0.0;
      for (unsigned b = j; b < gridDim.x; b += 8) t += p.partial[b * (3 + HM_C) + k];
      fold[k][j] = t;
    }
    __syncthreads();
    if (threadIdx.x < 3 + HM_C) {
      double t = 0.0;
      for (int j = 0; j < 8; ++j) t += fold[threadIdx.x][j];
      if (threadIdx.x < 3) p.sums[threadIdx.x] = t;
      else if (p.db) p.db[threadIdx.x - 3] = static_cast<float>(t);
    }
    if (threadIdx.x == 0) *p.counter = 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dW[16, 512] partials: out[c][d] = sum_rows coefn[row][c] * yhat[row][d].   One CTA per slab of rows, 8 warps x 64
// columns; per 16-row step a warp issues 4 ldmatrix.x4.trans (A = y_hat^T: m = column, k = row) and 8 MMAs.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SO2_ROWS = 32;                             // slab granularity (rows per CTA are a multiple of it)

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// 8 warps x HM_D/8 columns (4 or 6 m-tiles of 16); rows staged per iteration: 32 at D = 512, 16 at D = 768 (static smem < 48 KB)
template <int HM_D>
__global__ void __launch_bounds__(HM_THREADS, 2) skinny_outer_mma_kernel(const __nv_bfloat16* __restrict__ coefn,
                                                                         const __nv_bfloat16* __restrict__ yhat, int rows,
                                                                         int rows_per_cta, float* __restrict__ partial /*[grid][16*D]*/) {
  constexpr int STG = HM_D == 512 ? 32 : 16;             // rows staged per iteration
  constexpr int MT = HM_D / 128;                         // m-tiles (16 columns) per warp
  constexpr int SO2_PITCH = HM_D * 2 + 16;               // padded row pitch (bytes): ldmatrix rows 16 B apart in bank space
  __shared__ __align__(128) uint8_t s_y[STG * SO2_PITCH];               // [rows][D bf16 + pad]
  __shared__ __align__(16) __nv_bfloat16 s_c[HM_C][STG + 8];            // transposed coefficients [class][row]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & 3, r = lane >> 2;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  float acc[MT][2][4];                                    // [m-tile (16 columns)][n-tile (8 classes)][4]
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
  const uint32_t sy = smem_u32(s_y);
  for (int base = r0; base < r1; base += STG) {
    __syncthreads();
    // stage y_hat rows (zero beyond the slab) and the transposed coefficients
    for (int i = threadIdx.x; i < STG * (HM_D / 8); i += HM_THREADS) {
      const int rr = i / (HM_D / 8), ch = i % (HM_D / 8);
      const int row = base + rr;
      const uint4 v = row < r1 ? __ldg(reinterpret_cast<const uint4*>(yhat + static_cast<long long>(row) * HM_D) + ch) : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(s_y + rr * SO2_PITCH + ch * 16) = v;
    }
    for (int i = threadIdx.x; i < STG * HM_C; i += HM_THREADS) {
      const int rr = i >> 4, c = i & 15;   // HM_C == 16
      const int row = base + rr;
      s_c[c][rr] = row < r1 ? coefn[static_cast<long long>(row) * HM_C + c] : __float2bfloat16(0.f);
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < STG / 16; ++ks) {
      // B fragments: (k = row 16ks + 2q, +1 ; n = class r) and rows +8 ; second n-tile: class 8 + r
      uint32_t bfr[2][2];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        bfr[n][0] = *reinterpret_cast<const uint32_t*>(&s_c[8 * n + r][16 * ks + 2 * q]);
        bfr[n][1] = *reinterpret_cast<const uint32_t*>(&s_c[8 * n + r][16 * ks + 8 + 2 * q]);
      }
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        // A = y_hat^T tile: m = column (warp*16MT + 16m ..+16), k = row (16ks ..+16).  Stored [row][col]: four 8x8 blocks
        // (rows 0-7 | cols 0-7), (rows 0-7 | cols 8-15), (rows 8-15 | cols 0-7), (rows 8-15 | cols 8-15), transposed on load
        // -> a0 = (m 0-7, k 0-7), a1 = (m 8-15, k 0-7), a2 = (m 0-7, k 8-15), a3 = (m 8-15, k 8-15)
        const int col0 = warp * (16 * MT) + 16 * m;
        const int lrow = 16 * ks + (lane & 7) + ((lane >> 4) << 3);       // lanes 0-15: rows 0-7 ; 16-31: rows 8-15
        const int lcol = col0 + (((lane >> 3) & 1) << 3);                 // lanes 8-15, 24-31: cols +8
        uint32_t a[4];
        ldmatrix_x4_trans(a, sy + lrow * SO2_PITCH + lcol * 2);
#pragma unroll
        for (int n = 0; n < 2; ++n) hmma_16816(acc[m][n], a[0], a[1], a[2], a[3], bfr[n][0], bfr[n][1]);
      }
    }
  }
  // acc[m][n] = {(col 16m + r, class 8n + 2q), (col .., class +1), (col 16m + r + 8, class ..), (.., +1)}
  float* out = partial + static_cast<long long>(blockIdx.x) * (HM_C * HM_D);
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const int col = warp * (16 * MT) + 16 * m + r, c = 8 * n + 2 * q;
      out[c * HM_D + col] = acc[m][n][0];
      out[(c + 1) * HM_D + col] = acc[m][n][1];
      out[c * HM_D + col + 8] = acc[m][n][2];
      out[(c + 1) * HM_D + col + 8] = acc[m][n][3];
    }
}

// out[i] = scale * sum_parts partial[part][i]   (deterministic)
__global__ void __launch_bounds__(256) so2_reduce_kernel(const float* __restrict__ partial, int nparts, int n, float* __restrict__ out,
                                                         const float* __restrict__ scale, const float* __restrict__ db_raw,
                                                         float* __restrict__ db_out, int ndb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float sc = scale ? *scale : 1.0f;
  if (i < n) {
    float acc = 0.f;
    for (int part = 0; part < nparts; ++part) acc += partial[static_cast<long long>(part) * n + i];
    out[i] = acc * sc;
  }
  if (db_out && blockIdx.x == 0 && threadIdx.x < ndb) db_out[threadIdx.x] = db_raw[threadIdx.x] * sc;
}

static int hm_grid(long long rows) {
  const long long nblk = (rows + 15) / 16;
  return static_cast<int>(std::max<long long>(1, std::min<long long>((nblk + HM_WARPS - 1) / HM_WARPS, 2LL * num_sms())));
}
static int so2_rows_per_cta(long long rows) {
  // enough CTAs to fill the machine twice over, slabs a multiple of the staging depth
  const long long want = 2LL * num_sms();
  long long per = (rows + want - 1) / want;
  per = ((per + SO2_ROWS - 1) / SO2_ROWS) * SO2_ROWS;
  return static_cast<int>(std::max<long long>(per, SO2_ROWS));
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200clip_bce_heads_mma_workspace_bytes(long long rows) {
  const size_t a = static_cast<size_t>(hm_grid(rows)) * (3 + HM_C) * sizeof(double) + 256;
  const int per = so2_rows_per_cta(rows);
  const size_t b = static_cast<size_t>((rows + per - 1) / per) * HM_C * HM_MAXD * sizeof(float);
  return a + b + 256;
}

extern "C" int b200clip_bce_heads_mma_fwd(const void* yhat_bf16, const float* inv_norm, long long B, int D,
                                          const float* class_text, int c1, const float* fc_weight, const float* fc_bias, int c2,
                                          const float* labels, int label_cols, long long ld_labels, float temperature,
                                          const float* label_sum, double total_elems_text, double total_elems_fc, float* d_y,
                                          void* coefn_bf16, float* db_fc, double* sums, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  if ((D != 512 && D != 768) || c1 != HM_C || c2 != HM_C)
    return fail(B200_ERR_UNSUPPORTED, "bce_heads_mma: built for D=512/768 and 16+16 classes (got D=%d, %d+%d); use b200clip_bce_heads_fwd_bwd", D, c1, c2);
  B200_REQUIRE(B > 0 && yhat_bf16 && inv_norm && class_text && fc_weight && labels && label_sum && sums && temperature > 0.f,
               "bce_heads_mma: missing arguments");
  B200_REQUIRE(aligned16(yhat_bf16) && aligned16(class_text) && aligned16(fc_weight) && aligned16(d_y) && aligned16(coefn_bf16),
               "bce_heads_mma: pointers must be 16-byte aligned");
  B200_REQUIRE(label_cols > 0, "bce_heads_mma: label_cols must be positive");
  if (workspace_bytes < b200clip_bce_heads_mma_workspace_bytes(B)) return fail(B200_ERR_WORKSPACE, "bce_heads_mma: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = hm_grid(B);
  HeadsParams p{};
  p.yhat = static_cast<const __nv_bfloat16*>(yhat_bf16); p.inv_norm = inv_norm; p.cls = class_text; p.w = fc_weight; p.bias = fc_bias;
  p.labels = labels; p.label_cols = std::min(label_cols, HM_C); p.ld_labels = ld_labels; p.B = static_cast<int>(B);
  p.inv_tau = 1.0f / temperature; p.label_sum = label_sum; p.total_text = total_elems_text; p.total_fc = total_elems_fc;
  p.dy = d_y; p.coefn = static_cast<__nv_bfloat16*>(coefn_bf16); p.db = db_fc; p.sums = sums;
  p.partial = static_cast<double*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(grid) * (3 + HM_C) * sizeof(double));
  static SmemAttrOnce attr512, attr768;
  B200_CHECK_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s));
  if (D == 512) {
    B200_CHECK_CUDA(attr512.ensure(bce_heads_mma_kernel<512>, HmCfg<512>::SMEM));
    bce_heads_mma_kernel<512><<<grid, HM_THREADS, HmCfg<512>::SMEM, s>>>(p);
  } else {
    B200_CHECK_CUDA(attr768.ensure(bce_heads_mma_kernel<768>, HmCfg<768>::SMEM));
    bce_heads_mma_kernel<768><<<grid, HM_THREADS, HmCfg<768>::SMEM, s>>>(p);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// dW_fc[16, D] = *out_scale * coefn^T y_hat ;  db_out[16] = *out_scale * db_raw      (out_scale: optional device scalar)
extern "C" int b200clip_skinny_outer_mma(const void* coefn_bf16, const void* yhat_bf16, long long rows, int D, int C,
                                         const float* out_scale, float* out_w, const float* db_raw, float* db_out,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  if ((D != 512 && D != 768) || C != HM_C) return fail(B200_ERR_UNSUPPORTED, "skinny_outer_mma: built for D=512/768, C=16 (got D=%d C=%d)", D, C);
  B200_REQUIRE(rows > 0 && coefn_bf16 && yhat_bf16 && out_w, "skinny_outer_mma: missing arguments");
  B200_REQUIRE(aligned16(yhat_bf16) && aligned16(coefn_bf16), "skinny_outer_mma: pointers must be 16-byte aligned");
  if (workspace_bytes < b200clip_bce_heads_mma_workspace_bytes(rows)) return fail(B200_ERR_WORKSPACE, "skinny_outer_mma: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int per = so2_rows_per_cta(rows);
  const int grid = static_cast<int>((rows + per - 1) / per);
  float* partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(hm_grid(rows)) * (3 + HM_C) * sizeof(double) + 256);
  if (D == 512)
    skinny_outer_mma_kernel<512><<<grid, HM_THREADS, 0, s>>>(static_cast<const __nv_bfloat16*>(coefn_bf16),
                                                             static_cast<const __nv_bfloat16*>(yhat_bf16), static_cast<int>(rows), per, partial);
  else
    skinny_outer_mma_kernel<768><<<grid, HM_THREADS, 0, s>>>(static_cast<const __nv_bfloat16*>(coefn_bf16),
                                                             static_cast<const __nv_bfloat16*>(yhat_bf16), static_cast<int>(rows), per, partial);
  B200_LAUNCH_CHECK();
  so2_reduce_kernel<<<(HM_C * D + 255) / 256, 256, 0, s>>>(partial, grid, HM_C * D, out_w, out_scale, db_raw,
                                                           (db_raw && db_out) ? db_out : nullptr, HM_C);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
