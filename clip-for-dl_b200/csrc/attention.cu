// b200clip: MultiModalAttention (multimodal_attention/train.py:1069-1110) -- additive attention of every
// image over the C <= 16 class texts, without materialising the [B, C, D] tensor the reference expands:
//   ip = x Wi^T + bi  [B, D]      tp = t Wt^T + bt  [C, D]
//   s_bc = sum_d tanh(ip_bd + tp_cd) wa_d + ba        w = softmax_c(s)        e = ip + w tp        out = e Wo^T + bo
// The three [B, D] x [D, D] products run on the tcgen05 GEMM (gemm.cuh); the attention core is one row kernel per pass
// (D/128 warps per image, each owning 128 columns with its slice of tp and wa in registers, MUFU tanh, scores reduced so that
// lane c owns class c).  Backward recomputes tanh and accumulates the cross-row gradients (d tp [C, D], d wa [D], column sum
// of d ip) in register accumulators (each lane owns its columns: no atomics), block partials, deterministic final reduction.
// The text side (C rows) is tiny and stays in fp32 CUDA-core kernels.
// Bytes per image: forward 2 KB (ip f32) in, 1 KB (e bf16) + 64 B (w) out; backward 2 KB + 2 KB (ip, d e) in, 1 KB (d ip) out.
#include <algorithm>

#include "gemm.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {
int gemm_bf16(const void* a, const void* b, int a_mn, int b_mn, int M, int N, int K, long long lda, long long ldb,
              int epi, float alpha, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
              const void* resid, long long ld_res, const float* aux, long long ld_aux, int split_k, cudaStream_t stream,
              float drop_p = 0.f, unsigned int drop_seed = 0u, int aux_is_bf16 = 0, const unsigned int* drop_seed_dev = nullptr,
              float* splitk_ws = nullptr, size_t splitk_ws_bytes = 0);

constexpr int AT_MAXC = 16;
constexpr int AT_MAXD = 512;

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// sum over the 32 lanes of v[i] for every i; lane L returns the sum for index L (31 shuffles, halving exchange)
__device__ __forceinline__ float colsum32_at(float (&v)[32], int lane) {
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

// tp[c][o] = sum_j t[c][j] Wt[o][j] + bt[o]      (C x D outputs, fp32)
__global__ void __launch_bounds__(256) attn_text_proj_kernel(const float* __restrict__ t, const float* __restrict__ wt,
                                                             const float* __restrict__ bt, int C, int D, float* __restrict__ tp) {
  const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= C * D) return;
  const int c = warp / D, o = warp - c * D;
  float acc = 0.f;
  for (int j = lane; j < D; j += 32) acc += t[c * D + j] * wt[static_cast<long long>(o) * D + j];
  acc = warp_sum(acc);
  if (lane == 0) tp[c * D + o] = acc + bt[o];
}

struct AttnParams {
  const float* ip;       // [B, D]
  const float* tp;       // [C, D]
  const float* wa;       // [D]
  const float* ba;       // [1]
  int B, C, D;
  float* w;              // [B, C] attention weights
  __nv_bfloat16* e;      // fwd out: [B, D] ip + w tp (operand of the output GEMM)
  // backward
  const float* de;       // [B, D] gradient w.r.t. e
  const float* dw_up;    // [B, C] upstream gradient of the returned attention weights, or null
  __nv_bfloat16* dip;    // [B, D] gradient w.r.t. ip (operand of the dWi / dx GEMMs)
  float* partial;        // [grid][(C + 2) * D]: d tp rows, d wa, column sum of d ip
};

// Row kernels, round 2 layout: one CTA of D/128 warps walks PAIRS of rows; warp q owns columns [128q, 128q + 128) (lane l:
// the float4 at 128q + 4l), so tp (C x 4 values per lane), wa and -- in the backward pass -- the d tp / d wa / column-sum
// accumulators all live in REGISTERS: no shared-memory traffic in the inner loops (round 1 kept tp in shared memory and the
// accumulators in per-warp shared arrays: 6 LDS/STS.128 per (class, float4), 484 us at B = 32768; the MUFU floor is 58 us).
// Both kernels are ISSUE-bound (ncu: 69-76 % issue-active, profiles/r2_ncu_attention.txt), and a third of the instructions
// were per-row bookkeeping: the 32-wide column-sum exchange carried 16 classes + 16 zeros and the softmax used 5-level warp
// reductions.  Two rows per iteration fill the exchange (lanes 0-15: row A's classes, lanes 16-31: row B's) and run both
// softmaxes in the two half-warps.  The only cross-warp step is the per-pair score (forward) / d w (backward) reduction: 32
// floats per warp through shared memory and one __syncthreads per pair (double-buffered by parity).  The next pair's
// operands are prefetched before the current pair's arithmetic.
constexpr int AT_MAXW = AT_MAXD / 128;                   // warps per CTA at D = 512

__device__ __forceinline__ void at_load_tp(const AttnParams& p, int col, float4 (&tp)[AT_MAXC], float4& wa) {
#pragma unroll
  for (int c = 0; c < AT_MAXC; ++c) tp[c] = c < p.C ? ldf4(p.tp + c * p.D + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  wa = ldf4(p.wa + col);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float half_sum(float v) {      // sum over the 16 lanes of this half-warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// every warp returns, in lane l, the sum over the CTA's warps of `mine` (fixed order: deterministic)
__device__ __forceinline__ float at_cross_warp(float mine, float (*s_x)[AT_MAXW][32], int buf, int warp, int lane, int nwarps) {
  s_x[buf][warp][lane] = mine;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < nwarps; ++w) tot += s_x[buf][w][lane];
  return tot;
}

__global__ void __launch_bounds__(128, 4) attn_fwd_kernel(const AttnParams p) {
  __shared__ float s_x[2][AT_MAXW][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = p.C, D = p.D, col = warp * 128 + lane * 4;
  float4 tp[AT_MAXC], wa;
  at_load_tp(p, col, tp, wa);
  const float ba = *p.ba;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long B = p.B;
  long long ra = 2ll * blockIdx.x;
  float4 xa = ra < B ? ldf4(p.ip + ra * D + col) : z4, xb = ra + 1 < B ? ldf4(p.ip + (ra + 1) * D + col) : z4;
  for (int it = 0; ra < B; ra += 2ll * gridDim.x, ++it) {
    const long long na = ra + 2ll * gridDim.x;
    const float4 xan = na < B ? ldf4(p.ip + na * D + col) : z4, xbn = na + 1 < B ? ldf4(p.ip + (na + 1) * D + col) : z4;
    float part[32];
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) {
      part[c] = 0.f; part[16 + c] = 0.f;
      if (c < C) {                                                             // :1101
        part[c] = tanh_fast(xa.x + tp[c].x) * wa.x + tanh_fast(xa.y + tp[c].y) * wa.y + tanh_fast(xa.z + tp[c].z) * wa.z +
                  tanh_fast(xa.w + tp[c].w) * wa.w;
        part[16 + c] = tanh_fast(xb.x + tp[c].x) * wa.x + tanh_fast(xb.y + tp[c].y) * wa.y + tanh_fast(xb.z + tp[c].z) * wa.z +
                       tanh_fast(xb.w + tp[c].w) * wa.w;
      }
    }
    // lane l: class (l & 15) of row ra + (l >> 4)
    const float score = at_cross_warp(colsum32_at(part, lane), s_x, it & 1, warp, lane, nwarps) + ba;
    const bool act = (lane & 15) < C;
    const float m = half_max(act ? score : -INFINITY);
    const float ex = act ? expf(score - m) : 0.f;
    const float wgt = ex / half_sum(ex);                                      // :1102 softmax over the classes
    const long long my_row = ra + (lane >> 4);
    if (act && warp == 0 && my_row < B) p.w[my_row * C + (lane & 15)] = wgt;
    float4 atta = z4, attb = z4;
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) {
      const float wca = __shfl_sync(0xffffffffu, wgt, c), wcb = __shfl_sync(0xffffffffu, wgt, 16 + c);   // 0 for c >= C
      atta.x += wca * tp[c].x; atta.y += wca * tp[c].y; atta.z += wca * tp[c].z; atta.w += wca * tp[c].w;   // :1105
      attb.x += wcb * tp[c].x; attb.y += wcb * tp[c].y; attb.z += wcb * tp[c].z; attb.w += wcb * tp[c].w;
    }
    *reinterpret_cast<uint2*>(p.e + ra * D + col) =                           // :1108 image_proj + attended_features
        make_uint2(pack_bf16x2(xa.x + atta.x, xa.y + atta.y), pack_bf16x2(xa.z + atta.z, xa.w + atta.w));
    if (ra + 1 < B)
      *reinterpret_cast<uint2*>(p.e + (ra + 1) * D + col) =
          make_uint2(pack_bf16x2(xb.x + attb.x, xb.y + attb.y), pack_bf16x2(xb.z + attb.z, xb.w + attb.w));
    xa = xan; xb = xbn;
  }
}

// one (class, row) term of the backward pass: du = kc wa (1 - tanh^2), accumulated into d ip, d tp_c and d wa
__device__ __forceinline__ void at_bwd_term(const float4& x, const float4& g, const float4& t, const float4& wa, float kc, float wc,
                                            float4& dip, float4& acc_tp, float4& acc_wa) {
  const float4 th = make_float4(tanh_fast(x.x + t.x), tanh_fast(x.y + t.y), tanh_fast(x.z + t.z), tanh_fast(x.w + t.w));
  const float4 du = make_float4(kc * wa.x * (1.f - th.x * th.x), kc * wa.y * (1.f - th.y * th.y),
                                kc * wa.z * (1.f - th.z * th.z), kc * wa.w * (1.f - th.w * th.w));
  dip.x += du.x; dip.y += du.y; dip.z += du.z; dip.w += du.w;
  acc_tp.x += du.x + wc * g.x; acc_tp.y += du.y + wc * g.y; acc_tp.z += du.z + wc * g.z; acc_tp.w += du.w + wc * g.w;   // d tp_c
  acc_wa.x += kc * th.x; acc_wa.y += kc * th.y; acc_wa.z += kc * th.z; acc_wa.w += kc * th.w;                           // d wa
}

__global__ void __launch_bounds__(128, 2) attn_bwd_kernel(const AttnParams p) {
  __shared__ float s_x[2][AT_MAXW][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int C = p.C, D = p.D, col = warp * 128 + lane * 4;
  float4 tp[AT_MAXC], wa;
  at_load_tp(p, col, tp, wa);
  float4 acc_tp[AT_MAXC];
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < AT_MAXC; ++c) acc_tp[c] = z4;
  float4 acc_wa = z4, acc_cs = z4;
  const long long B = p.B;
  const bool act = (lane & 15) < C;
  auto ldw = [&](const float* src, long long ra) {          // lane l: element (l & 15) of row ra + (l >> 4), 0 outside
    const long long r = ra + (lane >> 4);
    return (src != nullptr && act && r < B) ? src[r * C + (lane & 15)] : 0.f;
  };
  long long ra = 2ll * blockIdx.x;
  float4 xa = ra < B ? ldf4(p.ip + ra * D + col) : z4, xb = ra + 1 < B ? ldf4(p.ip + (ra + 1) * D + col) : z4;
  float4 ga = ra < B ? ldf4(p.de + ra * D + col) : z4, gb = ra + 1 < B ? ldf4(p.de + (ra + 1) * D + col) : z4;
  float wgt = ldw(p.w, ra), up = ldw(p.dw_up, ra);
  for (int it = 0; ra < B; ra += 2ll * gridDim.x, ++it) {
    const long long na = ra + 2ll * gridDim.x;
    const float4 xan = na < B ? ldf4(p.ip + na * D + col) : z4, xbn = na + 1 < B ? ldf4(p.ip + (na + 1) * D + col) : z4;
    const float4 gan = na < B ? ldf4(p.de + na * D + col) : z4, gbn = na + 1 < B ? ldf4(p.de + (na + 1) * D + col) : z4;
    const float wn = ldw(p.w, na), un = ldw(p.dw_up, na);
    // d w_c = d e . tp_c  (+ upstream gradient of the returned weights)
    float part[32];
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) {
      part[c] = 0.f; part[16 + c] = 0.f;
      if (c < C) { part[c] = dot4(ga, tp[c]); part[16 + c] = dot4(gb, tp[c]); }
    }
    const float dwc = at_cross_warp(colsum32_at(part, lane), s_x, it & 1, warp, lane, nwarps) + up;
    const float sw = half_sum(wgt * dwc);                      // wgt = 0 on inactive lanes / missing rows
    const float ds = wgt * (dwc - sw);                         // softmax backward: lane l holds d s of (row l >> 4, class l & 15)
    float4 dipa = ga, dipb = gb;                               // e = ip + attended: direct path
#pragma unroll
    for (int c = 0; c < AT_MAXC; ++c) {
      if (c < C) {
        at_bwd_term(xa, ga, tp[c], wa, __shfl_sync(0xffffffffu, ds, c), __shfl_sync(0xffffffffu, wgt, c), dipa, acc_tp[c], acc_wa);
        at_bwd_term(xb, gb, tp[c], wa, __shfl_sync(0xffffffffu, ds, 16 + c), __shfl_sync(0xffffffffu, wgt, 16 + c), dipb, acc_tp[c], acc_wa);
      }
    }
    // column sum of d ip = d image_proj.bias (a missing row B has g = 0, wgt = 0: it contributes 0)
    acc_cs.x += dipa.x + dipb.x; acc_cs.y += dipa.y + dipb.y; acc_cs.z += dipa.z + dipb.z; acc_cs.w += dipa.w + dipb.w;
    *reinterpret_cast<uint2*>(p.dip + ra * D + col) = make_uint2(pack_bf16x2(dipa.x, dipa.y), pack_bf16x2(dipa.z, dipa.w));
    if (ra + 1 < B)
      *reinterpret_cast<uint2*>(p.dip + (ra + 1) * D + col) = make_uint2(pack_bf16x2(dipb.x, dipb.y), pack_bf16x2(dipb.z, dipb.w));
    xa = xan; xb = xbn; ga = gan; gb = gbn; wgt = wn; up = un;
  }
  float* out = p.partial + static_cast<long long>(blockIdx.x) * (C + 2) * D + col;
#pragma unroll
  for (int c = 0; c < AT_MAXC; ++c)
    if (c < C) *reinterpret_cast<float4*>(out + c * D) = acc_tp[c];
  *reinterpret_cast<float4*>(out + C * D) = acc_wa;
  *reinterpret_cast<float4*>(out + (C + 1) * D) = acc_cs;
}

// out[i] = sum_parts partial[part][i]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) attn_reduce_kernel(const float* __restrict__ partial, int nparts, int n, float* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
#pragma unroll 8
  for (int q = 0; q < nparts; ++q) a += partial[static_cast<long long>(q) * n + i];
  out[i] = a;
}

// text side (C rows, fp32): dWt[o][j] = sum_c dtp[c][o] t[c][j] ; dbt[o] = sum_c dtp[c][o] ; dt[c][j] = sum_o dtp[c][o] Wt[o][j]
__global__ void __launch_bounds__(256) attn_text_bwd_w_kernel(const float* __restrict__ dtp, const float* __restrict__ t, int C, int D,
                                                              float* __restrict__ dwt, float* __restrict__ dbt) {
  const long long idx = blockIdx.x * 256ll + threadIdx.x;
  if (idx >= static_cast<long long>(D) * D) return;
  const int o = static_cast<int>(idx / D), j = static_cast<int>(idx - static_cast<long long>(o) * D);
  float a = 0.f, b = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = dtp[c * D + o];
    a += d * t[c * D + j];
    b += d;
  }
  dwt[idx] = a;
  if (j == 0) dbt[o] = b;
}
// dt[c][j] = sum_o dtp[c][o] Wt[o][j]: block (j-block of 32, class c); warp s of the 8 sums the o-slice [s D/8, (s+1) D/8)
// (round 1: one thread per output walking all D rows of Wt -- 59 us of dependent loads for 16 x 512 outputs)
__global__ void __launch_bounds__(256) attn_text_bwd_x_kernel(const float* __restrict__ dtp, const float* __restrict__ wt, int C, int D,
                                                              float* __restrict__ dt) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int c = blockIdx.y, j = blockIdx.x * 32 + lane;
  const int o0 = s * (D / 8), o1 = o0 + D / 8;
  float a = 0.f;
#pragma unroll 8
  for (int o = o0; o < o1; ++o) a += dtp[c * D + o] * wt[static_cast<long long>(o) * D + j];
  red[s][lane] = a;
  __syncthreads();
  if (s == 0) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][lane];
    dt[c * D + j] = tot;
  }
}

static int at_split_for(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 255) / 256);
  const int kchunks = (K + 63) / 64;
  int s = num_sms() / tiles;
  if (s > kchunks) s = kchunks;
  return s < 1 ? 1 : s;
}
static int at_bwd_grid(long long B) {                      // 2 CTAs of D/128 warps per SM (register accumulators), row pairs
  return static_cast<int>(std::max<long long>(1, std::min<long long>((B + 1) / 2, 2LL * num_sms())));
}
static int check_attn(const char* who, long long B, int C, int D) {
  B200_REQUIRE(B > 0 && C > 0 && C <= AT_MAXC, "%s: need B > 0 and 0 < C <= %d (got B=%lld C=%d)", who, AT_MAXC, B, C);
  B200_REQUIRE(D > 0 && D % 128 == 0 && D <= AT_MAXD, "%s: D=%d must be a multiple of 128, <= %d", who, D, AT_MAXD);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

// x_bf16 [B, D] image features; t [C, D] class text features (fp32).  Saved for backward (caller tensors): ip, tp, w, e.
extern "C" int b200clip_attention_fwd(const void* x_bf16, const float* t, long long B, int C, int D, const void* wi_bf16,
                                      const float* bi, const float* wt, const float* bt, const float* wa, const float* ba,
                                      const void* wo_bf16, const float* bo, float* ip, float* tp, float* w, void* e_bf16,
                                      float* out, void* stream) {
  int rc = check_attn("attention_fwd", B, C, D);
  if (rc) return rc;
  B200_REQUIRE(x_bf16 && t && wi_bf16 && bi && wt && bt && wa && ba && wo_bf16 && bo && ip && tp && w && e_bf16 && out,
               "attention_fwd: missing arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if ((rc = gemm_bf16(x_bf16, wi_bf16, 0, 0, (int)B, D, D, D, D, EPI_STORE_F32, 1.0f, ip, D, nullptr, 0, bi, nullptr, 0, nullptr, 0, 1, s)))
    return rc;                                                                  // :1092 image_proj
  attn_text_proj_kernel<<<(C * D * 32 + 255) / 256, 256, 0, s>>>(t, wt, bt, C, D, tp);   // :1093 text_proj
  B200_LAUNCH_CHECK();
  AttnParams p{};
  p.ip = ip; p.tp = tp; p.wa = wa; p.ba = ba; p.B = (int)B; p.C = C; p.D = D; p.w = w; p.e = static_cast<__nv_bfloat16*>(e_bf16);
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((B + 1) / 2, 4LL * num_sms())));
  attn_fwd_kernel<<<grid, 32 * (D / 128), 0, s>>>(p);
  B200_LAUNCH_CHECK();
  return gemm_bf16(e_bf16, wo_bf16, 0, 0, (int)B, D, D, D, D, EPI_STORE_F32, 1.0f, out, D, nullptr, 0, bo, nullptr, 0, nullptr, 0, 1, s);   // :1108
}

extern "C" size_t b200clip_attention_bwd_workspace_bytes(long long B, int C, int D) {
  size_t n = 0;
  n += ((static_cast<size_t>(B) * D * 2) + 255) & ~size_t(255);                 // d_out bf16
  n += ((static_cast<size_t>(B) * D * 4) + 255) & ~size_t(255);                 // d_e f32
  n += ((static_cast<size_t>(B) * D * 2) + 255) & ~size_t(255);                 // d_ip bf16
  n += ((static_cast<size_t>(at_bwd_grid(B)) * (C + 2) * D * 4) + 255) & ~size_t(255);   // block partials
  n += ((static_cast<size_t>(C + 2) * D * 4) + 255) & ~size_t(255);             // reduced d tp | d wa | colsum(d ip)
  n += (b200clip_colsum_workspace_bytes(B, D) + 255) & ~size_t(255);
  n += (static_cast<size_t>(num_sms()) * 128 * 256 * sizeof(float) + 255) & ~size_t(255);   // deterministic split-K partial tiles
  return n + 1024;
}

// d_out [B, D] f32, d_w [B, C] f32 or null (gradient of the returned attention weights).  dx / dt optional.
extern "C" int b200clip_attention_bwd(const float* d_out, const float* d_w, const void* x_bf16, const float* t, long long B, int C,
                                      int D, const void* wi_bf16, const float* wt, const float* wa, const float* ba,
                                      const void* wo_bf16, const float* ip, const float* tp, const float* w, const void* e_bf16,
                                      float* dx, float* dt, float* dwi, float* dbi, float* dwt, float* dbt, float* dwa, float* dba,
                                      float* dwo, float* dbo, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_attn("attention_bwd", B, C, D);
  if (rc) return rc;
  B200_REQUIRE(d_out && x_bf16 && t && wi_bf16 && wt && wa && ba && wo_bf16 && ip && tp && w && e_bf16 && dwi && dbi && dwt && dbt &&
               dwa && dba && dwo && dbo, "attention_bwd: missing arguments");
  if (workspace_bytes < b200clip_attention_bwd_workspace_bytes(B, C, D)) return fail(B200_ERR_WORKSPACE, "attention_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto carve = [&](size_t bytes) { uint8_t* q = ws; ws += (bytes + 255) & ~size_t(255); return q; };
  void* dout_bf = carve(static_cast<size_t>(B) * D * 2);
  float* de = reinterpret_cast<float*>(carve(static_cast<size_t>(B) * D * 4));
  void* dip_bf = carve(static_cast<size_t>(B) * D * 2);
  const int grid = at_bwd_grid(B);
  float* partial = reinterpret_cast<float*>(carve(static_cast<size_t>(grid) * (C + 2) * D * 4));
  float* red = reinterpret_cast<float*>(carve(static_cast<size_t>(C + 2) * D * 4));
  const size_t cs_ws = b200clip_colsum_workspace_bytes(B, D);
  void* cs = carve(cs_ws);
  const size_t sk_ws = static_cast<size_t>(num_sms()) * 128 * 256 * sizeof(float);
  float* sk_work = reinterpret_cast<float*>(carve(sk_ws));

  if ((rc = b200clip_cast_f32_bf16(d_out, dout_bf, B * D, stream))) return rc;
  if ((rc = b200clip_colsum(d_out, 0, D, B, D, dbo, 0, cs, cs_ws, stream))) return rc;          // output_proj.bias
  if ((rc = gemm_bf16(dout_bf, e_bf16, 1, 1, D, D, (int)B, D, D, EPI_ATOMIC_F32, 1.0f, dwo, D, nullptr, 0, nullptr, nullptr, 0, nullptr,
                      0, at_split_for(D, D, (int)B), s, 0.f, 0u, 0, nullptr, sk_work, sk_ws)))
    return rc;                                                                  // dWo = d_out^T e
  if ((rc = gemm_bf16(dout_bf, wo_bf16, 0, 1, (int)B, D, D, D, D, EPI_STORE_F32, 1.0f, de, D, nullptr, 0, nullptr, nullptr, 0, nullptr, 0,
                      1, s)))
    return rc;                                                                  // d e = d_out Wo
  AttnParams p{};
  p.ip = ip; p.tp = tp; p.wa = wa; p.ba = ba; p.B = (int)B; p.C = C; p.D = D; p.w = const_cast<float*>(w); p.de = de; p.dw_up = d_w;
  p.dip = static_cast<__nv_bfloat16*>(dip_bf); p.partial = partial;
  attn_bwd_kernel<<<grid, 32 * (D / 128), 0, s>>>(p);
  B200_LAUNCH_CHECK();
  attn_reduce_kernel<<<((C + 2) * D + 255) / 256, 256, 0, s>>>(partial, grid, (C + 2) * D, red);
  B200_LAUNCH_CHECK();
  const float* dtp = red;                                                       // [C, D]
  B200_CHECK_CUDA(cudaMemcpyAsync(dwa, red + static_cast<size_t>(C) * D, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s));
  B200_CHECK_CUDA(cudaMemcpyAsync(dbi, red + static_cast<size_t>(C + 1) * D, static_cast<size_t>(D) * 4, cudaMemcpyDeviceToDevice, s));
  B200_CHECK_CUDA(cudaMemsetAsync(dba, 0, sizeof(float), s));    // a constant added to every score leaves the softmax unchanged
  // image side
  if ((rc = gemm_bf16(dip_bf, x_bf16, 1, 1, D, D, (int)B, D, D, EPI_ATOMIC_F32, 1.0f, dwi, D, nullptr, 0, nullptr, nullptr, 0, nullptr, 0,
                      at_split_for(D, D, (int)B), s, 0.f, 0u, 0, nullptr, sk_work, sk_ws)))
    return rc;
  if (dx) {
    if ((rc = gemm_bf16(dip_bf, wi_bf16, 0, 1, (int)B, D, D, D, D, EPI_STORE_F32, 1.0f, dx, D, nullptr, 0, nullptr, nullptr, 0, nullptr, 0,
                        1, s)))
      return rc;
  }
  // text side
  attn_text_bwd_w_kernel<<<static_cast<int>((static_cast<long long>(D) * D + 255) / 256), 256, 0, s>>>(dtp, t, C, D, dwt, dbt);
  B200_LAUNCH_CHECK();
  if (dt) {
    attn_text_bwd_x_kernel<<<dim3(D / 32, C), 256, 0, s>>>(dtp, wt, C, D, dt);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}
