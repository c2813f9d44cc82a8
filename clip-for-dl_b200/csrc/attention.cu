// b200clip: MultiModalAttention (multimodal_attention/train.py:1069-1110) -- additive attention of every
// image over the C <= 16 class texts, without materialising the [B, C, D] tensor the reference expands:
//   ip = x Wi^T + bi  [B, D]      tp = t Wt^T + bt  [C, D]
//   s_bc = sum_d tanh(ip_bd + tp_cd) wa_d + ba        w = softmax_c(s)        e = ip + w tp        out = e Wo^T + bo
// The three [B, D] x [D, D] products run on the tcgen05 GEMM (gemm.cuh); the attention core is one row kernel per pass
// (one warp per image, tp and wa resident in shared memory, MUFU tanh, scores reduced so that lane c owns class c).
// Backward recomputes tanh and accumulates the cross-row gradients (d tp [C, D], d wa [D], column sum of d ip) in per-warp
// shared-memory accumulators (each lane owns its columns: no atomics), block partials, deterministic final reduction.
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
constexpr int AT_V = AT_MAXD / 128;                      // float4 groups per lane
constexpr int AT_FWD_THREADS = 256;
constexpr int AT_BWD_THREADS = 128;                      // 4 warps: (C + 2) x D fp32 accumulators each in shared memory

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

__global__ void __launch_bounds__(AT_FWD_THREADS) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float at_smem[];          // tp [C][D] | wa [D]
  float* s_tp = at_smem;
  float* s_wa = at_smem + p.C * p.D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C, D = p.D, nv = D >> 7;
  for (int i = threadIdx.x; i < C * D; i += AT_FWD_THREADS) s_tp[i] = p.tp[i];
  for (int i = threadIdx.x; i < D; i += AT_FWD_THREADS) s_wa[i] = p.wa[i];
  __syncthreads();
  const float ba = *p.ba;
  for (long long row = blockIdx.x * (AT_FWD_THREADS / 32) + warp; row < p.B; row += static_cast<long long>(gridDim.x) * (AT_FWD_THREADS / 32)) {
    float4 x[AT_V];
#pragma unroll
    for (int i = 0; i < AT_V; ++i)
      if (i < nv) x[i] = ldf4(p.ip + row * D + i * 128 + lane * 4);
    float part[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      part[c] = 0.f;
      if (c < AT_MAXC && c < C) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < AT_V; ++i)
          if (i < nv) {
            const float4 t = ldf4(s_tp + c * D + i * 128 + lane * 4), wv = ldf4(s_wa + i * 128 + lane * 4);
            a += tanh_fast(x[i].x + t.x) * wv.x + tanh_fast(x[i].y + t.y) * wv.y + tanh_fast(x[i].z + t.z) * wv.z +
                 tanh_fast(x[i].w + t.w) * wv.w;                              // :1101
          }
        part[c] = a;
      }
    }
    const float score = colsum32_at(part, lane) + ba;                         // lane c: s_c
    const bool act = lane < C;
    const float m = warp_max(act ? score : -INFINITY);
    const float ex = act ? expf(score - m) : 0.f;
    const float wgt = ex / warp_sum(ex);                                      // :1102 softmax over the classes
    if (act) p.w[row * C + lane] = wgt;
    float4 att[AT_V];
#pragma unroll
    for (int i = 0; i < AT_V; ++i) att[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < C; ++c) {
      const float wc = __shfl_sync(0xffffffffu, wgt, c);
#pragma unroll
      for (int i = 0; i < AT_V; ++i)
        if (i < nv) {
          const float4 t = ldf4(s_tp + c * D + i * 128 + lane * 4);
          att[i].x += wc * t.x; att[i].y += wc * t.y; att[i].z += wc * t.z; att[i].w += wc * t.w;   // :1105
        }
    }
#pragma unroll
    for (int i = 0; i < AT_V; ++i)
      if (i < nv)                                                             // :1108 image_proj + attended_features
        *reinterpret_cast<uint2*>(p.e + row * D + i * 128 + lane * 4) =
            make_uint2(pack_bf16x2(x[i].x + att[i].x, x[i].y + att[i].y), pack_bf16x2(x[i].z + att[i].z, x[i].w + att[i].w));
  }
}

__global__ void __launch_bounds__(AT_BWD_THREADS, 1) attn_bwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float at_smem[];          // tp [C][D] | wa [D] | acc [4 warps][(C + 2)][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C, D = p.D, nv = D >> 7;
  float* s_tp = at_smem;
  float* s_wa = at_smem + C * D;
  float* s_acc = s_wa + D;
  float* my = s_acc + static_cast<size_t>(warp) * (C + 2) * D;
  for (int i = threadIdx.x; i < C * D; i += AT_BWD_THREADS) s_tp[i] = p.tp[i];
  for (int i = threadIdx.x; i < D; i += AT_BWD_THREADS) s_wa[i] = p.wa[i];
  for (int i = threadIdx.x; i < (AT_BWD_THREADS / 32) * (C + 2) * D; i += AT_BWD_THREADS) s_acc[i] = 0.f;
  __syncthreads();
  auto acc_add = [&](int r, int i, float4 v) {                 // lane-private columns: plain read-modify-write
    float4* q = reinterpret_cast<float4*>(my + r * D + i * 128 + lane * 4);
    float4 t = *q;
    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    *q = t;
  };
  for (long long row = blockIdx.x * (AT_BWD_THREADS / 32) + warp; row < p.B; row += static_cast<long long>(gridDim.x) * (AT_BWD_THREADS / 32)) {
    float4 x[AT_V], g[AT_V], dip[AT_V];
#pragma unroll
    for (int i = 0; i < AT_V; ++i)
      if (i < nv) {
        x[i] = ldf4(p.ip + row * D + i * 128 + lane * 4);
        g[i] = ldf4(p.de + row * D + i * 128 + lane * 4);
        dip[i] = g[i];                                         // e = ip + attended: direct path
      }
    const bool act = lane < C;
    const float wgt = act ? p.w[row * C + lane] : 0.f;
    // d w_c = d e . tp_c  (+ upstream gradient of the returned weights)
    float part[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      part[c] = 0.f;
      if (c < AT_MAXC && c < C) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < AT_V; ++i)
          if (i < nv) {
            const float4 t = ldf4(s_tp + c * D + i * 128 + lane * 4);
            a += g[i].x * t.x + g[i].y * t.y + g[i].z * t.z + g[i].w * t.w;
          }
        part[c] = a;
      }
    }
    float dwc = colsum32_at(part, lane);
    if (act && p.dw_up) dwc += p.dw_up[row * C + lane];
    const float sw = warp_sum(act ? wgt * dwc : 0.f);
    const float ds = act ? wgt * (dwc - sw) : 0.f;             // softmax backward: lane c holds d s_c
    for (int c = 0; c < C; ++c) {
      const float kc = __shfl_sync(0xffffffffu, ds, c), wc = __shfl_sync(0xffffffffu, wgt, c);
#pragma unroll
      for (int i = 0; i < AT_V; ++i)
        if (i < nv) {
          const float4 t = ldf4(s_tp + c * D + i * 128 + lane * 4), wv = ldf4(s_wa + i * 128 + lane * 4);
          const float4 th = make_float4(tanh_fast(x[i].x + t.x), tanh_fast(x[i].y + t.y), tanh_fast(x[i].z + t.z), tanh_fast(x[i].w + t.w));
          const float4 du = make_float4(kc * wv.x * (1.f - th.x * th.x), kc * wv.y * (1.f - th.y * th.y),
                                        kc * wv.z * (1.f - th.z * th.z), kc * wv.w * (1.f - th.w * th.w));
          dip[i].x += du.x; dip[i].y += du.y; dip[i].z += du.z; dip[i].w += du.w;
          acc_add(c, i, make_float4(du.x + wc * g[i].x, du.y + wc * g[i].y, du.z + wc * g[i].z, du.w + wc * g[i].w));   // d tp_c
          acc_add(C, i, make_float4(kc * th.x, kc * th.y, kc * th.z, kc * th.w));                                        // d wa
        }
    }
#pragma unroll
    for (int i = 0; i < AT_V; ++i)
      if (i < nv) {
        acc_add(C + 1, i, dip[i]);                             // column sum of d ip = gradient of image_proj.bias
        *reinterpret_cast<uint2*>(p.dip + row * D + i * 128 + lane * 4) =
            make_uint2(pack_bf16x2(dip[i].x, dip[i].y), pack_bf16x2(dip[i].z, dip[i].w));
      }
  }
  __syncthreads();
  float* out = p.partial + static_cast<long long>(blockIdx.x) * (C + 2) * D;
  for (int i = threadIdx.x; i < (C + 2) * D; i += AT_BWD_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < AT_BWD_THREADS / 32; ++w) a += s_acc[static_cast<size_t>(w) * (C + 2) * D + i];
    out[i] = a;
  }
}

// out[i] = sum_parts partial[part][i]   (fixed order: deterministic)
__global__ void __launch_bounds__(256) attn_reduce_kernel(const float* __restrict__ partial, int nparts, int n, float* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
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
__global__ void __launch_bounds__(256) attn_text_bwd_x_kernel(const float* __restrict__ dtp, const float* __restrict__ wt, int C, int D,
                                                              float* __restrict__ dt) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= C * D) return;
  const int c = idx / D, j = idx - c * D;
  float a = 0.f;
  for (int o = 0; o < D; ++o) a += dtp[c * D + o] * wt[static_cast<long long>(o) * D + j];
  dt[idx] = a;
}

static int at_split_for(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 255) / 256);
  const int kchunks = (K + 63) / 64;
  int s = num_sms() / tiles;
  if (s > kchunks) s = kchunks;
  return s < 1 ? 1 : s;
}
static int at_bwd_grid(long long B) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((B + 3) / 4, num_sms())));
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
  const size_t smem = static_cast<size_t>(C + 1) * D * sizeof(float);
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((B + 7) / 8, 4LL * num_sms())));
  attn_fwd_kernel<<<grid, AT_FWD_THREADS, smem, s>>>(p);
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
  const size_t smem = (static_cast<size_t>(C + 1) * D + static_cast<size_t>(AT_BWD_THREADS / 32) * (C + 2) * D) * sizeof(float);
  static SmemAttrOnce attr;
  B200_CHECK_CUDA(attr.ensure(attn_bwd_kernel, static_cast<int>((static_cast<size_t>(AT_MAXC + 1) * AT_MAXD + 4ull * (AT_MAXC + 2) * AT_MAXD) * 4)));
  attn_bwd_kernel<<<grid, AT_BWD_THREADS, smem, s>>>(p);
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
    attn_text_bwd_x_kernel<<<(C * D + 255) / 256, 256, 0, s>>>(dtp, wt, C, D, dt);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}
