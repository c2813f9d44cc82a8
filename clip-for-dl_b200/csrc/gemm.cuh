// b200clip: tcgen05/TMEM GEMM with TMA-fed 128B-swizzled smem tiles and fused epilogues.
//   D[M,N] = A * B  with bf16 operands, fp32 accumulation in TMEM.
// Operand storage (row-major global matrices, bf16):
//   A K-major : a[M][K]   (rows of A are contiguous in K)       A MN-major: a[K][M]  (i.e. A^T stored)
//   B K-major : b[N][K]   ("NT" GEMM, y = x W^T)                B MN-major: b[K][N]  ("NN" GEMM, y = x W)
// A work item is one 128 x BN output tile over a K range (split-K); persistent CTAs walk the item list.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (one elected thread), 2 = TMEM allocator, 4..11 = epilogue
// (warp_idx % 4 selects the TMEM lane quadrant a warp may read; warpgroup = column half).
#pragma once
#include "common.cuh"

namespace b200 {

enum GemmEpilogue : int {
  EPI_STORE_F32 = 0,        // out0 f32 [M,N]           = alpha * acc (+ bias[n])
  EPI_STORE_BF16 = 1,       // out0 bf16 [M,N]          = alpha * acc (+ bias[n])
  EPI_BIAS_GELU = 2,        // out0 bf16 = p = acc + bias ; out1 bf16 = gelu_erf(p)
  EPI_BIAS_RESID_F32 = 3,   // out0 f32  = acc + bias + resid(bf16)
  EPI_ATOMIC_F32 = 4,       // atomicAdd(out0 f32, alpha * acc)       (split-K reduction)
  EPI_GELU_BWD = 5,         // out0 bf16 = acc * gelu'(resid=p bf16) + aux(f32 [M,N])   (dp = dh*gelu'(p) + dz)
  EPI_RELU_BF16 = 6,        // out0 bf16 = dropout(relu(acc + bias))      (drop_p = 0: plain ReLU)
  EPI_RELU_BWD = 7,         // out0 bf16 = alpha * acc where resid(bf16) > 0, else 0      (d relu / d dropout of EPI_RELU_BF16)
};

struct GemmParams {
  int M, N, K;              // logical problem size
  int k_chunks;             // ceil(K / 64)
  int k_chunks_per_split;   // chunks handled by one split
  int splits;               // number of K splits (work items = m_tiles * n_tiles * splits)
  float alpha;
  void* out0;  long long ld0;
  void* out1;  long long ld1;
  const float* bias;        // [N] or nullptr
  const __nv_bfloat16* resid; long long ld_res;
  const float* aux;         long long ld_aux;
  int aux_is_bf16;          // EPI_GELU_BWD: aux points at bf16 data (dz as written for the GEMMs) instead of f32
  float drop_p;             // EPI_BIAS_RESID_F32: dropout on (acc + bias) before the residual add (0 = off)
  unsigned int drop_seed;
  const unsigned int* drop_seed_dev;   // optional device word added to drop_seed (a captured CUDA graph draws a fresh mask per replay)
  int split_rows;           // deterministic split-K: split z stores its partial tile at row offset z * split_rows of out0 (0 = off)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
#ifndef B200CLIP_GEMM_EPI_WARPS
#define B200CLIP_GEMM_EPI_WARPS 8
#endif
constexpr int GEMM_EPI_WARPS = B200CLIP_GEMM_EPI_WARPS;     // 8 or 16: lane quadrant x column half / quarter of the tile
constexpr int GEMM_THREADS = 128 + GEMM_EPI_WARPS * 32;     // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4.. epilogue
static_assert(GEMM_EPI_WARPS == 8 || GEMM_EPI_WARPS == 16, "epilogue warpgroups");

template <int BN>
constexpr int gemm_stage_bytes() { return GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2; }
template <int BN, int STAGES>
constexpr int gemm_smem_bytes() { return STAGES * gemm_stage_bytes<BN>() + 1024 /*align*/ + 256 /*barriers*/; }

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): one rcp + one ex2 + a degree-5 Horner chain instead of erff's
// two polynomial branches.  The GELU epilogues are ALU-bound (128 x 256 erf per tile against 6144 MMA clocks).  The
// backward epilogue shares exp(-x^2/2) between the cdf and the pdf.
__device__ __forceinline__ void phi_cdf_pdf(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x) * 0.70710678118654752f;                 // |x| / sqrt(2)
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  const float e = fast_exp2(-1.4426950408889634f * ax * ax);        // exp(-x^2 / 2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_erfc = 0.5f * poly * t * e;                      // 0.5 * erfc(|x| / sqrt 2)
  cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  pdf = 0.3989422804014327f * e;
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  float cdf, pdf;
  phi_cdf_pdf(x, cdf, pdf);
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad_f(float x) {
  // d/dx [x * Phi(x)] = Phi(x) + x * phi(x)
  float cdf, pdf;
  phi_cdf_pdf(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}

// one 32-column chunk of one output row: v = fp32 accumulator bits of columns [col, col+32)
template <int EPI>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&v)[32], int row, int col) {
  float f[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]) * p.alpha;
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(p.bias + col + i);
      f[i] += b4.x; f[i + 1] += b4.y; f[i + 2] += b4.z; f[i + 3] += b4.w;
    }
  }
  if constexpr (EPI == EPI_STORE_F32) {
    float* o = reinterpret_cast<float*>(p.out0) + static_cast<long long>(row) * p.ld0 + col;
#pragma unroll
    for (int i = 0; i < 32; i += 8) st_global_f32x8(o + i, f + i);
  } else if constexpr (EPI == EPI_ATOMIC_F32) {
    float* o = reinterpret_cast<float*>(p.out0) + static_cast<long long>(row) * p.ld0 + col;
#pragma unroll
    for (int i = 0; i < 32; i += 4)                 // 16-byte vector reductions: 8 instead of 32 L2 atomics per chunk
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(f[i]), "f"(f[i + 1]), "f"(f[i + 2]), "f"(f[i + 3])
                   : "memory");
  } else if constexpr (EPI == EPI_STORE_BF16 || EPI == EPI_RELU_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col;
    if constexpr (EPI == EPI_RELU_BF16) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
      if (p.drop_p > 0.f) {                       // nn.Dropout after the ReLU (MultiViewFusion, 0426/train.py:994)
        const unsigned int eff_seed = p.drop_seed + (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0u);
        const float sc = 1.0f / (1.0f - p.drop_p);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          f[i] = dropout_keep(eff_seed, static_cast<uint32_t>(row), static_cast<uint32_t>(col + i), static_cast<uint32_t>(p.N), p.drop_p) ? f[i] * sc : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 32; i += 16) st_global_bf16x16(o + i, f + i);
  } else if constexpr (EPI == EPI_BIAS_GELU) {
    __nv_bfloat16* o0 = reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col;
    __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(p.out1) + static_cast<long long>(row) * p.ld1 + col;
#pragma unroll
    for (int i = 0; i < 32; i += 16) {
      st_global_bf16x16(o0 + i, f + i);
      float g[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        // GELU is applied to the bf16-rounded p so that backward (which only sees the stored p) is consistent
        const float pr = __bfloat162float(__float2bfloat16_rn(f[i + t]));
        g[t] = gelu_erf_f(pr);
      }
      st_global_bf16x16(o1 + i, g);
    }
  } else if constexpr (EPI == EPI_BIAS_RESID_F32) {
    const __nv_bfloat16* rs = p.resid + static_cast<long long>(row) * p.ld_res + col;
    float* o = reinterpret_cast<float*>(p.out0) + static_cast<long long>(row) * p.ld0 + col;
    if (p.drop_p > 0.f) {                         // nn.Dropout between fc and the residual add (0426/train.py:93)
      const unsigned int eff_seed = p.drop_seed + (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0u);
      const float sc = 1.0f / (1.0f - p.drop_p);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        f[i] = dropout_keep(eff_seed, static_cast<uint32_t>(row), static_cast<uint32_t>(col + i), static_cast<uint32_t>(p.N), p.drop_p) ? f[i] * sc : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 32; i += 16) {
      float r[16];
      ld_global_bf16x16(rs + i, r);
#pragma unroll
      for (int t = 0; t < 16; ++t) f[i + t] += r[t];
    }
#pragma unroll
    for (int i = 0; i < 32; i += 8) st_global_f32x8(o + i, f + i);
  } else if constexpr (EPI == EPI_RELU_BWD) {
    const __nv_bfloat16* rs = p.resid + static_cast<long long>(row) * p.ld_res + col;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col;
#pragma unroll
    for (int i = 0; i < 32; i += 16) {
      float hv[16], g[16];
      ld_global_bf16x16(rs + i, hv);
#pragma unroll
      for (int t = 0; t < 16; ++t) g[t] = hv[t] > 0.f ? f[i + t] : 0.f;
      st_global_bf16x16(o + i, g);
    }
  } else if constexpr (EPI == EPI_GELU_BWD) {
    const __nv_bfloat16* rs = p.resid + static_cast<long long>(row) * p.ld_res + col;
    const float* ax = p.aux + static_cast<long long>(row) * p.ld_aux + col;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out0) + static_cast<long long>(row) * p.ld0 + col;
#pragma unroll
    for (int i = 0; i < 32; i += 16) {
      float pv[16], g[16];
      ld_global_bf16x16(rs + i, pv);
      float av[16];
      if (p.aux_is_bf16) {
        ld_global_bf16x16(reinterpret_cast<const __nv_bfloat16*>(p.aux) + static_cast<long long>(row) * p.ld_aux + col + i, av);
      } else {
        const u32x8 a0 = ld_global_v8(ax + i), a1 = ld_global_v8(ax + i + 8);
#pragma unroll
        for (int t = 0; t < 8; ++t) { av[t] = __uint_as_float(a0.v[t]); av[8 + t] = __uint_as_float(a1.v[t]); }
      }
#pragma unroll
      for (int t = 0; t < 16; ++t) g[t] = f[i + t] * gelu_erf_grad_f(pv[t]) + av[t];
      st_global_bf16x16(o + i, g);
    }
  }
}

// Persistent kernel: gridDim.x CTAs (<= one per SM) walk the list of (split, m-tile, n-tile) work items.  The fp32
// accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns), so the epilogue of item i (8 warps: two warpgroups, each
// owning half of the tile's columns) overlaps the TMA/MMA main loop of item i+1.
template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int EPI_COLS = BN / (GEMM_EPI_WARPS / 4);   // columns per epilogue warpgroup
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;          // 2
  uint64_t* acc_empty = acc_full + 2;               // 2 (one arrival per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int mn_tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * n_tiles;
  const int items = mn_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], GEMM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop; one elected lane issues) =====================
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), smem0 = smem_u32(smem);
    int s = 0;
    uint32_t ph = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int z = it / mn_tiles, mn = it - z * mn_tiles;
      const int m0 = (mn / n_tiles) * GEMM_BM, n0 = (mn % n_tiles) * BN;
      const int kc0 = z * p.k_chunks_per_split;
      const int nk = min(p.k_chunks, kc0 + p.k_chunks_per_split) - kc0;
      int k0 = kc0 * GEMM_BK;
      for (int i = 0; i < nk; ++i, k0 += GEMM_BK) {
        mbar_wait_a(empty0 + 8 * s, ph ^ 1);
        const uint32_t sa = smem0 + s * STAGE_BYTES, sb = sa + A_BYTES, fb = full0 + 8 * s;
        if (elect_one()) {
          mbar_arrive_expect_tx_a(fb, STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d_a(sa, &tmap_a, fb, k0, m0);                               // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int g = 0; g < GEMM_BM / 64; ++g) tma_load_2d_a(sa + g * 8192, &tmap_a, fb, m0 + g * 64, k0);   // {64 m, 64 k}
          }
          if constexpr (!B_MN) {
            tma_load_2d_a(sb, &tmap_b, fb, k0, n0);                               // box {64 k, BN n}
          } else {
#pragma unroll
            for (int g = 0; g < BN / 64; ++g) tma_load_2d_a(sb + g * 8192, &tmap_b, fb, n0 + g * 64, k0);        // {64 n, 64 k}
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: warp-uniform loop (descriptor words live in uniform registers), one
    // elected lane issues; one 32-bit add per operand per MMA =====================
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
    constexpr uint32_t A_STEP = (A_MN ? 2048 : 32) >> 4, B_STEP = (B_MN ? 2048 : 32) >> 4;
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    const uint32_t af0 = smem_u32(acc_full), ae0 = smem_u32(acc_empty);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), A_MN ? 8192 : 16);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem) + A_BYTES, B_MN ? 8192 : 16);
    int s = 0;
    uint32_t ph = 0;
    int local = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++local) {
      const int z = it / mn_tiles;
      const int kc0 = z * p.k_chunks_per_split;
      const int nk = min(p.k_chunks, kc0 + p.k_chunks_per_split) - kc0;
      const int buf = local & 1;
      mbar_wait_a(ae0 + 8 * buf, ((local >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * BN;
      for (int i = 0; i < nk; ++i) {
        mbar_wait_a(full0 + 8 * s, ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (STAGE_BYTES >> 4), b_lo = b_lo0 + s * (STAGE_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < GEMM_BK / 16; ++j)
            mma_ss_lo(d_tmem, a_lo + j * A_STEP, b_lo + j * B_STEP, idesc, (i | j) != 0);
          tc_commit_a(empty0 + 8 * s);                // frees the smem stage once these MMAs have read it
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (elect_one()) tc_commit_a(af0 + 8 * buf);    // accumulator complete
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global (2 warpgroups x half the columns) =====================
    const int q = warp & 3;                           // TMEM lane quadrant of this warp
    const int wg = (warp - 4) >> 2;
    const uint32_t af0 = smem_u32(acc_full), ae0 = smem_u32(acc_empty);
    int local = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++local) {
      const int z = it / mn_tiles, mn = it - z * mn_tiles;
      const int m0 = (mn / n_tiles) * GEMM_BM, n0 = (mn % n_tiles) * BN;
      const int buf = local & 1;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait_a(af0 + 8 * buf, (local >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = wg * EPI_COLS; c < (wg + 1) * EPI_COLS; c += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16) + c, v);
        tmem_ld_wait();
        if (c + 32 == (wg + 1) * EPI_COLS) {          // last TMEM read of this item: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(ae0 + 8 * buf);
        }
        const int col = n0 + c;
        if (!row_ok || col >= p.N) continue;          // N is a multiple of 32 (host-checked)
        gemm_epilogue_chunk<EPI>(p, v, row + z * p.split_rows, col);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
}

}  // namespace b200
