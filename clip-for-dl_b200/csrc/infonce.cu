// b200clip: fused symmetric InfoNCE (a-N, reference 0426/train.py:154-176) for L2-normalised bf16 embeddings.
//
// The B x B logit matrix S = I T^T / tau is never written to HBM.  With |S| <= 1/tau a fixed shift m = 1/tau
// makes one exponential E = exp(S - m) serve row sums, column sums and both softmax gradients
// (SURVEY.md 8a-N):   loss = m + (sum_i log r_i + sum_j log c_j)/(2B) - (sum_i S_ii)/B
//                      G    = E (1/r_i + 1/c_j)/(2B) - I/B ;  dI = G T / tau ;  dT = G^T I / tau
//
// Pass 1  nce_fwd_kernel : persistent CTAs walk 128x128 S tiles (X row block stationary in smem, Y streamed by TMA,
//                          tcgen05.mma into double-buffered TMEM); two epilogue warpgroups turn tiles into
//                          row sums (in-thread) and column sums (warp butterfly) -> r_part / c_part.
// Pass 2  nce_bwd_kernel : one CTA per (direction, 128-row block, D-half).  Per 32-column tile: recompute S
//                          (tcgen05, N=32), epilogue forms G (bf16) in swizzled smem, a second tcgen05.mma chain
//                          accumulates dX[:, half] += G * Y[:, half] in TMEM (Y tile reused as MN-major operand).
//                          Direction 0: X=I (rows), Y=T;  direction 1: X=T, Y=I  (G is symmetric under r<->c).
// Data-parallel use: I is the rank's local row block [b_loc, D] (global rows row0..), T holds all b_glob rows; column
// sums and dT are per-rank partials that the host combines (all-reduce / reduce-scatter).
#include "common.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

constexpr int NCE_D = 512;                 // shared embedding size (0426/config.py:30)
constexpr int NCE_KC = NCE_D / 64;         // 8 K-chunks of 64 bf16 (one 128-B swizzle row each)
constexpr int NCE_THREADS = 384;           // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue WG0, 8-11 WG1
constexpr int X_CHUNK_BYTES = 128 * 128;   // [128 rows x 64 bf16]
constexpr float LOG2E = 1.4426950408889634f;

// sum over the 32 lanes of v[i] for every i; lane L returns the sum for column L (31 shuffles, halving exchange)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
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

// ================================================================================================
// Pass 1: statistics
// ================================================================================================
struct NceFwdParams {
  int nrows;            // valid X rows (b_loc)
  int ncols;            // valid Y rows (b_glob)
  int num_col_tiles;    // ceil(ncols / BN)
  int total_tiles;      // row_blocks * num_col_tiles
  int tiles_per_cta;
  int r_slots;          // slots per row in r_part
  float k1, k2;         // e = exp2(s * k1 - k2)
  float* r_part;        // [r_slots][nrows_pad]   zero-initialised by the host
  float* c_part;        // [row_blocks][ncols]
  int nrows_pad;
};

template <int BN, int STAGES>
constexpr int nce_fwd_smem_bytes() {
  return NCE_KC * X_CHUNK_BYTES + STAGES * BN * 128 + 2 * 4 * BN * 4 + 512 + 1024;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(NCE_THREADS, 1)
nce_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
               const NceFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = sX + NCE_KC * X_CHUNK_BYTES;
  float* scratch = reinterpret_cast<float*>(sY + STAGES * BN * 128);        // [2 WG][4 warps][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 2 * 4 * BN);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 1;
  uint64_t* y_full = bars + 2;
  uint64_t* y_empty = y_full + STAGES;
  uint64_t* s_full = y_empty + STAGES;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * p.tiles_per_cta;
  const int t1 = min(p.total_tiles, t0 + p.tiles_per_cta);
  const int CT = p.num_col_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 128);
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
    if (lane == 0) {
      int y_it = 0, xcount = 0, cur_rb = -1;
      for (int t = t0; t < t1; ++t) {
        const int rb = t / CT, ct = t - rb * CT;
        if (rb != cur_rb) {
          mbar_wait(x_empty, (xcount & 1) ^ 1);
          mbar_arrive_expect_tx(x_full, NCE_KC * X_CHUNK_BYTES);
#pragma unroll
          for (int kc = 0; kc < NCE_KC; ++kc) tma_load_2d(sX + kc * X_CHUNK_BYTES, &tmap_x, x_full, kc * 64, rb * 128);
          ++xcount;
          cur_rb = rb;
        }
        for (int kc = 0; kc < NCE_KC; ++kc, ++y_it) {
          const int s = y_it % STAGES;
          mbar_wait(&y_empty[s], ((y_it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&y_full[s], BN * 128);
          tma_load_2d(sY + s * BN * 128, &tmap_y, &y_full[s], kc * 64, ct * BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
      int y_it = 0, xcount = 0, cur_rb = -1, n = 0;
      for (int t = t0; t < t1; ++t, ++n) {
        const int rb = t / CT;
        if (rb != cur_rb) {
          mbar_wait(x_full, xcount & 1);
          ++xcount;
          cur_rb = rb;
        }
        const int buf = n & 1;
        mbar_wait(&s_empty[buf], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kc = 0; kc < NCE_KC; ++kc, ++y_it) {
          const int s = y_it % STAGES;
          mbar_wait(&y_full[s], (y_it / STAGES) & 1);
          tc_fence_after();
          const uint32_t xa = smem_u32(sX + kc * X_CHUNK_BYTES);
          const uint32_t yb = smem_u32(sY + s * BN * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ss(d_tmem, desc_kmajor_sw128(xa + j * 32), desc_kmajor_sw128(yb + j * 32), idesc, (kc | j) ? 1u : 0u);
          tc_commit(&y_empty[s]);
        }
        tc_commit(&s_full[buf]);
        if (t + 1 == t1 || (t + 1) / CT != rb) tc_commit(x_empty);
      }
    }
  } else if (warp >= 4) {
    const int w = (warp - 4) >> 2;                 // epilogue warpgroup
    const int q = warp & 3;                        // TMEM lane quadrant
    const int tid_wg = threadIdx.x - 128 - w * 128;
    float* my_scratch = scratch + w * 4 * BN;
    float rsum = 0.f;
    int cur_rb = -1;
    int n = w;
    auto flush_rsum = [&](int rb) {
      const int row = rb * 128 + q * 32 + lane;
      const int first_cta = (rb * CT) / p.tiles_per_cta;
      const int slot = 2 * (static_cast<int>(blockIdx.x) - first_cta) + w;
      if (row < p.nrows) p.r_part[static_cast<long long>(slot) * p.nrows_pad + row] = rsum;
      rsum = 0.f;
    };
    for (int t = t0 + w; t < t1; t += 2, n += 2) {
      const int rb = t / CT, ct = t - rb * CT;
      if (rb != cur_rb) {
        if (cur_rb >= 0) flush_rsum(cur_rb);
        cur_rb = rb;
      }
      const int row = rb * 128 + q * 32 + lane;
      const bool row_ok = row < p.nrows;
      mbar_wait(&s_full[w], (n >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + w * BN + c, v);
        tmem_ld_wait();
        if (c + 32 == BN) {                         // last TMEM read of this tile: hand the buffer back
          tc_fence_before();
          mbar_arrive(&s_empty[w]);
        }
        const int col0 = ct * BN + c;
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x = fast_exp2(fmaf(__uint_as_float(v[i]), p.k1, -p.k2));
          e[i] = (row_ok && (col0 + i) < p.ncols) ? x : 0.f;
          rsum += e[i];
        }
        my_scratch[q * BN + c + lane] = warp_colsum32(e, lane);
      }
      named_bar_sync(1 + w, 128);
      {
        const int col = ct * BN + tid_wg;
        if (tid_wg < BN && col < p.ncols)
          p.c_part[static_cast<long long>(rb) * p.ncols + col] =
              (my_scratch[tid_wg] + my_scratch[BN + tid_wg]) + (my_scratch[2 * BN + tid_wg] + my_scratch[3 * BN + tid_wg]);
      }
      named_bar_sync(1 + w, 128);
    }
    if (cur_rb >= 0) flush_rsum(cur_rb);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BN);
}

// ================================================================================================
// Pass 2: gradients
// ================================================================================================
constexpr int BWD_BN = 32;                  // columns per S tile
constexpr int BWD_RA = 6;                   // ring A slots (Y K-chunks outside this CTA's D-half), 4 KB each
constexpr int BWD_RB = 12;                  // ring B slots (Y K-chunks inside the D-half; 3 tiles x 4 chunks)
constexpr int BWD_SLOT_BYTES = BWD_BN * 128;
constexpr int BWD_G_BYTES = 128 * 128;      // [128 rows x 64 bf16]: even tiles use K cols 0-31, odd tiles 32-63
constexpr int nce_bwd_smem_bytes() {
  return NCE_KC * X_CHUNK_BYTES + (BWD_RA + BWD_RB) * BWD_SLOT_BYTES + BWD_G_BYTES + 512 + 1024;
}

struct NceBwdParams {
  int nrows[2];             // valid X rows per direction
  int ncols[2];             // valid Y rows per direction
  int diag_off[2];          // diagonal: y column == x row + diag_off   (dir0: +row0, dir1: -row0)
  const float* row_stat[2]; // 0.5 / r or c for X rows
  const float* col_stat[2]; // 0.5 / c or r for Y rows
  float* out[2];            // dX [nrows, 512] f32
  float k1, k2;
  float out_scale;          // 1 / (B_glob * tau)
  const float* grad_scale;  // optional device scalar multiplied into out_scale (upstream dLoss)
};

__global__ void __launch_bounds__(NCE_THREADS, 1)
nce_bwd_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_y0,
               const __grid_constant__ CUtensorMap tmap_x1, const __grid_constant__ CUtensorMap tmap_y1,
               const NceBwdParams p) {
  const int dir = blockIdx.z;
  const int h = blockIdx.y;                         // D-half owned by this CTA
  const int rb = blockIdx.x;
  const int nrows = p.nrows[dir], ncols = p.ncols[dir];
  if (rb * 128 >= nrows) return;                    // uniform per CTA, before any barrier / allocation
  const CUtensorMap* tmap_x = dir == 0 ? &tmap_x0 : &tmap_x1;
  const CUtensorMap* tmap_y = dir == 0 ? &tmap_y0 : &tmap_y1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sA = sX + NCE_KC * X_CHUNK_BYTES;
  uint8_t* sB = sA + BWD_RA * BWD_SLOT_BYTES;
  uint8_t* sG = sB + BWD_RB * BWD_SLOT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + BWD_G_BYTES);
  uint64_t* x_full = bars;                          // 1
  uint64_t* a_full = bars + 1;                      // RA
  uint64_t* a_empty = a_full + BWD_RA;              // RA
  uint64_t* b_full = a_empty + BWD_RA;              // RB
  uint64_t* b_empty = b_full + BWD_RB;              // RB/4 (per tile group)
  uint64_t* s_full = b_empty + BWD_RB / 4;          // 2
  uint64_t* s_empty = s_full + 2;                   // 2
  uint64_t* g_full = s_empty + 2;                   // 2
  uint64_t* g_empty = g_full + 2;                   // 2
  uint64_t* acc_full = g_empty + 2;                 // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = (ncols + BWD_BN - 1) / BWD_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmap_x);
    tma_prefetch_desc(tmap_y);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < BWD_RA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < BWD_RB; ++s) mbar_init(&b_full[s], 1);
    for (int s = 0; s < BWD_RB / 4; ++s) mbar_init(&b_empty[s], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 128);
      mbar_init(&g_full[b], 128);
      mbar_init(&g_empty[b], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;              // 256 fp32 columns: dX[:, h*256 .. +256)
  const uint32_t tmem_s = tmem_base + 256;          // 2 x 32 columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(x_full, NCE_KC * X_CHUNK_BYTES);
#pragma unroll
      for (int kc = 0; kc < NCE_KC; ++kc) tma_load_2d(sX + kc * X_CHUNK_BYTES, tmap_x, x_full, kc * 64, rb * 128);
      int a_it = 0;
      for (int n = 0; n < nt; ++n) {
        const int grp = n % (BWD_RB / 4);
        for (int kc = 0; kc < NCE_KC; ++kc) {
          if ((kc >> 2) == h) {
            const int slot = grp * 4 + (kc & 3);
            if ((kc & 3) == 0) mbar_wait(&b_empty[grp], ((n / (BWD_RB / 4)) & 1) ^ 1);
            mbar_arrive_expect_tx(&b_full[slot], BWD_SLOT_BYTES);
            tma_load_2d(sB + slot * BWD_SLOT_BYTES, tmap_y, &b_full[slot], kc * 64, n * BWD_BN);
          } else {
            const int slot = a_it % BWD_RA;
            mbar_wait(&a_empty[slot], ((a_it / BWD_RA) & 1) ^ 1);
            mbar_arrive_expect_tx(&a_full[slot], BWD_SLOT_BYTES);
            tma_load_2d(sA + slot * BWD_SLOT_BYTES, tmap_y, &a_full[slot], kc * 64, n * BWD_BN);
            ++a_it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, BWD_BN, false, false);
      constexpr uint32_t idesc_g = make_idesc_bf16(128, 256, false, true);      // B = Y tile read MN-major
      int a_it = 0;
      auto issue_s = [&](int n) {
        const int buf = n & 1;
        const int grp = n % (BWD_RB / 4);
        mbar_wait(&s_empty[buf], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_s + buf * BWD_BN;
        for (int kc = 0; kc < NCE_KC; ++kc) {
          uint32_t yb;
          const bool in_half = (kc >> 2) == h;
          int slot;
          if (in_half) {
            slot = grp * 4 + (kc & 3);
            mbar_wait(&b_full[slot], (n / (BWD_RB / 4)) & 1);
            yb = smem_u32(sB + slot * BWD_SLOT_BYTES);
          } else {
            slot = a_it % BWD_RA;
            mbar_wait(&a_full[slot], (a_it / BWD_RA) & 1);
            yb = smem_u32(sA + slot * BWD_SLOT_BYTES);
          }
          tc_fence_after();
          const uint32_t xa = smem_u32(sX + kc * X_CHUNK_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ss(d_tmem, desc_kmajor_sw128(xa + j * 32), desc_kmajor_sw128(yb + j * 32), idesc_s, (kc | j) ? 1u : 0u);
          if (!in_half) {
            tc_commit(&a_empty[slot]);
            ++a_it;
          }
        }
        tc_commit(&s_full[buf]);
      };
      auto issue_g = [&](int n) {
        const int par = n & 1;
        const int grp = n % (BWD_RB / 4);
        mbar_wait(&g_full[par], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t ga = smem_u32(sG) + par * 64;
        const uint32_t yb = smem_u32(sB + grp * 4 * BWD_SLOT_BYTES);
#pragma unroll
        for (int jj = 0; jj < BWD_BN / 16; ++jj)
          mma_ss(tmem_acc, desc_kmajor_sw128(ga + jj * 32), desc_mnmajor_sw128(yb + jj * 2048, BWD_SLOT_BYTES), idesc_g,
                 (n | jj) ? 1u : 0u);
        tc_commit(&b_empty[grp]);
        tc_commit(&g_empty[par]);
      };
      mbar_wait(x_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int n = 0; n < nt; ++n) {
        if (n + 1 < nt) issue_s(n + 1);
        issue_g(n);
      }
      tc_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups =====================
    const int w = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row_l = q * 32 + lane;
    const int row = rb * 128 + row_l;
    const bool row_ok = row < nrows;
    const float rstat = row_ok ? p.row_stat[dir][row] : 0.f;
    const float* cstat = p.col_stat[dir];
    const int diag_col = row + p.diag_off[dir];
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    for (int n = w; n < nt; n += 2) {
      mbar_wait(&s_full[w], (n >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_x32(tmem_s + lane_base + w * BWD_BN, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[w]);
      const int col0 = n * BWD_BN;
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float cs[4];
        if (col0 + i + 3 < ncols) {
          const float4 c4 = *reinterpret_cast<const float4*>(cstat + col0 + i);
          cs[0] = c4.x; cs[1] = c4.y; cs[2] = c4.z; cs[3] = c4.w;
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) cs[t] = (col0 + i + t < ncols) ? cstat[col0 + i + t] : 0.f;
        }
        float g[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = col0 + i + t;
          const float e = fast_exp2(fmaf(__uint_as_float(v[i + t]), p.k1, -p.k2));
          float gv = e * (rstat + cs[t]);
          if (col == diag_col) gv -= 1.0f;               // G_ii = p_ii - 1 rounded as a whole: error relative to G_ii itself
          g[t] = (col < ncols) ? gv : 0.f;
        }
        packed[i / 2] = pack_bf16x2(g[0], g[1]);
        packed[i / 2 + 1] = pack_bf16x2(g[2], g[3]);
      }
      mbar_wait(&g_empty[w], ((n >> 1) & 1) ^ 1);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(sG + sw128_offset(row_l, w * 4 + c)) =
            make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]);
      fence_proxy_async_smem();
      mbar_arrive(&g_full[w]);
    }
    // final: dX[:, h*256 + w*128 .. +128) = acc * scale
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float scale = p.out_scale;
    if (p.grad_scale) scale *= *p.grad_scale;
    float* orow = p.out[dir] + static_cast<long long>(row) * NCE_D + h * 256 + w * 128;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem_acc + lane_base + w * 128 + c, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(orow + c + i) =
              make_float4(__uint_as_float(v[i]) * scale, __uint_as_float(v[i + 1]) * scale,
                          __uint_as_float(v[i + 2]) * scale, __uint_as_float(v[i + 3]) * scale);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================
// small reductions around the two passes
// ================================================================================================
// r[i] = sum_slots r_part ; c[j] = sum_rb c_part
__global__ void nce_reduce_stats_kernel(const float* __restrict__ r_part, int r_slots, int nrows_pad, int nrows,
                                        const float* __restrict__ c_part, int row_blocks, int ncols,
                                        float* __restrict__ r, float* __restrict__ c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) {
    float acc = 0.f;
    for (int s = 0; s < r_slots; ++s) acc += r_part[static_cast<long long>(s) * nrows_pad + i];
    r[i] = acc;
  }
  if (i < ncols) {
    float acc = 0.f;
    for (int b = 0; b < row_blocks; ++b) acc += c_part[static_cast<long long>(b) * ncols + i];
    c[i] = acc;
  }
}

// sums[0] = sum_i log r_i (local rows), sums[1] = sum_{j in [c_lo,c_hi)} log c_j, sums[2] = sum_i S_ii  (S_ii = I_i.T_{row0+i}/tau)
// also writes rinvh = 0.5 / r, cinvh = 0.5 / c.  Deterministic: per-block partials, last block folds them in order.
__global__ void __launch_bounds__(256) nce_loss_kernel(const __nv_bfloat16* __restrict__ I, const __nv_bfloat16* __restrict__ T,
                                                       int b_loc, int b_glob, int row0, float inv_tau,
                                                       const float* __restrict__ r, const float* __restrict__ c, int c_lo,
                                                       int c_hi, float* __restrict__ rinvh, float* __restrict__ cinvh,
                                                       double* __restrict__ partial /*[grid][3]*/,
                                                       unsigned int* __restrict__ counter, double* __restrict__ sums,
                                                       float* __restrict__ loss, float shift_m) {
  __shared__ double red[3][8];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a_logr = 0.0, a_logc = 0.0, a_diag = 0.0;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gsz = gridDim.x * blockDim.x;
  for (int i = gtid; i < b_loc; i += gsz) {
    const float ri = r[i];
    a_logr += static_cast<double>(logf(ri));
    rinvh[i] = 0.5f / ri;
  }
  for (int j = gtid; j < b_glob; j += gsz) {
    const float cj = c[j];
    cinvh[j] = 0.5f / cj;
    if (j >= c_lo && j < c_hi) a_logc += static_cast<double>(logf(cj));
  }
  // diagonal: one warp per row, 512 bf16 = 16 per lane
  for (int i = blockIdx.x * 8 + warp; i < b_loc; i += gridDim.x * 8) {
    const uint4* pi = reinterpret_cast<const uint4*>(I + static_cast<long long>(i) * NCE_D) + lane * 2;
    const uint4* pt = reinterpret_cast<const uint4*>(T + static_cast<long long>(row0 + i) * NCE_D) + lane * 2;
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint4 a = pi[k], b = pt[k];
      d += bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
           bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
    }
    d = warp_sum(d);
    if (lane == 0) a_diag += static_cast<double>(d * inv_tau);
  }
  double vals[3] = {a_logr, a_logc, a_diag};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
    partial[blockIdx.x * 3 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double s[3] = {0.0, 0.0, 0.0};
    for (unsigned b = 0; b < gridDim.x; ++b)
      for (int k = 0; k < 3; ++k) s[k] += partial[b * 3 + k];
    sums[0] = s[0]; sums[1] = s[1]; sums[2] = s[2];
    if (loss) *loss = static_cast<float>(static_cast<double>(shift_m) + (s[0] + s[1]) / (2.0 * b_glob) - s[2] / b_glob);
    *counter = 0;
  }
}

constexpr int FWD_BN = 128;
constexpr int FWD_STAGES = 5;

struct NceFwdPlan {
  int row_blocks, col_tiles, total_tiles, tiles_per_cta, grid, r_slots, nrows_pad;
  size_t r_part_bytes, c_part_bytes, total_bytes;
};
static NceFwdPlan plan_fwd(long long b_loc, long long b_glob) {
  NceFwdPlan pl{};
  pl.row_blocks = static_cast<int>((b_loc + 127) / 128);
  pl.col_tiles = static_cast<int>((b_glob + FWD_BN - 1) / FWD_BN);
  pl.total_tiles = pl.row_blocks * pl.col_tiles;
  const int sms = num_sms();
  pl.tiles_per_cta = (pl.total_tiles + sms - 1) / sms;
  if (pl.tiles_per_cta < 1) pl.tiles_per_cta = 1;
  pl.grid = (pl.total_tiles + pl.tiles_per_cta - 1) / pl.tiles_per_cta;
  pl.r_slots = 2 * ((pl.col_tiles + pl.tiles_per_cta - 1) / pl.tiles_per_cta + 1);
  pl.nrows_pad = pl.row_blocks * 128;
  pl.r_part_bytes = static_cast<size_t>(pl.r_slots) * pl.nrows_pad * sizeof(float);
  pl.c_part_bytes = static_cast<size_t>(pl.row_blocks) * b_glob * sizeof(float);
  pl.total_bytes = ((pl.r_part_bytes + 255) & ~size_t(255)) + ((pl.c_part_bytes + 255) & ~size_t(255)) + 4096 * 3 * sizeof(double) + 256;
  return pl;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200clip_infonce_workspace_bytes(long long b_loc, long long b_glob) {
  if (b_loc <= 0 || b_glob <= 0) return 0;
  return plan_fwd(b_loc, b_glob).total_bytes;
}

extern "C" int b200clip_infonce_fwd_stats(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob,
                                          float temperature, float* r, float* c_partial, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  B200_REQUIRE(D == NCE_D, "infonce: D=%d unsupported (kernels are built for D=%d)", D, NCE_D);
  B200_REQUIRE(b_loc > 0 && b_glob > 0 && b_loc <= b_glob, "infonce: need 0 < b_loc <= b_glob (got %lld, %lld)", b_loc, b_glob);
  B200_REQUIRE(temperature > 0.f, "infonce: temperature must be positive");
  B200_REQUIRE(b_glob < (1ll << 30), "infonce: batch too large");
  const NceFwdPlan pl = plan_fwd(b_loc, b_glob);
  if (workspace_bytes < pl.total_bytes) return fail(B200_ERR_WORKSPACE, "infonce_fwd: workspace %zu < %zu", workspace_bytes, pl.total_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* r_part = reinterpret_cast<float*>(ws);
  float* c_part = reinterpret_cast<float*>(ws + ((pl.r_part_bytes + 255) & ~size_t(255)));
  B200_CHECK_CUDA(cudaMemsetAsync(r_part, 0, pl.r_part_bytes, s));

  CUtensorMap tx, ty;
  int rc = make_tmap_bf16_2d(&tx, i_hat, b_loc, D, D, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&ty, t_hat, b_glob, D, D, 64, FWD_BN);
  if (rc) return rc;
  NceFwdParams p{};
  p.nrows = (int)b_loc; p.ncols = (int)b_glob; p.num_col_tiles = pl.col_tiles; p.total_tiles = pl.total_tiles;
  p.tiles_per_cta = pl.tiles_per_cta; p.r_slots = pl.r_slots; p.k1 = LOG2E / temperature; p.k2 = LOG2E / temperature;
  p.r_part = r_part; p.c_part = c_part; p.nrows_pad = pl.nrows_pad;
  auto kern = nce_fwd_kernel<FWD_BN, FWD_STAGES>;
  constexpr int smem = nce_fwd_smem_bytes<FWD_BN, FWD_STAGES>();
  static bool configured = false;
  if (!configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  kern<<<pl.grid, NCE_THREADS, smem, s>>>(tx, ty, p);
  B200_LAUNCH_CHECK();
  const int n = (int)std::max(b_loc, b_glob);
  nce_reduce_stats_kernel<<<(n + 255) / 256, 256, 0, s>>>(r_part, pl.r_slots, pl.nrows_pad, (int)b_loc, c_part, pl.row_blocks,
                                                        (int)b_glob, r, c_partial);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_infonce_loss(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob,
                                     long long row0, float temperature, const float* r, const float* c, long long c_lo,
                                     long long c_hi, float* rinvh, float* cinvh, double* sums, float* loss, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  B200_REQUIRE(D == NCE_D, "infonce: D=%d unsupported (kernels are built for D=%d)", D, NCE_D);
  B200_REQUIRE(b_loc > 0 && b_glob > 0 && row0 >= 0 && row0 + b_loc <= b_glob, "infonce_loss: bad row range");
  const NceFwdPlan pl = plan_fwd(b_loc, b_glob);
  if (workspace_bytes < pl.total_bytes) return fail(B200_ERR_WORKSPACE, "infonce_loss: workspace too small");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint8_t* tail = ws + ((pl.r_part_bytes + 255) & ~size_t(255)) + ((pl.c_part_bytes + 255) & ~size_t(255));
  double* partial = reinterpret_cast<double*>(tail);
  unsigned int* counter = reinterpret_cast<unsigned int*>(tail + 4096 * 3 * sizeof(double));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
  int grid = (int)std::min<long long>((b_glob + 255) / 256, 1024);
  nce_loss_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(i_hat), static_cast<const __nv_bfloat16*>(t_hat),
                                       (int)b_loc, (int)b_glob, (int)row0, 1.0f / temperature, r, c, (int)c_lo, (int)c_hi,
                                       rinvh, cinvh, partial, counter, sums, loss, 1.0f / temperature);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_infonce_bwd(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob,
                                    long long row0, float temperature, const float* rinvh, const float* cinvh,
                                    const float* grad_scale, float* d_i, float* d_t_partial, void* stream) {
  B200_REQUIRE(D == NCE_D, "infonce: D=%d unsupported (kernels are built for D=%d)", D, NCE_D);
  B200_REQUIRE(b_loc > 0 && b_glob > 0 && row0 >= 0 && row0 + b_loc <= b_glob, "infonce_bwd: bad row range");
  B200_REQUIRE(aligned16(d_i) && aligned16(d_t_partial) && aligned16(rinvh) && aligned16(cinvh), "infonce_bwd: unaligned pointer");
  CUtensorMap tx0, ty0, tx1, ty1;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tx0, i_hat, b_loc, D, D, 64, 128))) return rc;      // dir 0: X = I rows
  if ((rc = make_tmap_bf16_2d(&ty0, t_hat, b_glob, D, D, 64, BWD_BN))) return rc;  //        Y = T columns
  if ((rc = make_tmap_bf16_2d(&tx1, t_hat, b_glob, D, D, 64, 128))) return rc;     // dir 1: X = T rows
  if ((rc = make_tmap_bf16_2d(&ty1, i_hat, b_loc, D, D, 64, BWD_BN))) return rc;   //        Y = I columns
  NceBwdParams p{};
  p.nrows[0] = (int)b_loc;  p.ncols[0] = (int)b_glob; p.diag_off[0] = (int)row0;
  p.nrows[1] = (int)b_glob; p.ncols[1] = (int)b_loc;  p.diag_off[1] = -(int)row0;
  p.row_stat[0] = rinvh; p.col_stat[0] = cinvh; p.out[0] = d_i;
  p.row_stat[1] = cinvh; p.col_stat[1] = rinvh; p.out[1] = d_t_partial;
  p.k1 = LOG2E / temperature; p.k2 = LOG2E / temperature;
  p.out_scale = 1.0f / (static_cast<float>(b_glob) * temperature);
  p.grad_scale = grad_scale;
  constexpr int smem = nce_bwd_smem_bytes();
  static bool configured = false;
  if (!configured) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(nce_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>((b_glob + 127) / 128), 2, 2);
  nce_bwd_kernel<<<grid, NCE_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tx0, ty0, tx1, ty1, p);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
