// b200clip: fused symmetric InfoNCE (a-N, reference 0426/train.py:154-176) for L2-normalised bf16 embeddings.
//
// The B x B logit matrix S = I T^T / tau is never written to HBM.  With |S| <= 1/tau a fixed shift m = 1/tau
// makes one exponential E = exp(S - m) serve row sums, column sums and both softmax gradients
// (SURVEY.md 8a-N):   loss = m + (sum_i log r_i + sum_j log c_j)/(2B) - (sum_i S_ii)/B
//                      G    = E (1/r_i + 1/c_j)/(2B) - I/B ;  dI = G T / tau ;  dT = G^T I / tau
//
// Pass 1  nce_fwd2_kernel (round 2): a CTA PAIR (cluster 2x1x1) owns two stacked 128-row blocks and issues one
//                          tcgen05.mma.cta_group::2 of M = 256, N = 256 per k16 step; each CTA supplies its own rows of X and
//                          half of the Y tile, 16 epilogue warps per CTA turn the 128 x 256 fp32 tile into row sums
//                          (in-thread) and column sums (31-shuffle exchange) -> r_part / c_part.  D = 512: X stationary in
//                          smem; any other D = 64k <= 1024: X chunks streamed with Y.
//         nce_fwd_kernel : round 1's one-CTA-per-SM kernel (128 x 128 tiles), kept behind B200CLIP_FWD_VARIANT=1.
// Pass 2  nce_bwdc_kernel<NC> (infonce_bwd.cuh): cluster of NC = D/256 CTAs per (direction, 128-row block, column split); the
//                          CTAs own 256-wide slices of dX.  Per 32-column tile the owner CTA recomputes S (tcgen05, N = 32,
//                          X in TMEM), forms G (bf16) and sends it to its peers; every CTA accumulates
//                          dX[:, slice] += G * Y[:, slice] in TMEM (Y tile reused as MN-major operand).
//                          Direction 0: X=I (rows), Y=T;  direction 1: X=T, Y=I  (G is symmetric under r<->c).
//         nce_bwd4_kernel: round 1's D = 512 pair kernel, kept behind B200CLIP_BWD_VARIANT=4 (tests/test_gpu_variants.py).
// Data-parallel use: I is the rank's local row block [b_loc, D] (global rows row0..), T holds all b_glob rows; column
// sums and dT are per-rank partials that the host combines (all-reduce / reduce-scatter).
#include <stdlib.h>

#include "common.cuh"
#include "host.cuh"
#include "infonce_bwd.cuh"
#include "../../include/b200clip.h"

namespace b200 {

constexpr int NCE_D = 512;                 // shared embedding size (0426/config.py:30)
constexpr int NCE_KC = NCE_D / 64;         // 8 K-chunks of 64 bf16 (one 128-B swizzle row each)

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
  const __nv_bfloat16* x;   // X matrix (rows of the stationary operand), read by the TMEM-resident variant
};

// Issue loops below are WARP-UNIFORM (all 32 lanes run the loop, one elected lane issues TMA / tcgen05 instructions):
// descriptor words then live in uniform registers and an MMA costs one 32-bit add per operand.  The first version
// ran them under `if (lane == 0)`; ncu showed ~230 SASS instructions per 4 MMAs and the tensor pipe 14-33 % busy.
//
// XT = true (experimental, off by default -- see FWD_XT below): the stationary X row block lives in TMEM (256 columns of packed bf16 pairs) and is the A operand of
// tcgen05.mma in its TS form.  With X in shared memory every N = 128 MMA read 8 KB of operands per 64 clk -- exactly the
// 128 B/clk the shared-memory port delivers (tools/mma_rate.cu) -- while TMA wrote another 64 B/clk of Y into the same
// memory: the kernel was shared-memory-bound at ~56 % of the MMA rate.  TS form: operand reads drop to the Y tile
// (64 B/clk), the 128 KB of X leave shared memory (24-stage Y ring), S tiles are 128 x 64 (4 TMEM buffers x 64 columns).
// X is (re)loaded at every row-block switch by warps 0-3 (one TMEM lane quadrant each) straight from global memory.
template <int BN, int STAGES, int NWG, bool XT>
constexpr int nce_fwd_smem_bytes() {
  return (XT ? 0 : NCE_KC * X_CHUNK_BYTES) + STAGES * BN * 128 + NWG * 4 * BN * 4 + 512 + 1024;
}

template <int BN, int STAGES, int NWG, bool XT>
__global__ void __launch_bounds__(128 + NWG * 128, 1)
nce_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
               const NceFwdParams p) {
  // S buffers in TMEM: one per epilogue warpgroup, or (XT) two buffers shared by pairs of warpgroups that split the columns
  constexpr int NBUF = XT ? 2 : NWG;
  constexpr int WPB = NWG / NBUF;                 // warpgroups per buffer
  constexpr int CW = BN / WPB;                    // columns per warpgroup
  static_assert(!XT || NBUF * BN <= 256, "TMEM: 256 columns of X + the S buffers");
  static_assert(NWG % NBUF == 0 && CW % 32 == 0, "column split");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = sX + (XT ? 0 : NCE_KC * X_CHUNK_BYTES);
  float* scratch = reinterpret_cast<float*>(sY + STAGES * BN * 128);        // [NWG][4 warps][BN]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + NWG * 4 * BN);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + 1;
  uint64_t* y_full = bars + 2;
  uint64_t* y_empty = y_full + STAGES;
  uint64_t* s_full = y_empty + STAGES;          // NWG
  uint64_t* s_empty = s_full + NWG;             // NWG
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + NWG);
  constexpr int Y_BYTES = BN * 128;

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
    mbar_init(x_full, XT ? 128 : 1);
    mbar_init(x_empty, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&y_full[s], 1);
      mbar_init(&y_empty[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 128 * WPB);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, XT ? 512 : NWG * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_sbuf = tmem_base + (XT ? 256 : 0);      // S buffers; XT: columns [0, 256) hold X

  // XT: warps 0-3 refill the X operand in TMEM at a row-block switch (warp w owns TMEM lanes 32w .. 32w+31 = rows)
  auto load_x_to_tmem = [&](int rb, uint32_t xph) {
    mbar_wait_a(smem_u32(x_empty), xph ^ 1);                  // every MMA that read the previous row block has completed
    tc_fence_after();
    const int row = rb * 128 + warp * 32 + lane;
    const bool row_ok = row < p.nrows;
    const uint4* src = reinterpret_cast<const uint4*>(p.x + static_cast<long long>(row_ok ? row : 0) * NCE_D);
    const uint32_t dst = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
    for (int kc = 0; kc < NCE_KC; ++kc) {
      uint32_t xr[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 t = row_ok ? __ldg(src + kc * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
        xr[4 * i] = t.x; xr[4 * i + 1] = t.y; xr[4 * i + 2] = t.z; xr[4 * i + 3] = t.w;
      }
      tmem_st_x32(dst + kc * 32, xr);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive_a(smem_u32(x_full));
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t xf = smem_u32(x_full), xe = smem_u32(x_empty), yf0 = smem_u32(y_full), ye0 = smem_u32(y_empty);
    const uint32_t sx = smem_u32(sX), sy = smem_u32(sY);
    int s = 0, cur_rb = -1;
    uint32_t ph = 0, xph = 0;
    int rb = t0 / CT, ct = t0 - rb * CT;
    for (int t = t0; t < t1; ++t) {
      if (rb != cur_rb) {
        if constexpr (XT) {
          load_x_to_tmem(rb, xph);
        } else {
          mbar_wait_a(xe, xph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx_a(xf, NCE_KC * X_CHUNK_BYTES);
#pragma unroll
            for (int kc = 0; kc < NCE_KC; ++kc) tma_load_2d_a(sx + kc * X_CHUNK_BYTES, &tmap_x, xf, kc * 64, rb * 128);
          }
          __syncwarp();
        }
        xph ^= 1;
        cur_rb = rb;
      }
#pragma unroll 1
      for (int kc = 0; kc < NCE_KC; ++kc) {
        mbar_wait_a(ye0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx_a(yf0 + 8 * s, Y_BYTES);
          tma_load_2d_a(sy + s * Y_BYTES, &tmap_y, yf0 + 8 * s, kc * 64, ct * BN);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (++ct == CT) { ct = 0; ++rb; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, false, false);
    const uint32_t xf = smem_u32(x_full), xe = smem_u32(x_empty), yf0 = smem_u32(y_full), ye0 = smem_u32(y_empty);
    const uint32_t sf0 = smem_u32(s_full), se0 = smem_u32(s_empty);
    const uint32_t x_lo = desc_lo(smem_u32(sX), 16), y_lo = desc_lo(smem_u32(sY), 16);
    int s = 0, cur_rb = -1, buf = 0;
    uint32_t ph = 0, xph = 0, sph = 0;               // sph: phase of the s_empty ring (flips when buf wraps)
    int rb = t0 / CT, ct = t0 - rb * CT;
    for (int t = t0; t < t1; ++t) {
      if (rb != cur_rb) {
        if constexpr (XT) load_x_to_tmem(rb, xph);
        mbar_wait_a(xf, xph);
        xph ^= 1;
        cur_rb = rb;
      }
      mbar_wait_a(se0 + 8 * buf, sph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_sbuf + buf * BN;
#pragma unroll
      for (int kc = 0; kc < NCE_KC; ++kc) {
        mbar_wait_a(yf0 + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a = x_lo + kc * (X_CHUNK_BYTES >> 4), b = y_lo + s * (Y_BYTES >> 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if constexpr (XT) mma_ts_lo(d_tmem, tmem_base + kc * 32 + j * 8, b + 2 * j, idesc, (kc | j) != 0);
            else mma_ss_lo(d_tmem, a + 2 * j, b + 2 * j, idesc, (kc | j) != 0);
          }
          tc_commit_a(ye0 + 8 * s);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      const bool last_of_rb = (ct + 1 == CT) || (t + 1 == t1);
      if (elect_one()) {
        tc_commit_a(sf0 + 8 * buf);
        if (last_of_rb) tc_commit_a(xe);
      }
      __syncwarp();
      if (++buf == NBUF) { buf = 0; sph ^= 1; }
      if (++ct == CT) { ct = 0; ++rb; }
    }
  } else if (warp < 4) {
    // ===================== warps 2, 3 (XT): their TMEM lane quadrants of X at every row-block switch =====================
    if constexpr (XT) {
      int cur_rb = -1;
      uint32_t xph = 0;
      for (int t = t0; t < t1; ++t) {
        const int rb = t / CT;
        if (rb != cur_rb) {
          load_x_to_tmem(rb, xph);
          xph ^= 1;
          cur_rb = rb;
        }
      }
    }
  } else {
    // ===================== epilogue warpgroups: tile n of this CTA -> buffer n % NBUF; warpgroup w reads buffer w % NBUF,
    // columns [(w / NBUF) * CW, +CW) =====================
    const int w = (warp - 4) >> 2;                 // epilogue warpgroup
    const int q = warp & 3;                        // TMEM lane quadrant
    const int tid_wg = threadIdx.x - 128 - w * 128;
    const int bufw = w % NBUF, cpart = (w / NBUF) * CW;
    float* my_scratch = scratch + w * 4 * BN;
    const uint32_t sf = smem_u32(s_full) + 8 * bufw, se = smem_u32(s_empty) + 8 * bufw;
    const uint32_t t_addr = tmem_sbuf + (static_cast<uint32_t>(q * 32) << 16) + bufw * BN + cpart;
    float rsum = 0.f;
    int cur_rb = -1;
    uint32_t sph = 0;
    auto flush_rsum = [&](int rb) {
      const int row = rb * 128 + q * 32 + lane;
      const int first_cta = (rb * CT) / p.tiles_per_cta;
      const int slot = NWG * (static_cast<int>(blockIdx.x) - first_cta) + w;
      if (row < p.nrows) p.r_part[static_cast<long long>(slot) * p.nrows_pad + row] = rsum;
      rsum = 0.f;
    };
    for (int t = t0 + bufw; t < t1; t += NBUF) {
      const int rb = t / CT, ct = t - rb * CT;
      if (rb != cur_rb) {
        if (cur_rb >= 0) flush_rsum(cur_rb);
        cur_rb = rb;
      }
      const int row = rb * 128 + q * 32 + lane;
      const bool interior = (rb * 128 + 128 <= p.nrows) && (ct * BN + BN <= p.ncols);   // warp-uniform
      mbar_wait_a(sf, sph);
      sph ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < CW; c += 32) {
        uint32_t v[32];
        tmem_ld_x32(t_addr + c, v);
        tmem_ld_wait();
        if (c + 32 == CW) {                         // last TMEM read of this tile: hand the buffer back
          tc_fence_before();
          mbar_arrive_a(se);
        }
        float e[32];
        if (interior) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            e[i] = fast_exp2(fmaf(__uint_as_float(v[i]), p.k1, -p.k2));
            rsum += e[i];
          }
        } else {
          const bool row_ok = row < p.nrows;
          const int col0 = ct * BN + cpart + c;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = fast_exp2(fmaf(__uint_as_float(v[i]), p.k1, -p.k2));
            e[i] = (row_ok && (col0 + i) < p.ncols) ? x : 0.f;
            rsum += e[i];
          }
        }
        my_scratch[q * CW + c + lane] = warp_colsum32(e, lane);
      }
      named_bar_sync(1 + w, 128);
      {
        const int col = ct * BN + cpart + tid_wg;
        if (tid_wg < CW && col < p.ncols)
          p.c_part[static_cast<long long>(rb) * p.ncols + col] =
              (my_scratch[tid_wg] + my_scratch[CW + tid_wg]) + (my_scratch[2 * CW + tid_wg] + my_scratch[3 * CW + tid_wg]);
      }
      named_bar_sync(1 + w, 128);
    }
    if (cur_rb >= 0) flush_rsum(cur_rb);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, XT ? 512 : NWG * BN);
}

// ------------------------------------------------------------------------------------------------
// Pass 1, CTA-pair kernel (default).  nce_fwd_kernel above is shared-memory-port bound: per 64-clk N = 128 MMA the port
// serves 8 KB of operand reads plus 4 KB of TMA writes = 192 B/clk against 128 B/clk available (measured: tensor pipe 67 %).
// Here two CTAs (a cluster 2x1x1 = one TPC) own two stacked row blocks and issue ONE tcgen05.mma.cta_group::2 of M = 256,
// N = 256 per K16 step: each CTA supplies its own 128 rows of X and only HALF of the Y tile (128 of the 256 columns), so per
// 128-clk MMA its port serves 4 KB (A) + 4 KB (B half) + 4 KB (TMA writes of that half) = 96 B/clk -- under the limit, the
// pass is tensor-bound.  TMEM: two 256-column S buffers (all 512 columns) per CTA, each CTA's epilogue (16 warps: lane
// quadrant q x column quarter h) turns its own 128 x 256 fp32 tile into row / column exp-sums exactly like the 1-CTA kernel.
// Protocol: both CTAs' TMA loads complete on the LEADER's full barriers (cta_group::2 TMA, expect-tx armed by the leader for
// both halves); the leader's commits are multicast to both CTAs' empty / s_full barriers; the epilogue warps of both CTAs
// arrive on the leader's s_empty.  XSTAT = true (D = 512): the X row block is stationary in shared memory (128 KB);
// XSTAT = false (any D = 64 KC): X chunks are streamed next to the Y chunks (D = 768 would need 192 KB for a stationary X).
// ------------------------------------------------------------------------------------------------
constexpr int FWD2_BN = 256;               // S tile columns (= Y rows) per MMA
constexpr int FWD2_EPI_WG = 4;             // epilogue warpgroups: column quarters of the tile
constexpr int FWD2_THREADS = 128 + FWD2_EPI_WG * 128;
constexpr int FWD2_CHUNK = 128 * 128;      // [128 rows x 64 bf16]: one X chunk, or one CTA's half of a Y chunk
template <int KC, bool XSTAT>      // KC = 0: run-time K-chunk count (streamed X only)
__host__ __device__ constexpr int nce_fwd2_stages() { return XSTAT ? (227 * 1024 - KC * FWD2_CHUNK - 6 * 1024) / FWD2_CHUNK : 6; }
template <int KC, bool XSTAT>
constexpr int nce_fwd2_smem_bytes() {
  return (XSTAT ? KC * FWD2_CHUNK + nce_fwd2_stages<KC, XSTAT>() * FWD2_CHUNK : nce_fwd2_stages<KC, XSTAT>() * 2 * FWD2_CHUNK) +
         16 * 64 * 4 + 512 + 1024;
}

struct NceFwd2Params {
  int nrows, ncols;         // valid X rows (b_loc), valid Y rows (b_glob)
  int ct2;                  // column super-tiles: ceil(ncols / 256)
  int total;                // row super-blocks * ct2
  int tiles_per_cluster;
  float k1, k2;
  float* r_part;            // [r_slots][nrows_pad]  zero-initialised by the host
  float* c_part;            // [row_blocks(128)][ncols]
  int nrows_pad;
  int kc;                   // D / 64 (used when the kernel is instantiated with a run-time chunk count)
};

template <int KC_T, bool XSTAT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FWD2_THREADS, 1)
nce_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y, const NceFwd2Params p) {
  static_assert(!XSTAT || KC_T > 0, "a stationary X needs a compile-time chunk count");
  constexpr int STAGES = nce_fwd2_stages<KC_T, XSTAT>();
  constexpr int STAGE_BYTES = XSTAT ? FWD2_CHUNK : 2 * FWD2_CHUNK;       // streamed: [X chunk | Y half chunk]
  const int KC = KC_T > 0 ? KC_T : p.kc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sY = sX + (XSTAT ? KC_T * FWD2_CHUNK : 0);
  float* scratch = reinterpret_cast<float*>(sY + STAGES * STAGE_BYTES);   // [16 warps][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 16 * 64);
  uint64_t* x_full = bars;                     // leader: 1 arrival + both CTAs' X bytes
  uint64_t* x_empty = bars + 1;                // each CTA: 1 (multicast commit)
  uint64_t* y_full = bars + 2;                 // leader: 1 arrival + both CTAs' stage bytes
  uint64_t* y_empty = y_full + STAGES;         // each CTA: 1 (multicast commit)
  uint64_t* s_full = y_empty + STAGES;         // each CTA: 1 (multicast commit)            [2]
  uint64_t* s_empty = s_full + 2;              // leader: 2 CTAs x 16 epilogue warps        [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cr = cluster_ctarank();       // 0 = leader
  const int cl = static_cast<int>(blockIdx.x >> 1);
  const int t0 = cl * p.tiles_per_cluster;
  const int t1 = min(p.total, t0 + p.tiles_per_cluster);
  const int CT = p.ct2;

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
      mbar_init(&s_empty[b], 2 * FWD2_EPI_WG * 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();                          // barrier inits visible to the peer before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs: own X rows, own half of every Y tile) =====================
    const uint32_t xf_lead = mapa_u32(smem_u32(x_full), 0), yf0_lead = mapa_u32(smem_u32(y_full), 0);
    const uint32_t xf = smem_u32(x_full), xe = smem_u32(x_empty), yf0 = smem_u32(y_full), ye0 = smem_u32(y_empty);
    const uint32_t sx = smem_u32(sX), sy = smem_u32(sY);
    int s = 0, cur_rb = -1;
    uint32_t ph = 0, xph = 0;
    int rb = t0 / CT, ct = t0 - rb * CT;
    for (int t = t0; t < t1; ++t) {
      const int xrow = rb * 256 + static_cast<int>(cr) * 128, yrow = ct * FWD2_BN + static_cast<int>(cr) * 128;
      if constexpr (XSTAT) {
        if (rb != cur_rb) {
          mbar_wait_a(xe, xph ^ 1);
          if (elect_one()) {
            if (cr == 0) mbar_arrive_expect_tx_a(xf, 2 * KC_T * FWD2_CHUNK);
#pragma unroll
            for (int kc = 0; kc < KC_T; ++kc) tma_load_2d_2cta(sx + kc * FWD2_CHUNK, &tmap_x, xf_lead, kc * 64, xrow);
          }
          __syncwarp();
          xph ^= 1;
          cur_rb = rb;
        }
      }
#pragma unroll 1
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait_a(ye0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          if (cr == 0) mbar_arrive_expect_tx_a(yf0 + 8 * s, 2 * STAGE_BYTES);
          if constexpr (!XSTAT) tma_load_2d_2cta(sy + s * STAGE_BYTES, &tmap_x, yf0_lead + 8 * s, kc * 64, xrow);
          tma_load_2d_2cta(sy + s * STAGE_BYTES + (XSTAT ? 0 : FWD2_CHUNK), &tmap_y, yf0_lead + 8 * s, kc * 64, yrow);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (++ct == CT) { ct = 0; ++rb; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (cr == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, FWD2_BN, false, false);
      const uint32_t xf = smem_u32(x_full), xe = smem_u32(x_empty), yf0 = smem_u32(y_full), ye0 = smem_u32(y_empty);
      const uint32_t sf0 = smem_u32(s_full), se0 = smem_u32(s_empty);
      const uint32_t x_lo = desc_lo(smem_u32(sX), 16), y_lo = desc_lo(smem_u32(sY), 16);
      int s = 0, cur_rb = -1, buf = 0;
      uint32_t ph = 0, xph = 0, sph = 0;
      int rb = t0 / CT, ct = t0 - rb * CT;
      for (int t = t0; t < t1; ++t) {
        if constexpr (XSTAT) {
          if (rb != cur_rb) {
            mbar_wait_cluster_a(xf, xph);
            xph ^= 1;
            cur_rb = rb;
          }
        }
        mbar_wait_cluster_a(se0 + 8 * buf, sph ^ 1);        // both CTAs' epilogues have drained this S buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * FWD2_BN;
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait_cluster_a(yf0 + 8 * s, ph);              // both halves (and, streamed, both X chunks) have landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a = XSTAT ? x_lo + kc * (FWD2_CHUNK >> 4) : y_lo + s * (STAGE_BYTES >> 4);
            const uint32_t b = y_lo + s * (STAGE_BYTES >> 4) + (XSTAT ? 0 : (FWD2_CHUNK >> 4));
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_ss_lo_2cta(d_tmem, a + 2 * j, b + 2 * j, idesc, (kc | j) != 0);
            tc_commit_2cta_multicast_a(ye0 + 8 * s, 0x3);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        const bool last_of_rb = (ct + 1 == CT) || (t + 1 == t1);
        if (elect_one()) {
          tc_commit_2cta_multicast_a(sf0 + 8 * buf, 0x3);
          if (XSTAT && last_of_rb) tc_commit_2cta_multicast_a(xe, 0x3);
        }
        __syncwarp();
        if (++buf == 2) { buf = 0; sph ^= 1; }
        if (++ct == CT) { ct = 0; ++rb; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 16 warps = lane quadrant q (warp & 3) x column quarter h; every warp sees every tile =====
    const int h = (warp - 4) >> 2;
    const int q = warp & 3;
    const int tid_wg = threadIdx.x - 128 - h * 128;
    float* my_scratch = scratch + h * 4 * 64;                 // [4 quadrants][64 columns] of this column quarter
    const uint32_t sf0 = smem_u32(s_full);
    const uint32_t se0_lead = mapa_u32(smem_u32(s_empty), 0);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 64;
    float rsum = 0.f;
    int cur_rb = -1, buf = 0;
    uint32_t sph = 0;
    auto flush_rsum = [&](int rb) {
      const int row = rb * 256 + static_cast<int>(cr) * 128 + q * 32 + lane;
      const int first_cl = (rb * CT) / p.tiles_per_cluster;
      const int slot = FWD2_EPI_WG * (cl - first_cl) + h;
      if (row < p.nrows) p.r_part[static_cast<long long>(slot) * p.nrows_pad + row] = rsum;
      rsum = 0.f;
    };
    int rb = t0 / CT, ct = t0 - rb * CT;
    for (int t = t0; t < t1; ++t) {
      if (rb != cur_rb) {
        if (cur_rb >= 0) flush_rsum(cur_rb);
        cur_rb = rb;
      }
      const int row0 = rb * 256 + static_cast<int>(cr) * 128;
      const int row = row0 + q * 32 + lane;
      const int col_base = ct * FWD2_BN + h * 64;
      const bool interior = (row0 + 128 <= p.nrows) && (ct * FWD2_BN + FWD2_BN <= p.ncols);   // warp-uniform
      mbar_wait_a(sf0 + 8 * buf, sph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_x32(t_lane + buf * FWD2_BN + c, v);
        tmem_ld_wait();
        if (c == 32) {                              // last TMEM read of this tile: hand the buffer back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_a(se0_lead + 8 * buf);
        }
        float e[32];
        if (interior) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            e[i] = fast_exp2(fmaf(__uint_as_float(v[i]), p.k1, -p.k2));
            rsum += e[i];
          }
        } else {
          const bool row_ok = row < p.nrows;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float x = fast_exp2(fmaf(__uint_as_float(v[i]), p.k1, -p.k2));
            e[i] = (row_ok && (col_base + c + i) < p.ncols) ? x : 0.f;
            rsum += e[i];
          }
        }
        my_scratch[q * 64 + c + lane] = warp_colsum32(e, lane);
      }
      named_bar_sync(1 + h, 128);
      {
        const int col = col_base + tid_wg;
        const int rb128 = rb * 2 + static_cast<int>(cr);
        if (tid_wg < 64 && col < p.ncols && rb128 * 128 < p.nrows)
          p.c_part[static_cast<long long>(rb128) * p.ncols + col] =
              (my_scratch[tid_wg] + my_scratch[64 + tid_wg]) + (my_scratch[128 + tid_wg] + my_scratch[192 + tid_wg]);
      }
      named_bar_sync(1 + h, 128);
      if (++buf == 2) { buf = 0; sph ^= 1; }
      if (++ct == CT) { ct = 0; ++rb; }
    }
    if (cur_rb >= 0) flush_rsum(cur_rb);
  }

  tc_fence_before();
  cluster_sync_all();                          // the peer may still be arriving on / reading through this CTA's barriers
  if (warp == 2) tmem_dealloc_2cta(tmem_base, 512);
}

// ================================================================================================
// Pass 2: gradients
// ================================================================================================
constexpr int BWD_BN = 32;                  // columns per S tile
constexpr int BWD_TA = 2;                   // ring A depth in tiles (the 4 Y K-chunks outside this CTA's D-half)
constexpr int BWD_TB = 3;                   // ring B depth in tiles (the 4 Y K-chunks inside the D-half; live until dX MMA)
constexpr int BWD_SLOT_BYTES = BWD_BN * 128;            // one [32 x 64] bf16 chunk
constexpr int BWD_GROUP_BYTES = 4 * BWD_SLOT_BYTES;     // 4 chunks = half a Y tile
constexpr int BWD_G_BYTES = 128 * 128;      // [128 rows x 64 bf16]: even tiles use K cols 0-31, odd tiles 32-63
constexpr int BWD_G_TILE_BYTES = 128 * 64;  // version 4: one [128 rows x 32 bf16] G tile per slot (64B-swizzled, contiguous)
#ifndef B200CLIP_BWD4_NSB
#define B200CLIP_BWD4_NSB 1
#endif
constexpr int BWD4_NSB = B200CLIP_BWD4_NSB; // S buffers in TMEM (32 columns each): 1 leaves room for one more X chunk in TMEM
constexpr int BWD4_XS = BWD4_NSB;           // out-of-half X K-chunks kept in shared memory (SS MMAs); the other 4 - XS live in TMEM
constexpr int BWD4_TA = 4;                  // version 4: A groups (16 KB each) -- the smem X chunks moved to TMEM buy ring depth
constexpr int BWD4_TB = 5 + (4 - BWD4_XS) - 1;   // B groups
constexpr int BWD_THREADS = 384;            // warp 0 TMA, 1 S-MMA issuer, 2 TMEM alloc + dX-MMA issuer, 3 idle, 4-11 epilogue
constexpr int nce_bwd4_smem_bytes() {
  return BWD4_XS * X_CHUNK_BYTES + (BWD4_TA + BWD4_TB) * BWD_GROUP_BYTES + 2 * BWD_G_BYTES + 512 + 1024;
}


// debug instrumentation (build with B200CLIP_NCE_PROF=1 python build.py): cycles spent inside each wait, accumulated per
// role, and a per-tile timeline, for one cluster.  Compiled out of the product build.
#ifdef B200CLIP_NCE_PROF
#define NCE_PW(k, call)                      \
  do {                                       \
    if (prof_on) {                           \
      const long long t0_ = clock64();       \
      call;                                  \
      wacc[k] += clock64() - t0_;            \
    } else {                                 \
      call;                                  \
    }                                        \
  } while (0)
#define NCE_TS(n, k)                                                                             \
  do {                                                                                             \
    if (prof_on && lane == 0 && (n) >= 512 && (n) < 544) p.prof[128 + h * 256 + ((n)-512) * 8 + (k)] = clock64(); \
  } while (0)
#define NCE_PROF_BEGIN() long long wacc[4] = {0, 0, 0, 0}; const long long tstart_ = prof_on ? clock64() : 0
#define NCE_PROF_END(role)                                                              \
  do {                                                                                  \
    if (prof_on && lane == 0) {                                                         \
      long long* o_ = p.prof + ((h * 8 + (role)) * 8);                                  \
      o_[0] = clock64() - tstart_; o_[1] = wacc[0]; o_[2] = wacc[1]; o_[3] = wacc[2]; o_[4] = wacc[3]; \
    }                                                                                   \
  } while (0)
#else
#define NCE_PW(k, call) call
#define NCE_TS(n, k) do { } while (0)
#define NCE_PROF_BEGIN() do { } while (0)
#define NCE_PROF_END(role) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// Pass 2, CTA-pair kernel of round 1 (D = 512 only; kept behind B200CLIP_BWD_VARIANT=4 as the A/B baseline of
// nce_bwdc_kernel<2> in infonce_bwd.cuh, which is the same design generalised to D = 256 NC).  The two D-half CTAs of a row block form a cluster (1x2x1).
//  * Column tile n (32 columns) is OWNED by CTA (n & 1): only the owner recomputes S and forms G for it, so each logit is
//    exponentiated once per direction instead of once per D-half (executed tensor work per direction: S once + dX).
//  * The owner's epilogue warpgroup w = (own tile index & 1) owns G slot (owner, w) in BOTH CTAs: an 8 KB [128 x 32] bf16
//    tile, K-major with 64-byte rows and the 64B swizzle, so a warp's 32 rows are 2 KB contiguous.  Each warp writes its
//    rows with st.shared.v4, fence.proxy.async, arrives on the local g_full and sends the same 2 KB to the peer CTA with ONE
//    cp.async.bulk shared::cta -> shared::cluster that completes tx-bytes on the peer's g_full.  (The previous delivery,
//    2 x 4 st.async of 16 B per thread, cost one shared-memory port transaction per 16 B on BOTH CTAs: 1024 transactions
//    per tile pair, 4.08 ms; bulk delivery 3.35 ms.)
//  * g_full[slot]: own slots count the 4 producing warps; peer slots 1 arming arrival (the local dX issuer) + 8192 tx bytes.
//    g_empty[slot]: one multicast tcgen05.commit from each CTA's dX issuer (the slot may be refilled once BOTH CTAs have
//    read it).  One barrier per slot and one waiter per barrier: every waiter sees consecutive phases (a barrier shared by
//    two producers aliases parities).
//  * X lives in TMEM as the A operand of the S MMA (tcgen05.mma TS form) for 7 of its 8 K-chunks: the in-half range (128
//    columns) and the first 3 out-of-half chunks (96 columns); only the last chunk is read from shared memory.  Measured
//    (tools/mma_rate.cu): an SS MMA at N = 32 costs 40 clk (its 4 KB A tile at 128 B/clk of shared-memory bandwidth), a TS
//    MMA 17.8 clk (floor 16).  This leaves ONE 32-column S buffer: S(m+1) is issued once the epilogue has loaded S(m).
//  * Y traffic per CTA: in-half K-chunks of every tile (ring B, 7 tiles deep, also the MN-major B operand of the dX MMA) +
//    out-of-half chunks of its own tiles (ring A, 4 groups).
// TMEM: acc [0,256) | S NSB x 32 | X in-half K range (128 columns) | the first 4 - XS out-of-half X chunks (32 columns each).
// Tensor-pipe floor per tile pair (tools/mma_rate.cu, both issuers running): 1219 clk; the kernel runs ~1750 (B=32768):
// the rest is the S -> epilogue -> G -> dX dependency chain with two G slots per owner (wait-cycle breakdown:
// tools/nce_prof.py with a B200CLIP_NCE_PROF=1 build).
// Experiments that lost (kept in git history): 64-column tiles (rings too shallow), pairing own tiles / splitting the dX
// accumulator, polling test_wait instead of try_wait (no change), two S buffers with 2 X chunks in smem (no change).
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(1, 2, 1) __launch_bounds__(BWD_THREADS, 1)
nce_bwd4_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_y0,
                const __grid_constant__ CUtensorMap tmap_x1, const __grid_constant__ CUtensorMap tmap_y1,
                const NceBwdParams p) {
  const int dir = blockIdx.z;
  int h;                                            // D-half owned by this CTA == rank in the pair; read once (volatile asm:
  asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(h));   // the compiler otherwise re-reads the special register in the tile loop)
  const int nrows = dir ? p.nrows[1] : p.nrows[0];
  const int ncols = dir ? p.ncols[1] : p.ncols[0];
  // blockIdx.x = split * row_blocks + row block.  With few local rows (data parallel: b_loc << b_glob) direction 0 has few,
  // long row blocks; its columns are then cut into `nsplit` ranges handled by separate clusters (partial dX each).
  const int rbs = (nrows + 127) >> 7;
  const int split = static_cast<int>(blockIdx.x) / rbs;
  const int rb = static_cast<int>(blockIdx.x) - split * rbs;
  const int nsplit = dir ? p.nsplit[1] : p.nsplit[0];
  if (split >= nsplit) return;                      // uniform over the whole cluster (same blockIdx.x), before any barrier
  const CUtensorMap* tmap_x = dir == 0 ? &tmap_x0 : &tmap_x1;
  const CUtensorMap* tmap_y = dir == 0 ? &tmap_y0 : &tmap_y1;
  const uint32_t peer = static_cast<uint32_t>(h ^ 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;
  uint8_t* sA = sX + BWD4_XS * X_CHUNK_BYTES;       // sX holds only the LAST BWD4_XS chunks of the out-of-half K range of X
  uint8_t* sB = sA + BWD4_TA * BWD_GROUP_BYTES;
  uint8_t* sG = sB + BWD4_TB * BWD_GROUP_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 2 * BWD_G_BYTES);   // two G buffers: slot (owner, k) -> buffer k, K half owner
  uint64_t* x_full = bars;                          // 1
  uint64_t* a_full = bars + 1;                      // TA
  uint64_t* a_empty = a_full + BWD4_TA;              // TA
  uint64_t* b_full = a_empty + BWD4_TA;              // TB
  uint64_t* b_empty = b_full + BWD4_TB;              // TB
  uint64_t* s_full = b_empty + BWD4_TB;              // 2
  uint64_t* s_empty = s_full + 2;                   // 2
  uint64_t* g_full = s_empty + 2;                   // 4: slot = owner*2 + k ; 1 arming arrival + 8192 tx bytes (128 threads x 64 B)
  uint64_t* g_empty = g_full + 4;                   // 4: one multicast commit from each CTA of the pair
  uint64_t* acc_full = g_empty + 4;                 // 1
  uint64_t* xt_full = acc_full + 1;                 // 1: X in-half range stored to TMEM by warpgroup 0
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xt_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef B200CLIP_NCE_PROF
  const bool prof_on = p.prof != nullptr && dir == 0 && rb == 100;
#endif
  const int nt_all = (ncols + BWD_BN - 1) / BWD_BN; // all column tiles
  const int tile0 = static_cast<int>(static_cast<long long>(nt_all) * split / nsplit);           // this split's tile range
  const int nt = static_cast<int>(static_cast<long long>(nt_all) * (split + 1) / nsplit) - tile0;
  const int nown = (nt - h + 1) / 2;                // tiles owned by this CTA: n = 2m + h (n local to the split)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmap_x);
    tma_prefetch_desc(tmap_y);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < BWD4_TA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < BWD4_TB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 128);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&g_full[b], (b >> 1) == h ? 4 : 1);   // own slots: 4 epilogue warps arrive; peer slots: arming arrival + 8192 tx bytes
      mbar_init(&g_empty[b], 2);
    }
    mbar_init(acc_full, 1);
    mbar_init(xt_full, 256);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  cluster_sync_all();                               // barrier inits visible to the peer before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;
  const uint32_t tmem_s = tmem_base + 256;          // NSB x 32 columns
  const uint32_t tmem_x = tmem_s + BWD4_NSB * BWD_BN;   // 128 columns: X[:, h*256 .. +256) as packed bf16 pairs
  const uint32_t tmem_x2 = tmem_x + 128;            // 32 * (4 - XS) columns: the first 4 - XS out-of-half chunks of X

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t xf = smem_u32(x_full);
    const uint32_t af0 = smem_u32(a_full), ae0 = smem_u32(a_empty), bf0 = smem_u32(b_full), be0 = smem_u32(b_empty);
    const uint32_t sa = smem_u32(sA), sb = smem_u32(sB), sx = smem_u32(sX);
    if (elect_one()) {
      mbar_arrive_expect_tx_a(xf, BWD4_XS * X_CHUNK_BYTES);
#pragma unroll
      for (int kc = 0; kc < BWD4_XS; ++kc)
        tma_load_2d_a(sx + kc * X_CHUNK_BYTES, tmap_x, xf, (1 - h) * 256 + (4 - BWD4_XS + kc) * 64, rb * 128);
    }
    __syncwarp();
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    const int k_in = h * 256, k_out = (1 - h) * 256;
    NCE_PROF_BEGIN();
    for (int n = 0; n < nt; ++n) {
      NCE_PW(0, mbar_wait_a(be0 + 8 * ib, pb ^ 1));
      NCE_TS(n, 0);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(bf0 + 8 * ib, BWD_GROUP_BYTES);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_2d_a(sb + ib * BWD_GROUP_BYTES + c * BWD_SLOT_BYTES, tmap_y, bf0 + 8 * ib, k_in + c * 64, (tile0 + n) * BWD_BN);
      }
      __syncwarp();
      if (++ib == BWD4_TB) { ib = 0; pb ^= 1; }
      if ((n & 1) == h) {
        NCE_PW(1, mbar_wait_a(ae0 + 8 * ia, pa ^ 1));
        if (elect_one()) {
          mbar_arrive_expect_tx_a(af0 + 8 * ia, BWD_GROUP_BYTES);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tma_load_2d_a(sa + ia * BWD_GROUP_BYTES + c * BWD_SLOT_BYTES, tmap_y, af0 + 8 * ia, k_out + c * 64, (tile0 + n) * BWD_BN);
        }
        __syncwarp();
        if (++ia == BWD4_TA) { ia = 0; pa ^= 1; }
      }
    }
    NCE_PROF_END(0);
  } else if (warp == 1) {
    // ===================== S-MMA issuer (own tiles only) =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, BWD_BN, false, false);
    const uint32_t af0 = smem_u32(a_full), ae0 = smem_u32(a_empty), bf0 = smem_u32(b_full);
    const uint32_t sf0 = smem_u32(s_full), se0 = smem_u32(s_empty);
    const uint32_t x_out = desc_lo(smem_u32(sX), 16);
    const uint32_t a_lo0 = desc_lo(smem_u32(sA), 16), b_lo0 = desc_lo(smem_u32(sB), 16);
    mbar_wait_a(smem_u32(x_full), 0);
    mbar_wait_a(smem_u32(xt_full), 0);
    tc_fence_after();
    int ia = 0;
    uint32_t pa = 0;
    // ring B position of tile n: slot n % TB, phase (n / TB) & 1 -- tracked incrementally over ALL tiles
    int ib = 0;
    uint32_t pb = 0;
    if (h == 1) { ib = 1 % BWD4_TB; }                 // first own tile is n = 1
    NCE_PROF_BEGIN();
    for (int m = 0; m < nown; ++m) {
      const int buf = m & 1;                      // s_full / s_empty barrier pair (one per epilogue warpgroup)
      if (BWD4_NSB == 2) {
        NCE_PW(0, mbar_wait_a(se0 + 8 * buf, ((m >> 1) & 1) ^ 1));       // tile m-2 (same buffer) has been read
      } else if (m > 0) {
        NCE_PW(0, mbar_wait_a(se0 + 8 * (buf ^ 1), ((m - 1) >> 1) & 1));   // single buffer: tile m-1 has been read
      }
      NCE_PW(1, mbar_wait_a(bf0 + 8 * ib, pb));
      NCE_TS(2 * m + h, 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_s + (BWD4_NSB == 2 ? buf * BWD_BN : 0);
      if (elect_one()) {
        const uint32_t yb = b_lo0 + ib * (BWD_GROUP_BYTES >> 4);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ts_lo(d_tmem, tmem_x + c * 32 + j * 8, yb + c * (BWD_SLOT_BYTES >> 4) + 2 * j, idesc_s, (c | j) != 0);
      }
      __syncwarp();
      NCE_PW(2, mbar_wait_a(af0 + 8 * ia, pa));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ya = a_lo0 + ia * (BWD_GROUP_BYTES >> 4);
#pragma unroll
        for (int c = 0; c < 4 - BWD4_XS; ++c)       // out-of-half chunks whose X lives in TMEM
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ts_lo(d_tmem, tmem_x2 + c * 32 + j * 8, ya + c * (BWD_SLOT_BYTES >> 4) + 2 * j, idesc_s, true);
#pragma unroll
        for (int c = 4 - BWD4_XS; c < 4; ++c)       // the rest: X from shared memory
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ss_lo(d_tmem, x_out + (c - (4 - BWD4_XS)) * (X_CHUNK_BYTES >> 4) + 2 * j, ya + c * (BWD_SLOT_BYTES >> 4) + 2 * j, idesc_s, true);
        tc_commit_a(ae0 + 8 * ia);
        tc_commit_a(sf0 + 8 * buf);
      }
      __syncwarp();
      NCE_TS(2 * m + h, 2);
      if (++ia == BWD4_TA) { ia = 0; pa ^= 1; }
      // advance the ring-B cursor by two tiles
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (++ib == BWD4_TB) { ib = 0; pb ^= 1; }
    }
    NCE_PROF_END(1);
  } else if (warp == 2) {
    // ===================== dX-MMA issuer (every tile, this CTA's D-half) =====================
    constexpr uint32_t idesc_g = make_idesc_bf16(128, 256, false, true);
    const uint32_t gf0 = smem_u32(g_full), ge0 = smem_u32(g_empty), be0 = smem_u32(b_empty), bf0 = smem_u32(b_full);
    const uint32_t g_lo = desc_lo(smem_u32(sG), 16);
    const uint32_t y_lo0 = desc_lo(smem_u32(sB), BWD_SLOT_BYTES);
    int ib = 0;
    uint32_t pb = 0;
    NCE_PROF_BEGIN();
    for (int n = 0; n < nt; ++n) {
      const int par = n & 1, kbuf = (n >> 1) & 1;
      const int slot = par * 2 + kbuf;
      // peer-owned slot: arm for the four 2 KB bulk copies of the G tile; own slot: the epilogue warps arrive themselves
      if (par != h && elect_one()) mbar_arrive_expect_tx_a(gf0 + 8 * slot, BWD_G_TILE_BYTES);
      __syncwarp();
      NCE_PW(0, mbar_wait_a(bf0 + 8 * ib, pb));     // tiles owned by the peer were never waited on by the S issuer
      if (par == h) NCE_PW(1, mbar_wait_cluster_a(gf0 + 8 * slot, (n >> 2) & 1));
      else NCE_PW(2, mbar_wait_cluster_a(gf0 + 8 * slot, (n >> 2) & 1));
      NCE_TS(n, 5);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t yb = y_lo0 + ib * (BWD_GROUP_BYTES >> 4);
#pragma unroll
        for (int jj = 0; jj < BWD_BN / 16; ++jj)
          mma_ss_lo_ab(tmem_acc, g_lo + slot * (BWD_G_TILE_BYTES >> 4) + jj * 2, DESC_HI_SW64, yb + jj * (2048 >> 4), DESC_HI_SW128,
                       idesc_g, (n | jj) != 0);
        tc_commit_a(be0 + 8 * ib);
        tc_commit_multicast_a(ge0 + 8 * slot, 0x3);   // both CTAs' g_empty[slot]: the owner refills once BOTH have read it
      }
      __syncwarp();
      NCE_TS(n, 6);
      if (++ib == BWD4_TB) { ib = 0; pb ^= 1; }
    }
    NCE_PROF_END(2);
    if (elect_one()) tc_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups: own tile m -> warpgroup m & 1 (st.async G delivery) =====================
    const int w = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row_l = q * 32 + lane;
    const int row = rb * 128 + row_l;
    const bool row_ok = row < nrows;
    const float rstat = row_ok ? (dir ? p.row_stat[1] : p.row_stat[0])[row] : 0.f;
    const float* cstat = dir ? p.col_stat[1] : p.col_stat[0];
    const int diag_col = row + (dir ? p.diag_off[1] : p.diag_off[0]);
    const int warp_diag_lo = rb * 128 + q * 32 + (dir ? p.diag_off[1] : p.diag_off[0]);
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sf = smem_u32(s_full) + 8 * w, se = smem_u32(s_empty) + 8 * w;
    // this warpgroup owns G slot (h, w) in BOTH CTAs: a [128 rows x 32 bf16] K-major tile, 64 B rows, 64B swizzle (8 KB,
    // contiguous, so a warp's 32 rows are one 2 KB bulk copy)
    const uint32_t gf = smem_u32(g_full) + 8 * (2 * h + w), ge = smem_u32(g_empty) + 8 * (2 * h + w);
    const uint32_t gf_peer = mapa_u32(gf, peer);
    const uint32_t g_tile = smem_u32(sG) + (2 * h + w) * BWD_G_TILE_BYTES;
    const uint32_t g_row = g_tile + row_l * 64;
    const uint32_t g_warp = g_tile + q * 2048, g_warp_peer = mapa_u32(g_warp, peer);
    const uint32_t g_swz = static_cast<uint32_t>((row_l >> 1) & 3);
    {
      // warpgroup 0: X[row, h*256 .. +256) -> TMEM columns tmem_x .. +128 of this thread's lane (bf16 pairs, K ascending)
      // warpgroup 1: the first 4 - XS out-of-half chunks X[row, (1-h)*256 .. ) -> tmem_x2
      const int k0 = (w == 0) ? h * 256 : (1 - h) * 256;
      const int nch = (w == 0) ? 4 : 4 - BWD4_XS;
      const uint32_t dst = (w == 0) ? tmem_x : tmem_x2;
      const uint4* xsrc = reinterpret_cast<const uint4*>((dir ? p.xmat[1] : p.xmat[0]) + static_cast<long long>(row_ok ? row : 0) * NCE_D + k0);
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        uint32_t xr[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 t = row_ok ? __ldg(xsrc + c * 8 + i) : make_uint4(0u, 0u, 0u, 0u);
          xr[4 * i] = t.x; xr[4 * i + 1] = t.y; xr[4 * i + 2] = t.z; xr[4 * i + 3] = t.w;
        }
        tmem_st_x32(dst + lane_base + c * 32, xr);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(smem_u32(xt_full));
    }
    uint32_t ph = 0;
    NCE_PROF_BEGIN();
    for (int m = w; m < nown; m += 2) {
      const int n = 2 * m + h;
      const int col0 = (tile0 + n) * BWD_BN;
      const bool full_tile = col0 + BWD_BN <= ncols;
      const bool diag_tile = (col0 < warp_diag_lo + 32) && (col0 + BWD_BN > warp_diag_lo);
      float cs[32];
      if (full_tile) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(cstat + col0 + i));
          cs[i] = c4.x; cs[i + 1] = c4.y; cs[i + 2] = c4.z; cs[i + 3] = c4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) cs[i] = (col0 + i < ncols) ? cstat[col0 + i] : 0.f;
      }
      NCE_PW(0, mbar_wait_a(sf, ph));
      if (q == 0) NCE_TS(n, 3);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_x32(tmem_s + lane_base + (BWD4_NSB == 2 ? w * BWD_BN : 0), v);
      NCE_PW(2, tmem_ld_wait());
      tc_fence_before();
      mbar_arrive_a(se);
      uint32_t packed[16];
      if (full_tile && !diag_tile) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float g0 = fast_exp2(fmaf(__uint_as_float(v[i]), p.k1, -p.k2)) * (rstat + cs[i]);
          const float g1 = fast_exp2(fmaf(__uint_as_float(v[i + 1]), p.k1, -p.k2)) * (rstat + cs[i + 1]);
          packed[i / 2] = pack_bf16x2(g0, g1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float g[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int col = col0 + i + t;
            float gv = fast_exp2(fmaf(__uint_as_float(v[i + t]), p.k1, -p.k2)) * (rstat + cs[i + t]);
            if (col == diag_col) gv -= 1.0f;
            g[t] = (col < ncols) ? gv : 0.f;
          }
          packed[i / 2] = pack_bf16x2(g[0], g[1]);
        }
      }
      // slot (h, w) is reused every second own tile: wait until BOTH CTAs' dX MMAs of its previous use have read it
      NCE_PW(1, mbar_wait_cluster_a(ge, ((m >> 1) & 1) ^ 1));
#pragma unroll
      for (int c = 0; c < 4; ++c)
        st_shared_v4(g_row + ((static_cast<uint32_t>(c) ^ g_swz) << 4),
                     make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]));
      fence_proxy_async_smem();                       // generic stores -> visible to tcgen05.mma and the bulk copy (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(gf);                            // local dX issuer: 1 of 4 warps
        bulk_copy_s2s_cluster(g_warp_peer, g_warp, 2048, gf_peer);   // this warp's 32 rows to the peer CTA
      }
      if (q == 0) NCE_TS(n, 4);
      ph ^= 1;
    }
    if (q == 0) NCE_PROF_END(3 + w);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float scale = p.out_scale;
    if (p.grad_scale) scale *= *p.grad_scale;
    float* orow = (dir ? p.out[1] : p.out[0]) + split * (dir ? p.split_stride[1] : p.split_stride[0]) +
                  static_cast<long long>(row) * NCE_D + h * 256 + w * 128;
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem_acc + lane_base + w * 128 + c, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {           // 256-bit stores: full 32-byte sectors from a row-per-thread layout
          float f8[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) f8[t] = __uint_as_float(v[i + t]) * scale;
          st_global_f32x8(orow + c + i, f8);
        }
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();                               // the peer may still be writing into / arriving on this CTA's smem
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ================================================================================================
// small reductions around the two passes
// ================================================================================================
// r[i] = sum_slots r_part ; c[j] = sum_rb c_part.  Block = 64 columns x 4 row-block groups (c_part is 33 MB at B = 32768:
// one thread per column walking all 256 row blocks left the memory system idle); the 4 group sums are added in fixed order.
__global__ void __launch_bounds__(256) nce_reduce_stats_kernel(const float* __restrict__ r_part, int r_slots, int nrows_pad, int nrows,
                                                               const float* __restrict__ c_part, int row_blocks, int ncols,
                                                               float* __restrict__ r, float* __restrict__ c) {
  __shared__ float red[2][4][64];
  const int x = threadIdx.x & 63, y = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + x;
  float ra = 0.f, ca = 0.f;
  if (i < nrows)
    for (int s = y; s < r_slots; s += 4) ra += r_part[static_cast<long long>(s) * nrows_pad + i];
  if (i < ncols) {
    float c0 = 0.f, c1 = 0.f;                           // two independent chains (more loads in flight)
    int b = y;
    for (; b + 4 < row_blocks; b += 8) {
      c0 += c_part[static_cast<long long>(b) * ncols + i];
      c1 += c_part[static_cast<long long>(b + 4) * ncols + i];
    }
    if (b < row_blocks) c0 += c_part[static_cast<long long>(b) * ncols + i];
    ca = c0 + c1;
  }
  red[0][y][x] = ra;
  red[1][y][x] = ca;
  __syncthreads();
  if (y == 0) {
    if (i < nrows) r[i] = (red[0][0][x] + red[0][1][x]) + (red[0][2][x] + red[0][3][x]);
    if (i < ncols) c[i] = (red[1][0][x] + red[1][1][x]) + (red[1][2][x] + red[1][3][x]);
  }
}

// rinvh = 0.5 / r, cinvh = 0.5 / c: all the backward pass needs from the statistics.  Launched alone when the loss VALUE is
// computed off the critical path (b200clip_infonce_inv_stats; nce_loss_kernel then gets null rinvh / cinvh).
__global__ void __launch_bounds__(256) nce_inv_stats_kernel(const float* __restrict__ r, int b_loc, const float* __restrict__ c,
                                                            int b_glob, float* __restrict__ rinvh, float* __restrict__ cinvh) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < b_loc) rinvh[i] = 0.5f / r[i];
  if (i < b_glob) cinvh[i] = 0.5f / c[i];
}

// sums[0] = sum_i log r_i (local rows), sums[1] = sum_{j in [c_lo,c_hi)} log c_j, sums[2] = sum_i S_ii  (S_ii = I_i.T_{row0+i}/tau)
// also writes rinvh = 0.5 / r, cinvh = 0.5 / c.  Deterministic: per-block partials, last block folds them in order.
__global__ void __launch_bounds__(256) nce_loss_kernel(const __nv_bfloat16* __restrict__ I, const __nv_bfloat16* __restrict__ T,
                                                       int D, int b_loc, int b_glob, int row0, float inv_tau,
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
    if (rinvh) rinvh[i] = 0.5f / ri;
  }
  for (int j = gtid; j < b_glob; j += gsz) {
    const float cj = c[j];
    if (cinvh) cinvh[j] = 0.5f / cj;
    if (j >= c_lo && j < c_hi) a_logc += static_cast<double>(logf(cj));
  }
  // diagonal: one warp per row, D bf16 = D/8 16-byte pieces dealt round-robin to the lanes
  for (int i = blockIdx.x * 8 + warp; i < b_loc; i += gridDim.x * 8) {
    const uint4* pi = reinterpret_cast<const uint4*>(I + static_cast<long long>(i) * D);
    const uint4* pt = reinterpret_cast<const uint4*>(T + static_cast<long long>(row0 + i) * D);
    float d = 0.f;
    for (int k = lane; k < (D >> 3); k += 32) {
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
  if (is_last) {
    // the last block folds the per-block partials with all 256 threads (one thread walking ~1000 dependent L2 loads cost
    // more than the rest of the kernel); thread t takes blocks t, t+256, ..., then a fixed-order tree: deterministic
    __threadfence();
    __shared__ double fold[3][256];
    double s[3] = {0.0, 0.0, 0.0};
    for (unsigned b = threadIdx.x; b < gridDim.x; b += 256)
      for (int k = 0; k < 3; ++k) s[k] += partial[b * 3 + k];
    for (int k = 0; k < 3; ++k) fold[k][threadIdx.x] = s[k];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o)
        for (int k = 0; k < 3; ++k) fold[k][threadIdx.x] += fold[k][threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      sums[0] = fold[0][0]; sums[1] = fold[1][0]; sums[2] = fold[2][0];
      if (loss) *loss = static_cast<float>(static_cast<double>(shift_m) + (fold[0][0] + fold[1][0]) / (2.0 * b_glob) - fold[2][0] / b_glob);
      *counter = 0;
    }
  }
}

// Default: X in shared memory, 128 x 128 S tiles, 5-stage Y ring, 4 TMEM S buffers: 0.83 ms at B = 32768 (ncu: tensor pipe
// 67 % active; the 8 KB of SS operands per 64-clk MMA plus the TMA writes of Y oversubscribe the 128 B/clk smem port).
// The TMEM-resident-X variant (XT: TS MMAs, 12-stage ring, 2 S buffers shared by warpgroup pairs that split the columns)
// measured 0.80 ms stand-alone and 0.894 vs 0.938 ms inside the step (step 5.39 vs 5.42 ms; all 109 GPU tests pass with it) --
// B operand reads + TMA writes still fill the port (64 + 64 B/clk); a 2-CTA pair (cta_group::2) halving both is the real fix --
// so it stays optional (-DB200CLIP_FWD_XT=1).  A 64-column XT variant measured 1.42 ms.
#ifndef B200CLIP_FWD_XT
#define B200CLIP_FWD_XT 0
#endif
constexpr bool FWD_XT = B200CLIP_FWD_XT != 0;
constexpr int FWD_BN = 128;
constexpr int FWD_STAGES = FWD_XT ? 12 : 5;   // 16 KB Y chunks
constexpr int FWD_NWG = 4;                 // epilogue warpgroups = TMEM S buffers

struct NceFwdPlan {
  int row_blocks, col_tiles, total_tiles, tiles_per_cta, grid, r_slots, nrows_pad;
  size_t r_part_bytes, c_part_bytes, total_bytes;
  // CTA-pair kernel: 256 x 256 super-tiles walked by clusters of two CTAs
  int rb2, ct2, total2, tiles_per_cluster, grid2, r_slots2, nrows_pad2;
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
  pl.r_slots = FWD_NWG * ((pl.col_tiles + pl.tiles_per_cta - 1) / pl.tiles_per_cta + 1);
  pl.nrows_pad = pl.row_blocks * 128;
  pl.rb2 = static_cast<int>((b_loc + 255) / 256);
  pl.ct2 = static_cast<int>((b_glob + FWD2_BN - 1) / FWD2_BN);
  pl.total2 = pl.rb2 * pl.ct2;
  const int clusters = std::max(1, sms / 2);
  pl.tiles_per_cluster = std::max(1, (pl.total2 + clusters - 1) / clusters);
  pl.grid2 = 2 * ((pl.total2 + pl.tiles_per_cluster - 1) / pl.tiles_per_cluster);
  pl.r_slots2 = FWD2_EPI_WG * ((pl.ct2 + pl.tiles_per_cluster - 1) / pl.tiles_per_cluster + 1);
  pl.nrows_pad2 = pl.rb2 * 256;
  // one workspace serves either kernel: size it for the larger partial-sum buffer
  pl.r_part_bytes = std::max(static_cast<size_t>(pl.r_slots) * pl.nrows_pad, static_cast<size_t>(pl.r_slots2) * pl.nrows_pad2) * sizeof(float);
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

static bool nce_dim_supported(int D) { return D >= 64 && D <= 1024 && D % 64 == 0; }

extern "C" int b200clip_infonce_fwd_stats(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob,
                                          float temperature, float* r, float* c_partial, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  B200_REQUIRE(nce_dim_supported(D), "infonce: D=%d unsupported (need a multiple of 64, 64..1024)", D);
  B200_REQUIRE(b_loc > 0 && b_glob > 0 && b_loc <= b_glob, "infonce: need 0 < b_loc <= b_glob (got %lld, %lld)", b_loc, b_glob);
  B200_REQUIRE(temperature >= NCE_MIN_TAU, "infonce: temperature %g is below %g: the fixed-shift exponentials would underflow "
               "(CLIP clamps tau at 0.01)", temperature, NCE_MIN_TAU);
  B200_REQUIRE(b_glob < (1ll << 30), "infonce: batch too large");
  const NceFwdPlan pl = plan_fwd(b_loc, b_glob);
  if (workspace_bytes < pl.total_bytes) return fail(B200_ERR_WORKSPACE, "infonce_fwd: workspace %zu < %zu", workspace_bytes, pl.total_bytes);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* r_part = reinterpret_cast<float*>(ws);
  float* c_part = reinterpret_cast<float*>(ws + ((pl.r_part_bytes + 255) & ~size_t(255)));
  B200_CHECK_CUDA(cudaMemsetAsync(r_part, 0, pl.r_part_bytes, s));
  // B200CLIP_FWD_VARIANT=1: the single-CTA kernel (D = 512 only; kept for A/B measurements and as a cross-check in the tests)
  static const int variant = [] { const char* e = getenv("B200CLIP_FWD_VARIANT"); return (e && atoi(e) == 1) ? 1 : 2; }();
  CUtensorMap tx, ty;
  int rc;
  int r_slots, nrows_pad;
  if (variant == 1 && D == NCE_D) {
    if ((rc = make_tmap_bf16_2d(&tx, i_hat, b_loc, D, D, 64, 128))) return rc;
    if ((rc = make_tmap_bf16_2d(&ty, t_hat, b_glob, D, D, 64, FWD_BN))) return rc;
    NceFwdParams p{};
    p.nrows = (int)b_loc; p.ncols = (int)b_glob; p.num_col_tiles = pl.col_tiles; p.total_tiles = pl.total_tiles;
    p.tiles_per_cta = pl.tiles_per_cta; p.r_slots = pl.r_slots; p.k1 = LOG2E / temperature; p.k2 = nce_k2(temperature);
    p.r_part = r_part; p.c_part = c_part; p.nrows_pad = pl.nrows_pad; p.x = static_cast<const __nv_bfloat16*>(i_hat);
    auto kern = nce_fwd_kernel<FWD_BN, FWD_STAGES, FWD_NWG, FWD_XT>;
    constexpr int smem = nce_fwd_smem_bytes<FWD_BN, FWD_STAGES, FWD_NWG, FWD_XT>();
    static SmemAttrOnce attr;
    B200_CHECK_CUDA(attr.ensure(kern, smem));
    kern<<<pl.grid, 128 + FWD_NWG * 128, smem, s>>>(tx, ty, p);
    B200_LAUNCH_CHECK();
    r_slots = pl.r_slots; nrows_pad = pl.nrows_pad;
  } else {
    if ((rc = make_tmap_bf16_2d(&tx, i_hat, b_loc, D, D, 64, 128))) return rc;     // a CTA's own 128 rows of the 256-row super-block
    if ((rc = make_tmap_bf16_2d(&ty, t_hat, b_glob, D, D, 64, 128))) return rc;    // a CTA's half of the 256-column tile
    NceFwd2Params p{};
    p.nrows = (int)b_loc; p.ncols = (int)b_glob; p.ct2 = pl.ct2; p.total = pl.total2; p.tiles_per_cluster = pl.tiles_per_cluster;
    p.k1 = LOG2E / temperature; p.k2 = nce_k2(temperature);
    p.r_part = r_part; p.c_part = c_part; p.nrows_pad = pl.nrows_pad2; p.kc = D / 64;
    if (D == NCE_D) {
      constexpr int smem = nce_fwd2_smem_bytes<NCE_KC, true>();
      static SmemAttrOnce attr;
      B200_CHECK_CUDA(attr.ensure(nce_fwd2_kernel<NCE_KC, true>, smem));
      nce_fwd2_kernel<NCE_KC, true><<<pl.grid2, FWD2_THREADS, smem, s>>>(tx, ty, p);
    } else {
      constexpr int smem = nce_fwd2_smem_bytes<0, false>();
      static SmemAttrOnce attr;
      B200_CHECK_CUDA(attr.ensure(nce_fwd2_kernel<0, false>, smem));
      nce_fwd2_kernel<0, false><<<pl.grid2, FWD2_THREADS, smem, s>>>(tx, ty, p);
    }
    B200_LAUNCH_CHECK();
    r_slots = pl.r_slots2; nrows_pad = pl.nrows_pad2;
  }
  const int n = (int)std::max(b_loc, b_glob);
  nce_reduce_stats_kernel<<<(n + 63) / 64, 256, 0, s>>>(r_part, r_slots, nrows_pad, (int)b_loc, c_part, pl.row_blocks,
                                                        (int)b_glob, r, c_partial);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_infonce_loss(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob,
                                     long long row0, float temperature, const float* r, const float* c, long long c_lo,
                                     long long c_hi, float* rinvh, float* cinvh, double* sums, float* loss, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  B200_REQUIRE(nce_dim_supported(D), "infonce: D=%d unsupported (need a multiple of 64, 64..1024)", D);
  B200_REQUIRE(b_loc > 0 && b_glob > 0 && row0 >= 0 && row0 + b_loc <= b_glob, "infonce_loss: bad row range");
  const NceFwdPlan pl = plan_fwd(b_loc, b_glob);
  if (workspace_bytes < pl.total_bytes) return fail(B200_ERR_WORKSPACE, "infonce_loss: workspace too small");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint8_t* tail = ws + ((pl.r_part_bytes + 255) & ~size_t(255)) + ((pl.c_part_bytes + 255) & ~size_t(255));
  double* partial = reinterpret_cast<double*>(tail);
  unsigned int* counter = reinterpret_cast<unsigned int*>(tail + 4096 * 3 * sizeof(double));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  B200_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), s));
  int grid = (int)std::min<long long>((b_glob + 63) / 64, 1024);      // 8 rows of the diagonal per warp at B = 32768
  nce_loss_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(i_hat), static_cast<const __nv_bfloat16*>(t_hat),
                                       D, (int)b_loc, (int)b_glob, (int)row0, 1.0f / temperature, r, c, (int)c_lo, (int)c_hi,
                                       rinvh, cinvh, partial, counter, sums, loss, static_cast<float>(nce_shift(temperature)));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_infonce_inv_stats(const float* r, long long b_loc, const float* c, long long b_glob, float* rinvh,
                                          float* cinvh, void* stream) {
  B200_REQUIRE(b_loc > 0 && b_glob >= b_loc && r && c && rinvh && cinvh, "infonce_inv_stats: bad arguments");
  nce_inv_stats_kernel<<<static_cast<unsigned>((b_glob + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      r, (int)b_loc, c, (int)b_glob, rinvh, cinvh);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static long long* g_nce_prof = nullptr;
// debug only: device buffer of 128 int64 receiving the wait-cycle counters of one cluster of the next backward launches
extern "C" void b200clip_debug_set_nce_prof(void* buf) { g_nce_prof = static_cast<long long*>(buf); }

// Column splits of direction 0 (dI): with b_loc << b_glob its few row blocks are 1024-tile-long CTAs that outlive the rest of
// the grid; each split writes its own partial d_i (summed by the consumer, b200clip_l2norm_bwd's dy_partials).
extern "C" int b200clip_infonce_bwd_splits(long long b_loc, long long b_glob) {
  if (b_loc <= 0 || b_glob <= 0) return 1;
  // equal-length CTAs in both directions: direction 1's row blocks sweep b_loc columns, direction 0's b_glob / splits, and the
  // two launches share the SMs (7 waves of equal CTAs lose ~1 % to the tail; 2:1 mixes lose 10-15 %).  B200CLIP_BWD_SPLITS
  // overrides (A/B measurements).
  static const int forced = [] { const char* e = getenv("B200CLIP_BWD_SPLITS"); return e ? atoi(e) : 0; }();
  if (forced >= 1 && forced <= 8) return b_loc < b_glob ? forced : 1;
  // ... but never more than 4: every split writes (and the consumer re-reads) its own partial dI, and at 8 ranks 4 splits
  // measured 1057 us per step against 1099 us with 8 (256-tile CTAs amortise the per-CTA prologue / accumulator write-back).
  const long long s = (b_glob + b_loc / 2) / b_loc;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(4, s)));
}

template <int NC>
static int launch_bwdc(const CUtensorMap& tx0, const CUtensorMap& ty0, const CUtensorMap& tx1, const CUtensorMap& ty1,
                       const NceBwdParams& p, dim3 grid, cudaStream_t s) {
  static SmemAttrOnce attr;
  B200_CHECK_CUDA(attr.ensure(nce_bwdc_kernel<NC>, BcCfg<NC>::SMEM));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(BC_THREADS);
  cfg.dynamicSmemBytes = BcCfg<NC>::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = NC; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  B200_CHECK_CUDA(cudaLaunchKernelEx(&cfg, nce_bwdc_kernel<NC>, tx0, ty0, tx1, ty1, p));
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// directions: bit 0 = dI (direction 0), bit 1 = dT partial (direction 1).  The data-parallel step launches direction 1 first
// so that its reduce-scatter overlaps direction 0; a single-GPU step launches both at once (3).
extern "C" int b200clip_infonce_bwd(const void* i_hat, const void* t_hat, int D, long long b_loc, long long b_glob,
                                    long long row0, float temperature, const float* rinvh, const float* cinvh,
                                    const float* grad_scale, float* d_i, int d_i_splits, float* d_t_partial, int directions,
                                    void* stream) {
  B200_REQUIRE(D == 512 || D == 768, "infonce_bwd: D=%d unsupported (the backward kernel is built for D = 512 and 768)", D);
  B200_REQUIRE(b_loc > 0 && b_glob > 0 && row0 >= 0 && row0 + b_loc <= b_glob, "infonce_bwd: bad row range");
  B200_REQUIRE(directions >= 1 && directions <= 3, "infonce_bwd: directions must be 1 (dI), 2 (dT) or 3 (both)");
  B200_REQUIRE((!(directions & 1) || d_i) && (!(directions & 2) || d_t_partial), "infonce_bwd: missing output for a requested direction");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(d_i) & 31u) == 0 && (reinterpret_cast<uintptr_t>(d_t_partial) & 31u) == 0 &&
               aligned16(rinvh) && aligned16(cinvh), "infonce_bwd: d_i / d_t_partial must be 32-byte aligned, statistics 16-byte");
  B200_REQUIRE(temperature >= NCE_MIN_TAU, "infonce: temperature %g is below %g", temperature, NCE_MIN_TAU);
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
  p.xmat[0] = static_cast<const __nv_bfloat16*>(i_hat); p.xmat[1] = static_cast<const __nv_bfloat16*>(t_hat);
  p.row_stat[1] = cinvh; p.col_stat[1] = rinvh; p.out[1] = d_t_partial;
  p.k1 = LOG2E / temperature; p.k2 = nce_k2(temperature);
  p.out_scale = 1.0f / (static_cast<float>(b_glob) * temperature);
  p.grad_scale = grad_scale;
  B200_REQUIRE(d_i_splits >= 1 && d_i_splits <= 8, "infonce_bwd: d_i_splits=%d out of range", d_i_splits);
  p.nsplit[0] = d_i_splits; p.split_stride[0] = b_loc * static_cast<long long>(D);
  p.nsplit[1] = 1; p.split_stride[1] = 0;
  p.prof = g_nce_prof;
  p.dir_base = directions == 2 ? 1 : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long gx0 = ((b_loc + 127) / 128) * d_i_splits, gx1 = (b_glob + 127) / 128;
  const long long gx = directions == 3 ? std::max(gx0, gx1) : (directions == 1 ? gx0 : gx1);
  const unsigned gz = directions == 3 ? 2u : 1u;
  // B200CLIP_BWD_VARIANT=4: round 1's pair kernel (D = 512, both directions in one launch) for A/B measurements
  static const int variant = [] { const char* e = getenv("B200CLIP_BWD_VARIANT"); return (e && atoi(e) == 4) ? 4 : 5; }();
  if (variant == 4 && D == NCE_D && directions == 3) {
    static SmemAttrOnce attr4;
    B200_CHECK_CUDA(attr4.ensure(nce_bwd4_kernel, nce_bwd4_smem_bytes()));
    nce_bwd4_kernel<<<dim3(static_cast<unsigned>(gx), 2, 2), BWD_THREADS, nce_bwd4_smem_bytes(), s>>>(tx0, ty0, tx1, ty1, p);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  if (D == 512) return launch_bwdc<2>(tx0, ty0, tx1, ty1, p, dim3(static_cast<unsigned>(gx), 2, gz), s);
  return launch_bwdc<3>(tx0, ty0, tx1, ty1, p, dim3(static_cast<unsigned>(gx), 3, gz), s);
}

// the shift m (natural-log units) that b200clip_infonce_loss's sums are relative to: loss = m + (sums[0] + sums[1]) / (2B) - sums[2] / B
extern "C" double b200clip_infonce_shift(float temperature) { return temperature > 0.f ? nce_shift(temperature) : 0.0; }
