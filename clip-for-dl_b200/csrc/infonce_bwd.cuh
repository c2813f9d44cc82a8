// b200clip: pass 2 of the symmetric InfoNCE (gradients), cluster kernel generalised over the embedding width.
//
// D = NC * 256: the NC CTAs of a cluster (1 x NC x 1) own the NC 256-wide D-slices of one 128-row block of dX (a 128 x 256 fp32
// accumulator = 256 TMEM columns each).  NC = 2 is D = 512 (0426/config.py:30), NC = 3 is D = 768 (BASELINE.json configs[4]).
//  * Column tile n (32 columns) is OWNED by CTA n % NC: only the owner recomputes S = X Y_tile^T (tcgen05.mma, N = 32, full K = D)
//    and forms G (bf16); every CTA then runs acc += G . Y_tile[:, slice] (M128 N256 K32, the Y tile reused as an MN-major operand
//    from the same swizzled smem bytes).  Each logit is therefore recomputed once per direction, not once per D-slice.
//  * G delivery: the owner's epilogue warpgroup w = (own tile index & 1) owns G slot (owner, w) in EVERY CTA of the cluster: an
//    8 KB [128 x 32] bf16 tile, K-major with 64-byte rows and the 64B swizzle, so a warp's 32 rows are 2 KB contiguous.  Each warp
//    writes its rows with st.shared.v4, fence.proxy.async, arrives on the local g_full and sends the same 2 KB to every peer with
//    ONE cp.async.bulk shared::cta -> shared::cluster each, completing tx-bytes on that peer's g_full.
//  * g_full[slot]: own slots count the 4 producing warps; peer slots 1 arming arrival (the local dX issuer) + 8192 tx bytes.
//    g_empty[slot]: one multicast tcgen05.commit from every CTA's dX issuer (the slot is refilled once ALL CTAs have read it).
//  * X lives in TMEM as the A operand of the S MMA (TS form) for 7 of its 4*NC K-chunks: the in-slice range (4 chunks, 128
//    columns) and the first 3 out-of-slice chunks (96 columns); the remaining 4*NC - 7 chunks (1 for NC = 2, 5 for NC = 3) are
//    SS operands from shared memory.  Measured (tools/mma_rate.cu): an SS MMA at N = 32 costs 40 clk (its 4 KB A tile at 128 B/clk
//    of shared-memory bandwidth), a TS MMA 17.8 clk (floor 16).
//  * Y traffic per CTA: in-slice K-chunks of every tile (ring B, also the MN-major B operand of the dX MMA) + out-of-slice chunks
//    of its own tiles (ring A).
// TMEM: acc [0,256) | S 32 | X in-slice 128 | X out-of-slice 96  = 512 columns.
// Direction 0: X = I (local rows), Y = T -> dI;  direction 1: X = T, Y = I -> dT (G is symmetric under r <-> c).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int BC_BN = 32;                         // columns per S tile
constexpr int BC_SLOT = BC_BN * 128;              // one [32 x 64] bf16 chunk of a Y tile (4 KB)
constexpr int BC_BGROUP = 4 * BC_SLOT;            // ring B group: the 4 in-slice chunks of one tile (16 KB)
constexpr int BC_XCHUNK = 128 * 128;              // [128 rows x 64 bf16] chunk of X (16 KB)
constexpr int BC_GTILE = 128 * 64;                // one [128 x 32] bf16 G tile (8 KB, 64B-swizzled)
constexpr int BC_XT_OUT = 3;                      // out-of-slice X chunks kept in TMEM
constexpr int BC_THREADS = 384;                   // warp 0 TMA, 1 S-MMA issuer, 2 TMEM alloc + dX-MMA issuer, 3 idle, 4-11 epilogue

template <int NC> struct BcCfg {
  static constexpr int KC = 4 * NC;                       // K chunks of 64
  static constexpr int OUTC = KC - 4;                     // out-of-slice chunks
  static constexpr int XS = OUTC - BC_XT_OUT;             // out-of-slice X chunks in shared memory (SS operands)
  static constexpr int AGROUP = OUTC * BC_SLOT;           // ring A group: the out-of-slice chunks of one OWN tile
  static constexpr int TA = NC == 2 ? 4 : 1;              // ring A depth (groups)
  static constexpr int TB = NC == 2 ? 7 : 4;              // ring B depth (tiles)
  static constexpr int GSLOTS = 2 * NC;
  static constexpr int SMEM = XS * BC_XCHUNK + TA * AGROUP + TB * BC_BGROUP + GSLOTS * BC_GTILE + 512 + 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(1 + 2 * TA + 2 * TB + 4 + 2 * GSLOTS + 2 <= 60, "barrier area is 512 bytes");
};

struct NceBwdParams {
  int nrows[2];             // valid X rows per direction
  int ncols[2];             // valid Y rows per direction
  int diag_off[2];          // diagonal: y column == x row + diag_off   (dir0: +row0, dir1: -row0)
  const float* row_stat[2]; // 0.5 / r or c for X rows
  const float* col_stat[2]; // 0.5 / c or r for Y rows
  float* out[2];            // dX [nrows, D] f32
  const __nv_bfloat16* xmat[2];   // X matrices (for the TMEM-resident K range)
  float k1, k2;
  float out_scale;          // 1 / (B_glob * tau)
  const float* grad_scale;  // optional device scalar multiplied into out_scale (upstream dLoss)
  long long* prof;          // unused by this kernel (debug hook of the previous version)
  int nsplit[2];            // column splits per direction (each split is its own cluster, writes its own partial dX)
  long long split_stride[2];   // elements between the partial outputs of consecutive splits
  int dir_base;             // direction of blockIdx.z == 0 (launching ONE direction: gridDim.z = 1, dir_base = that direction)
};

template <int NC>
__global__ void __launch_bounds__(BC_THREADS, 1)
nce_bwdc_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_y0,
                const __grid_constant__ CUtensorMap tmap_x1, const __grid_constant__ CUtensorMap tmap_y1, const NceBwdParams p) {
  using C = BcCfg<NC>;
  constexpr int D = NC * 256;
  const int dir = static_cast<int>(blockIdx.z) + p.dir_base;
  int h;                                            // D-slice owned by this CTA == rank in the cluster; read once (volatile asm:
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

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX = smem;                               // the LAST XS out-of-slice chunks of X
  uint8_t* sA = sX + C::XS * BC_XCHUNK;
  uint8_t* sB = sA + C::TA * C::AGROUP;
  uint8_t* sG = sB + C::TB * BC_BGROUP;             // G slot (owner, k) at index owner * 2 + k
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + C::GSLOTS * BC_GTILE);
  uint64_t* x_full = bars;                          // 1
  uint64_t* a_full = bars + 1;                      // TA
  uint64_t* a_empty = a_full + C::TA;               // TA
  uint64_t* b_full = a_empty + C::TA;               // TB
  uint64_t* b_empty = b_full + C::TB;               // TB
  uint64_t* s_full = b_empty + C::TB;               // 2 (one per epilogue warpgroup)
  uint64_t* s_empty = s_full + 2;                   // 2
  uint64_t* g_full = s_empty + 2;                   // GSLOTS
  uint64_t* g_empty = g_full + C::GSLOTS;           // GSLOTS: one multicast commit from every CTA of the cluster
  uint64_t* acc_full = g_empty + C::GSLOTS;         // 1
  uint64_t* xt_full = acc_full + 1;                 // 1: X stored to TMEM by the two epilogue warpgroups
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xt_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt_all = (ncols + BC_BN - 1) / BC_BN;   // all column tiles
  const int tile0 = static_cast<int>(static_cast<long long>(nt_all) * split / nsplit);           // this split's tile range
  const int nt = static_cast<int>(static_cast<long long>(nt_all) * (split + 1) / nsplit) - tile0;
  const int nown = (nt - h + NC - 1) / NC;          // tiles owned by this CTA: n = NC * m + h (n local to the split)
  // K-chunk index (of 64 elements) of out-of-slice chunk o (cyclic order after this CTA's slice)
  auto out_chunk = [&](int o) { return ((h + 1) * 4 + o) % C::KC; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmap_x);
    tma_prefetch_desc(tmap_y);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(x_full, 1);
    for (int s = 0; s < C::TA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < C::TB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_empty[b], 128);
    }
    for (int b = 0; b < C::GSLOTS; ++b) {
      mbar_init(&g_full[b], (b >> 1) == h ? 4 : 1);   // own slots: 4 epilogue warps arrive; peer slots: arming arrival + 8192 tx bytes
      mbar_init(&g_empty[b], NC);
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
  cluster_sync_all();                               // barrier inits visible to the peers before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base;
  const uint32_t tmem_s = tmem_base + 256;          // 32 columns
  const uint32_t tmem_x = tmem_s + BC_BN;           // 128 columns: X[:, h*256 .. +256) as packed bf16 pairs
  const uint32_t tmem_x2 = tmem_x + 128;            // 96 columns: out-of-slice chunks 0..2 of X

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t xf = smem_u32(x_full);
    const uint32_t af0 = smem_u32(a_full), ae0 = smem_u32(a_empty), bf0 = smem_u32(b_full), be0 = smem_u32(b_empty);
    const uint32_t sa = smem_u32(sA), sb = smem_u32(sB), sx = smem_u32(sX);
    if (elect_one()) {
      mbar_arrive_expect_tx_a(xf, C::XS * BC_XCHUNK);
#pragma unroll
      for (int kc = 0; kc < C::XS; ++kc)
        tma_load_2d_a(sx + kc * BC_XCHUNK, tmap_x, xf, out_chunk(BC_XT_OUT + kc) * 64, rb * 128);
    }
    __syncwarp();
    int ia = 0, ib = 0, own = h;                    // own: tiles until the next own tile (n % NC == h)
    uint32_t pa = 0, pb = 0;
    const int k_in = h * 256;
    for (int n = 0; n < nt; ++n) {
      mbar_wait_a(be0 + 8 * ib, pb ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx_a(bf0 + 8 * ib, BC_BGROUP);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tma_load_2d_a(sb + ib * BC_BGROUP + c * BC_SLOT, tmap_y, bf0 + 8 * ib, k_in + c * 64, (tile0 + n) * BC_BN);
      }
      __syncwarp();
      if (++ib == C::TB) { ib = 0; pb ^= 1; }
      if (own == 0) {
        mbar_wait_a(ae0 + 8 * ia, pa ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx_a(af0 + 8 * ia, C::AGROUP);
#pragma unroll
          for (int o = 0; o < C::OUTC; ++o)
            tma_load_2d_a(sa + ia * C::AGROUP + o * BC_SLOT, tmap_y, af0 + 8 * ia, out_chunk(o) * 64, (tile0 + n) * BC_BN);
        }
        __syncwarp();
        if (++ia == C::TA) { ia = 0; pa ^= 1; }
        own = NC;
      }
      --own;
    }
  } else if (warp == 1) {
    // ===================== S-MMA issuer (own tiles only) =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, BC_BN, false, false);
    const uint32_t af0 = smem_u32(a_full), ae0 = smem_u32(a_empty), bf0 = smem_u32(b_full);
    const uint32_t sf0 = smem_u32(s_full), se0 = smem_u32(s_empty);
    const uint32_t x_out = desc_lo(smem_u32(sX), 16);
    const uint32_t a_lo0 = desc_lo(smem_u32(sA), 16), b_lo0 = desc_lo(smem_u32(sB), 16);
    mbar_wait_a(smem_u32(x_full), 0);
    mbar_wait_a(smem_u32(xt_full), 0);
    tc_fence_after();
    int ia = 0;
    uint32_t pa = 0;
    int ib = h % C::TB;                              // ring B position of tile n = NC * m + h: slot n % TB, phase (n / TB) & 1
    uint32_t pb = 0;
    for (int m = 0; m < nown; ++m) {
      const int buf = m & 1;                         // s_full / s_empty barrier pair (one per epilogue warpgroup)
      if (m > 0) mbar_wait_a(se0 + 8 * (buf ^ 1), ((m - 1) >> 1) & 1);   // single S buffer: tile m-1 has been read
      mbar_wait_a(bf0 + 8 * ib, pb);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t yb = b_lo0 + ib * (BC_BGROUP >> 4);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ts_lo(tmem_s, tmem_x + c * 32 + j * 8, yb + c * (BC_SLOT >> 4) + 2 * j, idesc_s, (c | j) != 0);
      }
      __syncwarp();
      mbar_wait_a(af0 + 8 * ia, pa);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ya = a_lo0 + ia * (C::AGROUP >> 4);
#pragma unroll
        for (int o = 0; o < BC_XT_OUT; ++o)          // out-of-slice chunks whose X lives in TMEM
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ts_lo(tmem_s, tmem_x2 + o * 32 + j * 8, ya + o * (BC_SLOT >> 4) + 2 * j, idesc_s, true);
#pragma unroll
        for (int o = BC_XT_OUT; o < C::OUTC; ++o)    // the rest: X from shared memory
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_ss_lo(tmem_s, x_out + (o - BC_XT_OUT) * (BC_XCHUNK >> 4) + 2 * j, ya + o * (BC_SLOT >> 4) + 2 * j, idesc_s, true);
        tc_commit_a(ae0 + 8 * ia);
        tc_commit_a(sf0 + 8 * buf);
      }
      __syncwarp();
      if (++ia == C::TA) { ia = 0; pa ^= 1; }
#pragma unroll
      for (int t = 0; t < NC; ++t)                   // advance the ring-B cursor by NC tiles
        if (++ib == C::TB) { ib = 0; pb ^= 1; }
    }
  } else if (warp == 2) {
    // ===================== dX-MMA issuer (every tile, this CTA's D-slice) =====================
    constexpr uint32_t idesc_g = make_idesc_bf16(128, 256, false, true);
    constexpr uint16_t all_ctas = static_cast<uint16_t>((1u << NC) - 1u);
    const uint32_t gf0 = smem_u32(g_full), ge0 = smem_u32(g_empty), be0 = smem_u32(b_empty), bf0 = smem_u32(b_full);
    const uint32_t g_lo = desc_lo(smem_u32(sG), 16);
    const uint32_t y_lo0 = desc_lo(smem_u32(sB), BC_SLOT);      // LBO = stride between 64-wide D groups
    int ib = 0, owner = 0, use = 0;                  // owner = n % NC; use = n / NC (the owner's own-tile index)
    uint32_t pb = 0;
    for (int n = 0; n < nt; ++n) {
      const int slot = owner * 2 + (use & 1);
      const uint32_t gpar = static_cast<uint32_t>(use >> 1) & 1u;   // each slot is used by every second own tile of its owner
      // peer-owned slot: arm for the four 2 KB bulk copies of the G tile; own slot: the epilogue warps arrive themselves
      if (owner != h && elect_one()) mbar_arrive_expect_tx_a(gf0 + 8 * slot, BC_GTILE);
      __syncwarp();
      mbar_wait_a(bf0 + 8 * ib, pb);                 // tiles owned by a peer were never waited on by the S issuer
      mbar_wait_cluster_a(gf0 + 8 * slot, gpar);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t yb = y_lo0 + ib * (BC_BGROUP >> 4);
#pragma unroll
        for (int jj = 0; jj < BC_BN / 16; ++jj)
          mma_ss_lo_ab(tmem_acc, g_lo + slot * (BC_GTILE >> 4) + jj * 2, DESC_HI_SW64, yb + jj * (2048 >> 4), DESC_HI_SW128, idesc_g,
                       (n | jj) != 0);
        tc_commit_a(be0 + 8 * ib);
        tc_commit_multicast_a(ge0 + 8 * slot, all_ctas);   // every CTA's g_empty[slot]: the owner refills once ALL have read it
      }
      __syncwarp();
      if (++ib == C::TB) { ib = 0; pb ^= 1; }
      if (++owner == NC) { owner = 0; ++use; }
    }
    if (elect_one()) tc_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue warpgroups: own tile m -> warpgroup m & 1 =====================
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
    // this warpgroup owns G slot (h, w) in EVERY CTA: a [128 rows x 32 bf16] K-major tile, 64 B rows, 64B swizzle (8 KB,
    // contiguous, so a warp's 32 rows are one 2 KB bulk copy)
    const uint32_t gf = smem_u32(g_full) + 8 * (2 * h + w), ge = smem_u32(g_empty) + 8 * (2 * h + w);
    const uint32_t g_tile = smem_u32(sG) + (2 * h + w) * BC_GTILE;
    const uint32_t g_row = g_tile + row_l * 64;
    const uint32_t g_warp = g_tile + q * 2048;
    uint32_t gf_peer[NC - 1], g_warp_peer[NC - 1];
#pragma unroll
    for (int pi = 0; pi < NC - 1; ++pi) {
      const uint32_t peer = static_cast<uint32_t>((h + 1 + pi) % NC);
      gf_peer[pi] = mapa_u32(gf, peer);
      g_warp_peer[pi] = mapa_u32(g_warp, peer);
    }
    const uint32_t g_swz = static_cast<uint32_t>((row_l >> 1) & 3);
    {
      // warpgroup 0: X[row, h*256 .. +256) -> TMEM columns tmem_x .. +128 of this thread's lane (bf16 pairs, K ascending)
      // warpgroup 1: out-of-slice chunks 0..2 of X -> tmem_x2
      const int nch = (w == 0) ? 4 : BC_XT_OUT;
      const uint32_t dst = (w == 0) ? tmem_x : tmem_x2;
      const __nv_bfloat16* xrow = (dir ? p.xmat[1] : p.xmat[0]) + static_cast<long long>(row_ok ? row : 0) * D;
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        const int chunk = (w == 0) ? h * 4 + c : out_chunk(c);
        const uint4* xsrc = reinterpret_cast<const uint4*>(xrow + chunk * 64);
        uint32_t xr[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 t = row_ok ? __ldg(xsrc + i) : make_uint4(0u, 0u, 0u, 0u);
          xr[4 * i] = t.x; xr[4 * i + 1] = t.y; xr[4 * i + 2] = t.z; xr[4 * i + 3] = t.w;
        }
        tmem_st_x32(dst + lane_base + c * 32, xr);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(smem_u32(xt_full));
    }
    uint32_t ph = 0;
    for (int m = w; m < nown; m += 2) {
      const int n = NC * m + h;
      const int col0 = (tile0 + n) * BC_BN;
      const bool full_tile = col0 + BC_BN <= ncols;
      const bool diag_tile = (col0 < warp_diag_lo + 32) && (col0 + BC_BN > warp_diag_lo);
      float cs[32];                                  // column statistics first: their L2 latency hides behind the wait for S
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
      mbar_wait_a(sf, ph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_x32(tmem_s + lane_base, v);
      tmem_ld_wait();
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
            if (col == diag_col) gv -= 1.0f;           // G_ii = p_ii - 1 rounded as a whole: error relative to G_ii itself
            g[t] = (col < ncols) ? gv : 0.f;
          }
          packed[i / 2] = pack_bf16x2(g[0], g[1]);
        }
      }
      // slot (h, w) is reused every second own tile: wait until ALL CTAs' dX MMAs of its previous use have read it
      mbar_wait_cluster_a(ge, ((m >> 1) & 1) ^ 1);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        st_shared_v4(g_row + ((static_cast<uint32_t>(c) ^ g_swz) << 4),
                     make_uint4(packed[c * 4], packed[c * 4 + 1], packed[c * 4 + 2], packed[c * 4 + 3]));
      fence_proxy_async_smem();                       // generic stores -> visible to tcgen05.mma and the bulk copies (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(gf);                            // local dX issuer: 1 of 4 warps
#pragma unroll
        for (int pi = 0; pi < NC - 1; ++pi)           // this warp's 32 rows to every peer CTA
          bulk_copy_s2s_cluster(g_warp_peer[pi], g_warp, 2048, gf_peer[pi]);
      }
      ph ^= 1;
    }
    // final: dX[:, h*256 + w*128 .. +128) = acc * scale
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float scale = p.out_scale;
    if (p.grad_scale) scale *= *p.grad_scale;
    float* orow = (dir ? p.out[1] : p.out[0]) + split * (dir ? p.split_stride[1] : p.split_stride[0]) +
                  static_cast<long long>(row) * D + h * 256 + w * 128;
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
  cluster_sync_all();                               // peers may still be writing into / arriving on this CTA's smem
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace b200
