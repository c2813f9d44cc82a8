// b200clip: zero-shot label scoring (a-Z).  Reference: 0426/disease_analysis.py:329-356 (softmax top-k),
// multimodal_attention/disease_analysis.py:345-413 (sigmoid(cos/0.5) >= thr), NB02 c41:27-32 / c44:24-36 (cosine
// argmax / sigmoid(cos) > 0.5) and the north-star 14 x (positive, negative) prompt shape.
//
// HBM-bound: N x 512 bf16 embeddings are read exactly once with 128-bit loads straight into tensor-core fragments
// (no shared-memory staging of X).  28*512 FMAs per embedding would make fp32 CUDA cores the bound (~0.4 ms for 1M
// rows vs 157 us of HBM time), so the <=32 dot products per row run on warp-level mma.sync.m16n8k16 (bf16 -> fp32):
// dot products are invariant under a permutation of K, so each lane's 16 contiguous bytes of a row ARE a valid
// A fragment for two k16 steps as long as the prompt fragments (prepared once per CTA in shared memory) use the
// same permutation.  Row norms, (pos,neg) differences, thresholds, argmax/top-k and bit-packing stay in registers
// with quad shuffles.
// Exactness: integer outputs must equal the exactly-rounded answer.  Rows whose decision margin is below a guard
// band are re-evaluated in fp64 on device (warp-cooperative), so fp32 accumulation order cannot flip a label.
#include "common.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

// embedding width: template parameter of the kernels (ZsCfg<DD>, 512 or 768)

constexpr int ZS_MAXP = 32;

struct ZsParams {
  const __nv_bfloat16* x; long long ldx; long long n;
  const __nv_bfloat16* prompts;      // [np, 512] bf16, already L2-normalised
  int np;                            // prompts (<= 32)
  int pair_mode;                     // 1: labels = np/2, score_l = l(2l) - l(2l+1) ; 0: labels = np, score = l
  int nlabels;
  int normalize_x;
  float inv_tau;
  float thr_logit[ZS_MAXP];          // label passes when score (>|>=) thr_logit
  int thr_inclusive;                 // 1: >= (multimodal Z2), 0: >
  float guard;                       // fp64 re-evaluation band on the score scale
  int topk;                          // 0..4
  int value_mode;                    // top-k values: 0 raw score, 1 softmax prob over labels, 2 sigmoid prob
  uint8_t* argmax;                   // [n] or null
  void* mask; int mask_is_u32;       // [n] u16 / u32 or null
  uint8_t* topk_idx;                 // [n, topk] or null
  float* topk_val;                   // [n, topk] or null
  float* scores;                     // [n, nlabels] f32 or null
  unsigned long long* guard_count;   // rows re-evaluated in fp64 (diagnostic) or null
  unsigned int* fix_count;           // deferred re-evaluation: number of listed rows (device counter, zeroed by the host) or null
  unsigned int* fix_rows;            // [fix_cap] row indices
  unsigned int fix_cap;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ bool passes(float s, float thr, int inclusive) { return inclusive ? (s >= thr) : (s > thr); }
__device__ __forceinline__ bool passes_d(double s, double thr, int inclusive) { return inclusive ? (s >= thr) : (s > thr); }

// Final decisions for one row from its label scores held by ONE thread (used by the fp64 path; scores in double).
__device__ void emit_row_fp64(const ZsParams& p, long long row, const double* sc) {
  const int L = p.nlabels;
  unsigned int mask = 0;
  int best = 0;
  for (int l = 0; l < L; ++l) {
    if (passes_d(sc[l], static_cast<double>(p.thr_logit[l]), p.thr_inclusive)) mask |= 1u << l;
    if (sc[l] > sc[best]) best = l;
  }
  if (p.argmax) p.argmax[row] = static_cast<uint8_t>(best);
  if (p.mask) {
    if (p.mask_is_u32) static_cast<uint32_t*>(p.mask)[row] = mask;
    else static_cast<uint16_t*>(p.mask)[row] = static_cast<uint16_t>(mask);
  }
  if (p.scores)
    for (int l = 0; l < L; ++l) p.scores[row * L + l] = static_cast<float>(sc[l]);
  if (p.topk > 0 && p.topk_idx) {
    double mx = sc[best], den = 0.0;
    if (p.value_mode == 1)
      for (int l = 0; l < L; ++l) den += exp(sc[l] - mx);
    unsigned int taken = 0;
    for (int k = 0; k < p.topk; ++k) {
      int b = -1;
      for (int l = 0; l < L; ++l)
        if (!((taken >> l) & 1u) && (b < 0 || sc[l] > sc[b])) b = l;
      if (b < 0) b = 0;
      taken |= 1u << b;
      p.topk_idx[row * p.topk + k] = static_cast<uint8_t>(b);
      if (p.topk_val) {
        double v = sc[b];
        if (p.value_mode == 1) v = exp(sc[b] - mx) / den;
        else if (p.value_mode == 2) v = 1.0 / (1.0 + exp(-sc[b]));
        p.topk_val[row * p.topk + k] = static_cast<float>(v);
      }
    }
  }
}

// Rows per consumer iteration: 8 (one MMA N-tile).  Round 2 measurement: with 16-row blocks and 12 consumer warps the kernel was
// LATENCY-bound (42 % issue utilisation, ~8 clk per instruction per warp, tensor pipe 36 %, DRAM 4.0 TB/s even with the fp64
// path switched off): 3 warps per scheduler cannot hide LDS / HMMA / shuffle latencies.  The ring needs one stage per consumer
// (see the static_assert), so more warps means smaller blocks: 24 consumers x 8-row (8 KB) stages = the same 192 KB in flight.
constexpr int ZS_ROWS = 8;                              // rows per block = per stage
// Per embedding width (512 = 0426/config.py:30 default, 768 = BASELINE.json configs[4]'s other width): the prompt fragment table
// (32 prompts x DD bf16) and the ring share the 227 KB of shared memory, so the wider rows get fewer, larger stages.
template <int DD> struct ZsCfg {
  static_assert(DD == 512 || DD == 768, "zero-shot kernel widths");
  static constexpr int KS = DD / 32;                    // 32-element steps per row (two k16 MMAs each)
  static constexpr int EPL = DD / 32;                   // elements per lane in the exact re-evaluation
  static constexpr int FRAG_BYTES = KS * 4 * 32 * 16;   // 32 KB / 48 KB
  static constexpr int CONSUMERS = DD == 512 ? 24 : 14; // consumer warps; the last warp of the CTA is the bulk-copy producer
  static constexpr int THREADS = (CONSUMERS + 1) * 32;
  static constexpr int PITCH = DD * 2;                  // dense rows: ONE bulk copy per block (small copies cost ~75 clk each)
  static constexpr int STAGE_BYTES = ZS_ROWS * PITCH;   // one 8-row block: 8 KB / 12 KB
  // a consumer may hold a claim at most CONSUMERS-1 blocks ahead of the oldest unconsumed block; with fewer stages than
  // consumers a claim could be two fills ahead of its stage's barrier and the parity wait would alias
  static constexpr int STAGES = CONSUMERS;
  static constexpr int SMEM_BYTES = FRAG_BYTES + STAGES * STAGE_BYTES + 2 * STAGES * 8 + 128;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void bulk_load_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
               : "memory");
}

// Rows are staged by a producer warp with asynchronous bulk copies (ONE copy per 8-row block: 8 KB at D = 512, 12 KB at 768)
// into a ring with one stage per consumer warp, so 170-190 KB of loads are in flight per SM independent of the consumers'
// registers and occupancy; the consumer warps claim blocks dynamically and read their MMA fragments straight from the rows.
// Exact re-evaluation of ONE row by a whole warp (rows whose fast-path decision margin was inside the guard band).
template <bool PAIR, int DD>
__device__ __forceinline__ void reevaluate_row(const ZsParams& p, long long row, int lane) {
  constexpr int EPL = ZsCfg<DD>::EPL, NW = EPL / 2, NV = EPL / 8;      // elements, 32-bit words, 16-byte loads per lane
  const int L = p.nlabels;
  // lane owns k in [EPL*lane, EPL*lane + EPL)
  const uint4* px = reinterpret_cast<const uint4*>(p.x + row * p.ldx) + lane * NV;
  uint32_t xw[NW];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 q = __ldg(px + v);
    xw[4 * v] = q.x; xw[4 * v + 1] = q.y; xw[4 * v + 2] = q.z; xw[4 * v + 3] = q.w;
  }
  double ss = 0.0;
#pragma unroll
  for (int jx = 0; jx < NW; ++jx) {
    const double a = static_cast<double>(bf16_lo(xw[jx])), b = static_cast<double>(bf16_hi(xw[jx]));
    ss += a * a + b * b;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const double kk = (p.normalize_x ? 1.0 / fmax(sqrt(ss), 1e-12) : 1.0) * static_cast<double>(p.inv_tau);
  // Dot products in double-float arithmetic on the fp32 pipe (B200's fp64 rate is 1/64: 28 x 16 DFMA per lane made this path
  // cost ~50 us of a 270 us kernel at 0.4 % flagged rows).  bf16 x bf16 products are EXACT in fp32 (8 + 8 significand
  // bits), and Knuth's TwoSum keeps the rounding error of every addition, so (hi, lo) carries the sum to ~2^-45: the
  // decisions made from it are those of the fp64 evaluation unless a margin is below ~1e-13 of the score scale.
  float xf[EPL];
#pragma unroll
  for (int jx = 0; jx < NW; ++jx) { xf[2 * jx] = bf16_lo(xw[jx]); xf[2 * jx + 1] = bf16_hi(xw[jx]); }
  double l[ZS_MAXP];
  // 4 prompts per trip: their TwoSum chains (16 dependent additions each) and L2 round trips are independent, so one row
  // costs ~7 chain latencies instead of 28 (the fix-up kernel's duration IS one row's latency: every warp has 1-2 rows)
#pragma unroll 4
  for (int c = 0; c < p.np; ++c) {
    // the prompts themselves (28 KB, L1/L2-resident) -- the fragment table is in MMA register order
    const uint4* pp = reinterpret_cast<const uint4*>(p.prompts + static_cast<long long>(c) * DD) + lane * NV;
    uint32_t pw[NW];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const uint4 q = __ldg(pp + v);
      pw[4 * v] = q.x; pw[4 * v + 1] = q.y; pw[4 * v + 2] = q.z; pw[4 * v + 3] = q.w;
    }
    float hi = 0.f, lo = 0.f;
#pragma unroll
    for (int jx = 0; jx < EPL; ++jx) {
      const float pv = (jx & 1) ? bf16_hi(pw[jx >> 1]) : bf16_lo(pw[jx >> 1]);
      const float prod = __fmul_rn(xf[jx], pv);                         // exact
      const float s = __fadd_rn(hi, prod);                              // TwoSum(hi, prod)
      const float bb = __fsub_rn(s, hi);
      lo = __fadd_rn(lo, __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(prod, bb)));
      hi = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                                  // TwoSum tree over the lanes
      const float oh = __shfl_xor_sync(0xffffffffu, hi, o), ol = __shfl_xor_sync(0xffffffffu, lo, o);
      const float s = __fadd_rn(hi, oh);
      const float bb = __fsub_rn(s, hi);
      lo = __fadd_rn(__fadd_rn(lo, ol), __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(oh, bb)));
      hi = s;
    }
    l[c] = (static_cast<double>(hi) + static_cast<double>(lo)) * kk;
  }
  if (lane == 0) {
    double sc[ZS_MAXP];
    for (int lab = 0; lab < L; ++lab) sc[lab] = PAIR ? (l[2 * lab] - l[2 * lab + 1]) : l[lab];
    emit_row_fp64(p, row, sc);
    if (p.guard_count) atomicAdd(p.guard_count, 1ull);
  }
}

// Second pass: one warp per listed row.  Inline, the re-evaluation cost 43 us of a 263 us kernel at 0.4 % flagged rows (28
// dependent L2 round trips per row on ONE warp of an SM whose other warps wait on the same instruction cache); here the
// rows are independent work items of a full grid.
template <bool PAIR, int DD>
__global__ void __launch_bounds__(256) zeroshot_fixup_kernel(const ZsParams p) {
  const int lane = threadIdx.x & 31;
  const unsigned int nfix = min(*p.fix_count, p.fix_cap);
  for (unsigned int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < nfix; i += gridDim.x * 8)
    reevaluate_row<PAIR, DD>(p, static_cast<long long>(p.fix_rows[i]), lane);
}

// PAIR / TOPK are compile-time: the decision code below is branch-heavy and the common call (pair mode or plain labels, no
// top-k) should not carry the other modes' instructions.
//
// Operand roles (round 2, from two ncu source-page captures: the kernel was ISSUE-bound, 2458 then 1924 warp instructions per
// 16-row block at 53 % issue utilisation against 3.5 TB/s of DRAM reads; 36 % of them IMAD.MOV):
//   * the PROMPTS are the A operand (M = 16 prompts per tile, two tiles), the EMBEDDINGS the B operand (N = 8 rows): a B
//     fragment is two registers holding consecutive K elements of ONE row, i.e. exactly the halves of the 16 bytes a lane
//     loads, so no register shuffling is needed; the A fragments come from a table laid out in register order.  (With X as the
//     A operand a fragment interleaves registers of rows r and r + 8: ptxas rebuilt that quad with 4 moves per MMA.)
//   * dot products are invariant under a permutation of K, so a lane's 16 contiguous bytes of a row serve two k16 steps as
//     long as the prompt table uses the same permutation: logical k pairs (2t, 2t+1 | 2t+8, 2t+9) of step j <-> elements
//     32s + 8t + 4j + (0,1 | 2,3).
//   * prompt -> M index: tile mt holds prompts 16mt + 2g (row g) and 16mt + 2g + 1 (row g + 8), so the (positive, negative)
//     prompts of a label -- or two neighbouring labels -- meet in ONE lane: c0/c1 = prompt 16mt+2g x rows (2t, 2t+1),
//     c2/c3 = prompt 16mt+2g+1 x the same rows.
//   * row norms: diagonal of the Gram blocks X[0..15] . X[0..7]^T and X[0..15] . X[8..15]^T (one extra MMA pair per k16 with X
//     as BOTH operands; the A quad of rows (r, r+8) costs 4 moves per k16, the only ones left).
template <bool PAIR, bool TOPK, int DD>
__global__ void __launch_bounds__(ZsCfg<DD>::THREADS, 1) zeroshot_kernel(const ZsParams p) {
  using Cfg = ZsCfg<DD>;
  constexpr int ZS_STAGES = Cfg::STAGES, ZS_STAGE_BYTES = Cfg::STAGE_BYTES, ZS_PITCH = Cfg::PITCH, ZS_CONSUMERS = Cfg::CONSUMERS;
  constexpr int ZS_THREADS2 = Cfg::THREADS, ZS_D = DD;
  extern __shared__ __align__(128) uint8_t zs_smem[];
  // A fragments in consumption order: frag[((s*2 + mt)*2 + j)*32 + lane] = {P[pr0][kb..+1], P[pr1][kb..+1], P[pr0][kb+2..+3],
  // P[pr1][kb+2..+3]}, pr0 = 16mt + 2(lane/4), pr1 = pr0 + 1, kb = 32s + 8(lane%4) + 4j
  uint4* s_frag = reinterpret_cast<uint4*>(zs_smem);
  uint8_t* ring = zs_smem + Cfg::FRAG_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + ZS_STAGES * ZS_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + ZS_STAGES;
  int* next_claim = reinterpret_cast<int*>(empty_bar + ZS_STAGES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = lane & 3, g = lane >> 2;
  for (int i = threadIdx.x; i < Cfg::KS * 4 * 32; i += ZS_THREADS2) {
    const int l = i & 31, j = (i >> 5) & 1, mt = (i >> 6) & 1, s = i >> 7;
    const int pr0 = 16 * mt + 2 * (l >> 2), kb = 32 * s + 8 * (l & 3) + 4 * j;
    uint2 v0 = make_uint2(0u, 0u), v1 = make_uint2(0u, 0u);
    if (pr0 < p.np) v0 = *reinterpret_cast<const uint2*>(p.prompts + static_cast<long long>(pr0) * ZS_D + kb);
    if (pr0 + 1 < p.np) v1 = *reinterpret_cast<const uint2*>(p.prompts + static_cast<long long>(pr0 + 1) * ZS_D + kb);
    s_frag[i] = make_uint4(v0.x, v1.x, v0.y, v1.y);
  }
  if (threadIdx.x == 0) {
    *next_claim = 0;
    for (int s = 0; s < ZS_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const long long nblocks = (p.n + ZS_ROWS - 1) / ZS_ROWS;
  const int L = p.nlabels;
  const uint32_t ring_a = smem_u32(ring), full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);

  if (warp == ZS_CONSUMERS) {
    // ===================== producer =====================
    int st = 0;
    uint32_t ph = 0;
    for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
      mbar_wait_a(empty_a + 8 * st, ph ^ 1);
      const long long row0 = blk * ZS_ROWS;
      const int valid = static_cast<int>(min(static_cast<long long>(ZS_ROWS), p.n - row0));
      if (lane == 0) mbar_arrive_expect_tx_a(full_a + 8 * st, static_cast<uint32_t>(valid) * ZS_D * 2);
      __syncwarp();
      if (p.ldx == ZS_D) {
        if (lane == 0)
          bulk_load_g2s(ring_a + st * ZS_STAGE_BYTES, p.x + row0 * p.ldx, static_cast<uint32_t>(valid) * ZS_D * 2, full_a + 8 * st);
      } else if (lane < valid) {
        bulk_load_g2s(ring_a + st * ZS_STAGE_BYTES + lane * ZS_PITCH, p.x + (row0 + lane) * p.ldx, ZS_D * 2, full_a + 8 * st);
      }
      if (++st == ZS_STAGES) { st = 0; ph ^= 1; }
    }
    return;
  }

  // ===================== consumers: blocks are claimed dynamically (a warp busy with the rare fp64 re-evaluation must not
  // stall the ring for the others); block i of this CTA lives in stage i % ZS_STAGES =====================
  constexpr int CNT = PAIR ? 2 : 4;                       // label scores per (lane, row)
  // the labels this lane decides are the same for every block: ids, validity and thresholds are loop invariants
  int id[CNT];
  float thr_l[CNT];
  bool lab_ok[CNT];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    if (PAIR) {
      id[mt] = 8 * mt + g;
    } else {
      id[2 * mt] = 16 * mt + 2 * g;
      id[2 * mt + 1] = 16 * mt + 2 * g + 1;
    }
  }
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    lab_ok[i] = id[i] < L;
    thr_l[i] = lab_ok[i] ? p.thr_logit[id[i]] : INFINITY;
  }
  const bool incl = p.thr_inclusive != 0;
  const bool want_mask = p.mask != nullptr, rank_matters = (p.argmax != nullptr) || TOPK;
  const float guard = p.guard, inv_tau = p.inv_tau;
  const bool normalize = p.normalize_x != 0;
  for (;;) {
    long long it = 0;
    if (lane == 0) it = atomicAdd(next_claim, 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    const long long blk = static_cast<long long>(blockIdx.x) + it * gridDim.x;
    if (blk >= nblocks) break;
    const int st = static_cast<int>(it % ZS_STAGES);
    const uint32_t ph = static_cast<uint32_t>((it / ZS_STAGES) & 1);
    mbar_wait_a(full_a + 8 * st, ph);
    // this lane's 16 bytes of row g at every 32-element step (B fragments: .x,.y = first k16, .z,.w = second); rows past the
    // end of the matrix (last block only) read the last valid row of the stage: their results are never written
    const int nvalid = static_cast<int>(min(static_cast<long long>(ZS_ROWS), p.n - blk * ZS_ROWS));
    const uint4* pa = reinterpret_cast<const uint4*>(ring + st * ZS_STAGE_BYTES + min(g, nvalid - 1) * ZS_PITCH) + t;
    float acc[2][4];                                      // [prompt tile mt][c0..c3]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;
    float na[4] = {0.f, 0.f, 0.f, 0.f};                   // Gram block X . X^T for the row norms (A rows 8..15 are zero)
#pragma unroll 4
    for (int s = 0; s < Cfg::KS; ++s) {
      const uint4 xa = pa[s * 4];
      if (normalize) {
        mma_bf16_16816(na, xa.x, 0u, xa.y, 0u, xa.x, xa.y);
        mma_bf16_16816(na, xa.z, 0u, xa.w, 0u, xa.z, xa.w);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint4 f0 = s_frag[((s * 2 + mt) * 2 + 0) * 32 + lane];
        const uint4 f1 = s_frag[((s * 2 + mt) * 2 + 1) * 32 + lane];
        mma_bf16_16816(acc[mt], f0.x, f0.y, f0.z, f0.w, xa.x, xa.y);
        mma_bf16_16816(acc[mt], f1.x, f1.y, f1.z, f1.w, xa.z, xa.w);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_a(empty_a + 8 * st);       // stage consumed (smem reads above are complete: values are in registers)

    // decisions: this lane holds, for the 2 rows 2t + j, the scores of labels (PAIR: 8mt + g | single: 16mt + 2g, +1);
    // a row's labels are spread over the 8 lanes with the same t (xor 4, 8, 16)
    unsigned int flag_rows = 0;                           // bit j: this row needs the fp64 re-evaluation
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const long long row = blk * ZS_ROWS + 2 * t + j;
      const bool ok = row < p.n;
      // ||row||^2: diagonal entry C[i][i] (i = 2t + j) of the Gram block lives in lane 4i + i/2 = 9t + 4j, register j
      const float ss = __shfl_sync(0xffffffffu, na[j], 9 * t + 4 * j);
      // 1 / max(||x||, 1e-12) / tau  (rsqrt: the rounding of this scale only matters inside the guard band)
      const float kk = (normalize ? rsqrtf(fmaxf(ss, 1e-24f)) : 1.0f) * inv_tau;
      float sc[CNT];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (PAIR) {
          sc[mt] = (acc[mt][j] - acc[mt][2 + j]) * kk;
        } else {
          sc[2 * mt] = acc[mt][j] * kk;
          sc[2 * mt + 1] = acc[mt][2 + j] * kk;
        }
      }
      unsigned int mask = 0;
      bool near_thr = false;
      float best = -INFINITY;
      int best_id = 1 << 20;
#pragma unroll
      for (int i = 0; i < CNT; ++i) {
        const float s_i = lab_ok[i] ? sc[i] : -INFINITY;          // labels past L never pass, never win
        mask |= ((s_i > thr_l[i]) || (incl && s_i == thr_l[i])) ? (1u << id[i]) : 0u;
        near_thr |= fabsf(s_i - thr_l[i]) < guard;
        if (s_i > best) { best = s_i; best_id = id[i]; }          // ids ascend with i: the first maximum wins ties
      }
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        mask |= __shfl_xor_sync(0xffffffffu, mask, o);
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
        if (ob > best || (ob == best && oi < best_id)) { best = ob; best_id = oi; }
      }
      bool near_top = false;
      if (rank_matters) {
#pragma unroll
        for (int i = 0; i < CNT; ++i) near_top |= lab_ok[i] && id[i] != best_id && (best - sc[i]) < guard;
      }
      if (TOPK && p.topk > 1) {
        // with top-k > 1 every adjacent gap matters; be conservative: any two labels of the row closer than the guard
#pragma unroll
        for (int a = 0; a < CNT; ++a)
#pragma unroll
          for (int b = a + 1; b < CNT; ++b)
            if (lab_ok[a] && lab_ok[b]) near_top |= fabsf(sc[a] - sc[b]) < guard;
#pragma unroll
        for (int o = 4; o < 32; o += 4)                    // the seven other lanes of this row
#pragma unroll
          for (int b = 0; b < CNT; ++b) {
            const float other = __shfl_xor_sync(0xffffffffu, sc[b], o);
            const int oid = __shfl_xor_sync(0xffffffffu, id[b], o);
#pragma unroll
            for (int a = 0; a < CNT; ++a)
              if (lab_ok[a] && oid < L) near_top |= fabsf(sc[a] - other) < guard;
          }
      }
      unsigned int fl = ((near_thr && want_mask) || near_top) ? 1u : 0u;
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
      if (fl) flag_rows |= 1u << j;
      const bool emit = ok && !fl;
      if (emit && p.scores) {
#pragma unroll
        for (int i = 0; i < CNT; ++i)
          if (lab_ok[i]) p.scores[row * L + id[i]] = sc[i];
      }
      if (TOPK && p.topk_idx) {                           // warp-uniform branch: shuffles run on all lanes
        // top-k of the row by repeated arg-max with removal over its 8 lanes
        const float mx = best;
        float den = 0.f;
        if (p.value_mode == 1) {
#pragma unroll
          for (int i = 0; i < CNT; ++i)
            if (lab_ok[i]) den += expf(sc[i] - mx);
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        }
        unsigned int taken = 0;
        for (int k = 0; k < p.topk; ++k) {
          float bv = -INFINITY;
          int bi = 1 << 20;
#pragma unroll
          for (int i = 0; i < CNT; ++i)
            if (lab_ok[i] && !((taken >> id[i]) & 1u) && (sc[i] > bv || (sc[i] == bv && id[i] < bi))) { bv = sc[i]; bi = id[i]; }
#pragma unroll
          for (int o = 4; o <= 16; o <<= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > bv || (ob == bv && oi < bi)) { bv = ob; bi = oi; }
          }
          if (bi >= L) bi = 0;
          taken |= 1u << bi;
          if (emit && g == 0) {
            p.topk_idx[row * p.topk + k] = static_cast<uint8_t>(bi);
            if (p.topk_val) {
              float v = bv;
              if (p.value_mode == 1) v = expf(bv - mx) / den;
              else if (p.value_mode == 2) v = 1.0f / (1.0f + expf(-bv));
              p.topk_val[row * p.topk + k] = v;
            }
          }
        }
      }
      if (emit && g == 0) {
        if (p.argmax) p.argmax[row] = static_cast<uint8_t>(best_id);
        if (p.mask) {
          if (p.mask_is_u32) static_cast<uint32_t*>(p.mask)[row] = mask;
          else static_cast<uint16_t*>(p.mask)[row] = static_cast<uint16_t>(mask);
        }
      }
    }

    // fp64 re-evaluation of flagged rows, one row at a time by the whole warp (rare: < 1 % of rows)
    unsigned int rows_flagged = 0;                        // bit i: row blk*8+i
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const unsigned int bal = __ballot_sync(0xffffffffu, ((flag_rows >> j) & 1u) && g == 0) & 0xFu;   // lanes 0..3 = t
#pragma unroll
      for (int tt = 0; tt < 4; ++tt)
        if ((bal >> tt) & 1u) rows_flagged |= 1u << (2 * tt + j);
    }
    while (rows_flagged) {
      const int i = __ffs(rows_flagged) - 1;
      rows_flagged &= rows_flagged - 1;
      const long long row = blk * ZS_ROWS + i;
      if (row >= p.n) continue;
      if (p.fix_count != nullptr) {                       // deferred: list the row for zeroshot_fixup_kernel
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(p.fix_count, 1u);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        if (slot < p.fix_cap) {
          if (lane == 0) p.fix_rows[slot] = static_cast<unsigned int>(row);
          continue;
        }
      }
      reevaluate_row<PAIR, DD>(p, row, lane);             // no list (or list full): re-evaluate in place
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200clip_zeroshot_workspace_bytes(long long n) {
  // room for every row (a pathological batch where all margins are inside the guard band); 4 bytes per row
  return n <= 0 ? 0 : 256 + static_cast<size_t>(n) * 4;
}

extern "C" int b200clip_zeroshot_score(const void* x_bf16, long long ldx, long long n, const void* prompts_bf16, int np,
                                       int D, int pair_mode, int normalize_x, float temperature,
                                       const float* thr_logit_host, int thr_inclusive, float guard, int topk,
                                       int value_mode, uint8_t* argmax, void* mask, int mask_is_u32, uint8_t* topk_idx,
                                       float* topk_val, float* scores, unsigned long long* guard_count, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  B200_REQUIRE(D == 512 || D == 768, "zeroshot: D=%d unsupported (the kernel is built for D = 512 and 768)", D);
  B200_REQUIRE(n >= 0 && np > 0 && np <= ZS_MAXP, "zeroshot: need 0 < np <= %d", ZS_MAXP);
  B200_REQUIRE(!pair_mode || np % 2 == 0, "zeroshot: pair mode needs an even number of prompts");
  B200_REQUIRE(temperature > 0.f && guard >= 0.f && topk >= 0 && topk <= 4, "zeroshot: bad scalar arguments");
  B200_REQUIRE(aligned16(x_bf16) && aligned16(prompts_bf16) && ldx % 8 == 0, "zeroshot: operands must be 16-byte aligned");
  const int L = pair_mode ? np / 2 : np;
  B200_REQUIRE(mask == nullptr || mask_is_u32 || L <= 16, "zeroshot: u16 mask holds at most 16 labels");
  B200_REQUIRE(thr_logit_host != nullptr || mask == nullptr, "zeroshot: thresholds required when a mask is requested");
  if (n == 0) return B200_OK;
  ZsParams p{};
  p.x = static_cast<const __nv_bfloat16*>(x_bf16); p.ldx = ldx; p.n = n;
  p.prompts = static_cast<const __nv_bfloat16*>(prompts_bf16); p.np = np; p.pair_mode = pair_mode; p.nlabels = L;
  p.normalize_x = normalize_x; p.inv_tau = 1.0f / temperature;
  for (int i = 0; i < ZS_MAXP; ++i) p.thr_logit[i] = (thr_logit_host && i < L) ? thr_logit_host[i] : INFINITY;
  p.thr_inclusive = thr_inclusive; p.guard = guard; p.topk = topk; p.value_mode = value_mode;
  p.argmax = argmax; p.mask = mask; p.mask_is_u32 = mask_is_u32; p.topk_idx = topk_idx; p.topk_val = topk_val;
  p.scores = scores; p.guard_count = guard_count;
  // optional workspace: [count u32 | pad | row indices u32 ...] -> flagged rows are re-evaluated by a second kernel
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (workspace != nullptr && workspace_bytes >= 256 + 4 && n < (1ll << 32)) {
    p.fix_count = static_cast<unsigned int*>(workspace);
    p.fix_rows = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + 256);
    p.fix_cap = static_cast<unsigned int>(std::min<size_t>((workspace_bytes - 256) / 4, 0xFFFFFFFFull));
    B200_CHECK_CUDA(cudaMemsetAsync(p.fix_count, 0, 4, s));
  }
  const long long nblk = (n + ZS_ROWS - 1) / ZS_ROWS;
  const int grid = static_cast<int>(std::min<long long>(nblk, static_cast<long long>(num_sms())));
  static SmemAttrOnce attr[8];
#define B200_ZS_CASE(PAIR, TOPK, DD, IDX)                                                                       \
  if ((pair_mode != 0) == PAIR && (topk > 0) == TOPK && D == DD) {                                              \
    B200_CHECK_CUDA(attr[IDX].ensure(zeroshot_kernel<PAIR, TOPK, DD>, ZsCfg<DD>::SMEM_BYTES));                  \
    zeroshot_kernel<PAIR, TOPK, DD><<<grid, ZsCfg<DD>::THREADS, ZsCfg<DD>::SMEM_BYTES, s>>>(p);                 \
  }
  B200_ZS_CASE(false, false, 512, 0)
  B200_ZS_CASE(false, true, 512, 1)
  B200_ZS_CASE(true, false, 512, 2)
  B200_ZS_CASE(true, true, 512, 3)
  B200_ZS_CASE(false, false, 768, 4)
  B200_ZS_CASE(false, true, 768, 5)
  B200_ZS_CASE(true, false, 768, 6)
  B200_ZS_CASE(true, true, 768, 7)
#undef B200_ZS_CASE
  B200_LAUNCH_CHECK();
  if (p.fix_count != nullptr) {
    const int fgrid = 2 * num_sms();
    if (D == 512) {
      if (pair_mode) zeroshot_fixup_kernel<true, 512><<<fgrid, 256, 0, s>>>(p);
      else zeroshot_fixup_kernel<false, 512><<<fgrid, 256, 0, s>>>(p);
    } else {
      if (pair_mode) zeroshot_fixup_kernel<true, 768><<<fgrid, 256, 0, s>>>(p);
      else zeroshot_fixup_kernel<false, 768><<<fgrid, 256, 0, s>>>(p);
    }
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}
