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

constexpr int ZS_D = 512;

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
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float sumsq8(uint4 v) {
  float s = 0.f;
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = bf16_lo(w[i]), b = bf16_hi(w[i]);
    s = fmaf(a, a, s);
    s = fmaf(b, b, s);
  }
  return s;
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

constexpr int ZS_CONSUMERS = 12;                        // consumer warps (3 per SM sub-partition); the last warp is the bulk-copy producer
constexpr int ZS_THREADS2 = (ZS_CONSUMERS + 1) * 32;
constexpr int ZS_PITCH = ZS_D * 2;                      // dense rows: ONE 16 KB bulk copy per block (small copies cost ~75 clk each)
constexpr int ZS_STAGE_BYTES = 16 * ZS_PITCH;           // one 16-row block
constexpr int ZS_STAGES = 12;
// a consumer may hold a claim at most ZS_CONSUMERS-1 blocks ahead of the oldest unconsumed block; with fewer stages than
// consumers a claim could be two fills ahead of its stage's barrier and the parity wait would alias
static_assert(ZS_STAGES >= ZS_CONSUMERS, "ring must have at least as many stages as consumer warps");
constexpr int ZS_SMEM_BYTES = 16 * 4 * 32 * 16 + ZS_STAGES * ZS_STAGE_BYTES + 2 * ZS_STAGES * 8 + 128;

__device__ __forceinline__ void bulk_load_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(bar)
               : "memory");
}

// Rows are staged by a producer warp with asynchronous bulk copies (one 16 KB copy per 16-row block) into
// a 10-stage ring, so ~160 KB of loads are in flight per SM independent of the consumers' registers and occupancy;
// the 8 consumer warps take 16-row blocks round-robin and read their MMA fragments from the padded rows.
// PAIR / TOPK are compile-time: the decision code below is branch-heavy and the common call (pair mode or plain labels, no
// top-k) should not carry the other modes' instructions (ncu, round 2: the kernel was ISSUE-bound -- 2458 warp instructions
// per 16-row block at 53 % issue utilisation against 3.5 TB/s of DRAM reads).
template <bool PAIR, bool TOPK>
__global__ void __launch_bounds__(ZS_THREADS2, 1) zeroshot_kernel(const ZsParams p) {
  extern __shared__ __align__(128) uint8_t zs_smem[];
  // prompt fragments in consumption order: frag[(s*4 + t)*32 + lane] = P[8t + lane/4][32s + 8(lane%4) .. +7]
  uint4* s_frag = reinterpret_cast<uint4*>(zs_smem);
  uint8_t* ring = zs_smem + 16 * 4 * 32 * 16;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + ZS_STAGES * ZS_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + ZS_STAGES;
  int* next_claim = reinterpret_cast<int*>(empty_bar + ZS_STAGES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & 3, r = lane >> 2;
  for (int i = threadIdx.x; i < 16 * 4 * 32; i += ZS_THREADS2) {
    const int l = i & 31, t = (i >> 5) & 3, s = i >> 7;
    const int n = 8 * t + (l >> 2), k0 = 32 * s + 8 * (l & 3);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n < p.np) v = *reinterpret_cast<const uint4*>(p.prompts + static_cast<long long>(n) * ZS_D + k0);
    s_frag[i] = v;
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

  const long long nblocks16 = (p.n + 15) / 16;
  const int L = p.nlabels;
  const uint32_t ring_a = smem_u32(ring), full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);

  if (warp == ZS_CONSUMERS) {
    // ===================== producer =====================
    int st = 0;
    uint32_t ph = 0;
    for (long long blk = blockIdx.x; blk < nblocks16; blk += gridDim.x) {
      mbar_wait_a(empty_a + 8 * st, ph ^ 1);
      const long long row0 = blk * 16;
      const int valid = static_cast<int>(min(16ll, p.n - row0));
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
  // stall the ring for the other seven); block i of this CTA lives in stage i % ZS_STAGES =====================
  for (;;) {
    long long it = 0;
    if (lane == 0) it = atomicAdd(next_claim, 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    const long long blk = static_cast<long long>(blockIdx.x) + it * gridDim.x;
    if (blk >= nblocks16) break;
    const int st = static_cast<int>(it % ZS_STAGES);
    const uint32_t ph = static_cast<uint32_t>((it / ZS_STAGES) & 1);
    const long long row_a = blk * 16 + r, row_b = row_a + 8;
    const bool ok_a = row_a < p.n, ok_b = row_b < p.n;
    mbar_wait_a(full_a + 8 * st, ph);
    const uint4* pa = reinterpret_cast<const uint4*>(ring + st * ZS_STAGE_BYTES + r * ZS_PITCH) + q;
    const uint4* pb = reinterpret_cast<const uint4*>(ring + st * ZS_STAGE_BYTES + (r + 8) * ZS_PITCH) + q;
    float acc[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
    // Row norms on the tensor cores too: with B[k][n] = X[n][k] the B fragment of lane (g = lane/4, q) IS that lane's own A data
    // of row g, so one extra MMA per k16 step accumulates the Gram block X[0..15] . X[0..7]^T (and one more X . X[8..15]^T);
    // ||row g||^2 is its diagonal: C[g][g] = c[g & 1] of lane 4g + g/2, C'[g+8][g] = c[2 + (g & 1)] of the same lane.  The
    // unpack-and-FMA version cost 768 of the 2458 instructions per block (256 FFMA + 512 shift/mask).
    float na[4] = {0.f, 0.f, 0.f, 0.f}, nb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int s = 0; s < 16; ++s) {
      const uint4 xa = ok_a ? pa[s * 4] : make_uint4(0u, 0u, 0u, 0u);
      const uint4 xb = ok_b ? pb[s * 4] : make_uint4(0u, 0u, 0u, 0u);
      if (p.normalize_x) {
        mma_bf16_16816(na, xa.x, xb.x, xa.y, xb.y, xa.x, xa.y);
        mma_bf16_16816(na, xa.z, xb.z, xa.w, xb.w, xa.z, xa.w);
        mma_bf16_16816(nb, xa.x, xb.x, xa.y, xb.y, xb.x, xb.y);
        mma_bf16_16816(nb, xa.z, xb.z, xa.w, xb.w, xb.z, xb.w);
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint4 f = s_frag[(s * 4 + t) * 32 + lane];
        mma_bf16_16816(acc[t], xa.x, xb.x, xa.y, xb.y, f.x, f.y);
        mma_bf16_16816(acc[t], xa.z, xb.z, xa.w, xb.w, f.z, f.w);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_a(empty_a + 8 * st);       // stage consumed (smem reads above are complete: values are in registers)
    // row norms: the diagonal entries live in lane 4r + r/2 of each quad
    const int nsrc = (lane & 28) | (r >> 1);
    const float ss_a = __shfl_sync(0xffffffffu, (r & 1) ? na[1] : na[0], nsrc);
    const float ss_b = __shfl_sync(0xffffffffu, (r & 1) ? nb[3] : nb[2], nsrc);
    const float ka = (p.normalize_x ? 1.0f / fmaxf(sqrtf(ss_a), 1e-12f) : 1.0f) * p.inv_tau;
    const float kb = (p.normalize_x ? 1.0f / fmaxf(sqrtf(ss_b), 1e-12f) : 1.0f) * p.inv_tau;

    // label scores held by this lane.  pair mode: label 4t+q from prompts (8t+2q, 8t+2q+1) -> 4 labels per row;
    // single mode: labels 8t+2q, 8t+2q+1 -> 8 labels per row.
    unsigned int flag_bits = 0;                           // bit0: row a near a decision boundary, bit1: row b
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const float kk = which ? kb : ka;
      float sc[8];
      int id[8];
      int cnt;
      if (PAIR) {
        cnt = 4;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          sc[t] = (acc[t][which * 2] - acc[t][which * 2 + 1]) * kk;
          id[t] = 4 * t + q;
        }
#pragma unroll
        for (int t = 4; t < 8; ++t) { sc[t] = 0.f; id[t] = 1 << 20; }
      } else {
        cnt = 8;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          sc[2 * t] = acc[t][which * 2] * kk;          id[2 * t] = 8 * t + 2 * q;
          sc[2 * t + 1] = acc[t][which * 2 + 1] * kk;  id[2 * t + 1] = 8 * t + 2 * q + 1;
        }
      }
      unsigned int mask = 0;
      bool near_thr = false;
      float best = -INFINITY;
      int best_id = 1 << 20;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t < cnt && id[t] < L) {
          const float thr = p.thr_logit[id[t]];
          if (passes(sc[t], thr, p.thr_inclusive)) mask |= 1u << id[t];
          near_thr |= fabsf(sc[t] - thr) < p.guard;
          if (sc[t] > best || (sc[t] == best && id[t] < best_id)) { best = sc[t]; best_id = id[t]; }
        }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        mask |= __shfl_xor_sync(0xffffffffu, mask, o);
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_id, o);
        if (ob > best || (ob == best && oi < best_id)) { best = ob; best_id = oi; }
      }
      bool near_top = false;
      const bool rank_matters = (p.argmax != nullptr) || TOPK;
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t < cnt && id[t] < L && id[t] != best_id) near_top |= (best - sc[t]) < p.guard;
      // with top-k > 1 every adjacent gap matters; be conservative: any two labels closer than the guard
      if (TOPK && p.topk > 1) {
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = a + 1; b < 8; ++b)
            if (a < cnt && b < cnt && id[a] < L && id[b] < L) near_top |= fabsf(sc[a] - sc[b]) < p.guard;
        // cross-lane pairs: compare against the three other lanes' scores
#pragma unroll
        for (int o = 1; o <= 3; ++o)
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            const float other = __shfl_xor_sync(0xffffffffu, sc[b], o);
            const int oid = __shfl_xor_sync(0xffffffffu, id[b], o);
#pragma unroll
            for (int a = 0; a < 8; ++a)
              if (a < cnt && id[a] < L && oid < L) near_top |= fabsf(sc[a] - other) < p.guard;
          }
      }
      unsigned int fl = (near_thr && p.mask != nullptr) || (near_top && rank_matters) ? 1u : 0u;
      fl |= __shfl_xor_sync(0xffffffffu, fl, 1);
      fl |= __shfl_xor_sync(0xffffffffu, fl, 2);
      const long long row = which ? row_b : row_a;
      const bool ok = which ? ok_b : ok_a;
      if (fl) flag_bits |= 1u << which;
      const bool emit = ok && !fl;
      if (emit && p.scores) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (t < cnt && id[t] < L) p.scores[row * L + id[t]] = sc[t];
      }
      if (TOPK && p.topk_idx) {                          // warp-uniform branch: shuffles run on all lanes
        // quad-cooperative top-k by repeated arg-max with removal
        const float mx = best;
        float den = 0.f;
        if (p.value_mode == 1) {
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t < cnt && id[t] < L) den += expf(sc[t] - mx);
          den += __shfl_xor_sync(0xffffffffu, den, 1);
          den += __shfl_xor_sync(0xffffffffu, den, 2);
        }
        unsigned int taken = 0;
        for (int k = 0; k < p.topk; ++k) {
          float bv = -INFINITY;
          int bi = 1 << 20;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t < cnt && id[t] < L && !((taken >> id[t]) & 1u) && (sc[t] > bv || (sc[t] == bv && id[t] < bi))) {
              bv = sc[t]; bi = id[t];
            }
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > bv || (ob == bv && oi < bi)) { bv = ob; bi = oi; }
          }
          if (bi >= L) bi = 0;
          taken |= 1u << bi;
          if (emit && q == 0) {
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
      if (emit && q == 0) {
        if (p.argmax) p.argmax[row] = static_cast<uint8_t>(best_id);
        if (p.mask) {
          if (p.mask_is_u32) static_cast<uint32_t*>(p.mask)[row] = mask;
          else static_cast<uint16_t*>(p.mask)[row] = static_cast<uint16_t>(mask);
        }
      }
    }

    // fp64 re-evaluation of flagged rows, one row at a time by the whole warp (rare: ~1-2 % of rows)
    unsigned int rows_flagged = 0;                        // bit i: row blk*16+i
    {
      const unsigned int ba = __ballot_sync(0xffffffffu, (flag_bits & 1u) && q == 0);
      const unsigned int bb = __ballot_sync(0xffffffffu, (flag_bits & 2u) && q == 0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if ((ba >> (4 * i)) & 1u) rows_flagged |= 1u << i;
        if ((bb >> (4 * i)) & 1u) rows_flagged |= 1u << (i + 8);
      }
    }
    while (rows_flagged) {
      const int i = __ffs(rows_flagged) - 1;
      rows_flagged &= rows_flagged - 1;
      const long long row = blk * 16 + i;
      if (row >= p.n) continue;
      // lane owns k in [16*lane, 16*lane+16)
      const uint4* px = reinterpret_cast<const uint4*>(p.x + row * p.ldx) + lane * 2;
      const uint4 x0 = __ldg(px), x1 = __ldg(px + 1);
      const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      double xs[16];
      double ss = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xs[2 * j] = static_cast<double>(bf16_lo(xw[j]));
        xs[2 * j + 1] = static_cast<double>(bf16_hi(xw[j]));
        ss += xs[2 * j] * xs[2 * j] + xs[2 * j + 1] * xs[2 * j + 1];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const double kk = (p.normalize_x ? 1.0 / fmax(sqrt(ss), 1e-12) : 1.0) * static_cast<double>(p.inv_tau);
      double l[ZS_MAXP];
      for (int c = 0; c < p.np; ++c) {
        // P[c][16*lane .. +16) from the fragment table: s = lane/2, 8-element groups 2*(lane&1) and 2*(lane&1)+1
        const int fbase = ((lane >> 1) * 4 + (c >> 3)) * 32 + (c & 7) * 4 + (lane & 1) * 2;
        const uint4 p0 = s_frag[fbase], p1 = s_frag[fbase + 1];
        const uint32_t pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        double d = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d += xs[2 * j] * static_cast<double>(bf16_lo(pw[j])) + xs[2 * j + 1] * static_cast<double>(bf16_hi(pw[j]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        l[c] = d * kk;
      }
      if (lane == 0) {
        double sc[ZS_MAXP];
        for (int lab = 0; lab < L; ++lab) sc[lab] = PAIR ? (l[2 * lab] - l[2 * lab + 1]) : l[lab];
        emit_row_fp64(p, row, sc);
        if (p.guard_count) atomicAdd(p.guard_count, 1ull);
      }
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200clip_zeroshot_score(const void* x_bf16, long long ldx, long long n, const void* prompts_bf16, int np,
                                       int D, int pair_mode, int normalize_x, float temperature,
                                       const float* thr_logit_host, int thr_inclusive, float guard, int topk,
                                       int value_mode, uint8_t* argmax, void* mask, int mask_is_u32, uint8_t* topk_idx,
                                       float* topk_val, float* scores, unsigned long long* guard_count, void* stream) {
  B200_REQUIRE(D == ZS_D, "zeroshot: D=%d unsupported (kernel is built for D=%d)", D, ZS_D);
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
  const long long nblk16 = (n + 15) / 16;
  const int grid = static_cast<int>(std::min<long long>(nblk16, static_cast<long long>(num_sms())));
  static SmemAttrOnce attr[4];
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define B200_ZS_CASE(PAIR, TOPK, IDX)                                                         \
  if ((pair_mode != 0) == PAIR && (topk > 0) == TOPK) {                                       \
    B200_CHECK_CUDA(attr[IDX].ensure(zeroshot_kernel<PAIR, TOPK>, ZS_SMEM_BYTES));            \
    zeroshot_kernel<PAIR, TOPK><<<grid, ZS_THREADS2, ZS_SMEM_BYTES, s>>>(p);                  \
  }
  B200_ZS_CASE(false, false, 0)
  B200_ZS_CASE(false, true, 1)
  B200_ZS_CASE(true, false, 2)
  B200_ZS_CASE(true, true, 3)
#undef B200_ZS_CASE
  B200_LAUNCH_CHECK();
  return B200_OK;
}
