// b200clip: step edges around the head.
//
//  * multilabel metrics  -- calculate_multilabel_metrics (0426/train.py:251-302) and the in-loop accuracy counters of
//    train_epoch / validate (0426/train.py:437-447: per-sample accuracy mean, per-class accuracies).  The reference issues
//    ~20 tiny torch kernels and 7 + C `.item()` host syncs per call; here ONE pass over [B, C] (C <= 32: lane = class)
//    produces every counter and a fixed-order fold writes 7 + C doubles that the host reads with one copy.
//  * prompt-mean pooling -- per-disease mean of the L2-normalised prompt features (get_text_features_with_findings,
//    0426/disease_analysis.py:486-497: F.normalize(text_projector(cls)).mean(0, keepdim), concatenated over diseases); the
//    optional re-normalisation is what a cosine head needs downstream (0426/train.py:966-971 normalises single prompts).
//
// HBM-bound row kernels: coalesced loads (one warp reads one [C] row = 64-128 B; prompt rows with 128-bit loads),
// warp-shuffle/ballot reductions, integer counters wherever the quantity is a count (exact, order-independent).
#include "common.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

constexpr int MET_THREADS = 256;
constexpr int MET_WARPS = MET_THREADS / 32;
// per-warp partial record: [0] sum of per-row (eq/C*100) (double), [1] sum of per-row F1 (double),
// [2] exact-match rows, [3] top-1 hit seen (0/1), [4] top-3 hit rows, [5] total eq count, [8..40) per-class eq counts
constexpr int MET_REC = 40;

__global__ void __launch_bounds__(MET_THREADS) multilabel_metrics_kernel(const float* __restrict__ pred, long long ldp,
                                                                         const float* __restrict__ labels, long long ldl,
                                                                         long long B, int C, float threshold,
                                                                         double* __restrict__ partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gw = static_cast<long long>(blockIdx.x) * MET_WARPS + warp;
  const long long nw = static_cast<long long>(gridDim.x) * MET_WARPS;
  const bool act = lane < C;
  const unsigned cmask = C >= 32 ? 0xffffffffu : ((1u << C) - 1u);
  double s_acc = 0.0, s_f1 = 0.0;
  unsigned long long n_exact = 0, n_top3 = 0, n_eq = 0, cls_eq = 0;
  unsigned any_top1 = 0;
  const int k3 = C < 3 ? C : 3;
  for (long long row = gw; row < B; row += nw) {
    const float p = act ? pred[row * ldp + lane] : -INFINITY;
    const float y = act ? labels[row * ldl + lane] : 0.f;
    const float pl = (p > threshold) ? 1.f : 0.f;                 // (predictions > 0.5).float()   :261
    const bool eq = act && (pl == y);
    const unsigned eqb = __ballot_sync(0xffffffffu, eq) & cmask;
    const int neq = __popc(eqb);
    s_acc += static_cast<double>((static_cast<float>(neq) / static_cast<float>(C)) * 100.f);   // :264 per-row mean * 100
    n_eq += neq;
    cls_eq += eq ? 1u : 0u;
    n_exact += (neq == C) ? 1u : 0u;
    // top-k by repeated arg-max (first index wins ties, like torch.argmax / topk on equal values)
    unsigned taken = 0;
    bool hit3 = false;
    for (int k = 0; k < k3; ++k) {
      float v = (act && !((taken >> lane) & 1u)) ? p : -INFINITY;
      const float mx = warp_max(v);
      unsigned cand = __ballot_sync(0xffffffffu, act && !((taken >> lane) & 1u) && (v == mx || (mx != mx)));
      if (cand == 0) cand = __ballot_sync(0xffffffffu, act && !((taken >> lane) & 1u));     // all-NaN row: take the first free class
      const int best = __ffs(cand) - 1;
      taken |= 1u << best;
      const float yb = __shfl_sync(0xffffffffu, y, best);
      const bool hit = (yb == 1.f);
      if (k == 0 && hit) any_top1 = 1u;                           // :277  torch.any(..., dim=0) over the batch
      hit3 |= hit;
    }
    n_top3 += hit3 ? 1u : 0u;
    // sample-level F1 in fp32, term by term as :283-289
    const float pred_pos = static_cast<float>(__popc(__ballot_sync(0xffffffffu, act && pl == 1.f)));
    const float true_pos = warp_sum(act ? y : 0.f);
    const float correct = warp_sum(act ? pl * y : 0.f);
    const float precision = correct / (pred_pos + 1e-8f);
    const float recall = correct / (true_pos + 1e-8f);
    const float f1 = 2.f * precision * recall / (precision + recall + 1e-8f);
    s_f1 += static_cast<double>(f1);
  }
  double* rec = partial + gw * MET_REC;
  if (lane == 0) {
    rec[0] = s_acc; rec[1] = s_f1; rec[2] = static_cast<double>(n_exact); rec[3] = static_cast<double>(any_top1);
    rec[4] = static_cast<double>(n_top3); rec[5] = static_cast<double>(n_eq); rec[6] = 0.0; rec[7] = 0.0;
  }
  rec[8 + lane] = static_cast<double>(cls_eq);
}

// one block: fixed-order fold of the per-warp records -> out[0..7) metrics in the reference's dict order
// (sample_acc, label_acc, hamming_score, exact_match, top1_acc, top3_acc, f1_score), out[7..7+C) per-class accuracy (%)
__global__ void __launch_bounds__(64) multilabel_metrics_finalize_kernel(const double* __restrict__ partial, long long nrec,
                                                                         long long B, int C, double* __restrict__ out) {
  __shared__ double s[MET_REC];
  const int t = threadIdx.x;
  if (t < MET_REC) {
    double a = 0.0;
    for (long long r = 0; r < nrec; ++r) {
      const double v = partial[r * MET_REC + t];
      a = (t == 3) ? fmax(a, v) : a + v;
    }
    s[t] = a;
  }
  __syncthreads();
  if (t == 0) {
    const double b = static_cast<double>(B);
    double lab = 0.0;
    for (int c = 0; c < C; ++c) {
      const double acc = s[8 + c] / b * 100.0;
      out[7 + c] = acc;
      lab += acc;
    }
    out[0] = s[0] / b;
    out[1] = lab / C;
    out[2] = s[5] / (b * C) * 100.0;
    out[3] = s[2] / b * 100.0;
    out[4] = s[3] * 100.0;
    out[5] = s[4] / b * 100.0;
    out[6] = s[1] / b * 100.0;
  }
}

// out[d, :] = mean over the prompts p of disease d of normalize(feats[p, :]) (optionally re-normalised); the prompts of
// disease d are rows offsets[d] .. offsets[d+1]-1 (ragged: prompts.get(disease, [default]) gives 1..P rows).  One CTA per
// disease, prompts in index order (fixed summation order: deterministic).
__global__ void __launch_bounds__(256) prompt_mean_pool_kernel(const float* __restrict__ feats, const int* __restrict__ offsets,
                                                               int D, float eps, int renormalize, float* __restrict__ out) {
  extern __shared__ float acc[];                       // [D]
  __shared__ float red[8];
  const int d = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p0 = offsets[d], p1 = offsets[d + 1];
  for (int i = threadIdx.x; i < D; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  for (int pidx = p0; pidx < p1; ++pidx) {
    const float* row = feats + static_cast<long long>(pidx) * D;
    float ss = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { const float v = row[i]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float inv = 1.f / fmaxf(sqrtf(tot), eps);
    for (int i = threadIdx.x; i < D; i += blockDim.x) acc[i] += row[i] * inv;
    __syncthreads();
  }
  const float invp = p1 > p0 ? 1.f / static_cast<float>(p1 - p0) : 0.f;
  float ss = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) { const float v = acc[i] * invp; acc[i] = v; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float inv = renormalize ? 1.f / fmaxf(sqrtf(tot), eps) : 1.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) out[static_cast<long long>(d) * D + i] = acc[i] * inv;
}

static int metrics_grid(long long B) {
  const long long want = (B + MET_WARPS - 1) / MET_WARPS;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(want, 2ll * num_sms())));
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200clip_multilabel_metrics_workspace_bytes(long long B) {
  if (B <= 0) return 0;
  return static_cast<size_t>(metrics_grid(B)) * MET_WARPS * MET_REC * sizeof(double);
}

extern "C" int b200clip_multilabel_metrics(const float* predictions, long long ld_pred, const float* labels,
                                           long long ld_labels, long long B, int C, float threshold, double* out,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(B > 0 && C >= 1 && C <= 32, "multilabel_metrics: need B > 0 and 1 <= C <= 32 (got %lld, %d)", B, C);
  B200_REQUIRE(predictions && labels && out && ld_pred >= C && ld_labels >= C, "multilabel_metrics: bad pointers / strides");
  const size_t need = b200clip_multilabel_metrics_workspace_bytes(B);
  if (workspace_bytes < need || !workspace) return fail(B200_ERR_WORKSPACE, "multilabel_metrics: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = metrics_grid(B);
  double* partial = static_cast<double*>(workspace);
  multilabel_metrics_kernel<<<grid, MET_THREADS, 0, s>>>(predictions, ld_pred, labels, ld_labels, B, C, threshold, partial);
  B200_LAUNCH_CHECK();
  multilabel_metrics_finalize_kernel<<<1, 64, 0, s>>>(partial, static_cast<long long>(grid) * MET_WARPS, B, C, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int b200clip_prompt_mean_pool(const float* prompt_features, const int* offsets, int num_diseases, int D, float eps,
                                         int renormalize, float* out, void* stream) {
  B200_REQUIRE(prompt_features && offsets && out && num_diseases > 0 && D > 0 && D <= 8192, "prompt_mean_pool: bad arguments");
  prompt_mean_pool_kernel<<<num_diseases, 256, D * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      prompt_features, offsets, D, eps, renormalize, out);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
