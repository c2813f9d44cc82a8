// b200clip: zero-shot post-processing on the device (multimodal_attention/zero_shot_predict.py:66-213).
// The reference runs these steps as python loops over samples with .cpu() round trips:
//   (1) dynamic per-label thresholds (:112-159): positive / negative score statistics per label, a 20-point np.linspace grid
//       between max(0.1, neg_mean - neg_std) and min(0.9, pos_mean + pos_std), binary F1 (sklearn, zero_division = 0) per grid
//       point, first best wins; 0.8 / 0.2 when a label has no positives / negatives;
//   (2) weighted two-view merge (:183-213) of the per-view prediction lists of predict_zero_shot (disease_analysis.py:361-413).
// Here: (1) = one statistics pass (double accumulators, fixed-order block partials) + one histogram pass (for every score the
// number of grid points it reaches; integer shared-memory atomics, so the counts are exact and order-independent) + a tiny
// finalisation; (2) = one thread per sample.  Scores are float32 sigmoid outputs; statistics and the merged comparison run in
// double like numpy / python floats, the per-view threshold test in float32 like the torch comparison it restates.
// Bytes: (1) reads scores + labels twice (2 x 8 B per entry), (2) reads 8 B per (sample, label) and writes 5 B.
#include <algorithm>

#include "common.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

constexpr int ZP_MAXL = 32;
constexpr int ZP_GRID = 20;                                  // np.linspace(lo, hi, 20)
constexpr int ZP_THREADS = 256;

// per-label moments: [L][6] = n_pos, sum_pos, sumsq_pos, n_neg, sum_neg, sumsq_neg
__global__ void __launch_bounds__(ZP_THREADS) zp_moments_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                                                                long long N, int L, double* __restrict__ partial /*[grid][L][6]*/) {
  __shared__ double red[ZP_THREADS / 32][ZP_MAXL][6];
  const int l = threadIdx.x & 31, sub = threadIdx.x >> 5;
  double a[6] = {0, 0, 0, 0, 0, 0};
  if (l < L)
    for (long long r = blockIdx.x * (ZP_THREADS / 32) + sub; r < N; r += static_cast<long long>(gridDim.x) * (ZP_THREADS / 32)) {
      const double s = static_cast<double>(scores[r * L + l]);
      const int o = labels[r * L + l] == 1.0f ? 0 : (labels[r * L + l] == 0.0f ? 3 : -1);   // scores[labels == 1] / [labels == 0]
      if (o >= 0) { a[o] += 1.0; a[o + 1] += s; a[o + 2] += s * s; }
    }
  for (int k = 0; k < 6; ++k) red[sub][l][k] = a[k];
  __syncthreads();
  for (int i = threadIdx.x; i < L * 6; i += ZP_THREADS) {
    const int ll = i / 6, k = i - ll * 6;
    double t = 0.0;
    for (int w = 0; w < ZP_THREADS / 32; ++w) t += red[w][ll][k];
    partial[(static_cast<long long>(blockIdx.x) * L + ll) * 6 + k] = t;
  }
}

// fold the block partials (fixed order) and lay out the threshold grid per label; flag[l]: 0 search, 1 no positives, 2 no negatives
__global__ void zp_grid_kernel(const double* __restrict__ partial, int nparts, int L, double* __restrict__ grid /*[L][20]*/,
                               int* __restrict__ flag, double* __restrict__ npos_out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  double m[6] = {0, 0, 0, 0, 0, 0};
  for (int q = 0; q < nparts; ++q)
    for (int k = 0; k < 6; ++k) m[k] += partial[(static_cast<long long>(q) * L + l) * 6 + k];
  npos_out[l] = m[0];
  if (m[0] == 0.0) { flag[l] = 1; return; }                  // :121-124
  if (m[3] == 0.0) { flag[l] = 2; return; }                  // :127-130
  flag[l] = 0;
  const double pos_mean = m[1] / m[0], neg_mean = m[4] / m[3];
  const double pos_std = sqrt(fmax(m[2] / m[0] - pos_mean * pos_mean, 0.0));      // np.std: population standard deviation
  const double neg_std = sqrt(fmax(m[5] / m[3] - neg_mean * neg_mean, 0.0));
  const double lo = fmax(0.1, neg_mean - neg_std), hi = fmin(0.9, pos_mean + pos_std);   // :143-144
  const double step = (hi - lo) / (ZP_GRID - 1);             // np.linspace: arange(num) * step + start, last point = stop
  for (int t = 0; t < ZP_GRID; ++t) grid[l * ZP_GRID + t] = (t == ZP_GRID - 1) ? hi : t * step + lo;
}

// hist[c][l][k]: number of class-c entries of label l whose score reaches exactly k grid points (grid ascending or not: the
// count is taken point by point, as (scores >= threshold) is in the reference)
__global__ void __launch_bounds__(ZP_THREADS) zp_hist_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                                                             long long N, int L, const double* __restrict__ grid,
                                                             const int* __restrict__ flag, unsigned int* __restrict__ hist /*[2][L][20]*/) {
  __shared__ double s_grid[ZP_MAXL][ZP_GRID];
  __shared__ unsigned int s_hist[2][ZP_MAXL][ZP_GRID];       // [class][label][grid point]: predictions == 1 at that point
  for (int i = threadIdx.x; i < L * ZP_GRID; i += ZP_THREADS) s_grid[i / ZP_GRID][i % ZP_GRID] = grid[i];
  for (int i = threadIdx.x; i < 2 * ZP_MAXL * ZP_GRID; i += ZP_THREADS) (&s_hist[0][0][0])[i] = 0u;
  __syncthreads();
  const int l = threadIdx.x & 31, sub = threadIdx.x >> 5;
  if (l < L && flag[l] == 0)
    for (long long r = blockIdx.x * (ZP_THREADS / 32) + sub; r < N; r += static_cast<long long>(gridDim.x) * (ZP_THREADS / 32)) {
      const double s = static_cast<double>(scores[r * L + l]);
      const float y = labels[r * L + l];
      if (y != 1.0f && y != 0.0f) continue;                  // labels are {0, 1}
      const int c = y == 1.0f ? 0 : 1;
      for (int t = 0; t < ZP_GRID; ++t)
        if (s >= s_grid[l][t]) atomicAdd(&s_hist[c][l][t], 1u);      // :147
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * ZP_MAXL * ZP_GRID; i += ZP_THREADS) {
    const unsigned int v = (&s_hist[0][0][0])[i];
    const int c = i / (ZP_MAXL * ZP_GRID), rem = i - c * ZP_MAXL * ZP_GRID, ll = rem / ZP_GRID, t = rem - ll * ZP_GRID;
    if (v && ll < L) atomicAdd(&hist[(c * L + ll) * ZP_GRID + t], v);
  }
}

__global__ void zp_pick_kernel(const unsigned int* __restrict__ hist, const double* __restrict__ grid, const int* __restrict__ flag,
                               const double* __restrict__ npos, int L, double* __restrict__ thresholds, double* __restrict__ best_f1_out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  if (flag[l] == 1) { thresholds[l] = 0.8; if (best_f1_out) best_f1_out[l] = 0.0; return; }
  if (flag[l] == 2) { thresholds[l] = 0.2; if (best_f1_out) best_f1_out[l] = 0.0; return; }
  double best_f1 = 0.0, best_thr = 0.5;                      // :139-140
  for (int t = 0; t < ZP_GRID; ++t) {
    const double tp = hist[(0 * L + l) * ZP_GRID + t], fp = hist[(1 * L + l) * ZP_GRID + t], fn = npos[l] - tp;
    const double den = 2.0 * tp + fp + fn;
    const double f1 = den == 0.0 ? 0.0 : 2.0 * tp / den;     // sklearn f1_score, zero_division = 0
    if (f1 > best_f1) { best_f1 = f1; best_thr = grid[l * ZP_GRID + t]; }      // :149-151
  }
  thresholds[l] = best_thr;
  if (best_f1_out) best_f1_out[l] = best_f1;
}

// one thread per sample: per-view lists (disease_analysis.py:372-410 with per-label thresholds and top_k = None), weighted
// maximum, per-label filter, fall back to the single best label (:183-213)
__global__ void __launch_bounds__(ZP_THREADS) zp_merge_kernel(const float* __restrict__ prob /*[N][2][L]*/, const double* __restrict__ thr,
                                                              long long N, int L, double w0, double w1, uint8_t* __restrict__ pred,
                                                              float* __restrict__ merged) {
  __shared__ double s_thr[ZP_MAXL];
  __shared__ float s_thr32[ZP_MAXL];
  if (threadIdx.x < L) { s_thr[threadIdx.x] = thr[threadIdx.x]; s_thr32[threadIdx.x] = static_cast<float>(thr[threadIdx.x]); }
  __syncthreads();
  const long long i = blockIdx.x * static_cast<long long>(ZP_THREADS) + threadIdx.x;
  if (i >= N) return;
  const float* p0 = prob + i * 2 * L;
  const float* p1 = p0 + L;
  double ds[ZP_MAXL];                                         // weighted score per label, < 0: not in the dict
  int order_key[ZP_MAXL];                                     // insertion position (ties in the fallback max: first inserted wins)
  int n_ins = 0;
#pragma unroll
  for (int l = 0; l < ZP_MAXL; ++l) { ds[l] = -1.0; order_key[l] = 1 << 30; }
  for (int v = 0; v < 2; ++v) {
    const float* p = v ? p1 : p0;
    const double w = v ? w1 : w0;                             // :190
    int npass = 0, top = 0;
    for (int l = 0; l < L; ++l) {
      if (p[l] >= s_thr32[l]) ++npass;                        // float32 comparison, as the torch expression (:377)
      if (p[l] > p[top]) top = l;                             // torch.topk(1): largest value, lowest index on ties
    }
    for (int l = 0; l < L; ++l) {
      const bool in_list = npass > 0 ? (p[l] >= s_thr32[l]) : (l == top);      // :393-410 with top_k = None: the single best
      if (!in_list) continue;
      const double s = static_cast<double>(p[l]) * w;
      if (ds[l] < 0.0) { ds[l] = 0.0; order_key[l] = n_ins++; }                // :192-193
      ds[l] = fmax(ds[l], s);                                                  // :194
    }
  }
  int kept = 0, best = -1;
  for (int l = 0; l < L; ++l) {
    const bool in = ds[l] >= 0.0;
    const bool keep = in && ds[l] >= s_thr[l];                                 // :199-202 (python floats: double)
    kept += keep;
    if (in && (best < 0 || ds[l] > ds[best] || (ds[l] == ds[best] && order_key[l] < order_key[best]))) best = l;
  }
  for (int l = 0; l < L; ++l) {
    const bool in = ds[l] >= 0.0;
    const bool on = kept > 0 ? (in && ds[l] >= s_thr[l]) : (l == best);        // :205-208
    pred[i * L + l] = on ? 1 : 0;
    if (merged) merged[i * L + l] = on ? static_cast<float>(ds[l]) : 0.f;
  }
}

static int zp_grid_blocks(long long N) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((N + 7) / 8, 2LL * num_sms())));
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200clip_zs_thresholds_workspace_bytes(long long N, int L) {
  const size_t part = static_cast<size_t>(zp_grid_blocks(N)) * L * 6 * sizeof(double);
  return part + static_cast<size_t>(L) * ZP_GRID * sizeof(double) + static_cast<size_t>(L) * (sizeof(int) + sizeof(double)) +
         2ull * L * ZP_GRID * sizeof(unsigned int) + 1024;
}

// scores [N, L] f32 (per-sample maximum over the views of the sigmoid scores), labels [N, L] f32 in {0, 1}
extern "C" int b200clip_zs_dynamic_thresholds(const float* scores, const float* labels, long long N, int L, double* thresholds,
                                              double* best_f1, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(N >= 0 && L > 0 && L <= ZP_MAXL && thresholds && (N == 0 || (scores && labels)), "zs_dynamic_thresholds: bad arguments (L <= %d)", ZP_MAXL);
  if (workspace_bytes < b200clip_zs_thresholds_workspace_bytes(N, L)) return fail(B200_ERR_WORKSPACE, "zs_dynamic_thresholds: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nb = zp_grid_blocks(N);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  double* partial = reinterpret_cast<double*>(ws); ws += static_cast<size_t>(nb) * L * 6 * sizeof(double);
  double* grid = reinterpret_cast<double*>(ws); ws += static_cast<size_t>(L) * ZP_GRID * sizeof(double);
  double* npos = reinterpret_cast<double*>(ws); ws += static_cast<size_t>(L) * sizeof(double);
  unsigned int* hist = reinterpret_cast<unsigned int*>(ws); ws += 2ull * L * ZP_GRID * sizeof(unsigned int);
  int* flag = reinterpret_cast<int*>(ws);
  B200_CHECK_CUDA(cudaMemsetAsync(hist, 0, 2ull * L * ZP_GRID * sizeof(unsigned int), s));
  zp_moments_kernel<<<nb, ZP_THREADS, 0, s>>>(scores, labels, N, L, partial);
  B200_LAUNCH_CHECK();
  zp_grid_kernel<<<1, 32, 0, s>>>(partial, nb, L, grid, flag, npos);
  B200_LAUNCH_CHECK();
  zp_hist_kernel<<<nb, ZP_THREADS, 0, s>>>(scores, labels, N, L, grid, flag, hist);
  B200_LAUNCH_CHECK();
  zp_pick_kernel<<<1, 32, 0, s>>>(hist, grid, flag, npos, L, thresholds, best_f1);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// prob_views [N, 2, L] f32 sigmoid scores of the two views; pred [N, L] u8; merged [N, L] f32 (weighted score of kept labels) or null
extern "C" int b200clip_zs_merge_views(const float* prob_views, const double* thresholds, long long N, int L, double w0, double w1,
                                       uint8_t* pred, float* merged, void* stream) {
  B200_REQUIRE(N >= 0 && L > 0 && L <= ZP_MAXL && thresholds && pred && (N == 0 || prob_views), "zs_merge_views: bad arguments (L <= %d)", ZP_MAXL);
  if (N == 0) return B200_OK;
  zp_merge_kernel<<<static_cast<int>((N + ZP_THREADS - 1) / ZP_THREADS), ZP_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      prob_views, thresholds, N, L, w0, w1, pred, merged);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
