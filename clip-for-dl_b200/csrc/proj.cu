// b200clip: the projection block of the CLIP head (a-P1 / a-P2, reference 0426/train.py:73-116)
//   p = x W1^T + b1 ; h = GELU_erf(p) ; f = h W2^T + b2 ; z = f + p ; y = LayerNorm(z) [; yhat = y / ||y||]
// as a short chain of tcgen05 GEMMs with fused epilogues (gemm.cuh) and the row kernels (rowops.cu).
// Forward : GEMM1 (+b1, GELU -> p,h bf16)  GEMM2 (+b2 +p -> z f32)  LayerNorm(+L2-norm)
// Backward: LN-bwd -> dz ; dW2 = dz^T h (split-K) ; dp = (dz W2) * gelu'(p) + dz (fused epilogue) ;
//           dW1 = dp^T x (split-K) ; dx = dp W1 ; bias grads = column sums.
// Dropout (p=0.1 in train mode, 0426/train.py:93): a counter-based keep-mask hash(seed, row, col) applied in the second GEMM's
// epilogue and regenerated in the LayerNorm-backward kernel (no mask is stored); drop_p = 0 in eval mode.
#include "gemm.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {
int gemm_bf16(const void* a, const void* b, int a_mn, int b_mn, int M, int N, int K, long long lda, long long ldb,
              int epi, float alpha, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
              const void* resid, long long ld_res, const float* aux, long long ld_aux, int split_k, cudaStream_t stream,
              float drop_p = 0.f, unsigned int drop_seed = 0u, int aux_is_bf16 = 0, const unsigned int* drop_seed_dev = nullptr,
              float* splitk_ws = nullptr, size_t splitk_ws_bytes = 0);
}
using namespace b200;

static int split_for(int M, int N, int K) {
  const int tiles = ((M + 127) / 128) * ((N + 255) / 256);
  const int kchunks = (K + 63) / 64;
  int s = num_sms() / tiles;                       // floor: tiles * s <= #SMs, i.e. ONE wave of the persistent GEMM
  if (s > kchunks) s = kchunks;
  return s < 1 ? 1 : s;
}

extern "C" int b200clip_proj_fwd(const void* x_bf16, long long B, int E, int D, const void* w1_bf16, const float* b1,
                                 const void* w2_bf16, const float* b2, const float* gamma, const float* beta,
                                 float ln_eps, float drop_p, unsigned int drop_seed, const unsigned int* drop_seed_dev,
                                 void* p_bf16, void* h_bf16,
                                 float* z_f32, float* y_f32, void* yhat_bf16, float* mean, float* rstd, float* inv_norm,
                                 void* stream) {
  B200_REQUIRE(B > 0 && E > 0 && D > 0, "proj_fwd: empty problem");
  B200_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "proj_fwd: dropout probability must be in [0,1)");
  B200_REQUIRE(E % 8 == 0 && D % 128 == 0 && D <= 1024, "proj_fwd: need E %% 8 == 0 and D %% 128 == 0, D <= 1024 (E=%d D=%d)", E, D);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = gemm_bf16(x_bf16, w1_bf16, 0, 0, (int)B, D, E, E, E, EPI_BIAS_GELU, 1.0f, p_bf16, D, h_bf16, D, b1, nullptr, 0,
                     nullptr, 0, 1, s);
  if (rc) return rc;
  rc = gemm_bf16(h_bf16, w2_bf16, 0, 0, (int)B, D, D, D, D, EPI_BIAS_RESID_F32, 1.0f, z_f32, D, nullptr, 0, b2, p_bf16, D,
                 nullptr, 0, 1, s, drop_p, drop_seed, 0, drop_seed_dev);
  if (rc) return rc;
  return b200clip_layernorm_fwd(z_f32, gamma, beta, y_f32, yhat_bf16, mean, rstd, inv_norm, B, D, ln_eps, 1e-12f, stream);
}

// fp32 partial tiles of the deterministic split-K weight-gradient GEMMs.  split_for keeps tiles * splits <= #SMs and a
// tile is at most 128 x 256, so splits * M * N * 4 <= #SMs * 128 KB (19.4 MB) whatever the shape.
static size_t splitk_ws_bytes() { return static_cast<size_t>(num_sms()) * 128 * 256 * sizeof(float); }

extern "C" size_t b200clip_proj_bwd_workspace_bytes(long long B, int E, int D) {
  size_t n = 0;
  n += (splitk_ws_bytes() + 255) & ~size_t(255);
  n += static_cast<size_t>(B) * D * 4;                      // dz f32
  n += static_cast<size_t>(B) * D * 2;                      // dz bf16
  n += static_cast<size_t>(B) * D * 2;                      // dp bf16
  n += b200clip_layernorm_bwd_workspace_bytes(B, D);
  n += b200clip_colsum_workspace_bytes(B, D);
  return n + 1024;
}

// dy == nullptr selects the fused entry: the gradient arrives w.r.t. the L2-normalised output (dyhat, possibly as partial
// sums) and the L2-norm backward runs inside the LayerNorm-backward kernel (b200clip_layernorm_l2_bwd).
extern "C" int b200clip_proj_bwd(const float* dy, const float* dyhat, int dyhat_partials, const void* yhat_bf16,
                                 const float* inv_norm, const float* addend, const float* addend_scale,
                                 const void* x_bf16, long long B, int E, int D, const void* w1_bf16,
                                 const void* w2_bf16, const float* gamma, const void* p_bf16, const void* h_bf16,
                                 const float* z_f32, const float* mean, const float* rstd, float drop_p,
                                 unsigned int drop_seed, const unsigned int* drop_seed_dev, float* dx_f32, void* dx_bf16,
                                 float* dw1, float* db1, float* dw2, float* db2, float* dgamma, float* dbeta, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  B200_REQUIRE(B > 0 && E % 8 == 0 && D % 128 == 0 && D <= 1024, "proj_bwd: bad shape B=%lld E=%d D=%d", B, E, D);
  if (workspace_bytes < b200clip_proj_bwd_workspace_bytes(B, E, D)) return fail(B200_ERR_WORKSPACE, "proj_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto carve = [&](size_t bytes) { uint8_t* p = ws; ws += (bytes + 255) & ~size_t(255); return p; };
  float* dz = reinterpret_cast<float*>(carve(static_cast<size_t>(B) * D * 4));
  void* dz_bf = carve(static_cast<size_t>(B) * D * 2);
  void* dp_bf = carve(static_cast<size_t>(B) * D * 2);
  const size_t ln_ws = b200clip_layernorm_bwd_workspace_bytes(B, D);
  void* ln_work = carve(ln_ws);
  const size_t cs_ws = b200clip_colsum_workspace_bytes(B, D);
  void* cs_work = carve(cs_ws);
  const size_t sk_ws = splitk_ws_bytes();
  float* sk_work = reinterpret_cast<float*>(carve(sk_ws));

  // Without dropout the fc branch and the residual branch see the same dz: the f32 copy (67 MB written + read at B = 32768)
  // is skipped and the residual add in the GELU-backward epilogue reads the bf16 copy the GEMMs use anyway.
  const bool need_f32_dz = drop_p > 0.f;
  B200_REQUIRE(dy != nullptr || (dyhat != nullptr && yhat_bf16 != nullptr && inv_norm != nullptr), "proj_bwd: need dy, or dyhat + yhat + inv_norm");
  int rc = dy ? b200clip_layernorm_bwd(dy, z_f32, mean, rstd, gamma, need_f32_dz ? dz : nullptr, dz_bf, dgamma, dbeta, db2, 0, B, D,
                                       drop_p, drop_seed, drop_seed_dev, ln_work, ln_ws, stream)
              : b200clip_layernorm_l2_bwd(dyhat, dyhat_partials, yhat_bf16, inv_norm, 1e-12f, addend, addend_scale, z_f32, mean,
                                          rstd, gamma, need_f32_dz ? dz : nullptr, dz_bf, dgamma, dbeta, db2, 0, B, D, drop_p,
                                          drop_seed, drop_seed_dev, ln_work, ln_ws, stream);
  if (rc) return rc;
  // Two chains from here: the data chain dz -> dp -> dx on the caller's stream, the parameter-gradient chain (dW2, db1, dW1)
  // on a fork lane.  At per-rank batches of a few thousand rows every kernel here is a fraction of a wave, so the chains run
  // side by side (B = 4096: 105 -> 65 us per side); at large batches they simply queue.  B200CLIP_PROJ_BWD_FORK=0 disables.
  static const bool fork_enabled = [] { const char* e = getenv("B200CLIP_PROJ_BWD_FORK"); return !(e && atoi(e) == 0); }();
  ForkLane* lane = fork_enabled ? acquire_fork_lane() : nullptr;
  cudaStream_t sa = lane ? lane->aux : s;
  if (lane) B200_CHECK_CUDA(lane->link(s, sa, 0));                                   // aux sees dz
  // dW2[o][j] = sum_b dz[b][o] h[b][j]
  if ((rc = gemm_bf16(dz_bf, h_bf16, 1, 1, D, D, (int)B, D, D, EPI_ATOMIC_F32, 1.0f, dw2, D, nullptr, 0, nullptr, nullptr, 0,
                      nullptr, 0, split_for(D, D, (int)B), sa, 0.f, 0u, 0, nullptr, sk_work, sk_ws)))
    return rc;
  // dp = (dz W2) * gelu'(p) + dz
  if ((rc = gemm_bf16(dz_bf, w2_bf16, 0, 1, (int)B, D, D, D, D, EPI_GELU_BWD, 1.0f, dp_bf, D, nullptr, 0, nullptr, p_bf16, D,
                      need_f32_dz ? dz : reinterpret_cast<const float*>(dz_bf), D, 1, s, 0.f, 0u, need_f32_dz ? 0 : 1)))
    return rc;
  if (lane) B200_CHECK_CUDA(lane->link(s, sa, 1));                                   // aux sees dp
  if ((rc = b200clip_colsum(dp_bf, 1, D, B, D, db1, 0, cs_work, cs_ws, sa))) return rc;
  // dW1[o][e] = sum_b dp[b][o] x[b][e]
  if ((rc = gemm_bf16(dp_bf, x_bf16, 1, 1, D, E, (int)B, D, E, EPI_ATOMIC_F32, 1.0f, dw1, E, nullptr, 0, nullptr, nullptr, 0,
                      nullptr, 0, split_for(D, E, (int)B), sa, 0.f, 0u, 0, nullptr, sk_work, sk_ws)))
    return rc;
  if (dx_f32 || dx_bf16) {                         // input gradient in the caller's dtype (bf16 inputs get bf16 grads directly)
    B200_REQUIRE(E % 32 == 0, "proj_bwd: dx needs E %% 32 == 0");
    if ((rc = gemm_bf16(dp_bf, w1_bf16, 0, 1, (int)B, E, D, D, E, dx_f32 ? EPI_STORE_F32 : EPI_STORE_BF16, 1.0f,
                        dx_f32 ? static_cast<void*>(dx_f32) : dx_bf16, E, nullptr, 0, nullptr, nullptr, 0, nullptr, 0, 1, s)))
      return rc;
  }
  if (lane) B200_CHECK_CUDA(lane->link(sa, s, 2));                                   // join
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// MultiViewFusion (0426/train.py:988-1000): cat[frontal, lateral] -> Linear(2D, D) -> ReLU -> Dropout(0.2)
// -> Linear(D, D).  x_bf16 = the concatenated views [B, 2D]; h = dropout(relu(x W0^T + b0)) is saved for backward (the masked,
// rescaled activations: h > 0 <=> the unit was positive AND kept).
// ------------------------------------------------------------------------------------------------------------------
extern "C" int b200clip_fusion_fwd(const void* x_bf16, long long B, int D, const void* w0_bf16, const float* b0,
                                   const void* w3_bf16, const float* b3, float drop_p, unsigned int drop_seed, void* h_bf16,
                                   float* y_f32, void* stream) {
  B200_REQUIRE(B > 0 && D > 0 && D % 128 == 0 && D <= 1024, "fusion_fwd: bad shape B=%lld D=%d", B, D);
  B200_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "fusion_fwd: dropout probability must be in [0,1)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = gemm_bf16(x_bf16, w0_bf16, 0, 0, (int)B, D, 2 * D, 2 * D, 2 * D, EPI_RELU_BF16, 1.0f, h_bf16, D, nullptr, 0, b0, nullptr, 0,
                     nullptr, 0, 1, s, drop_p, drop_seed);
  if (rc) return rc;
  return gemm_bf16(h_bf16, w3_bf16, 0, 0, (int)B, D, D, D, D, EPI_STORE_F32, 1.0f, y_f32, D, nullptr, 0, b3, nullptr, 0, nullptr, 0, 1, s);
}

extern "C" size_t b200clip_fusion_bwd_workspace_bytes(long long B, int D) {
  size_t n = 0;
  n += 2 * (((static_cast<size_t>(B) * D * 2) + 255) & ~size_t(255));      // dy bf16, dh bf16
  n += 2 * ((b200clip_colsum_workspace_bytes(B, D) + 255) & ~size_t(255));
  n += (splitk_ws_bytes() + 255) & ~size_t(255);
  return n + 1024;
}

extern "C" int b200clip_fusion_bwd(const float* dy, const void* x_bf16, long long B, int D, const void* w0_bf16,
                                   const void* w3_bf16, const void* h_bf16, float drop_p, float* dx_f32, float* dw0, float* db0,
                                   float* dw3, float* db3, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(B > 0 && D > 0 && D % 128 == 0 && D <= 1024, "fusion_bwd: bad shape B=%lld D=%d", B, D);
  if (workspace_bytes < b200clip_fusion_bwd_workspace_bytes(B, D)) return fail(B200_ERR_WORKSPACE, "fusion_bwd: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto carve = [&](size_t bytes) { uint8_t* p = ws; ws += (bytes + 255) & ~size_t(255); return p; };
  void* dy_bf = carve(static_cast<size_t>(B) * D * 2);
  void* dh_bf = carve(static_cast<size_t>(B) * D * 2);
  const size_t cs_ws = b200clip_colsum_workspace_bytes(B, D);
  void* cs0 = carve(cs_ws);
  void* cs1 = carve(cs_ws);
  const size_t sk_ws = splitk_ws_bytes();
  float* sk_work = reinterpret_cast<float*>(carve(sk_ws));
  int rc;
  if ((rc = b200clip_cast_f32_bf16(dy, dy_bf, B * D, stream))) return rc;
  if ((rc = b200clip_colsum(dy, 0, D, B, D, db3, 0, cs0, cs_ws, stream))) return rc;           // db3 = column sums of dy
  // dW3[o][j] = sum_b dy[b][o] h[b][j]
  if ((rc = gemm_bf16(dy_bf, h_bf16, 1, 1, D, D, (int)B, D, D, EPI_ATOMIC_F32, 1.0f, dw3, D, nullptr, 0, nullptr, nullptr, 0,
                      nullptr, 0, split_for(D, D, (int)B), s, 0.f, 0u, 0, nullptr, sk_work, sk_ws)))
    return rc;
  // dh = (dy W3) / (1 - p) where the unit was positive and kept
  if ((rc = gemm_bf16(dy_bf, w3_bf16, 0, 1, (int)B, D, D, D, D, EPI_RELU_BWD, 1.0f / (1.0f - drop_p), dh_bf, D, nullptr, 0, nullptr,
                      h_bf16, D, nullptr, 0, 1, s)))
    return rc;
  if ((rc = b200clip_colsum(dh_bf, 1, D, B, D, db0, 0, cs1, cs_ws, stream))) return rc;
  // dW0[o][e] = sum_b dh[b][o] x[b][e]
  if ((rc = gemm_bf16(dh_bf, x_bf16, 1, 1, D, 2 * D, (int)B, D, 2 * D, EPI_ATOMIC_F32, 1.0f, dw0, 2 * D, nullptr, 0, nullptr, nullptr, 0,
                      nullptr, 0, split_for(D, 2 * D, (int)B), s, 0.f, 0u, 0, nullptr, sk_work, sk_ws)))
    return rc;
  if (dx_f32) {
    // d frontal | d lateral as two contiguous [B, D] halves (dx_f32 is [2][B][D]): the callers hand them to autograd as they
    // are -- a [B, 2D] result would cost two strided 64 MB copies at B = 32768 to split
    for (int v = 0; v < 2; ++v)
      if ((rc = gemm_bf16(dh_bf, static_cast<const __nv_bfloat16*>(w0_bf16) + static_cast<size_t>(v) * D, 0, 1, (int)B, D, D, D, 2 * D,
                          EPI_STORE_F32, 1.0f, dx_f32 + static_cast<size_t>(v) * B * D, D, nullptr, 0, nullptr, nullptr, 0, nullptr, 0, 1, s)))
        return rc;
  }
  return B200_OK;
}

extern "C" int b200clip_version(void) { return 100; }
extern "C" unsigned long long b200clip_launch_count(void) { return b200::launch_counter().load(); }
extern "C" const char* b200clip_last_error_string(void) { return b200::last_error().c_str(); }
