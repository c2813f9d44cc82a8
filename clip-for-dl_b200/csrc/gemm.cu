// b200clip: C-ABI entry point for the tcgen05 GEMM (used by the projection block and by the parity tests).
#include <algorithm>

#include "gemm.cuh"
#include "host.cuh"
#include "../../include/b200clip.h"

namespace b200 {

// out[i] = sum_z partial[z * total + i], z ascending (fixed order: deterministic)
__global__ void __launch_bounds__(256) sum_splits_kernel(const float* __restrict__ partial, int splits, long long total,
                                                         float* __restrict__ out) {
  for (long long i = (blockIdx.x * 256ll + threadIdx.x) * 4; i < total; i += static_cast<long long>(gridDim.x) * 1024) {
    float4 a = *reinterpret_cast<const float4*>(partial + i);
    for (int z = 1; z < splits; ++z) {
      const float4 b = *reinterpret_cast<const float4*>(partial + z * total + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(out + i) = a;
  }
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int splits,
                       cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<BN, STAGES, A_MN, B_MN, EPI>;
  constexpr int smem = gemm_smem_bytes<BN, STAGES>();
  static SmemAttrOnce attr;          // per instantiation, per device
  B200_CHECK_CUDA(attr.ensure(kern, smem));
  GemmParams q = p;
  q.splits = splits;
  const int items = ((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + BN - 1) / BN) * splits;
  kern<<<std::min(items, num_sms()), GEMM_THREADS, smem, stream>>>(ta, tb, q);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

int gemm_bf16(const void* a, const void* b, int a_mn, int b_mn, int M, int N, int K, long long lda, long long ldb,
              int epi, float alpha, void* out0, long long ld0, void* out1, long long ld1, const float* bias,
              const void* resid, long long ld_res, const float* aux, long long ld_aux, int split_k,
              cudaStream_t stream, float drop_p, unsigned int drop_seed, int aux_is_bf16, const unsigned int* drop_seed_dev,
              float* splitk_ws, size_t splitk_ws_bytes) {
  B200_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  // the epilogues use 256-bit global accesses: 32-byte aligned bases, leading dimensions a multiple of 16 elements
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
  B200_REQUIRE(out0 != nullptr && al32(out0) && al32(out1) && al32(resid) && al32(aux), "gemm: out0/out1/resid/aux must be 32-byte aligned");
  B200_REQUIRE(ld0 % 16 == 0 && ld1 % 16 == 0 && ld_res % 16 == 0 && ld_aux % 16 == 0, "gemm: output/resid/aux leading dimensions must be multiples of 16 elements");
  // 128 x 256 tiles unless they would leave SMs idle: a [4096 x 512] output is only 64 such tiles for 148 SMs, and the kernel
  // is then one wave of half-empty hardware (data-parallel ranks and BASELINE configs[1] live there); 128 x 128 tiles double
  // the CTA count at the price of re-streaming the A tile twice (L2-resident at these sizes)
  // (split-K launches already size their split count to one wave of 128 x 256 tiles)
  const int tiles256 = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + 255) / 256);
  const int BN = (N % 256 == 0 && (split_k > 1 || tiles256 >= num_sms())) ? 256 : 128;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.k_chunks = (K + GEMM_BK - 1) / GEMM_BK;
  if (split_k < 1) split_k = 1;
  if (split_k > p.k_chunks) split_k = p.k_chunks;
  B200_REQUIRE(split_k == 1 || epi == EPI_ATOMIC_F32, "gemm: split-K needs the atomic epilogue");
  p.k_chunks_per_split = (p.k_chunks + split_k - 1) / split_k;
  const int splits = (p.k_chunks + p.k_chunks_per_split - 1) / p.k_chunks_per_split;
  // Deterministic split-K: with a workspace every split stores its fp32 partial tile (plain stores) and a second kernel adds
  // the partials in split order -- bit-reproducible, unlike red.global.add whose order depends on CTA scheduling.
  const bool det_split = epi == EPI_ATOMIC_F32 && splitk_ws != nullptr && ld0 == N &&
                         splitk_ws_bytes >= static_cast<size_t>(splits) * M * N * sizeof(float);
  float* final_out = static_cast<float*>(out0);
  if (det_split) {
    epi = EPI_STORE_F32;
    out0 = splitk_ws;
  }
  p.alpha = alpha;
  p.out0 = out0; p.ld0 = ld0; p.out1 = out1; p.ld1 = ld1; p.bias = bias;
  p.resid = reinterpret_cast<const __nv_bfloat16*>(resid); p.ld_res = ld_res; p.aux = aux; p.ld_aux = ld_aux;
  p.drop_p = drop_p; p.drop_seed = drop_seed; p.aux_is_bf16 = aux_is_bf16; p.drop_seed_dev = drop_seed_dev;
  p.split_rows = det_split ? M : 0;

  CUtensorMap ta, tb;
  int rc;
  if (!a_mn) rc = make_tmap_bf16_2d(&ta, a, M, K, lda, 64, GEMM_BM);       // a[M][K], box 128 rows x 64 k
  else       rc = make_tmap_bf16_2d(&ta, a, K, M, lda, 64, 64);            // a[K][M], box 64 k-rows x 64 m
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_bf16_2d(&tb, b, N, K, ldb, 64, BN);            // b[N][K]
  else       rc = make_tmap_bf16_2d(&tb, b, K, N, ldb, 64, 64);            // b[K][N]
  if (rc) return rc;

  auto finish = [&](int rc_launch) -> int {
    if (rc_launch || !det_split) return rc_launch;
    const long long total = static_cast<long long>(M) * N;       // multiple of 4 (N % 32 == 0)
    const int grid = static_cast<int>(std::min<long long>((total / 4 + 255) / 256, 4ll * num_sms()));
    sum_splits_kernel<<<grid, 256, 0, stream>>>(splitk_ws, splits, total, final_out);
    B200_LAUNCH_CHECK();
    return B200_OK;
  };
#define B200_GEMM_CASE(AMN, BMN, E)                                                                     \
  if (a_mn == AMN && b_mn == BMN && epi == E) {                                                         \
    return finish(BN == 256 ? launch_gemm<256, 4, AMN, BMN, E>(ta, tb, p, splits, stream)               \
                            : launch_gemm<128, 6, AMN, BMN, E>(ta, tb, p, splits, stream));             \
  }
  B200_GEMM_CASE(0, 0, EPI_STORE_F32)
  B200_GEMM_CASE(0, 0, EPI_STORE_BF16)
  B200_GEMM_CASE(0, 0, EPI_BIAS_GELU)
  B200_GEMM_CASE(0, 0, EPI_BIAS_RESID_F32)
  B200_GEMM_CASE(0, 0, EPI_RELU_BF16)
  B200_GEMM_CASE(0, 1, EPI_STORE_F32)
  B200_GEMM_CASE(0, 1, EPI_STORE_BF16)
  B200_GEMM_CASE(0, 1, EPI_GELU_BWD)
  B200_GEMM_CASE(0, 1, EPI_RELU_BWD)
  B200_GEMM_CASE(1, 1, EPI_STORE_F32)
  B200_GEMM_CASE(1, 1, EPI_ATOMIC_F32)
  B200_GEMM_CASE(1, 0, EPI_STORE_F32)
#undef B200_GEMM_CASE
  return fail(B200_ERR_UNSUPPORTED, "gemm: no kernel for a_mn=%d b_mn=%d epilogue=%d", a_mn, b_mn, epi);
}

}  // namespace b200

extern "C" int b200clip_gemm_bf16(const void* a, const void* b, int a_mn_major, int b_mn_major, int M, int N, int K,
                                  long long lda, long long ldb, int epilogue, float alpha, void* out0, long long ld0,
                                  void* out1, long long ld1, const float* bias, const void* resid, long long ld_res,
                                  const float* aux, long long ld_aux, int split_k, void* stream) {
  return b200::gemm_bf16(a, b, a_mn_major, b_mn_major, M, N, K, lda, ldb, epilogue, alpha, out0, ld0, out1, ld1, bias,
                         resid, ld_res, aux, ld_aux, split_k, static_cast<cudaStream_t>(stream), 0.f, 0u, 0, nullptr, nullptr, 0);
}
