// b200clip: host-side helpers shared by the C-ABI translation units (error reporting, TMA descriptor encoding).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

namespace b200 {

// error codes returned across the C ABI (0 = ok)
enum : int {
  B200_OK = 0,
  B200_ERR_INVALID = -1,     // shape / alignment / dtype contract violated (rejected before any launch)
  B200_ERR_CUDA = -2,        // a CUDA runtime/driver call failed
  B200_ERR_WORKSPACE = -3,   // caller's workspace is too small
  B200_ERR_UNSUPPORTED = -4, // valid request the kernels do not cover (never a silent fallback)
};

inline std::string& last_error() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}
#define B200_CHECK_CUDA(expr)                                                                    \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::b200::fail(::b200::B200_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)
#define B200_REQUIRE(cond, ...)                                            \
  do {                                                                     \
    if (!(cond)) return ::b200::fail(::b200::B200_ERR_INVALID, __VA_ARGS__); \
  } while (0)
// every kernel launch is followed by exactly one B200_LAUNCH_CHECK(): it also feeds b200clip_launch_count()
inline std::atomic<unsigned long long>& launch_counter() {
  static std::atomic<unsigned long long> n{0};
  return n;
}
#define B200_LAUNCH_CHECK()                              \
  do {                                                   \
    ::b200::launch_counter().fetch_add(1);               \
    B200_CHECK_CUDA(cudaGetLastError());                 \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time dependency on libcuda
// (it must load on a CPU-only box for the symbol-export test).
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 2-D bf16 row-major matrix [rows x cols] with row pitch `ld` elements; box = [box_rows x box_cols], 128B swizzle
// (box_cols * 2 bytes must be <= 128). Out-of-bounds elements read as zero.
inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_cols, uint32_t box_rows) {
  // The driver entry point needs a current context on THIS thread; PyTorch's autograd engine calls us from worker
  // threads where only its own runtime instance has bound one.  A no-op runtime call binds the primary context.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaError_t e = cudaFree(nullptr);
    if (e != cudaSuccess) return fail(B200_ERR_CUDA, "cudaFree(0) (context bind) failed: %s", cudaGetErrorString(e));
    ctx_bound = true;
  }
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (driver too old?)");
  if (!aligned16(base)) return fail(B200_ERR_INVALID, "TMA base pointer must be 16-byte aligned");
  if ((ld * 2) % 16 != 0) return fail(B200_ERR_INVALID, "TMA row pitch (%llu elems) must be a multiple of 8", (unsigned long long)ld);
  if (box_cols * 2 > 128 || box_rows > 256) return fail(B200_ERR_INVALID, "TMA box too large");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return B200_OK;
}

// The fixed log-sum-exp shift of the flash InfoNCE, in log2 units (csrc/infonce.cu, head_loss_finalize in smallc.cu).
// E = exp2(s k1 - k2) with k1 = log2(e) / tau and cosines s in [-1, 1].  k2 = k1 (shift m = 1/tau) keeps E <= 1, but at small
// tau a whole row can flush to zero: exp2 flushes below 2^-126, i.e. when every cosine of the row is below 1 - 126/k1 (0.125 at
// tau = 0.01, CLIP's lower clamp -- normal early in training).  So the shift is lowered by off = clamp(k1 - 40, 0, 88):
// E <= 2^off, sums of < 2^30 terms stay below 2^118, and a row only underflows if all its cosines are below 1 - (126 + off)/k1
// (-0.49 at tau = 0.01).  tau >= 0.008 is required; tau >= 0.036 (k1 <= 40) gives off = 0: bit-identical to plain m = 1/tau.
constexpr float NCE_LOG2E = 1.4426950408889634f;
constexpr float NCE_MIN_TAU = 0.008f;
inline float nce_k2(float temperature) {
  const float k1 = NCE_LOG2E / temperature;
  const float off = k1 - 40.f < 0.f ? 0.f : (k1 - 40.f > 88.f ? 88.f : k1 - 40.f);
  return k1 - off;
}
inline double nce_shift(float temperature) { return static_cast<double>(nce_k2(temperature)) / static_cast<double>(NCE_LOG2E); }

// Per-device immutable attribute cache (SURVEY.md 8b "Threading"): SM counts, and which (kernel, device) pairs already carry
// their opt-in dynamic shared-memory limit -- cudaFuncSetAttribute is a PER-DEVICE setting, so a process that touches a second
// GPU must set it again there.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
inline int num_sms() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}
// One flag array per call site (a `static SmemAttrOnce` next to the launch): sets MaxDynamicSharedMemorySize once per device.
struct SmemAttrOnce {
  std::once_flag once[kMaxDevices];
  cudaError_t err[kMaxDevices] = {};
  template <typename F>
  cudaError_t ensure(F* func, int bytes) {
    const int dev = current_device();
    std::call_once(once[dev], [&] { err[dev] = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); });
    return err[dev];
  }
};


// A fork lane = one auxiliary stream + events, for entry points that run two independent kernel chains concurrently
// (b200clip_proj_bwd: weight-gradient GEMMs beside the dz -> dp -> dx chain).  fork(): aux waits for everything enqueued on
// the caller's stream so far; join(): the caller's stream waits for aux.  Event record / wait are capturable, so inside a
// CUDA-graph capture the lane's work becomes a parallel branch of the same graph.  Lanes are created once per device (all of
// them at the first use, i.e. in the un-captured warm-up call) and handed out round-robin, so two concurrent callers (the
// text-side and image-side chains run on two streams) get different lanes.
struct ForkLane {
  cudaStream_t aux = nullptr;
  cudaEvent_t ev[4] = {};
  cudaError_t link(cudaStream_t from, cudaStream_t to, int e) {       // `to` waits for what `from` holds now
    cudaError_t rc = cudaEventRecord(ev[e], from);
    return rc != cudaSuccess ? rc : cudaStreamWaitEvent(to, ev[e], 0);
  }
};
inline ForkLane* acquire_fork_lane() {
  constexpr int kLanes = 8;
  static ForkLane lanes[kMaxDevices][kLanes];
  static std::once_flag once[kMaxDevices];
  static bool ok[kMaxDevices] = {};
  static std::atomic<unsigned> next[kMaxDevices];
  const int dev = current_device();
  std::call_once(once[dev], [&] {
    bool good = true;
    for (auto& l : lanes[dev]) {
      good = good && cudaStreamCreateWithFlags(&l.aux, cudaStreamNonBlocking) == cudaSuccess;
      for (auto& e : l.ev) good = good && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    }
    ok[dev] = good;
  });
  if (!ok[dev]) return nullptr;
  return &lanes[dev][next[dev].fetch_add(1u, std::memory_order_relaxed) % kLanes];
}

}  // namespace b200
