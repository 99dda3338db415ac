// Runtime plumbing of the C ABI: status/last-error, architecture gate, TMA descriptor encoding, and the
// runtime-dispatched GEMM entry (llb_gemm_bf16).
#include <stdarg.h>

#include <mutex>

#include "llb_gemm.cuh"

namespace llb {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 1;
  }
  return sms;
}

int require_sm100() {
  static int status = 1;  // 1 = unknown
  if (status == 1) {
    int dev = 0, major = 0, minor = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (e != cudaSuccess) return fail(LLB_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
    if (major != 10)
      return fail(LLB_ERR_ARCH, "llamole_b200 is sm_100a-only (tcgen05/TMEM/TMA); device %d is sm_%d%d and there is no fallback path",
                  dev, major, minor);
    status = LLB_OK;
  }
  return status;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map_bf16(CUtensorMap* out, const void* ptr, int rows, int cols, int ld_elems, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(LLB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(LLB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for rows=%d cols=%d ld=%d box_rows=%d ptr=%p", (int)r, rows,
                cols, ld_elems, box_rows, ptr);
  return LLB_OK;
}

// Pick the N tile: fewest "tile-column units" per SM wave, with a mild preference for the wide tile.
static int pick_bn(int M, int N) {
  const int sms = num_sms();
  const int mt = ceil_div(M, GEMM_BM);
  int best = 256;
  double best_cost = 1e30;
  const int cand[3] = {256, 128, 64};
  const double ineff[3] = {1.0, 1.08, 1.3};
  for (int i = 0; i < 3; ++i) {
    const int tiles = mt * ceil_div(N, cand[i]);
    const double cost = (double)ceil_div(tiles, sms) * cand[i] * ineff[i];
    if (cost < best_cost) {
      best_cost = cost;
      best = cand[i];
    }
  }
  return best;
}

template <int ACT, bool F32>
static int gemm_dispatch_bn(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int M,
                            int N, int K, cudaStream_t stream, GemmCounters* ctr) {
  EpiBiasAct<ACT, F32> epi{C, bias, ldc};
  switch (pick_bn(M, N)) {
    case 256: return launch_gemm<256, 8>(A, lda, W, ldw, M, N, K, epi, stream, ctr);
    case 128: return launch_gemm<128, 8>(A, lda, W, ldw, M, N, K, epi, stream, ctr);
    default: return launch_gemm<64, 8>(A, lda, W, ldw, M, N, K, epi, stream, ctr);
  }
}

template <bool F32>
static int gemm_dispatch_act(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int M,
                             int N, int K, int act, cudaStream_t stream, GemmCounters* ctr) {
  switch (act) {
    case LLB_ACT_NONE: return gemm_dispatch_bn<LLB_ACT_NONE, F32>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, ctr);
    case LLB_ACT_GELU: return gemm_dispatch_bn<LLB_ACT_GELU, F32>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, ctr);
    case LLB_ACT_SILU: return gemm_dispatch_bn<LLB_ACT_SILU, F32>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, ctr);
    case LLB_ACT_SOFTSIGN:
      return gemm_dispatch_bn<LLB_ACT_SOFTSIGN, F32>(A, lda, W, ldw, bias, C, ldc, M, N, K, stream, ctr);
  }
  return fail(LLB_ERR_INVALID, "gemm: unknown activation %d", act);
}

int gemm_bias_act(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int M, int N,
                  int K, int act, bool out_f32, cudaStream_t stream, GemmCounters* ctr) {
  return out_f32 ? gemm_dispatch_act<true>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, stream, ctr)
                 : gemm_dispatch_act<false>(A, lda, W, ldw, bias, C, ldc, M, N, K, act, stream, ctr);
}

}  // namespace llb

extern "C" {

const char* llb_last_error(void) { return llb::g_last_error.c_str(); }

int llb_version(void) { return 100; }

int llb_arch_check(int device) {
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e != cudaSuccess) return llb::fail(LLB_ERR_CUDA, "device %d: %s", device, cudaGetErrorString(e));
  if (major != 10)
    return llb::fail(LLB_ERR_ARCH, "device %d is sm_%d%d; llamole_b200 runs on sm_100 (B200) only and has no fallback", device,
                     major, minor);
  return LLB_OK;
}

int llb_gemm_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int M, int N,
                  int K, int act, int out_fp32, llb_stream_t stream) {
  LLB_TRY(llb::require_sm100());
  LLB_CHECK_ARG(A && W && C, "llb_gemm_bf16: null operand");
  return llb::gemm_bias_act(A, lda, W, ldw, bias, C, ldc, M, N, K, act, out_fp32 != 0, (cudaStream_t)stream, nullptr);
}

}  // extern "C"
