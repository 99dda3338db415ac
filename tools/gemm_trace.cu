// Debug probe (not part of the product): per-tile timeline of the CTA-pair GEMM and single-stage knock-out
// experiments.  Build & run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DLLB_GEMM_TRACE -I. tools/gemm_trace.cu \
//        llamole_b200/csrc/llb_runtime.cu -lcuda -o /tmp/gemm_trace && /tmp/gemm_trace [trace]
// Experiment mask: 1 = no MMA issue, 2 = no epilogue, 4 = no epilogue math, 8 = no epilogue store.
#include <vector>
#include "../llamole_b200/csrc/llb_gemm.cuh"
using namespace llb;

static __nv_bfloat16 *A, *W, *C;
static float* bias;

static int run(int M, int N, int K, int act) {
  if (act == 1) return launch_gemm<256>(A, K, W, K, M, N, K, EpiBiasAct<LLB_ACT_GELU, false>{C, N, bias}, 0);
  if (act == 9) return launch_gemm<256>(A, K, W, K, M, N, K, EpiBiasAct<9, false>{C, N, bias}, 0);
  if (act == 10) return launch_gemm<256>(A, K, W, K, M, N, K, EpiBiasAct<10, false>{C, N, bias}, 0);
  return launch_gemm<256>(A, K, W, K, M, N, K, EpiBiasAct<LLB_ACT_NONE, false>{C, N, bias}, 0);
}

static float time_ms(int M, int N, int K, int act, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) run(M, N, K, act);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) run(M, N, K, act);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / iters;
}

int main(int argc, char** argv) {
  const bool want_trace = argc > 1;
  const size_t maxMK = (size_t)204800 * 4096;
  cudaMalloc(&A, maxMK * 2), cudaMalloc(&W, (size_t)4096 * 4096 * 2), cudaMalloc(&C, maxMK * 2), cudaMalloc(&bias, 4096 * 4);
  cudaMemset(A, 0x11, maxMK * 2), cudaMemset(W, 0x11, (size_t)4096 * 4096 * 2), cudaMemset(bias, 0, 4096 * 4);
  struct Shape { const char* name; int M, N, K, act; };
  const Shape shapes[] = {{"fc1", 204800, 4096, 1024, 1}, {"fc1-nomufu", 204800, 4096, 1024, 9}, {"fc1-8ffma", 204800, 4096, 1024, 10},
                          {"fc1-noact", 204800, 4096, 1024, 0}, {"fc2", 204800, 1024, 4096, 0}};
  const int masks[] = {0, 4, 8, 12, 2, 1, 3};
  const char* mask_name[] = {"full", "no-math", "no-store", "ld-only-epi", "no-epilogue", "no-mma", "loads-only"};
  printf("%-10s", "shape");
  for (auto n : mask_name) printf(" %12s", n);
  printf("   (ms | TFLOP/s-equivalent)\n");
  for (const Shape& s : shapes) {
    printf("%-10s", s.name);
    for (int mi = 0; mi < 7; ++mi) {
      cudaMemcpyToSymbol(g_gemm_exp, &masks[mi], sizeof(int));
      const float ms = time_ms(s.M, s.N, s.K, s.act, 20);
      printf(" %5.3f|%5.0f ", ms, 2.0 * s.M * s.N * s.K / ms / 1e9);
    }
    printf("\n");
    fflush(stdout);
  }
  int zero = 0;
  cudaMemcpyToSymbol(g_gemm_exp, &zero, sizeof(int));
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  if (!want_trace) return 0;
  const int tiles = 64;
  long long* trace;
  cudaMalloc(&trace, tiles * 2 * 16 * 8);
  for (const Shape& s : shapes) {
    cudaMemset(trace, 0, tiles * 2 * 16 * 8);
    cudaMemcpyToSymbol(g_gemm_trace, &trace, sizeof(trace));
    run(s.M, s.N, s.K, s.act);
    cudaDeviceSynchronize();
    std::vector<long long> h(tiles * 2 * 16);
    cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost);
    auto T = [&](int t, int cta, int sl) { return h[((size_t)t * 2 + cta) * 16 + sl]; };
    printf("%s: tile | leader: prod_wait_empty  mma_wait_tmem_empty  mma_wait_full  mma_issue_span  tile_period  epi_wait  epi_work | peer: prod_wait epi_wait epi_work\n", s.name);
    for (int t = 4; t < 12; ++t)
      printf("%3d | %6lld %6lld %6lld %6lld %6lld %6lld %6lld | %6lld %6lld %6lld\n", t, T(t, 0, 0), T(t, 0, 3) - T(t, 0, 2), T(t, 0, 4),
             T(t, 0, 5) - T(t, 0, 3), T(t, 0, 7) - T(t - 1, 0, 7), T(t, 0, 7) - T(t, 0, 6), T(t, 0, 8) - T(t, 0, 7), T(t, 1, 0),
             T(t, 1, 7) - T(t, 1, 6), T(t, 1, 8) - T(t, 1, 7));
  }
  return 0;
}
