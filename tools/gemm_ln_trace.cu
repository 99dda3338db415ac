// Debug probe (not part of the product): phase timeline of the fused GEMM + LayerNorm CTA-pair kernel (GraphDiT block tails).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DLLB_GEMM_TRACE -I. \
//        tools/gemm_ln_trace.cu llamole_b200/csrc/llb_gemm_ln.cu llamole_b200/csrc/llb_runtime.cu llamole_b200/csrc/llb_rowops.cu -lcuda -o tools/gemm_ln_trace.bin
#include <vector>
#include "../llamole_b200/csrc/llb_gemm_ln.cuh"
using namespace llb;
extern "C" void llb_gln_set_trace(long long* p);
extern "C" void llb_gln_set_exp(int m);
int main() {
  const int M = 204800, N = 1024;
  __nv_bfloat16 *A, *W, *xb;
  float *x, *mod, *bias;
  int32_t* grp;
  cudaMalloc(&A, (size_t)M * 4096 * 2), cudaMalloc(&W, (size_t)N * 4096 * 2), cudaMalloc(&xb, (size_t)M * N * 2);
  cudaMalloc(&x, (size_t)M * N * 4), cudaMalloc(&mod, (size_t)2049 * 6 * N * 4), cudaMalloc(&bias, N * 4), cudaMalloc(&grp, M * 4);
  cudaMemset(A, 0x11, (size_t)M * 4096 * 2), cudaMemset(W, 0x11, (size_t)N * 4096 * 2), cudaMemset(x, 0, (size_t)M * N * 4);
  cudaMemset(mod, 0, (size_t)2049 * 6 * N * 4), cudaMemset(bias, 0, N * 4);
  std::vector<int32_t> g(M);
  for (int r = 0; r < M; ++r) g[r] = r < M / 2 ? r / 50 : 2048;
  cudaMemcpy(grp, g.data(), M * 4, cudaMemcpyHostToDevice);
  long long* trace;
  cudaMalloc(&trace, 64 * 16 * 8);
  void* ws; const size_t wsb = gemm_ln_pair_workspace_bytes(); cudaMalloc(&ws, wsb);
  const int nexp = 3;
  const int exps[nexp] = {0, 3, 16};
  for (int xi = 0; xi < nexp; ++xi)
  for (int K : {1024, 4096}) {
    llb_gln_set_exp(exps[xi]);
    printf("=== pair kernel, knock-out mask %d (1 no x store, 2 no xb store, 4 no residual reload, 8 no pass 2, 16 empty epilogue)\n", exps[xi]);
    GemmLnArgs e{bias, grp, mod, mod + N, mod + 2 * N, 6 * N, x, N, xb, N};
    auto launch = [&]() { return launch_gemm_ln_pair(A, K, W, K, M, N, K, e, ws, wsb, 0, nullptr); };
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i)
      if (launch()) { printf("launch failed: %s\n", llb_last_error()); return 1; }
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    printf("K=%d: %.3f ms  %.0f TFLOP/s  (%s)\n", K, ms, 2.0 * M * N * K / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaMemset(trace, 0, 64 * 16 * 8);
    llb_gln_set_trace(trace);
    launch();
    cudaDeviceSynchronize();
    llb_gln_set_trace(nullptr);
    std::vector<long long> h(64 * 16);
    cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost);
    auto T = [&](int t, int s) { return h[(size_t)t * 16 + s]; };
    printf("tile | epi: phaseA  wait_tmem  pass1  stats_wait  pass2  period | mma: wait_empty  span | pass2: tmem+math  store_read_wait  res_wait  add+store\n");
    for (int t = 5; t < 9; ++t)
      printf("%3d | %6lld %6lld %6lld %6lld %6lld %6lld | %6lld %6lld | %6lld %6lld %6lld %6lld\n", t, T(t, 1) - T(t, 0), T(t, 2) - T(t, 1), T(t, 3) - T(t, 2),
             T(t, 4) - T(t, 3), T(t, 5) - T(t, 4), T(t, 0) - T(t - 1, 0), T(t, 9) - T(t, 8), T(t, 10) - T(t, 9), T(t, 11), T(t, 12), T(t, 13), T(t, 14));
  }
  return 0;
}
