// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
// A (activations) and W (nn.Linear weight, (out,in)) are both K-major bf16, so both operands go through TMA
// with the 128-byte swizzle straight into the UMMA shared-memory layout.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      TMA producer: fills a ring of STAGES {A 128x64, W BNx64} tiles, arms full[] with expect_tx
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage and
//               tcgen05.commit's the stage's empty[] barrier; after the last k-block commits tmem_full[acc]
//   warp 2      TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4..   epilogue: tcgen05.ld their lane quarter, run the fused epilogue functor, store to global,
//               then release the accumulator stage (tmem_empty[acc]) so the next tile's MMAs overlap
// Tiles are walked n-fastest so that the CTAs in flight share A tiles through L2 and W stays L2-resident.
#pragma once
#include "llb_common.cuh"

namespace llb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = 2 * BN;  // 512 / 256 / 128
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment slack
};

// Epilogue functor contract:
//   static constexpr int CHUNK (32 or 64): consecutive columns handed over per call
//   __device__ void operator()(int row, int col0, const float* acc, int M, int N) const
//     row < M guaranteed; columns col0 .. col0+CHUNK-1 may exceed N (functor guards).
template <int BN, int EPI_WARPS, class Epi>
__global__ void __launch_bounds__(128 + 32 * EPI_WARPS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                    int K, Epi epi) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * GEMM_BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          tma_load_2d(smA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * GEMM_BK, m0);
          tma_load_2d(smB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * GEMM_BK, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(smB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            umma_bf16(d_tmem, umma_desc_k128(a_addr + k * 32), umma_desc_k128(b_addr + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const int ew = warp - 4;
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    constexpr int COL_GROUPS = EPI_WARPS / 4;   // warps sharing a quarter split the columns
    constexpr int COLS_PER_WARP = BN / COL_GROUPS;
    const int cg = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * GEMM_BM;
      const int n0 = (tile % num_n) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + cg * COLS_PER_WARP;
      float v[Epi::CHUNK];
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += Epi::CHUNK) {
        if (n0 + cg * COLS_PER_WARP + c < N) {   // warp-uniform
          tmem_ld32(t_row + c, v);
          if (Epi::CHUNK == 64) tmem_ld32(t_row + c + 32, v + 32);
          tmem_ld_wait();
          if (row < M) epi(row, n0 + cg * COLS_PER_WARP + c, v, M, N);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
// 2-D bf16 tensor map: inner dim = K (contiguous), outer dim = rows; box = 64 x box_rows; 128B swizzle.
int make_tensor_map_bf16(CUtensorMap* out, const void* ptr, int rows, int cols, int ld_elems, int box_rows);

struct GemmCounters {
  int64_t launches = 0;
};

template <int BN, int EPI_WARPS, class Epi>
int launch_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const Epi& epi,
                cudaStream_t stream, GemmCounters* ctr = nullptr) {
  using Cfg = GemmCfg<BN>;
  if (M <= 0 || N <= 0) return LLB_OK;
  LLB_CHECK_ARG(K > 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K=%d lda=%d ldw=%d (ld must be a multiple of 8)", K, lda, ldw);
  LLB_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                "gemm: operands must be 16-byte aligned");
  CUtensorMap tmA, tmB;
  LLB_TRY(make_tensor_map_bf16(&tmA, A, M, K, lda, GEMM_BM));
  LLB_TRY(make_tensor_map_bf16(&tmB, W, N, K, ldw, BN));
  auto kern = gemm_tcgen05_kernel<BN, EPI_WARPS, Epi>;
  static bool configured = false;  // per template instantiation
  if (!configured) {
    LLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ceil_div(M, GEMM_BM) * ceil_div(N, BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, 128 + 32 * EPI_WARPS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, M, N, K, epi);
  LLB_CUDA_OK(cudaGetLastError());
  if (ctr) ctr->launches++;
  return LLB_OK;
}

// ------------------------------------------------------------------------------------------------
// Epilogues
// ------------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == LLB_ACT_GELU) return gelu_erf(x);
  if (ACT == LLB_ACT_SILU) return silu(x);
  if (ACT == LLB_ACT_SOFTSIGN) return softsign(x);
  return x;
}

// C = act(acc + bias) -> bf16 or fp32 row-major.
template <int ACT, bool OUT_F32>
struct EpiBiasAct {
  static constexpr int CHUNK = 32;
  void* C;
  const float* bias;  // may be null
  int ldc;
  __device__ __forceinline__ void operator()(int row, int col0, const float* acc, int M, int N) const {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = col0 + i;
      float b = (bias != nullptr && c < N) ? __ldg(bias + c) : 0.0f;
      v[i] = apply_act<ACT>(acc[i] + b);
    }
    if (OUT_F32) {
      float* out = reinterpret_cast<float*>(C) + (size_t)row * ldc + col0;
      if (col0 + 32 <= N && (ldc & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(out + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
        for (int i = 0; i < 32; ++i)
          if (col0 + i < N) out[i] = v[i];
      }
    } else {
      __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(C) + (size_t)row * ldc + col0;
      if (col0 + 32 <= N && (ldc & 7) == 0) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(out + i) = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                                          pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
      } else {
        for (int i = 0; i < 32; ++i)
          if (col0 + i < N) out[i] = __float2bfloat16(v[i]);
      }
    }
  }
};

// Generic runtime-dispatched GEMM used by the small / non-critical linears.
int gemm_bias_act(const void* A, int lda, const void* W, int ldw, const float* bias, void* C, int ldc, int M, int N,
                  int K, int act, bool out_f32, cudaStream_t stream, GemmCounters* ctr = nullptr);

}  // namespace llb
