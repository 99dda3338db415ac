// Fused   x += gate[g] * ( LN(A . W^T + bias) * (1 + scale[g]) + shift[g] )   ->  x (fp32, in place) and xb (bf16)
//
// This is the tail of both halves of a GraphDiT block (transformer.py:143-144: attention projection and MLP fc2, each
// followed by a non-affine LayerNorm, the adaLN modulation, the gate and the residual add).  Unfused, the GEMM writes
// y (bf16) and a row kernel re-reads it together with x: 14 bytes per element of HBM traffic and a memory-bound
// kernel between two GEMMs.  Here the LayerNorm runs on the fp32 accumulators in TMEM: 10 bytes per element, all of
// it overlapped with the next tile's MMAs.
//
// A full output row (N = H columns) has to be visible to normalise it, so a CLUSTER of CL = N / 256 CTAs owns one
// 128-row tile: CTA r computes columns [256 r, 256 r + 256) with the same pipeline as gemm_tcgen05_kernel<256>
// (TMA producer warp, single-thread tcgen05.mma issuer, 2 TMEM accumulator stages, 8 epilogue warps).  The epilogue
//   pass 1  tcgen05.ld -> per-row sum / sum of squares over the CTA's 256 columns; every CTA pushes its partials into
//           all CTAs of the cluster through distributed shared memory (st.shared::cluster) and arrives on their
//           mbarrier (release.cluster); nobody waits on a cluster-wide barrier, so the producer / MMA warps of the
//           four SMs keep streaming.
//   pass 2  tcgen05.ld again (TMEM reads are cheap) -> normalise, modulate with the tile's (<= 4) modulation rows staged
//           in shared memory, gate -> transpose through a swizzled per-warp staging tile -> coalesced 128-byte
//           row segments: residual read (prefetched before the TMEM load), x store, bf16 xb store.
#pragma once
#include "llb_gemm.cuh"

namespace llb {

constexpr int GLN_BN = 256;
#ifndef GLN_STAGES_OVERRIDE
constexpr int GLN_STAGES = 3;
#else
constexpr int GLN_STAGES = GLN_STAGES_OVERRIDE;
#endif
constexpr int GLN_MAX_CL = 4;
constexpr int GLN_PAIR_MAX_GROUPS = 18;   // 4-pair groups of the CTA-pair variant (72 of the B200's 74 pairs)
constexpr int GLN_MAX_GROUPS = 4;   // modulation rows staged per tile (a 128-row tile of 50-atom molecules touches <= 4)

struct GemmLnArgs {
  const float* bias;            // (N) or null
  const int32_t* row_group;     // (M) modulation row of each token row, non-decreasing runs
  const float* shift;           // modulation vectors: v[g * mod_ld + col]
  const float* scale;
  const float* gate;
  int mod_ld;
  float* x;                     // (M, ldx) fp32 residual stream, updated in place
  int ldx;
  __nv_bfloat16* xb;            // (M, ldxb) bf16 copy of the new x
  int ldxb;
};

// N must be CL * 256 with CL in 1..4 (hidden sizes 256 / 512 / 768 / 1024); callers fall back to GEMM + row kernel otherwise.
inline bool gemm_ln_supported(int N, int K) { return N % GLN_BN == 0 && N / GLN_BN >= 1 && N / GLN_BN <= GLN_MAX_CL && K % 8 == 0; }
int gemm_ln_mode();       // env LLB_FUSED_LN, 0..3 (default 3): see dit_forward
inline bool gemm_ln_enabled() { return gemm_ln_mode() != 0; }
int launch_gemm_ln(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmLnArgs& e, cudaStream_t stream,
                   GemmCounters* ctr);

// Tail of a GIN layer fused into the second linear of the node MLP (graph_encoder/model.py:131-149, graph_predictor/model.py:318-348):
//   h[r,:] = gate_g * act( LN(A[r,:] . W^T + bias) * s1 + sh ) + h[r,:] + addvec_g      g = row_group[r] (graph of node r)
// with (s1, sh) = (gamma, beta) per column (encoder: affine LayerNorm) or (1 + scale_g, shift_g) per graph (predictor: text-adaLN);
// gate / addvec optional.  h (fp32, in place) and hb (its bf16 copy) are written through TMA; optionally the per-graph column
// MAXIMUM of the new h is accumulated into pool_max (order-preserving uint encoding, atomicMax: order-independent, hence
// deterministic) -- the next layer's virtual-node update pools exactly this matrix (model.py:148).
struct GinTailArgs {
  const float* bias;          // (N)
  const int32_t* row_group;   // (M) graph id per node row, non-decreasing
  const float* gamma;         // (N) or null
  const float* beta;          // (N) or null
  const float* shift;         // (graphs, mod_ld) or null
  const float* scale;
  const float* gate;
  int mod_ld;
  const float* addvec;        // (graphs, addvec_ld) or null
  int addvec_ld;
  int act;                    // LLB_ACT_NONE or LLB_ACT_GELU
  float* x;                   // (M, ldx) fp32 node features, updated in place
  int ldx;
  __nv_bfloat16* xb;          // (M, ldxb)
  int ldxb;
  uint32_t* pool_max;         // (graphs, pool_ld) encoded running maxima (cleared to 0 by the caller) or null
  int pool_ld;
  int a_f16;                  // A and W are IEEE fp16 instead of bf16
};
// N = 768 (three CTA pairs per 256-row block) or 1024 (four); needs the same exchange workspace as launch_gemm_ln_pair.
inline bool gin_tail_supported(int N, int K) { return (N == 3 * GLN_BN || N == 4 * GLN_BN) && K % 8 == 0; }
int launch_gin_tail(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GinTailArgs& t, void* sync_ws, size_t sync_bytes,
                    cudaStream_t stream, GemmCounters* ctr);
// order-preserving encoding of a float for unsigned atomicMax (0 = below every float)
__host__ __device__ inline uint32_t float_order_enc(uint32_t bits) { return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u); }
__host__ __device__ inline uint32_t float_order_dec(uint32_t e) { return (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e; }

// CTA-pair variant for N = 1024 (see llb_gemm_ln.cu): needs a caller-provided exchange workspace (any contents; the launch
// clears its counters with a stream-ordered memset) that no other launch uses concurrently.
size_t gemm_ln_pair_workspace_bytes();
int launch_gemm_ln_pair(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const GemmLnArgs& e, void* sync_ws,
                        size_t sync_bytes, cudaStream_t stream, GemmCounters* ctr);

}  // namespace llb
