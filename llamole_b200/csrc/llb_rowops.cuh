// Row-wise fused kernels shared by the GraphDiT and GIN paths (HBM-bound; one warp per row).
#pragma once
#include "llb_common.cuh"

namespace llb {

// out = resid + gate[g] * act( LN(in) * gamma + beta  ->  * (1 + scale[g]) + shift[g] ) + addvec[g]
// every stage optional; g = row_group[row] (or row).  Writes fp32 and/or bf16, optionally a second copy
// `dup_rows` rows further down (the unconditional half of the CFG batch shares the token embedding).
struct RowLnArgs {
  const void* in = nullptr;
  int in_ld = 0;
  bool in_bf16 = false;
  int in_parts = 1;          // the input row is the sum of in_parts pieces, in_part_stride elements apart (split-K partials),
  int in_part_stride = 0;    // added in index order
  int rows = 0, width = 0;
  bool normalize = true;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  const int32_t* row_group = nullptr;
  const float* shift = nullptr;
  const float* scale = nullptr;
  const float* gate = nullptr;
  int mod_ld = 0;
  int act = LLB_ACT_NONE;
  const float* resid = nullptr;
  int resid_ld = 0;
  const float* addvec = nullptr;  // (groups, width) fp32, added last
  int addvec_ld = 0;
  float* out_f32 = nullptr;
  int out_f32_ld = 0;
  __nv_bfloat16* out_bf16 = nullptr;
  int out_bf16_ld = 0;
  int dup_rows = 0;
  bool l2_normalize = false;  // out = v / ||v|| instead of LN (GraphCLIP head)
  int prof_slot = LLB_PROF_LN_MOD_RES;
};
int launch_row_ln(const RowLnArgs& a, cudaStream_t stream);

// fp32 (rows, cols) -> bf16 (rows, dst_ld) with zero padding up to pad_cols.
int launch_f32_to_bf16(const float* src, int src_ld, __nv_bfloat16* dst, int dst_ld, int rows, int cols, int pad_cols,
                       cudaStream_t stream);
// out[r, o] = act(sum_k in[r,k] W[o,k] + b[o]) in fp32 on CUDA cores (tiny set-up linears only).
int launch_linear_f32(const float* in, int in_ld, const float* W, int w_ld, const float* b, float* out, int out_ld,
                      int rows, int out_f, int in_f, int act, cudaStream_t stream);

}  // namespace llb
