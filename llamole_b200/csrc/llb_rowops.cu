#include "llb_rowops.cuh"

namespace llb {

namespace {

struct RowLnDev {
  const void* in;
  int in_ld, in_bf16, rows, width, normalize;
  const float *gamma, *beta;
  const int32_t* row_group;
  const float *shift, *scale, *gate;
  int mod_ld, act;
  const float* resid;
  int resid_ld;
  const float* addvec;
  int addvec_ld;
  float* out_f32;
  int out_f32_ld;
  __nv_bfloat16* out_bf16;
  int out_bf16_ld, dup_rows, l2_normalize;
};

__device__ __forceinline__ float4 load4(const void* base, int is_bf16, size_t idx) {
  if (is_bf16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    return make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
}

__device__ __forceinline__ float act_rt(float x, int act) {
  switch (act) {
    case LLB_ACT_GELU: return gelu_erf(x);
    case LLB_ACT_SILU: return silu(x);
    case LLB_ACT_SOFTSIGN: return softsign(x);
  }
  return x;
}

// One warp per row; the row is re-read from L1/L2 for the second and third sweep (<= 16 KB per row).
__global__ void __launch_bounds__(256) row_ln_kernel(RowLnDev a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.rows) return;
  const size_t in_off = (size_t)row * a.in_ld;
  const int W = a.width;
  float mean = 0.f, rstd = 1.f;
  if (a.normalize || a.l2_normalize) {
    float s = 0.f;
    if (a.normalize) {
      for (int c = lane * 4; c < W; c += 128) {
        const float4 v = load4(a.in, a.in_bf16, in_off + c);
        s += (v.x + v.y) + (v.z + v.w);
      }
      mean = warp_sum(s) / (float)W;
    }
    float q = 0.f;
    for (int c = lane * 4; c < W; c += 128) {
      const float4 v = load4(a.in, a.in_bf16, in_off + c);
      const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    q = warp_sum(q);
    rstd = a.l2_normalize ? rsqrtf(q) : rsqrtf(q / (float)W + 1e-5f);
  }
  const int g = a.row_group ? a.row_group[row] : row;
  const float* sh = a.shift ? a.shift + (size_t)g * a.mod_ld : nullptr;
  const float* sc = a.scale ? a.scale + (size_t)g * a.mod_ld : nullptr;
  const float* gt = a.gate ? a.gate + (size_t)g * a.mod_ld : nullptr;
  const float* av = a.addvec ? a.addvec + (size_t)g * a.addvec_ld : nullptr;
  for (int c = lane * 4; c < W; c += 128) {
    const float4 v4 = load4(a.in, a.in_bf16, in_off + c);
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.resid) {
      const float4 r4 = *reinterpret_cast<const float4*>(a.resid + (size_t)row * a.resid_ld + c);
      r[0] = r4.x, r[1] = r4.y, r[2] = r4.z, r[3] = r4.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = (v[i] - mean) * rstd;
      if (a.gamma) x = x * __ldg(a.gamma + c + i) + __ldg(a.beta + c + i);
      if (sc) x = x * (1.0f + __ldg(sc + c + i)) + __ldg(sh + c + i);
      x = act_rt(x, a.act);
      if (gt) x *= __ldg(gt + c + i);
      if (a.resid) x += r[i];
      if (av) x += __ldg(av + c + i);
      v[i] = x;
    }
    if (a.out_f32) {
      const float4 o = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(a.out_f32 + (size_t)row * a.out_f32_ld + c) = o;
      if (a.dup_rows) *reinterpret_cast<float4*>(a.out_f32 + (size_t)(row + a.dup_rows) * a.out_f32_ld + c) = o;
    }
    if (a.out_bf16) {
      const uint2 o = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
      *reinterpret_cast<uint2*>(a.out_bf16 + (size_t)row * a.out_bf16_ld + c) = o;
      if (a.dup_rows) *reinterpret_cast<uint2*>(a.out_bf16 + (size_t)(row + a.dup_rows) * a.out_bf16_ld + c) = o;
    }
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, int src_ld, __nv_bfloat16* __restrict__ dst, int dst_ld,
                                   int rows, int cols, int pad_cols) {
  const size_t total = (size_t)rows * pad_cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / pad_cols), c = (int)(i % pad_cols);
    dst[(size_t)r * dst_ld + c] = __float2bfloat16(c < cols ? src[(size_t)r * src_ld + c] : 0.0f);
  }
}

// One warp per output element (set-up only: timestep-embedding table and friends).
__global__ void __launch_bounds__(256) linear_f32_kernel(const float* __restrict__ in, int in_ld, const float* __restrict__ W,
                                                         int w_ld, const float* __restrict__ b, float* __restrict__ out,
                                                         int out_ld, int rows, int out_f, int in_f, int act) {
  const size_t widx = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (widx >= (size_t)rows * out_f) return;
  const int r = (int)(widx / out_f), o = (int)(widx % out_f);
  const float* x = in + (size_t)r * in_ld;
  const float* w = W + (size_t)o * w_ld;
  float s = 0.f;
  for (int k = lane; k < in_f; k += 32) s = fmaf(x[k], w[k], s);
  s = warp_sum(s);
  if (lane == 0) out[(size_t)r * out_ld + o] = act_rt(s + (b ? b[o] : 0.0f), act);
}

}  // namespace

int launch_row_ln(const RowLnArgs& a, cudaStream_t stream) {
  if (a.rows <= 0) return LLB_OK;
  LLB_CHECK_ARG(a.width % 4 == 0 && a.in_ld % 4 == 0, "row_ln: width %d / ld %d must be multiples of 4", a.width, a.in_ld);
  RowLnDev d{a.in, a.in_ld, a.in_bf16 ? 1 : 0, a.rows, a.width, a.normalize ? 1 : 0, a.gamma, a.beta, a.row_group,
             a.shift, a.scale, a.gate, a.mod_ld, a.act, a.resid, a.resid_ld, a.addvec, a.addvec_ld, a.out_f32,
             a.out_f32_ld, a.out_bf16, a.out_bf16_ld, a.dup_rows, a.l2_normalize ? 1 : 0};
  row_ln_kernel<<<ceil_div(a.rows, 8), 256, 0, stream>>>(d);
  LLB_CUDA_OK(cudaGetLastError());
  return LLB_OK;
}

int launch_f32_to_bf16(const float* src, int src_ld, __nv_bfloat16* dst, int dst_ld, int rows, int cols, int pad_cols,
                       cudaStream_t stream) {
  if (rows <= 0) return LLB_OK;
  const size_t total = (size_t)rows * pad_cols;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  f32_to_bf16_kernel<<<blocks, 256, 0, stream>>>(src, src_ld, dst, dst_ld, rows, cols, pad_cols);
  LLB_CUDA_OK(cudaGetLastError());
  return LLB_OK;
}

int launch_linear_f32(const float* in, int in_ld, const float* W, int w_ld, const float* b, float* out, int out_ld,
                      int rows, int out_f, int in_f, int act, cudaStream_t stream) {
  if (rows <= 0) return LLB_OK;
  const size_t warps = (size_t)rows * out_f;
  linear_f32_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(in, in_ld, W, w_ld, b, out, out_ld, rows, out_f, in_f, act);
  LLB_CUDA_OK(cudaGetLastError());
  return LLB_OK;
}

}  // namespace llb
