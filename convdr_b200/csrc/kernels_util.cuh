// kernels_util.cuh — ingest-side kernels: bf16 shadow + norm bound, synthetic rows, query prep.
#pragma once
#include "common.cuh"

namespace b2f {

// Upper bound of a row norm^2 computed in fp32 (any summation order): inflate by 2^-12 so the
// stored value is >= the true sum of squares.
__device__ __forceinline__ float norm2_upper(float ss) { return ss * (1.0f + 2.44140625e-4f); }

__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}

// One warp per row: write the bf16 shadow row (round-to-nearest, K-block-major tiled layout — 16
// lanes fill one 128-byte K-block segment) and fold the row's norm^2 bound
// into *maxnorm2_bits (float bits; valid because the values are non-negative).
// HBM traffic per row: 3072 B read + 1536 B written.
__global__ void __launch_bounds__(256) convert_rows_kernel(const float* __restrict__ x32,
                                                           __nv_bfloat16* __restrict__ x16,
                                                           int64_t row0, int64_t n_rows,
                                                           unsigned int* __restrict__ maxnorm2_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float wmax = 0.f;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(x32 + (row0 + r) * kD);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      float4 v = ldg_stream_f4(src + lane + 32 * i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      if (x16) *reinterpret_cast<uint2*>(x16 + shadow_index(row0 + r, 4 * (lane + 32 * i))) = pack_bf16x4(v);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    wmax = fmaxf(wmax, norm2_upper(ss));
  }
  if (lane == 0 && wmax > 0.f) atomicMax(maxnorm2_bits, __float_as_uint(wmax));
}

// Synthetic rows (see include/b2f.h b2f_add_synthetic).  One warp per row; lane l produces
// float4 chunks l, l+32, ..., l+160; chunk c of row r is Philox4x32-10(counter = (r_lo, r_hi, c, 0),
// key = (seed_lo ^ stream_lo, seed_hi ^ stream_hi ^ 0x5eed)).  Writes fp32, the bf16 shadow and
// the norm bound in one pass (no re-read).
__global__ void __launch_bounds__(256) synth_rows_kernel(float* __restrict__ x32,
                                                         __nv_bfloat16* __restrict__ x16,
                                                         int64_t dst_row0, int64_t first_row,
                                                         int64_t n_rows, uint32_t k0, uint32_t k1,
                                                         float norm,
                                                         unsigned int* __restrict__ maxnorm2_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float wmax = 0.f;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const uint64_t row = static_cast<uint64_t>(first_row + r);
    int comp[kF4PerLane][4];
    int ss = 0;
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      U4 c;
      c.x = static_cast<uint32_t>(row);
      c.y = static_cast<uint32_t>(row >> 32);
      c.z = static_cast<uint32_t>(lane + 32 * i);
      c.w = 0u;
      U4 o = philox4x32_10(c, k0, k1);
      comp[i][0] = synth_component(o.x);
      comp[i][1] = synth_component(o.y);
      comp[i][2] = synth_component(o.z);
      comp[i][3] = synth_component(o.w);
      ss += comp[i][0] * comp[i][0] + comp[i][1] * comp[i][1] + comp[i][2] * comp[i][2] +
            comp[i][3] * comp[i][3];
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    // ss <= 768 * 510^2 < 2^31.  IEEE sqrt and division: bit-identical to the host restatement.
    const float inv = (ss > 0) ? __fdiv_rn(norm, __fsqrt_rn(static_cast<float>(ss))) : 0.f;
    float4* dst32 = reinterpret_cast<float4*>(x32 + (dst_row0 + r) * kD);
    float fs = 0.f;
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      float4 v;
      v.x = __fmul_rn(static_cast<float>(comp[i][0]), inv);
      v.y = __fmul_rn(static_cast<float>(comp[i][1]), inv);
      v.z = __fmul_rn(static_cast<float>(comp[i][2]), inv);
      v.w = __fmul_rn(static_cast<float>(comp[i][3]), inv);
      fs += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      dst32[lane + 32 * i] = v;
      if (x16) *reinterpret_cast<uint2*>(x16 + shadow_index(dst_row0 + r, 4 * (lane + 32 * i))) = pack_bf16x4(v);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, s);
    wmax = fmaxf(wmax, norm2_upper(fs));
  }
  if (lane == 0 && wmax > 0.f) atomicMax(maxnorm2_bits, __float_as_uint(wmax));
}

// Query preparation: one warp per (padded) query row.  Writes the bf16 copy used by the tensor
// path (zero rows for q >= nq) and an upper bound of ||q||.
__global__ void __launch_bounds__(128) prep_queries_kernel(const float* __restrict__ q32, int nq,
                                                           int nq_pad,
                                                           __nv_bfloat16* __restrict__ q16,
                                                           float* __restrict__ qnorm) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq_pad) return;
  uint2* dst = reinterpret_cast<uint2*>(q16 + static_cast<int64_t>(q) * kD);
  if (q >= nq) {
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) dst[lane + 32 * i] = make_uint2(0u, 0u);
    return;
  }
  const float4* src = reinterpret_cast<const float4*>(q32 + static_cast<int64_t>(q) * kD);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kF4PerLane; ++i) {
    float4 v = __ldg(src + lane + 32 * i);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    dst[lane + 32 * i] = pack_bf16x4(v);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  if (lane == 0) qnorm[q] = __fsqrt_ru(norm2_upper(ss));
}

}  // namespace b2f
