// kernels_util.cuh — ingest-side kernels: bf16 shadow + norm bound, synthetic rows, query prep.
#pragma once
#include "common.cuh"

namespace b2f {

// Upper bound of a row norm^2 computed in fp32 (any summation order): inflate by 2^-12 so the
// stored value is >= the true sum of squares.
__device__ __forceinline__ float norm2_upper(float ss) { return ss * (1.0f + 2.44140625e-4f); }

// Squared bf16 rounding error of four components: (x - bf16_rn(x)) is exact in fp32 (the difference
// of two floats within half a bf16 ulp of each other), so only the squares and the sum round.
__device__ __forceinline__ float bf16_err2(float4 v) {
  const float dx = v.x - __bfloat162float(__float2bfloat16_rn(v.x));
  const float dy = v.y - __bfloat162float(__float2bfloat16_rn(v.y));
  const float dz = v.z - __bfloat162float(__float2bfloat16_rn(v.z));
  const float dw = v.w - __bfloat162float(__float2bfloat16_rn(v.w));
  return dx * dx + dy * dy + dz * dz + dw * dw;
}

__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}

// One warp per row: write the bf16 shadow row (round-to-nearest, K-block-major tiled layout — 16
// lanes fill one 128-byte K-block segment) and fold the row's norm^2 bound into maxnorm2_bits[0]
// and the squared norm of its bf16 rounding error ||p - bf16(p)||^2 into maxnorm2_bits[1]
// (float bits; valid because the values are non-negative).  The second bound is what makes the
// prefilter margin data-dependent and ~2x tighter than the worst case 2^-9 * ||p||.
// HBM traffic per row: 3072 B read + 1536 B written.
__global__ void __launch_bounds__(256) convert_rows_kernel(const float* __restrict__ x32,
                                                           __nv_bfloat16* __restrict__ x16,
                                                           int64_t row0, int64_t n_rows,
                                                           unsigned int* __restrict__ maxnorm2_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float wmax = 0.f, emax = 0.f;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(x32 + (row0 + r) * kD);
    float ss = 0.f, es = 0.f;
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      float4 v = ldg_stream_f4(src + lane + 32 * i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      es += bf16_err2(v);
      if (x16) *reinterpret_cast<uint2*>(x16 + shadow_index(row0 + r, 4 * (lane + 32 * i))) = pack_bf16x4(v);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, s);
      es += __shfl_xor_sync(0xffffffffu, es, s);
    }
    wmax = fmaxf(wmax, norm2_upper(ss));
    emax = fmaxf(emax, norm2_upper(es));
  }
  if (lane == 0 && wmax > 0.f) atomicMax(maxnorm2_bits, __float_as_uint(wmax));
  if (lane == 0 && emax > 0.f) atomicMax(maxnorm2_bits + 1, __float_as_uint(emax));
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
  float wmax = 0.f, emax = 0.f;
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
    float fs = 0.f, es = 0.f;
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      float4 v;
      v.x = __fmul_rn(static_cast<float>(comp[i][0]), inv);
      v.y = __fmul_rn(static_cast<float>(comp[i][1]), inv);
      v.z = __fmul_rn(static_cast<float>(comp[i][2]), inv);
      v.w = __fmul_rn(static_cast<float>(comp[i][3]), inv);
      fs += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      es += bf16_err2(v);
      dst32[lane + 32 * i] = v;
      if (x16) *reinterpret_cast<uint2*>(x16 + shadow_index(dst_row0 + r, 4 * (lane + 32 * i))) = pack_bf16x4(v);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      fs += __shfl_xor_sync(0xffffffffu, fs, s);
      es += __shfl_xor_sync(0xffffffffu, es, s);
    }
    wmax = fmaxf(wmax, norm2_upper(fs));
    emax = fmaxf(emax, norm2_upper(es));
  }
  if (lane == 0 && wmax > 0.f) atomicMax(maxnorm2_bits, __float_as_uint(wmax));
  if (lane == 0 && emax > 0.f) atomicMax(maxnorm2_bits + 1, __float_as_uint(emax));
}

// Query preparation: one warp per (padded) query row.  Writes the bf16 copy used by the tensor
// path (zero rows for q >= nq), an upper bound of ||q|| and one of ||q - bf16(q)||.
__global__ void __launch_bounds__(128) prep_queries_kernel(const float* __restrict__ q32, int nq,
                                                           int nq_pad,
                                                           __nv_bfloat16* __restrict__ q16,
                                                           float* __restrict__ qnorm,
                                                           float* __restrict__ qerr) {
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
  float ss = 0.f, es = 0.f;
#pragma unroll
  for (int i = 0; i < kF4PerLane; ++i) {
    float4 v = __ldg(src + lane + 32 * i);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    es += bf16_err2(v);
    dst[lane + 32 * i] = pack_bf16x4(v);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, s);
    es += __shfl_xor_sync(0xffffffffu, es, s);
  }
  if (lane == 0) {
    qnorm[q] = __fsqrt_ru(norm2_upper(ss));
    qerr[q] = __fsqrt_ru(norm2_upper(es));
  }
}

}  // namespace b2f
