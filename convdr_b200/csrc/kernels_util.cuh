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

// Column means of rows [row0, row0 + n_rows) in two deterministic steps (fixed summation order, so the
// same rows always give the same centre and therefore the same shadow bits):
//   col_sum_partial_kernel : block b sums rows b, b + gridDim.x, ... (192 threads, one float4 column chunk each)
//   col_mean_final_kernel  : sums the partials in block order and divides by n_rows
__global__ void __launch_bounds__(kRowF4) col_sum_partial_kernel(const float* __restrict__ x32, int64_t row0,
                                                                 int64_t n_rows, float* __restrict__ partial) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const float4 v = ldg_stream_f4(reinterpret_cast<const float4*>(x32 + (row0 + r) * kD) + threadIdx.x);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(partial + static_cast<int64_t>(blockIdx.x) * kD)[threadIdx.x] = acc;
}
__global__ void __launch_bounds__(256) col_mean_final_kernel(const float* __restrict__ partial, int n_partials,
                                                             int64_t n_rows, float* __restrict__ mu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= kD) return;
  float acc = 0.f;
  for (int b = 0; b < n_partials; ++b) acc += partial[static_cast<int64_t>(b) * kD + c];
  mu[c] = acc / static_cast<float>(n_rows);
}

// One warp per row: write the bf16 shadow row of the CENTRED vector d = fl32(x - mu) (round-to-nearest,
// K-block-major tiled layout — 16 lanes fill one 128-byte K-block segment) and fold into maxnorm2_bits
// (float bits; valid because the values are non-negative):
//   [0] an upper bound of max ||x||^2            (uncentred: the fp32 scan engine's margin)
//   [1] an upper bound of max ||(x - mu) - bf16(d)||^2, the rounding error of the shadow row INCLUDING the
//       rounding of the fp32 subtraction (|d - (x - mu)| <= 2^-24 |d| per component): what makes the
//       prefilter margin data-dependent and ~2x tighter than the worst case 2^-8 * ||d||
//   [2] an upper bound of max ||x - mu||^2
// mu is the shard's centre (all zeros when centring is off: d = x exactly, the bounds are the uncentred
// ones).  Inner products against centred rows differ from the true ones by the per-query constant q.mu,
// so every threshold comparison of a pass can live in centred space; the exact rescoring reads x32.
// HBM traffic per row: 3072 B read + 1536 B written.
__global__ void __launch_bounds__(256) convert_rows_kernel(const float* __restrict__ x32,
                                                           __nv_bfloat16* __restrict__ x16,
                                                           int64_t row0, int64_t n_rows,
                                                           unsigned int* __restrict__ maxnorm2_bits,
                                                           const float* __restrict__ mu) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  float4 m4[kF4PerLane];
#pragma unroll
  for (int i = 0; i < kF4PerLane; ++i) m4[i] = __ldg(reinterpret_cast<const float4*>(mu) + lane + 32 * i);
  float wmax = 0.f, emax = 0.f, cmax = 0.f;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(x32 + (row0 + r) * kD);
    float ss = 0.f, es = 0.f, cs = 0.f;
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      const float4 v = ldg_stream_f4(src + lane + 32 * i);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      float4 d;
      d.x = __fsub_rn(v.x, m4[i].x); d.y = __fsub_rn(v.y, m4[i].y); d.z = __fsub_rn(v.z, m4[i].z); d.w = __fsub_rn(v.w, m4[i].w);
      cs += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
      es += bf16_err2(d);
      if (x16) *reinterpret_cast<uint2*>(x16 + shadow_index(row0 + r, 4 * (lane + 32 * i))) = pack_bf16x4(d);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, s);
      es += __shfl_xor_sync(0xffffffffu, es, s);
      cs += __shfl_xor_sync(0xffffffffu, cs, s);
    }
    const float cn = __fsqrt_ru(norm2_upper(cs));                         // >= ||d||
    const float e = __fadd_ru(__fsqrt_ru(norm2_upper(es)), __fmul_ru(cn, 6.0e-8f));   // + 2^-24 ||d|| (subtraction rounding)
    wmax = fmaxf(wmax, norm2_upper(ss));
    emax = fmaxf(emax, __fmul_ru(e, e));
    cmax = fmaxf(cmax, __fmul_ru(__fmul_ru(cn, cn), 1.0000002f));         // ||x - mu|| <= (1 + 2^-24) ||d||
  }
  if (lane == 0 && wmax > 0.f) atomicMax(maxnorm2_bits, __float_as_uint(wmax));
  if (lane == 0 && emax > 0.f) atomicMax(maxnorm2_bits + 1, __float_as_uint(emax));
  if (lane == 0 && cmax > 0.f) atomicMax(maxnorm2_bits + 2, __float_as_uint(cmax));
}

// Synthetic rows (see include/b2f.h b2f_add_synthetic).  One warp per row; lane l produces
// float4 chunks l, l+32, ..., l+160; chunk c of row r is Philox4x32-10(counter = (r_lo, r_hi, c, 0),
// key = (seed_lo ^ stream_lo, seed_hi ^ stream_hi ^ 0x5eed)).  With mean_shift = M > 0 every component t
// is shifted by sign_t * M before the normalisation, sign_t = +-1 from the low bit of word t of
// Philox(counter = (0xffffffff, 0xffffffff, c, 0x6d65616e), key = (seed_lo, seed_hi ^ 0x5eed)) — one fixed
// direction per seed, shared by every stream (passages and queries), which gives LayerNorm-like
// embeddings with a common mean (cos(p, p') ~ M^2 / (M^2 + 148^2); M = 443 -> 0.9).  Integer arithmetic up
// to the final scale, so CPU and GPU agree bit for bit.  Writes the fp32 rows only; the shadow and the
// norm bounds come from the same ingest pass as for real rows (convert_rows_kernel).
__global__ void __launch_bounds__(256) synth_rows_kernel(float* __restrict__ x32,
                                                         int64_t dst_row0, int64_t first_row,
                                                         int64_t n_rows, uint32_t k0, uint32_t k1,
                                                         float norm, int mean_shift, uint32_t km0, uint32_t km1) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  int shift[kF4PerLane][4];
#pragma unroll
  for (int i = 0; i < kF4PerLane; ++i) {
    shift[i][0] = shift[i][1] = shift[i][2] = shift[i][3] = 0;
    if (mean_shift) {
      U4 c;
      c.x = 0xffffffffu; c.y = 0xffffffffu; c.z = static_cast<uint32_t>(lane + 32 * i); c.w = 0x6d65616eu;
      const U4 o = philox4x32_10(c, km0, km1);
      shift[i][0] = (o.x & 1u) ? mean_shift : -mean_shift;
      shift[i][1] = (o.y & 1u) ? mean_shift : -mean_shift;
      shift[i][2] = (o.z & 1u) ? mean_shift : -mean_shift;
      shift[i][3] = (o.w & 1u) ? mean_shift : -mean_shift;
    }
  }
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const uint64_t row = static_cast<uint64_t>(first_row + r);
    int comp[kF4PerLane][4];
    long long ss = 0;
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      U4 c;
      c.x = static_cast<uint32_t>(row);
      c.y = static_cast<uint32_t>(row >> 32);
      c.z = static_cast<uint32_t>(lane + 32 * i);
      c.w = 0u;
      U4 o = philox4x32_10(c, k0, k1);
      comp[i][0] = synth_component(o.x) + shift[i][0];
      comp[i][1] = synth_component(o.y) + shift[i][1];
      comp[i][2] = synth_component(o.z) + shift[i][2];
      comp[i][3] = synth_component(o.w) + shift[i][3];
      ss += static_cast<long long>(comp[i][0] * comp[i][0] + comp[i][1] * comp[i][1]) +
            static_cast<long long>(comp[i][2] * comp[i][2] + comp[i][3] * comp[i][3]);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    // ss <= 768 * (510 + 4096)^2 < 2^34: converted to fp32 with one rounding, like the host restatement.
    // IEEE sqrt and division: bit-identical to the host restatement.
    const float inv = (ss > 0) ? __fdiv_rn(norm, __fsqrt_rn(static_cast<float>(ss))) : 0.f;
    float4* dst32 = reinterpret_cast<float4*>(x32 + (dst_row0 + r) * kD);
#pragma unroll
    for (int i = 0; i < kF4PerLane; ++i) {
      float4 v;
      v.x = __fmul_rn(static_cast<float>(comp[i][0]), inv);
      v.y = __fmul_rn(static_cast<float>(comp[i][1]), inv);
      v.z = __fmul_rn(static_cast<float>(comp[i][2]), inv);
      v.w = __fmul_rn(static_cast<float>(comp[i][3]), inv);
      dst32[lane + 32 * i] = v;
    }
  }
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
