// kernels_umma_ss.cuh — tensor-core scoring engine, "smem-stationary queries" variant (sm_100a).
//
//   passages : bf16 shadow tiles  HBM --TMA prefetch--> L2 --TMA (128B swizzle)--> smem ring : A operand
//   queries  : bf16, resident in shared memory for the whole launch (split across the CTA pair): B
//   scores   : tcgen05.mma cta_group::2, M = 256 passage rows, N = padded query count (16..192)
//   select   : tcgen05.ld --> one PASSAGE ROW per epilogue thread, per-query thresholds in smem,
//              survivors appended with one batched atomic per 16-query chunk
//
// Compared with the TMEM-stationary variant (kernels_umma.cuh) the MMA shape follows the query
// count exactly (N = 176 for 173 queries instead of M = 256), i.e. 31 % fewer tensor cycles per row
// — what matters when the GPU is power-capped and the other variant becomes tensor-bound — at the
// price of the queries occupying 135 KB of each CTA's shared memory, which leaves a 5-stage ring.
// The ring is too shallow to cover HBM latency, so the producer prefetches the next tile into L2.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "kernels_umma.cuh"

namespace b2f {

constexpr int kSsTileRowsCta = 128;
constexpr int kSsTileRows = 256;                 // per CTA pair = MMA M
constexpr int kSsStageBytes = kSsTileRowsCta * kBlockK * 2;  // 16 KB: 128 rows of one K-block
constexpr int kSsMaxStages = 8;
constexpr int kSsMaxQ = 192;
constexpr int kSsAccStride = 256;                // TMEM columns per accumulator stage
constexpr int kSsTailBytes = 2048;               // barriers + tmem pointer + thresholds
constexpr int kSsSmemLimit = 232448;

struct UmmaSsArgs {
  int64_t n_rows;             // valid rows of the shard
  int tile_begin, tile_end;   // pair tiles of 256 rows
  int n_cols;                 // MMA N: padded query count, multiple of 16, 16..192
  int nq;                     // valid queries (<= n_cols)
  int stages;                 // smem ring depth
  int prefetch;               // D > 0: keep the TMA L2 prefetch D tiles ahead of the ring; 0: off
  int dense;                  // 1: store every score at slot (row - dense_row0)
  int64_t dense_row0;
  uint64_t* cand;             // [nq][C] flat lists
  int* cnt;                   // [nq]
  int C;
  const float* tau;           // [nq]
  int* ovf;                   // [nq]
  int* err;
};

inline int umma_ss_q_bytes(int n_cols) { return kNumKBlocks * (n_cols / 2) * 128; }
inline int umma_ss_stages(int n_cols) {
  int s = (kSsSmemLimit - kSsTailBytes - 1024 - umma_ss_q_bytes(n_cols)) / kSsStageBytes;
  return s > kSsMaxStages ? kSsMaxStages : s;
}
inline int umma_ss_smem_bytes(int n_cols, int stages) {
  return umma_ss_q_bytes(n_cols) + stages * kSsStageBytes + kSsTailBytes + 1024;
}

__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1) : "memory");
}

// ------------------------------------------------------------------------------------------
// The kernel.  Grid = 2 * (number of CTA pairs), cluster (2,1,1), 256 threads:
//   warp 0 lane 0 : TMA producer (both CTAs: own 128 rows; own half of the queries once)
//   warp 1 lane 0 : MMA issuer (leader CTA only)
//   warp 2        : TMEM allocation / release
//   warps 4..7    : epilogue — TMEM lane quarter (warp % 4), one passage row per thread
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kUmmaThreads, 1)
    umma_ss_score_select_kernel(const __grid_constant__ CUtensorMap tmap_p,
                                const __grid_constant__ CUtensorMap tmap_pf,
                                const __grid_constant__ CUtensorMap tmap_q, const UmmaSsArgs a) {
  extern __shared__ unsigned char umma_smem_raw[];
  const uint32_t raw = smem_u32(umma_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 1024-byte alignment for the 128B swizzle atoms
  unsigned char* base_ptr = umma_smem_raw + (base - raw);

  const int n_half = a.n_cols >> 1;
  const uint32_t q_kblock_bytes = static_cast<uint32_t>(n_half) * 128u;
  const uint32_t q_bytes = kNumKBlocks * q_kblock_bytes;
  const uint32_t smem_q = base;
  const uint32_t smem_a = base + q_bytes;
  const uint32_t tail = smem_a + static_cast<uint32_t>(a.stages) * kSsStageBytes;
  const uint32_t bar_full = tail;                      // [kSsMaxStages]
  const uint32_t bar_empty = tail + 8 * kSsMaxStages;    // [kSsMaxStages]
  const uint32_t bar_qfull = tail + 16 * kSsMaxStages;
  const uint32_t bar_tfull = bar_qfull + 8;            // [2]
  const uint32_t bar_tempty = bar_tfull + 16;          // [2]
  unsigned char* tail_ptr = base_ptr + q_bytes + static_cast<uint32_t>(a.stages) * kSsStageBytes;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tail_ptr + 16 * kSsMaxStages + 8 + 16 + 16);
  float* tau_s = reinterpret_cast<float*>(tail_ptr + 256);  // [kSsMaxQ]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_p);
    prefetch_tmap(&tmap_pf);
    prefetch_tmap(&tmap_q);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kSsMaxStages; ++s) {
      mbar_init(bar_full + 8 * s, 2);   // leader's expect_tx arrive + peer's remote arrive
      mbar_init(bar_empty + 8 * s, 1);  // one multicast commit
    }
    mbar_init(bar_qfull, 2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);   // one multicast commit
      mbar_init(bar_tempty + 8 * s, 8);  // 4 epilogue warps x 2 CTAs (leader's copy is the one used)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_ptr_s)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kSsMaxQ; i += kUmmaThreads)
    tau_s[i] = (i < a.nq && !a.dense) ? a.tau[i] : INFINITY;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_s);

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    const uint32_t qfull_leader = mapa_u32(bar_qfull, 0);
    if (leader) mbar_arrive_expect_tx(bar_qfull, 2u * q_bytes);
    else mbar_arrive_cluster(qfull_leader);
    for (int kb = 0; kb < kNumKBlocks; ++kb)
      tma_load_2d_2sm(smem_q + kb * q_kblock_bytes, &tmap_q, qfull_leader, kb * kBlockK,
                      static_cast<int>(cta_rank) * n_half, kHintEvictLast);
    uint32_t stage = 0, phase = 0;
    // L2 prefetch: with the queries resident in shared memory only ~80 KB of the ring are left, too
    // little to cover HBM latency at full bandwidth.  The producer therefore asks the TMA unit to
    // pull the NEXT tile of this CTA into L2 (12 contiguous 16 KB chunks) while it loads the current
    // one, so the ring only has to cover L2 latency.
    // a.prefetch = D > 0: stay D tiles ahead.  The D tiles after the first are requested up front; then
    // every ring stage of tile i also requests one 16 KB chunk of tile i + D, so the requests are spread
    // evenly over the stream instead of arriving in bursts of 12.
    auto prefetch_chunk = [&](int tile, int j) {
      if (tile >= a.tile_end) return;
      const int t32 = (tile * 2 + static_cast<int>(cta_rank)) * (kSsTileRowsCta / kShadowTileRows);
      tma_prefetch_2d(&tmap_pf, 0, (t32 * kNumKBlocks + j * 4) * kShadowTileRows);
    };
    const int pfd = a.prefetch;
    for (int d = 0; d < pfd; ++d)
      for (int j = 0; j < kNumKBlocks; ++j) prefetch_chunk(a.tile_begin + pair + d * npairs, j);
    for (int tile = a.tile_begin + pair; tile < a.tile_end; tile += npairs) {
      // shadow layout (common.cuh): this CTA's 128 rows are 4 consecutive 32-row tiles; K-block kb
      // of each is a contiguous 4 KB piece -> 4 TMA boxes per 16 KB stage
      const int t32 = (tile * 2 + static_cast<int>(cta_rank)) * (kSsTileRowsCta / kShadowTileRows);
      for (int kb = 0; kb < kNumKBlocks; ++kb) {
        if (pfd) prefetch_chunk(tile + pfd * npairs, kb);
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u, a.err);
        const uint32_t full_leader = mapa_u32(bar_full + 8 * stage, 0);
        if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2u * kSsStageBytes);
        else mbar_arrive_cluster(full_leader);
#pragma unroll
        for (int j = 0; j < kSsTileRowsCta / kShadowTileRows; ++j)
          tma_load_2d_2sm(smem_a + stage * kSsStageBytes + j * (kShadowTileRows * 128), &tmap_p, full_leader, 0,
                          ((t32 + j) * kNumKBlocks + kb) * kShadowTileRows, kHintEvictFirst);
        if (++stage == static_cast<uint32_t>(a.stages)) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA; the whole warp waits, one elected lane issues) =====
    const uint32_t idesc = umma_idesc_bf16(256, a.n_cols);
    const bool elected = elect_one();
    mbar_wait(bar_qfull, 0, a.err);
    tc_fence_after();
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int tile = a.tile_begin + pair; tile < a.tile_end; tile += npairs, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(bar_tempty + 8 * as, aph ^ 1u, a.err);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * kSsAccStride;
      for (int kb = 0; kb < kNumKBlocks; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase, a.err);
        tc_fence_after();
        if (elected) {
          const uint64_t adesc = umma_desc_sw128(smem_a + stage * kSsStageBytes);
          const uint64_t bdesc = umma_desc_sw128(smem_q + kb * q_kblock_bytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)  // UMMA K = 16 bf16 = 32 bytes = 2 descriptor units
            umma_bf16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(bar_empty + 8 * stage);  // frees this smem stage in both CTAs
          if (kb == kNumKBlocks - 1) umma_commit_pair(bar_tfull + 8 * as);  // accumulator ready
        }
        __syncwarp();
        if (++stage == static_cast<uint32_t>(a.stages)) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: filter + append =====================
    const int ew = warp & 3;
    const uint32_t tempty_leader0 = mapa_u32(bar_tempty, 0);
    const uint32_t lt_mask = (1u << lane) - 1u;
    int it = 0;
    for (int tile = a.tile_begin + pair; tile < a.tile_end; tile += npairs, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * as, aph, a.err);
      tc_fence_after();
      const int64_t row = static_cast<int64_t>(tile) * kSsTileRows + cta_rank * kSsTileRowsCta + ew * 32 + lane;
      const bool row_ok = row < a.n_rows;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * kSsAccStride;
      for (int c0 = 0; c0 < a.n_cols; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(taddr + c0, v);
        tmem_ld_wait();
        if (a.dense) {
          const int64_t slot = row - a.dense_row0;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int q = c0 + j;
            if (q < a.nq)
              a.cand[static_cast<int64_t>(q) * a.C + slot] =
                  row_ok ? pack_cand(__uint_as_float(v[j]), static_cast<uint32_t>(row)) : 0ull;
          }
        } else {
          uint32_t m = 0;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            m |= (__uint_as_float(v[j]) >= tau_s[c0 + j]) ? (1u << j) : 0u;
          if (!row_ok) m = 0;
          const uint32_t any = __reduce_or_sync(0xffffffffu, m);
          if (any) {
            // Rare: some row of this warp beat a threshold in this 16-query chunk.  Reserve the
            // slots of ALL affected queries first (one atomic per query, all in flight together),
            // then write — one L2 round trip per chunk instead of one per query.
            int slot0[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              slot0[j] = 0;
              if ((any >> j) & 1u) {
                const uint32_t b = __ballot_sync(0xffffffffu, (m >> j) & 1u);
                if (lane == __ffs(b) - 1) slot0[j] = atomicAdd(a.cnt + (c0 + j), __popc(b));
              }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if ((any >> j) & 1u) {
                const bool pass = (m >> j) & 1u;
                const uint32_t b = __ballot_sync(0xffffffffu, pass);
                const int s0 = __shfl_sync(0xffffffffu, slot0[j], __ffs(b) - 1);
                if (pass) {
                  const int q = c0 + j;
                  const int slot = s0 + __popc(b & lt_mask);
                  if (slot < a.C)
                    a.cand[static_cast<int64_t>(q) * a.C + slot] = pack_cand(__uint_as_float(v[j]), static_cast<uint32_t>(row));
                  else
                    a.ovf[q] = 1;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader0 + 8 * as);
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace b2f
