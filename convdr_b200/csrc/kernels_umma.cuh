// kernels_umma.cuh — the tensor-core scoring engine (sm_100a): a fused score + select kernel.
//
//   HBM (bf16 shadow rows) --TMA, 128B swizzle--> smem ring --tcgen05.mma cta_group::2--> TMEM
//   TMEM --tcgen05.ld--> registers --threshold filter--> candidate lists (global atomics, rare)
//
// One CTA pair (2 SMs) owns a tile of 256 passage rows (128 per CTA, the MMA M dimension) and all
// queries of the pass (MMA N <= 192, split in halves across the pair's shared memory, resident for
// the whole launch).  K = 768 is walked in 12 blocks of 64 bf16 (one 128-byte swizzle span).
// The [rows x queries] score tile never leaves the SM: accumulators live in TMEM (2 stages of 256
// columns) and the epilogue keeps only scores >= the per-query threshold.
//
// Replaces the arithmetic of `index.search` (reference drivers/run_convdr_inference.py:182) —
// FAISS's `nq >= 20` branch: blocked sgemm + per-row heap (upstream utils/distances.cpp), and
// FAISS-GPU's cuBLAS GEMM + BlockSelect that materialise score tiles in HBM.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace b2f {

constexpr int kUmmaThreads = 256;
constexpr int kBlockK = 64;                    // bf16 per K block = 128 bytes
constexpr int kNumKBlocks = kD / kBlockK;      // 12
constexpr int kTileRowsCta = 128;
constexpr int kTileRows = 256;                 // per CTA pair
constexpr int kStageBytes = kTileRowsCta * kBlockK * 2;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kUmmaMaxQ = 192;
constexpr int kAccStride = 256;                // TMEM columns per accumulator stage
constexpr int kTmemCols = 512;
constexpr int kUmmaTailBytes = 2048;           // barriers + tmem pointer + thresholds
constexpr int kSmemLimit = 232448;             // 227 KB opt-in maximum per CTA

struct UmmaArgs {
  int64_t n_rows;             // valid rows of the shard
  int tile_begin, tile_end;   // pair tiles of 256 rows
  int n_cols;                 // MMA N: padded query count, multiple of 16, 16..192
  int nq;                     // valid queries (<= n_cols)
  int stages;                 // smem ring depth (2..8)
  int dense;                  // 1: store every score at slot (row - dense_row0)
  int64_t dense_row0;
  uint64_t* cand;             // [nq][C]
  int* cnt;                   // [nq]
  int C;
  const float* tau;           // [nq]
  int* ovf;                   // [nq]
  int* err;                   // device error flag (barrier timeout)
};

inline int umma_q_bytes(int n_cols) { return kNumKBlocks * (n_cols / 2) * 128; }
inline int umma_stages(int n_cols) {
  int s = (kSmemLimit - kUmmaTailBytes - 1024 /*alignment slack*/ - umma_q_bytes(n_cols)) / kStageBytes;
  return s > kMaxStages ? kMaxStages : s;
}
inline int umma_smem_bytes(int n_cols, int stages) {
  return umma_q_bytes(n_cols) + stages * kStageBytes + kUmmaTailBytes + 1024;
}

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Remote arrive on a barrier of the peer CTA.  Deliberately the plain form (default .release.cta):
// `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR, which cost ~1300 cycles per TMA stage
// when the peer's producer executed it (profiles/r01: the whole pipeline ran at that pace).  The
// data these barriers guard travels through the async proxy (TMA complete_tx) or is ordered by
// tcgen05.fence::before_thread_sync, not by this arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Bounded wait: a protocol bug must end in a trap (reported as a CUDA error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if ((spin & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 6000000000ll) {  // ~3 s at 2 GHz
        if (err) atomicExch(err, 1);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* tmap,
                                                uint32_t leader_bar_cluster, int c0, int c1,
                                                uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar_cluster), "r"(c0), "r"(c1),
        "l"(hint)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once) on the barrier at this smem offset in BOTH CTAs of the pair when all previously
// issued MMAs have completed.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ uint32_t tmem_ld_x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand descriptor (PTX ISA "tcgen05 shared memory descriptor"):
// start address >> 4 in bits [0,14); stride-dimension byte offset (8 rows * 128 B = 1024) >> 4 in
// bits [32,46); descriptor version 1 in bits [46,48); swizzle mode 2 (128B) in bits [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu) | (static_cast<uint64_t>(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::f16: D = fp32 (bits [4,6) = 1), A = B = bf16 (bits [7,10), [10,13)
// = 1), both K-major (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

constexpr uint64_t kHintEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kHintEvictLast = 0x14F0000000000000ull;

// ------------------------------------------------------------------------------------------
// The kernel.  Grid = 2 * (number of CTA pairs), cluster (2,1,1), 256 threads:
//   warp 0 lane 0 : TMA producer (both CTAs: own 128 rows; own half of the queries once)
//   warp 1 lane 0 : MMA issuer (leader CTA only)
//   warp 2        : TMEM allocation / release
//   warps 4..7    : epilogue — TMEM lane quarter (warp % 4), one passage row per thread
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kUmmaThreads, 1)
    umma_score_select_kernel(const __grid_constant__ CUtensorMap tmap_p,
                             const __grid_constant__ CUtensorMap tmap_q, const UmmaArgs a) {
  extern __shared__ unsigned char umma_smem_raw[];
  const uint32_t raw = smem_u32(umma_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 1024-byte alignment for the 128B swizzle atoms
  unsigned char* base_ptr = umma_smem_raw + (base - raw);

  const int n_half = a.n_cols >> 1;
  const uint32_t q_kblock_bytes = static_cast<uint32_t>(n_half) * 128u;
  const uint32_t q_bytes = kNumKBlocks * q_kblock_bytes;
  const uint32_t smem_q = base;
  const uint32_t smem_a = base + q_bytes;
  const uint32_t tail = smem_a + static_cast<uint32_t>(a.stages) * kStageBytes;
  const uint32_t bar_full = tail;                      // [kMaxStages]
  const uint32_t bar_empty = tail + 8 * kMaxStages;    // [kMaxStages]
  const uint32_t bar_qfull = tail + 16 * kMaxStages;
  const uint32_t bar_tfull = bar_qfull + 8;            // [2]
  const uint32_t bar_tempty = bar_tfull + 16;          // [2]
  unsigned char* tail_ptr = base_ptr + q_bytes + static_cast<uint32_t>(a.stages) * kStageBytes;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tail_ptr + 16 * kMaxStages + 8 + 16 + 16);
  float* tau_s = reinterpret_cast<float*>(tail_ptr + 256);  // [kUmmaMaxQ]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_p);
    prefetch_tmap(&tmap_q);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
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
  for (int i = threadIdx.x; i < kUmmaMaxQ; i += kUmmaThreads)
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
    for (int tile = a.tile_begin + pair; tile < a.tile_end; tile += npairs) {
      // shadow layout: CTA tile t = 2*tile + rank, K-block kb -> 128 consecutive 128-byte rows
      const int blk0 = (tile * 2 + static_cast<int>(cta_rank)) * kNumKBlocks;
      for (int kb = 0; kb < kNumKBlocks; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1u, a.err);
        const uint32_t full_leader = mapa_u32(bar_full + 8 * stage, 0);
        if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, 2u * kStageBytes);
        else mbar_arrive_cluster(full_leader);
        tma_load_2d_2sm(smem_a + stage * kStageBytes, &tmap_p, full_leader, 0, (blk0 + kb) * kTileRowsCta,
                        kHintEvictFirst);
        if (++stage == static_cast<uint32_t>(a.stages)) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA) =====================
    const uint32_t idesc = umma_idesc_bf16(256, a.n_cols);
    mbar_wait(bar_qfull, 0, a.err);
    tc_fence_after();
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int tile = a.tile_begin + pair; tile < a.tile_end; tile += npairs, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      mbar_wait(bar_tempty + 8 * as, aph ^ 1u, a.err);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * kAccStride;
      for (int kb = 0; kb < kNumKBlocks; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase, a.err);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_a + stage * kStageBytes);
        const uint64_t bdesc = umma_desc_sw128(smem_q + kb * q_kblock_bytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)  // UMMA K = 16 bf16 = 32 bytes = 2 descriptor units
          umma_bf16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_pair(bar_empty + 8 * stage);  // frees this smem stage in both CTAs
        if (kb == kNumKBlocks - 1) umma_commit_pair(bar_tfull + 8 * as);  // accumulator ready
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
      const int64_t row = static_cast<int64_t>(tile) * kTileRows + cta_rank * kTileRowsCta + ew * 32 + lane;
      const bool row_ok = row < a.n_rows;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * kAccStride;
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
