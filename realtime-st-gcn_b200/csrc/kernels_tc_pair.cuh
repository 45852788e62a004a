// Temporal convolution on CTA PAIRS (tcgen05 cta_group::2): two SMs of a TPC run ONE M = 256
// MMA per instruction -- each CTA contributes its own 128-row tile (its own item) as the A
// operand and HALF of every weight tile as the B operand -- so per CTA the weight tiles cost half
// the shared memory (twice the stages in flight), half the shared-memory read bandwidth per MMA
// and half the L2 traffic.  ncu + cycle counters on the single-CTA kernel showed
// exactly those three as its limits: N <= 128 MMAs shared-memory bound, C = 256 weight streaming
// starved with 3 stages.
//
// Protocol (all barriers live at the same shared-memory offsets in both CTAs):
//   * every CTA runs its own TMA producers (input window of ITS item; rows rank*C/2.. of each
//     weight tile) and its own 8 epilogue warps on its own TMEM;
//   * both CTAs' TMA loads (cta_group::2 form) signal the LEADER's full barriers, which expect
//     the bytes of both halves;
//   * the leader's warp 1 issues the MMAs for the pair and releases stages / publishes
//     accumulators in BOTH CTAs with multicast commits;
//   * epilogue warps of both CTAs return accumulator buffers to the leader's barrier.
#pragma once
#include "kernels_tc.cuh"

namespace stgcn {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.
// Default (.release.cta) semantics on purpose: what the epilogue hands back is a TMEM accumulator buffer, ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync, not generic-proxy memory.  The .release.cluster form
// compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the arrive, i.e. every epilogue warp waited once per
// item for its own output stores to become visible GPU-wide (ncu: 0.9 "membar" stall cycles per issued instruction
// in k_tcn_tc2p<64>).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t mapa_cta(uint32_t addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(cta));
  return remote;
}
// arrive + expect `bytes` on a barrier given by its shared::cluster address (the leader's)
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, completion bytes are signalled
// on the barrier `bar_cluster` (a shared::cluster address: the leader's barrier)
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap *m, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap *m, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// NOTE: default (.acquire.cta) semantics on purpose.  With .acquire.cluster every successful wait
// is followed by CCTL.IVALL -- an invalidation of the whole L1 -- issued from the MMA warp's wait
// loop; ncu showed it wiping the epilogue's parameter tables (L1 hit rate 45-60 %).  The waiter
// (MMA issuer) consumes only async-proxy data (TMA -> shared memory, tcgen05 -> TMEM), ordered by
// the barrier completion itself and tcgen05.fence, as in CUTLASS's ClusterBarrier::wait.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// the MMA issuer's waits: same bounded, hardware-suspended wait as everywhere else (a pure spin of
// the 32 lanes took issue slots from the epilogue warps sharing the scheduler, and power)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_bf16_2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all previously issued MMAs completed) on the barrier at this offset in both CTAs
__device__ __forceinline__ void umma_commit_2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

// SETS: sets of eight epilogue warps.  With two sets (items of two tiles) each set normalises one
// tile, so the item's epilogue takes half as long -- the epilogue is latency-bound, not issue-bound.
template <int C, int SETS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (3 + 4 * kEpiNH * SETS), 1)
    k_tcn_tc2p(const __grid_constant__ CUtensorMap tm_u0, const __grid_constant__ CUtensorMap tm_u1,
               const __grid_constant__ CUtensorMap tm_w, const TcnTc2Params p) {
  constexpr int kBHalf = (C / 2) * 128;           // this CTA's half of a weight tile: [C/2 rows][64 ch] bf16
  constexpr int kTmemCols = 512;
  const int TB = p.tb;
  const int buf_cols = p.NT * C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int S = p.b_stages;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + 2 * p.a_stage_bytes;
  const uint32_t sPart = sB + S * kBHalf;
  const uint32_t sPatch = sPart + SETS * kPartBytes;
  const uint32_t sBar = sPatch + SETS * kPatchTotal;
  const uint32_t bFullA = sBar, bEmptyA = sBar + 16, bPeerA = sBar + 32, bTmemFull = sBar + 48, bTmemEmpty = sBar + 64;
  const uint32_t bFullB = sBar + 80, bEmptyB = bFullB + 8 * S, bPeerB = bEmptyB + 8 * S;
  const uint32_t sTmemPtr = bPeerB + 8 * S;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  float *s_part = reinterpret_cast<float *>(gen_base + (sPart - smem_base));
  uint8_t *s_patch = gen_base + (sPatch - smem_base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int KC = C / 64;
  const int RT = p.FT * p.V;
  const int iters = (p.items + 2 * npairs - 1 - 2 * pair) / (2 * npairs);   // same for both CTAs of the pair
  const bool dbg = (p.epi.debug & 4) && blockIdx.x == 0;
  long long d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_u0);
    tma_prefetch_desc(&tm_u1);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bFullA + 8 * i, 2);                     // one arrive.expect_tx per CTA of the pair (leader's is used)
      mbar_init(bEmptyA + 8 * i, 1);
      mbar_init(bPeerA + 8 * i, 1);
      mbar_init(bTmemFull + 8 * i, 1);
      mbar_init(bTmemEmpty + 8 * i, 2 * 4 * kEpiNH * SETS);   // the epilogue warps of BOTH CTAs
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(bFullB + 8 * i, 2);
      mbar_init(bEmptyB + 8 * i, 1);
      mbar_init(bPeerB + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(sTmemPtr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ---- input-window producer for this CTA's own items ----
    if (lane == 0) {
      int as = 0, a_ph = 0;
      for (int it = 0; it < iters; ++it) {
        const int item = 2 * (pair + it * npairs) + (int)rank;
        const bool valid = item < p.items;
        const int n = valid ? item / p.groups_per_trial : 0;
        const int f0 = valid ? (item - n * p.groups_per_trial) * p.NT * p.FT : 0;
        const int n_ld = valid ? n : 0x3fffff;     // out-of-range trial: TMA zero-fills (dummy item of an odd tail)
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap) {
            mbar_wait(bEmptyA + 8 * as, a_ph ^ 1);
            const uint32_t lbar = mapa_cta(bFullA + 8 * as, 0);        // the LEADER's full barrier
            mbar_expect_tx_cluster(lbar, (uint32_t)(p.load_bytes[0] + (p.n_loads > 1 ? p.load_bytes[1] : 0)));
            tma_load_5d_2sm(sA + as * p.a_stage_bytes + p.load_row[0] * 128, &tm_u0, lbar, kc * 64, 0,
                            f0 + p.load_f0[0], n_ld, ap);
            if (p.n_loads > 1)
              tma_load_5d_2sm(sA + as * p.a_stage_bytes + p.load_row[1] * 128, &tm_u1, lbar, kc * 64, 0,
                              f0 + p.load_f0[1], n_ld, ap);
            as ^= 1;
            if (as == 0) a_ph ^= 1;
          }
      }
    }
  } else if (warp == 2) {
    // ---- weight producer: this CTA's half (rows rank*C/2 ..) of every weight tile ----
    if (lane == 0) {
      int bs = 0, b_ph = 0;
      for (int it = 0; it < iters; ++it)
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap) {
            const int nb = (ap == 0) ? p.planes : 1;
            for (int j = 0; j < p.G; ++j)
              for (int bp = 0; bp < nb; ++bp) {
                mbar_wait(bEmptyB + 8 * bs, b_ph ^ 1);
                const uint32_t lbar = mapa_cta(bFullB + 8 * bs, 0);
                mbar_expect_tx_cluster(lbar, kBHalf);
                tma_load_4d_2sm(sB + bs * kBHalf, &tm_w, lbar, kc * 64, (int)rank * (C / 2), j, bp);
                if (++bs == S) { bs = 0; b_ph ^= 1; }
              }
          }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(256, C);
    int a_s = 0, a_ph = 0, b_s = 0, b_ph = 0, buf = 0, t_ph = 0;
    if (rank == 0) {
      // ---- leader: MMA issuer for the pair ----
      DbgTimer tall(dbg);
      for (int it = 0; it < iters; ++it) {
        { DbgTimer tm(dbg); mbar_wait_cluster(bTmemEmpty + 8 * buf, t_ph ^ 1); tm.stop(d0); }
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * buf_cols;
        uint32_t acc = 0;
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap) {
            { DbgTimer tm(dbg); mbar_wait_cluster(bFullA + 8 * a_s, a_ph); tm.stop(d1); }
            tc_fence_after();
            const uint32_t a_lo0 = umma_desc_lo(sA + a_s * p.a_stage_bytes);
            const int nb = (ap == 0) ? p.planes : 1;
            for (int j = 0; j < p.G; ++j) {
              const uint32_t a_tap = a_lo0 + (uint32_t)(p.tap_row[j] * 8);
              for (int bp = 0; bp < nb; ++bp) {
                { DbgTimer tm(dbg); mbar_wait_cluster(bFullB + 8 * b_s, b_ph); tm.stop(d2); }
                tc_fence_after();
                if (elect_one()) {
                  const uint32_t b_lo = umma_desc_lo(sB + b_s * kBHalf);
                  for (int m = 0; m < p.NT; ++m) {
                    const uint32_t a_lo = a_tap + (uint32_t)(m * RT * 8);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_bf16_2(tacc + m * C, umma_desc_join(a_lo + 2 * k), umma_desc_join(b_lo + 2 * k), idesc,
                                  acc | (uint32_t)k);
                  }
                  umma_commit_2(bEmptyB + 8 * b_s);
                }
                __syncwarp();
                acc = 1;
                if (++b_s == S) { b_s = 0; b_ph ^= 1; }
              }
            }
            if (elect_one()) umma_commit_2(bEmptyA + 8 * a_s);
            __syncwarp();
            a_s ^= 1;
            if (a_s == 0) a_ph ^= 1;
          }
        if (elect_one()) umma_commit_2(bTmemFull + 8 * buf);
        __syncwarp();
        if (++buf == TB) { buf = 0; t_ph ^= 1; }
      }
      tall.stop(d3);
      if (lane == 0) {
        dbg_flush(dbg, 0, d0); dbg_flush(dbg, 1, d1); dbg_flush(dbg, 2, d2); dbg_flush(dbg, 3, d3);
        dbg_flush(dbg, 9, d4); dbg_flush(dbg, 11, iters);
      }
    }
  } else {
    // ---- epilogue (both CTAs, own TMEM, own item) ----
    const int ew = warp - 3;                       // 0 .. 8*SETS-1
    const int set = ew >> 3;                       // SETS == 2: tile of the item this set normalises
    const int q = warp & 3;
    const int h = (ew & 7) >> 2;
    const int r = q * 32 + lane;
    const int fr = r / p.V, w = r - fr * p.V;
    float *s_part_set = s_part + set * (kPartBytes / 4);
    uint8_t *patch = s_patch + ew * kPatchBytes;
    const int m_first = SETS == 2 ? set : 0, m_last = SETS == 2 ? set + 1 : p.NT;
    int buf = 0, t_ph = 0, par = 0;
    for (int it = 0; it < iters; ++it) {
      const int item = 2 * (pair + it * npairs) + (int)rank;
      const bool valid = item < p.items;
      const int n = valid ? item / p.groups_per_trial : 0;
      const int f0 = valid ? (item - n * p.groups_per_trial) * p.NT * p.FT : 0;
      if (valid && p.epi.res && r < RT) {
        for (int m = m_first; m < m_last; ++m) {
          const int t = f0 + m * p.FT + fr;
          if (t < p.T_out) {
            const char *rp = reinterpret_cast<const char *>(
                p.epi.res + (n * p.res_trial_rows + (long long)t * p.V + w) * C + h * (C / kEpiNH));
#pragma unroll
            for (int o = 0; o < (C / kEpiNH) * 4; o += 128) prefetch_l2(rp + o);
          }
        }
      } else if (valid && p.epi.res_hi && r < RT) {
        for (int m = m_first; m < m_last; ++m) {
          const int t = f0 + m * p.FT + fr;
          if (t < p.T_out) {
            const long long ro = (n * p.res_trial_rows + (long long)t * p.V + w) * C + h * (C / kEpiNH);
            const char *rh = reinterpret_cast<const char *>(p.epi.res_hi + ro);
            const char *rl = reinterpret_cast<const char *>(p.epi.res_lo ? p.epi.res_lo + ro : p.epi.res_hi + ro);
#pragma unroll
            for (int o = 0; o < (C / kEpiNH) * 2; o += 128) {
              prefetch_l2(rh + o);
              prefetch_l2(rl + o);
            }
          }
        }
      }
      mbar_wait(bTmemFull + 8 * buf, t_ph);
      tc_fence_after();
#pragma unroll 1
      for (int m = m_first; m < m_last; ++m, par ^= 1) {
        const int t = f0 + m * p.FT + fr;
        const bool row_ok = valid && (r < RT) && (t < p.T_out);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * buf_cols + m * C);
        const long long row = ((long long)n * p.T_out + t) * p.V + w;
        const long long row_res = n * p.res_trial_rows + (long long)t * p.V + w;
        ln_epilogue_tile<C, kEpiNH, true>(p.epi, taddr, r, RT, p.V, fr, w, row_ok, row_res, row, s_part_set, par, h,
                                          patch, 1 + set);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(bTmemEmpty + 8 * buf, 0);   // leader's barrier (local when rank == 0)
      if (++buf == TB) { buf = 0; t_ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // neither CTA may leave while the pair's MMAs / arrives are in flight
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

template <int C>
int launch_tcn_pair(const CUtensorMap &tm_u0, const CUtensorMap &tm_u1, const CUtensorMap &tm_wh,
                    const TcnTc2Params &p, int grid, int smem, int sets, cudaStream_t st) {
  if constexpr (C <= 128) {
    if (sets == 2) {
      STGCN_CUDA_OK(cudaFuncSetAttribute(k_tcn_tc2p<C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      k_tcn_tc2p<C, 2><<<grid, 32 * (3 + 8 * kEpiNH), smem, st>>>(tm_u0, tm_u1, tm_wh, p);
      return 0;
    }
  }
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_tcn_tc2p<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_tcn_tc2p<C, 1><<<grid, kTcn2Threads, smem, st>>>(tm_u0, tm_u1, tm_wh, p);
  return 0;
}

}  // namespace tc
}  // namespace stgcn
