// tcgen05 / TMEM / TMA kernels of the ST-GCN forward path (sm_100a).
//
// Arithmetic modes (STGCN_MATH_*):
//   BF16X3: every fp32 operand is split into two bf16 planes, x = hi + lo with
//           hi = bf16(x), lo = bf16(x - hi).  The product is accumulated in fp32 TMEM as
//           hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-18 relative), which keeps
//           the result inside the 1e-4 fp32-parity bound at 3 bf16 MMAs per product.
//   BF16  : single bf16 MMA (hi planes only).
//
// Shared-memory operand layout is the canonical K-major SWIZZLE_128B UMMA layout: rows of
// 64 bf16 (128 B), 8-row groups 1024 B apart, written by TMA with CU_TENSOR_MAP_SWIZZLE_128B.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace stgcn {
namespace tc {

// --------------------------------------------------------------------------- //
// PTX wrappers
// --------------------------------------------------------------------------- //
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware may park the thread until the phase completes
// (or the hint expires) instead of returning immediately.
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap, never as a hung GPU.  Waiters back off
// with nanosleep so that spinning warps do not steal issue slots from the working warps
// (ncu on the first persistent kernel showed ~40% of all issued instructions were wait spins).
template <int kSleepNs = 0>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  // the suspend-time hint lets the hardware park the waiting thread until the phase completes
  // instead of spinning: waiting warps then neither take issue slots from the working warps
  // nor burn power (the pure-spin version ran into the 1 kW power cap at ~1.76 GHz)
  while (!mbar_try_wait_hint(bar, parity, 4000u)) {
    if (kSleepNs > 0) __nanosleep(kSleepNs);
    if ((++spins & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ll) {  // ~2 s at 2 GHz
        printf("stgcn_b200: mbarrier timeout (block %d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x,
               blockIdx.y, threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// store 32 lanes x 16 consecutive fp32 columns (thread i of the warp writes lane base_lane + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
      "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
      "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
      "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
      "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major, 1) | [32,46) SBO>>4 (1024 B)
//   [46,48) version = 1 | [49,52) base offset = 0 (all starts are 1024-B-atom aligned + k*32 B)
//   [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same descriptor split into its constant upper half and the address-dependent lower half,
// so that an issue loop advances a descriptor with one 32-bit add (32 B along K = +2).
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO | version | SWIZZLE_128B
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16);
}
__device__ __forceinline__ uint64_t umma_desc_join(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }

// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Measurement aid (STGCN_DEBUG & 4): cycles spent per pipeline role / wait, summed by CTA 0.
//  0 MMA wait TMEM-empty   1 MMA wait A-full     2 MMA wait B-full      3 MMA thread total
//  4 A/x producer wait     5 B producer wait     6 epilogue wait TMEM   7 epilogue tile work
//  8 transform wait x-full 9 transform wait A-empty 10 transform compute 11 items (CTA 0)
__device__ unsigned long long g_dbg[16];
struct DbgTimer {
  long long t0;
  bool on;
  __device__ __forceinline__ DbgTimer(bool enable) : t0(0), on(enable) {
    if (on) t0 = clock64();
  }
  __device__ __forceinline__ void stop(long long &acc) {
    if (on) acc += clock64() - t0;
  }
};
__device__ __forceinline__ void dbg_flush(bool on, int cat, long long v) {
  if (on) atomicAdd(&g_dbg[cat], (unsigned long long)v);
}

// Streaming 16-B global load: no L1 allocation, so that the per-CTA LayerNorm parameter tables
// (re-read for every tile) stay L1-resident next to the row streams (residual, FIFO state).
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// --------------------------------------------------------------------------- //
// operand preparation
// --------------------------------------------------------------------------- //
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// temporal conv weight (c_out, c_in, G) fp32 -> Wp[plane][tap][c_out][c_in] bf16 (hi, lo)
__global__ void k_pack_tcn_w_bf16(const float *__restrict__ w, __nv_bfloat16 *__restrict__ wp, int c_out,
                                  int c_in, int G) {
  const long long total = (long long)G * c_out * c_in;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = (int)(i % c_in);
  const long long r = i / c_in;
  const int co = (int)(r % c_out);
  const int j = (int)(r / c_out);
  __nv_bfloat16 hi, lo;
  split_bf16(w[((long long)co * c_in + ci) * G + j], hi, lo);
  wp[i] = hi;
  wp[total + i] = lo;
}

// --------------------------------------------------------------------------- //
// Shared LayerNorm epilogue of the persistent kernels.  One thread owns one accumulator row
// r = (frame r / V, joint r % V) of a 128-row tile held in TMEM (C fp32 columns):
//   y = LN_{C,V}(acc + bias) * g + b  [+ res]  [relu]  -> fp32 rows or split-bf16 planes.
// The (C,V) statistics of a frame are reduced across its V rows through `s_part`
// (float[2][128] in shared memory) and named barrier 1 (the 128 epilogue threads).
// --------------------------------------------------------------------------- //
struct EpiParams {
  // Per-(joint, channel) tables use the layout [C/4][V][4]: the float4 of channel group g and joint
  // w is at (g*V + w)*4, so the 32 rows of a warp (consecutive joints) read ~4 cache lines per load
  // instruction instead of 32.
  const float *bias;             // bias_sw = 1: [C/4][V][4] table; bias_sw = 0: plain [C] vector
  int bias_sw;
  const float *n_wT, *n_bT;      // LayerNorm affine in [C/4][V][4] (k_transpose_affine)
  const float *res;              // fp32 [rows][C] added after the norm, or null
  const __nv_bfloat16 *res_hi, *res_lo;   // the same residual as bf16 hi/lo planes [rows][C] (res == null)
  float *out_f32;                // fp32 [rows][C] or null
  __nv_bfloat16 *out_hi, *out_lo;  // split-bf16 planes [rows][C] or null
  int out_T, out_t0;             // frames per trial of the plane buffer and frame offset of t = 0
                                 // (T-split: the planes carry halo frames on both sides); 0,0 = same as rows
  int relu;
  int raw;                       // 1: no norm at all -- out_f32 = acc + bias (RT step: the state update and the
                                 // LayerNorm run in the streaming kernel k_rt_update)
  float eps;
  int debug;                     // measurement aid: 1 = skip the epilogue body, 2 = skip transform math
  // RT-ST-GCN continual step (rtstgcn.py:611-625): per-stream ring FIFO and running accumulators.
  // fifo [F][rows][C], acc [S][rows][C] (rows = streams * V), one frame counter per stream.
  float *rt_fifo, *rt_acc;
  const int *rt_counter;
  int rt_F, rt_S;
  long long rt_slot;             // elements per slot = rows * C
};

// Statistics in ONE pass over TMEM: each row accumulates sum / sum of squares of (x - shift) with
// shift = its first element (so the squares do not cancel), giving a partial mean and M2; the
// partials of a frame (V rows x NH column groups) are then merged exactly (Chan et al.):
//   M2 = sum M2_p + CH * sum (m_p - mean)^2.
// NH epilogue warps share one TMEM lane quarter, each owning C/NH of the columns (`h`), so that
// 128*NH threads keep loads in flight.  `s_part` is float[2][2][NH][128]; `tile_parity` alternates
// the half used so one named barrier per tile suffices.
constexpr int kEpiNH = 2;                       // column groups (epilogue warps per TMEM lane quarter)
constexpr int kEpiThreads = 128 * kEpiNH;
constexpr int kPartBytes = 2 * 2 * kEpiNH * 128 * 4;
constexpr int kPatchPitch = 144;                           // 128 B of data + 16 B: conflict-free own-row access
constexpr int kPatchBytes = 32 * kPatchPitch;              // one epilogue warp's staging patch
constexpr int kPatchTotal = 4 * kEpiNH * kPatchBytes;     // ST-GCN kernels: one patch per epilogue warp
constexpr int kPatchTotalRt = 2 * kPatchTotal;            // RT kernel: two (cp.async double buffering)

template <int C, int NH>
__device__ __forceinline__ void frame_stats(const float *sp, int fr, int V, float eps, float &mean, float &rstd) {
  // one pass over the frame's NH*V partials, shifted by the first partial mean so that the
  // sum of squares does not cancel; three independent accumulation chains
  constexpr int CH = C / NH;
  const float ref = sp[fr * V];
  float sd = 0.f, sdd = 0.f, sm2 = 0.f;
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    const float *pm = sp + hh * 128 + fr * V;
    const float *pq = sp + (NH + hh) * 128 + fr * V;
#pragma unroll 5
    for (int j = 0; j < V; ++j) {
      const float d = pm[j] - ref;
      sd += d;
      sdd = fmaf(d, d, sdd);
      sm2 += pq[j];
    }
  }
  const float inv_n = 1.f / (float)(NH * V);
  mean = ref + sd * inv_n;
  const float tq = sm2 + (float)CH * fmaxf(sdd - sd * sd * inv_n, 0.f);
  rstd = 1.f / sqrtf(tq * (1.f / (float)(V * C - 1)) + eps);
}

// The same statistics computed by the warp as a whole (16 <= V <= 32, so the warp's 32 rows span at
// most three frames): lane j reads the partials of row j of a frame, three shuffle reductions merge
// them.  Every warp that touches a frame runs the identical reduction, so all rows of a frame get
// bit-identical statistics.  ~1/3 of the dependent-chain length of the per-thread loop above.
template <int C, int NH>
__device__ __forceinline__ void frame_stats_warp(const float *sp, int r0, int fr, int V, int RT, float eps,
                                                 float &mean, float &rstd) {
  constexpr int CH = C / NH;
  const int lane = threadIdx.x & 31;
  const int f_first = r0 / V;
  const float inv_n = 1.f / (float)(NH * V);
  const float inv_cv = 1.f / (float)(V * C - 1);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int f = f_first + k;
    if (f * V >= RT || f * V > r0 + 31) break;          // warp-uniform
    float d[NH], sm2 = 0.f;
    const bool on = lane < V;
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      d[hh] = on ? sp[hh * 128 + f * V + lane] : 0.f;
      sm2 += on ? sp[(NH + hh) * 128 + f * V + lane] : 0.f;
    }
    const float ref = __shfl_sync(0xffffffffu, d[0], 0);
    float sd = 0.f, sdd = 0.f;
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      const float x = on ? d[hh] - ref : 0.f;
      sd += x;
      sdd = fmaf(x, x, sdd);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
      sdd += __shfl_xor_sync(0xffffffffu, sdd, o);
      sm2 += __shfl_xor_sync(0xffffffffu, sm2, o);
    }
    const float m = ref + sd * inv_n;
    const float tq = sm2 + (float)CH * fmaxf(sdd - sd * sd * inv_n, 0.f);
    const float rs = 1.f / sqrtf(tq * inv_cv + eps);
    if (fr == f) {
      mean = m;
      rstd = rs;
    }
  }
}

// Second half of both epilogues: publish this thread's partial statistics, merge the frame's
// partials, then normalise / add the residual / activate the accumulator rows (re-read from TMEM)
// and write them out.  `add_bias`: the accumulator still lacks the bias (ST-GCN); the RT path has
// already folded it into the stashed accumulator.
// Pass 2 works in 32-column super-chunks through this warp's shared-memory patch [32 rows][144 B]:
// the residual block is loaded and the output block stored COOPERATIVELY (8 lanes per row, whole
// 128-B lines, 4 rows per instruction) instead of 32 different rows per instruction; in between
// every lane touches only its own patch row.
// kPR: compile the bf16-plane residual path (temporal kernels of the opt-in graph-conv v3 path only:
// in the graph-conv / RT kernels it only costs registers)
template <int C, int NH, bool kPR>
__device__ __forceinline__ void epi_finish(const EpiParams &e, uint32_t taddr, int r, int RT, int V, int fr, int w,
                                           bool row_ok, long long row, long long row_o, float *s_part,
                                           int tile_parity, int h, uint8_t *patch, float shift, float s1, float s2,
                                           bool add_bias, bool relu_mid, int bar_id = 1) {
  constexpr int CH = C / NH;
  const int c0 = h * CH;
  const int pstep = e.bias_sw ? V : 1;
  const float4 *bias4 = reinterpret_cast<const float4 *>(e.bias) + (c0 >> 2) * pstep + (e.bias_sw ? w : 0);
  float *sp = s_part + tile_parity * (2 * NH * 128);
  float v[16];
  const float m_r = shift + s1 * (1.f / (float)CH);
  const float M2_r = fmaxf(s2 - s1 * s1 * (1.f / (float)CH), 0.f);
  sp[h * 128 + r] = row_ok ? m_r : 0.f;
  sp[(NH + h) * 128 + r] = row_ok ? M2_r : 0.f;
  const bool fdbg = add_bias && (e.debug & 4) && blockIdx.x == 0 && r == 0 && h == 0;
  const long long tf0 = fdbg ? clock64() : 0;
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(128 * NH) : "memory");   // the 4*NH warps working on this tile
  float mean = 0.f, rstd = 0.f;
  if (V >= 16 && V <= 32) {
    if (!(e.debug & 64)) frame_stats_warp<C, NH>(sp, r - (threadIdx.x & 31), fr, V, RT, e.eps, mean, rstd);
  } else if (r < RT && !(e.debug & 64)) {
    frame_stats<C, NH>(sp, fr, V, e.eps, mean, rstd);
  }
  const float nmr = -mean * rstd;
  const long long tf1 = fdbg ? clock64() : 0;
  const float4 *nw4 = reinterpret_cast<const float4 *>(e.n_wT) + (c0 >> 2) * V + w;
  const float4 *nb4 = reinterpret_cast<const float4 *>(e.n_bT) + (c0 >> 2) * V + w;
  const int lane = threadIdx.x & 31;
  const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
  const long long row0 = __shfl_sync(0xffffffffu, row, 0);        // rows of a warp are contiguous
  const long long rowo0 = __shfl_sync(0xffffffffu, row_o, 0);
  uint8_t *mine = patch + lane * kPatchPitch;
  const bool use_res = (e.res != nullptr || (kPR && e.res_hi != nullptr)) && !(e.debug & 8);
  const bool res_planes = kPR && e.res == nullptr;
#ifdef STGCN_EPI_DEBUG   // fine-grained pass-2 phase timers (cost registers: off by default)
  long long tq = fdbg ? clock64() : 0, q_res = 0, q_tm = 0, q_math = 0, q_st = 0;
#define EPI_STAMP(acc) if (fdbg) { const long long n_ = clock64(); acc += n_ - tq; tq = n_; }
#else
#define EPI_STAMP(acc)
#endif
#pragma unroll 1
  for (int sb = 0; sb < CH; sb += 32) {
    if (use_res && res_planes) {
      // patch row, per 16-column half h: [hi: 32 B][lo: 32 B] at h * 64 (a half's output never
      // overwrites the other half's unread residual, whatever the residual / output formats)
      uint4 tr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pc = i * 32 + lane, rr = pc >> 3, qq = pc & 7;
        tr[i] = make_uint4(0u, 0u, 0u, 0u);
        const __nv_bfloat16 *src = (qq < 4) ? e.res_hi : e.res_lo;
        if (((okmask >> rr) & 1) && src)
          tr[i] = *reinterpret_cast<const uint4 *>(src + (row0 + rr) * C + c0 + sb + (qq & 3) * 8);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pc = i * 32 + lane, rr = pc >> 3, qq = pc & 7, q3 = qq & 3;
        *reinterpret_cast<uint4 *>(patch + rr * kPatchPitch + (q3 >> 1) * 64 + (q3 & 1) * 16 + (qq < 4 ? 0 : 32)) = tr[i];
      }
      __syncwarp();
      if (e.out_f32) {
        // fp32 output reuses the patch row as 32 floats: turn this lane's own row into hi + lo first
        uint4 h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          h[i] = *reinterpret_cast<const uint4 *>(mine + (i >> 1) * 64 + (i & 1) * 16);
          l[i] = *reinterpret_cast<const uint4 *>(mine + (i >> 1) * 64 + 32 + (i & 1) * 16);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t hw[4] = {h[i].x, h[i].y, h[i].z, h[i].w}, lw[4] = {l[i].x, l[i].y, l[i].z, l[i].w};
          float o[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&hw[j]));
            const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&lw[j]));
            o[2 * j] = a.x + b.x;
            o[2 * j + 1] = a.y + b.y;
          }
          *reinterpret_cast<float4 *>(mine + i * 32) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4 *>(mine + i * 32 + 16) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    } else if (use_res) {
      float4 tr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pc = i * 32 + lane, rr = pc >> 3, qq = pc & 7;
        tr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((okmask >> rr) & 1)
          tr[i] = *reinterpret_cast<const float4 *>(e.res + (row0 + rr) * C + c0 + sb + qq * 4);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int pc = i * 32 + lane, rr = pc >> 3, qq = pc & 7;
        *reinterpret_cast<float4 *>(patch + rr * kPatchPitch + qq * 16) = tr[i];
      }
      __syncwarp();
    }
    EPI_STAMP(q_res)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int cb = sb + half * 16;
      if (half == 1) { EPI_STAMP(q_math) }
      tmem_ld16(taddr + c0 + cb, v);
      EPI_STAMP(q_tm)
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pi = (cb >> 2) + i;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (add_bias) b4 = __ldg(bias4 + pi * pstep);
          const float4 g4 = __ldg(nw4 + pi * V);
          const float4 o4 = __ldg(nb4 + pi * V);
          // ((v + b) - mean) * rstd * g + o as two FMAs on (v + b): x * rstd + (-mean * rstd), then * g + o
          v[4 * i] = fmaf(fmaf(v[4 * i] + b4.x, rstd, nmr), g4.x, o4.x);
          v[4 * i + 1] = fmaf(fmaf(v[4 * i + 1] + b4.y, rstd, nmr), g4.y, o4.y);
          v[4 * i + 2] = fmaf(fmaf(v[4 * i + 2] + b4.z, rstd, nmr), g4.z, o4.z);
          v[4 * i + 3] = fmaf(fmaf(v[4 * i + 3] + b4.w, rstd, nmr), g4.w, o4.w);
        }
        if (relu_mid) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (use_res && res_planes && !e.out_f32) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint4 h4 = *reinterpret_cast<const uint4 *>(mine + half * 64 + i * 16);
            const uint4 l4 = *reinterpret_cast<const uint4 *>(mine + half * 64 + 32 + i * 16);
            const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&hw[j]));
              const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&lw[j]));
              v[8 * i + 2 * j] += a.x + b.x;
              v[8 * i + 2 * j + 1] += a.y + b.y;
            }
          }
        } else if (use_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 r4 = *reinterpret_cast<const float4 *>(mine + half * 64 + i * 16);
            v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
          }
        }
        if (e.relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (e.out_f32) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4 *>(mine + half * 64 + i * 16) =
                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // packed conversions (one F2FP per pair; the scalar form compiles to two slow-pipe F2F + PRMT)
            const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            const float2 hf = __bfloat1622float2(hh);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
            hi[i] = *reinterpret_cast<const uint32_t *>(&hh);
            lo[i] = *reinterpret_cast<const uint32_t *>(&ll);
          }
          uint4 *ph = reinterpret_cast<uint4 *>(mine + half * 64);
          ph[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          ph[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          uint4 *pl = reinterpret_cast<uint4 *>(mine + half * 64 + 32);
          pl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          pl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    }
    __syncwarp();
    EPI_STAMP(q_math)
    if (!(e.debug & 32)) {
      if (e.out_f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pc = i * 32 + lane, rr = pc >> 3, qq = pc & 7;
          if ((okmask >> rr) & 1)
            *reinterpret_cast<float4 *>(e.out_f32 + (row0 + rr) * C + c0 + sb + qq * 4) =
                *reinterpret_cast<const float4 *>(patch + rr * kPatchPitch + qq * 16);
        }
      } else if (e.out_hi) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pc = i * 32 + lane, rr = pc >> 2, qq = pc & 3;
          if ((okmask >> rr) & 1) {
            const uint8_t *src = patch + rr * kPatchPitch + (qq >> 1) * 64 + (qq & 1) * 16;
            *reinterpret_cast<uint4 *>(e.out_hi + (rowo0 + rr) * C + c0 + sb + qq * 8) =
                *reinterpret_cast<const uint4 *>(src);
            if (e.out_lo)
              *reinterpret_cast<uint4 *>(e.out_lo + (rowo0 + rr) * C + c0 + sb + qq * 8) =
                  *reinterpret_cast<const uint4 *>(src + 32);
          }
        }
      }
    }
    __syncwarp();
  }
  if (fdbg) {
#ifdef STGCN_EPI_DEBUG
    q_st = (clock64() - tf1) - q_res - q_tm - q_math;
    atomicAdd(&g_dbg[8], (unsigned long long)q_res);
    atomicAdd(&g_dbg[9], (unsigned long long)q_tm);
    atomicAdd(&g_dbg[10], (unsigned long long)q_math);
    atomicAdd(&g_dbg[15], (unsigned long long)q_st);
#endif
    atomicAdd(&g_dbg[13], (unsigned long long)(tf1 - tf0));
    atomicAdd(&g_dbg[14], (unsigned long long)(clock64() - tf1));
  }
}

// Raw epilogue: out_f32 = acc + bias, through the warp's patch so that rows leave as whole 128-B lines.
template <int C, int NH>
__device__ __forceinline__ void raw_epilogue_tile(const EpiParams &e, uint32_t taddr, int V, int w, bool row_ok,
                                                  long long row, int h, uint8_t *patch) {
  constexpr int CH = C / NH;
  const int c0 = h * CH;
  const int pstep = e.bias_sw ? V : 1;
  const float4 *bias4 = reinterpret_cast<const float4 *>(e.bias) + (c0 >> 2) * pstep + (e.bias_sw ? w : 0);
  const int lane = threadIdx.x & 31;
  const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
  const long long row0 = __shfl_sync(0xffffffffu, row, 0);
  uint8_t *mine = patch + lane * kPatchPitch;
  float v[16];
#pragma unroll 1
  for (int sb = 0; sb < CH; sb += 32) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int cb = sb + half * 16;
      tmem_ld16(taddr + c0 + cb, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b4 = __ldg(bias4 + ((cb >> 2) + i) * pstep);
        *reinterpret_cast<float4 *>(mine + half * 64 + i * 16) =
            make_float4(v[4 * i] + b4.x, v[4 * i + 1] + b4.y, v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int pc = i * 32 + lane, rr = pc >> 3, qq = pc & 7;
      if ((okmask >> rr) & 1)
        *reinterpret_cast<float4 *>(e.out_f32 + (row0 + rr) * C + c0 + sb + qq * 4) =
            *reinterpret_cast<const float4 *>(patch + rr * kPatchPitch + qq * 16);
    }
    __syncwarp();
  }
}

// ST-GCN epilogue: y = LN_{C,V}(acc + bias) * g + b [+ res] [relu].
template <int C, int NH, bool kPR = false>
__device__ __forceinline__ void ln_epilogue_tile(const EpiParams &e, uint32_t taddr, int r, int RT, int V, int fr,
                                                 int w, bool row_ok, long long row, long long row_o,
                                                 float *s_part, int tile_parity, int h, uint8_t *patch,
                                                 int bar_id = 1) {
  if (e.debug & 1) return;
  constexpr int CH = C / NH;
  static_assert(CH % 32 == 0, "epilogue works in 32-column super-chunks");
  const int c0 = h * CH;
  // float4 index of channel group g: table -> g*V + w, plain vector -> g
  const int pstep = e.bias_sw ? V : 1;
  const float4 *bias4 = reinterpret_cast<const float4 *>(e.bias) + (c0 >> 2) * pstep + (e.bias_sw ? w : 0);
  float v[16];
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
  const bool pdbg = (e.debug & 4) && blockIdx.x == 0 && r == 0 && h == 0;
  const long long tp0 = pdbg ? clock64() : 0;
#pragma unroll 1
  for (int cb = 0; cb < CH; cb += 16) {
    tmem_ld16(taddr + c0 + cb, v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = __ldg(bias4 + ((cb >> 2) + i) * pstep);
      v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
    }
    if (cb == 0) shift = v[0];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
  }
  if (pdbg) atomicAdd(&g_dbg[12], (unsigned long long)(clock64() - tp0));
  epi_finish<C, NH, kPR>(e, taddr, r, RT, V, fr, w, row_ok, row, row_o, s_part, tile_parity, h, patch, shift, s1, s2,
                         true, false, bar_id);
}

// RT-ST-GCN epilogue: the accumulator row holds z_t (graph-convolved frame, before bias).  Per
// element, in the reference's order (rtstgcn.py:611-612, 621):
//   acc <- (acc + z_t) + (-fifo[slot]);  fifo[slot] <- z_t;  o = acc
// then out = relu( relu(LN_{C,V}(o)) + res ) (res optional; rtstgcn.py:548-553).  The updated
// accumulator row is stashed back into the TMEM columns it came from (tcgen05.st) so that the
// LayerNorm statistics need no second trip to HBM.  `b` is the stream index of this row; streams
// may sit at different ring positions (independent resets), so slot indices travel by shuffle.
// The FIFO slot and accumulator rows move through the warp's patch ([64 B fifo | 64 B acc] per
// row and 16-column chunk) so that global loads and stores cover whole 64-B row segments.
template <int C, int NH>
__device__ __forceinline__ void rt_epilogue_tile(const EpiParams &e, uint32_t taddr, int r, int RT, int V, int fr,
                                                 int w, bool row_ok, long long row, int b, float *s_part,
                                                 int tile_parity, int h, uint8_t *patch) {
  if (e.debug & 1) return;
  constexpr int CH = C / NH;
  const int c0 = h * CH;
  const int pstep = e.bias_sw ? V : 1;
  const float4 *bias4 = reinterpret_cast<const float4 *>(e.bias) + (c0 >> 2) * pstep + (e.bias_sw ? w : 0);
  float v[16];
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
  const int lane = threadIdx.x & 31;
  const int cnt = row_ok ? __ldg(e.rt_counter + b) : 0;
  const int my_fi = cnt % e.rt_F, my_ai = cnt % e.rt_S;
  const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
  const long long row0 = __shfl_sync(0xffffffffu, row, 0);
  // This lane's four cooperative pieces (row rr = pc >> 2, 16-B piece qq = pc & 3) of every chunk:
  // slot indices of the piece's row (streams may sit at different ring positions) and addresses.
  const float *gsrc_f[4], *gsrc_a[4];
  uint32_t pdst[4];
  bool pok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pc = i * 32 + lane, rr = pc >> 2, qq = pc & 3;
    const int fi = __shfl_sync(0xffffffffu, my_fi, rr), ai = __shfl_sync(0xffffffffu, my_ai, rr);
    const long long off = (row0 + rr) * C + c0 + qq * 4;
    gsrc_f[i] = e.rt_fifo + (long long)fi * e.rt_slot + off;
    gsrc_a[i] = e.rt_acc + (long long)ai * e.rt_slot + off;
    pdst[i] = (uint32_t)(rr * kPatchPitch + qq * 16);
    pok[i] = (okmask >> rr) & 1;
  }
  // The FIFO slot / accumulator chunks stream global -> patch with cp.async, one chunk ahead
  // (two patches per warp), so their latency overlaps the arithmetic and stores of the previous chunk.
  auto issue = [&](int cb, int buf) {
    const uint32_t pb = smem_u32(patch + buf * kPatchBytes);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (pok[i]) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pb + pdst[i]), "l"(gsrc_f[i] + cb) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(pb + pdst[i] + 64), "l"(gsrc_a[i] + cb) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const bool tdbg = (e.debug & 4) && blockIdx.x == 0 && r == 0 && h == 0;
  long long tA = 0, tB = 0, tC = 0, tl = tdbg ? clock64() : 0;
  issue(0, 0);
  int buf = 0;
#pragma unroll 1
  for (int cb = 0; cb < CH; cb += 16, buf ^= 1) {
    if (cb + 16 < CH) {
      issue(cb + 16, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    if (tdbg) { const long long n_ = clock64(); tA += n_ - tl; tl = n_; }
    uint8_t *pw = patch + buf * kPatchBytes;
    uint8_t *mine = pw + lane * kPatchPitch;
    tmem_ld16(taddr + c0 + cb, v);
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b4 = __ldg(bias4 + ((cb >> 2) + i) * pstep);
        const float4 old = *reinterpret_cast<const float4 *>(mine + i * 16);
        float4 a = *reinterpret_cast<const float4 *>(mine + 64 + i * 16);
        const float4 z = make_float4(v[4 * i] + b4.x, v[4 * i + 1] + b4.y, v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w);
        a.x = (a.x + z.x) + (-old.x);
        a.y = (a.y + z.y) + (-old.y);
        a.z = (a.z + z.z) + (-old.z);
        a.w = (a.w + z.w) + (-old.w);
        *reinterpret_cast<float4 *>(mine + i * 16) = z;
        *reinterpret_cast<float4 *>(mine + 64 + i * 16) = a;
        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    if (cb == 0) shift = v[0];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
    tmem_st16(taddr + c0 + cb, v);
    __syncwarp();
    if (tdbg) { const long long n_ = clock64(); tB += n_ - tl; tl = n_; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (pok[i]) {
        *reinterpret_cast<float4 *>(const_cast<float *>(gsrc_f[i]) + cb) = *reinterpret_cast<const float4 *>(pw + pdst[i]);
        *reinterpret_cast<float4 *>(const_cast<float *>(gsrc_a[i]) + cb) =
            *reinterpret_cast<const float4 *>(pw + pdst[i] + 64);
      }
    __syncwarp();
    if (tdbg) { const long long n_ = clock64(); tC += n_ - tl; tl = n_; }
  }
  if (tdbg) {
    atomicAdd(&g_dbg[12], (unsigned long long)tA);
    atomicAdd(&g_dbg[13], (unsigned long long)tB);
    atomicAdd(&g_dbg[14], (unsigned long long)tC);
  }
  epi_finish<C, NH, false>(e, taddr, r, RT, V, fr, w, row_ok, row, row, s_part, tile_parity, h, patch, shift, s1, s2,
                           false, true);
  if (tdbg) atomicAdd(&g_dbg[15], (unsigned long long)(clock64() - tl));
}

// --------------------------------------------------------------------------- //
// Temporal convolution v2: persistent, unpadded rows, stride 1 or 2.
//
// Measured on B200 (tools/probes/umma_row_offset.cu): a K-major SWIZZLE_128B descriptor whose
// start address is offset by ANY number of 128-B rows (base_offset field 0) addresses the rows
// TMA wrote -- the swizzle is a function of the absolute shared-memory address.  So the input
// window is staged densely (V rows per frame, no padding to 32) and a temporal tap is a shift
// of V rows.  A 128-row UMMA tile carries FT = floor(128 / V) whole frames (125 of 128 rows
// useful for V = 25); the few spill rows belong to the next tile and are ignored.
//
// Each CTA is persistent over "items" (NT consecutive tiles of one trial).  TMA producer, MMA
// issuer and epilogue run as a pipeline across items: the producer prefetches the next item's
// input window / weight tiles while the tensor core works, and (C <= 128) the accumulators
// are double buffered in TMEM so the LayerNorm epilogue of item i overlaps the MMAs of i+1.
// Stride 2: even and odd input frames are staged as two dense windows through two tensor maps
// (frame stride 2), so every tap again reads a contiguous window.
// --------------------------------------------------------------------------- //
constexpr int kTcn2Threads = 32 * (3 + 4 * kEpiNH);   // warps: 0 A producer, 1 MMA, 2 B producer, 3.. epilogue

struct TcnTc2Params {
  int T_out, V, G, planes;
  int FT, NT, tb;             // frames per 128-row tile, tiles per item, TMEM accumulator buffers
  int groups_per_trial, items;
  int a_stage_bytes, b_stages;
  int n_loads;                // TMA loads per A stage (1: stride 1, 2: stride 2 even/odd windows)
  int load_f0[2];             // first frame of the window relative to the item's first output frame
  int load_row[2];            // destination row in the stage
  int load_bytes[2];
  int tap_row[16];            // first stage row of tap j for tile 0
  long long res_trial_rows;   // rows between consecutive trials of the residual (set by the launcher: T_out*V;
                              // V when the trials are overlapping windows of one shared frame sequence)
  EpiParams epi;
};

template <int C>
__global__ void __launch_bounds__(kTcn2Threads, 1)
    k_tcn_tc2(const __grid_constant__ CUtensorMap tm_u0, const __grid_constant__ CUtensorMap tm_u1,
              const __grid_constant__ CUtensorMap tm_w, const TcnTc2Params p) {
  constexpr int kBBytes = C * 128;
  constexpr int kTmemCols = 512;                  // one CTA per SM: take all of TMEM
  const int TB = p.tb;                            // accumulator buffers of NT tiles each (NT*C*TB <= 512)
  const int buf_cols = p.NT * C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int S = p.b_stages;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + 2 * p.a_stage_bytes;
  const uint32_t sPart = sB + S * kBBytes;                    // float[2][2][NH][128]
  const uint32_t sPatch = sPart + kPartBytes;                 // epilogue staging patches
  const uint32_t sBar = sPatch + kPatchTotal;
  const uint32_t bFullA = sBar, bEmptyA = sBar + 16, bTmemFull = sBar + 32, bTmemEmpty = sBar + 48;
  const uint32_t bFullB = sBar + 64, bEmptyB = bFullB + 8 * S;
  const uint32_t sTmemPtr = bEmptyB + 8 * S;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  float *s_part = reinterpret_cast<float *>(gen_base + (sPart - smem_base));
  uint8_t *s_patch = gen_base + (sPatch - smem_base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = C / 64;
  const int RT = p.FT * p.V;                      // useful rows per tile
  const bool dbg = (p.epi.debug & 4) && blockIdx.x == 0;
  long long d0 = 0, d1 = 0, d2 = 0, d3 = 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_u0);
    tma_prefetch_desc(&tm_u1);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bFullA + 8 * i, 1);
      mbar_init(bEmptyA + 8 * i, 1);
      mbar_init(bTmemFull + 8 * i, 1);
      mbar_init(bTmemEmpty + 8 * i, 4 * kEpiNH);  // one arrive per epilogue warp
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(bFullB + 8 * i, 1);
      mbar_init(bEmptyB + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(sTmemPtr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ---- input-window producer (runs ahead by the 2-deep A ring, independent of the weights) ----
    if (lane == 0) {
      int as = 0, a_ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.groups_per_trial;
        const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;   // first output frame
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap) {
            { DbgTimer tm(dbg); mbar_wait(bEmptyA + 8 * as, a_ph ^ 1); tm.stop(d0); }
            mbar_expect_tx(bFullA + 8 * as, (uint32_t)(p.load_bytes[0] + (p.n_loads > 1 ? p.load_bytes[1] : 0)));
            tma_load_5d(sA + as * p.a_stage_bytes + p.load_row[0] * 128, &tm_u0, bFullA + 8 * as, kc * 64, 0,
                        f0 + p.load_f0[0], n, ap);
            if (p.n_loads > 1)
              tma_load_5d(sA + as * p.a_stage_bytes + p.load_row[1] * 128, &tm_u1, bFullA + 8 * as, kc * 64, 0,
                          f0 + p.load_f0[1], n, ap);
            as ^= 1;
            if (as == 0) a_ph ^= 1;
          }
      }
      dbg_flush(dbg, 4, d0);
    }
  } else if (warp == 2) {
    // ---- weight-tile producer ----
    if (lane == 0) {
      int bs = 0, b_ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x)
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap) {
            const int nb = (ap == 0) ? p.planes : 1;
            for (int j = 0; j < p.G; ++j)
              for (int bp = 0; bp < nb; ++bp) {
                { DbgTimer tm(dbg); mbar_wait(bEmptyB + 8 * bs, b_ph ^ 1); tm.stop(d0); }
                mbar_expect_tx(bFullB + 8 * bs, kBBytes);
                tma_load_4d(sB + bs * kBBytes, &tm_w, bFullB + 8 * bs, kc * 64, 0, j, bp);
                if (++bs == S) { bs = 0; b_ph ^= 1; }
              }
          }
      dbg_flush(dbg, 5, d0);
    }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp walks the schedule (uniform control flow); one elected lane
    // issues the MMAs and commits ----
    constexpr uint32_t idesc = umma_idesc_bf16(128, C);
    int a_s = 0, a_ph = 0, b_s = 0, b_ph = 0, buf = 0, t_ph = 0, it = 0;
    DbgTimer tall(dbg);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      { DbgTimer tm(dbg); mbar_wait(bTmemEmpty + 8 * buf, t_ph ^ 1); tm.stop(d0); }
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * buf_cols;
      uint32_t acc = 0;
      for (int kc = 0; kc < KC; ++kc)
        for (int ap = 0; ap < p.planes; ++ap) {
          { DbgTimer tm(dbg); mbar_wait(bFullA + 8 * a_s, a_ph); tm.stop(d1); }
          tc_fence_after();
          const uint32_t a_lo0 = umma_desc_lo(sA + a_s * p.a_stage_bytes);
          const int nb = (ap == 0) ? p.planes : 1;
          for (int j = 0; j < p.G; ++j) {
            const uint32_t a_tap = a_lo0 + (uint32_t)(p.tap_row[j] * 8);        // rows * 128 B >> 4
            for (int bp = 0; bp < nb; ++bp) {
              { DbgTimer tm(dbg); mbar_wait(bFullB + 8 * b_s, b_ph); tm.stop(d2); }
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_lo = umma_desc_lo(sB + b_s * kBBytes);
                for (int m = 0; m < p.NT; ++m) {
                  const uint32_t a_lo = a_tap + (uint32_t)(m * RT * 8);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(tacc + m * C, umma_desc_join(a_lo + 2 * k), umma_desc_join(b_lo + 2 * k), idesc,
                              acc | (uint32_t)k);
                }
                umma_commit(bEmptyB + 8 * b_s);
              }
              __syncwarp();
              acc = 1;
              if (++b_s == S) { b_s = 0; b_ph ^= 1; }
            }
          }
          if (elect_one()) umma_commit(bEmptyA + 8 * a_s);
          __syncwarp();
          a_s ^= 1;
          if (a_s == 0) a_ph ^= 1;
        }
      if (elect_one()) umma_commit(bTmemFull + 8 * buf);
      __syncwarp();
      if (++buf == TB) { buf = 0; t_ph ^= 1; }
    }
    tall.stop(d3);
    if (lane == 0) {
      dbg_flush(dbg, 0, d0); dbg_flush(dbg, 1, d1); dbg_flush(dbg, 2, d2); dbg_flush(dbg, 3, d3);
      dbg_flush(dbg, 11, it);
    }
  } else {
    // ---- epilogue: thread <-> accumulator row r = 32*(warp&3) + lane = (frame r / V, joint r % V);
    // the kEpiNH warps that share a TMEM lane quarter split the channel range between them ----
    const int q = warp & 3;
    const int h = (warp - 3) >> 2;
    const int r = q * 32 + lane;
    const int fr = r / p.V, w = r - fr * p.V;
    int buf = 0, t_ph = 0, par = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int n = item / p.groups_per_trial;
      const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;
      if (p.epi.res && r < RT) {
        // pull this thread's residual rows towards L2 while the MMAs of the item are still running
        for (int m = 0; m < p.NT; ++m) {
          const int t = f0 + m * p.FT + fr;
          if (t < p.T_out) {
            const char *rp = reinterpret_cast<const char *>(
                p.epi.res + (n * p.res_trial_rows + (long long)t * p.V + w) * C + h * (C / kEpiNH));
#pragma unroll
            for (int o = 0; o < (C / kEpiNH) * 4; o += 128) prefetch_l2(rp + o);
          }
        }
      }
      { DbgTimer tm(dbg); mbar_wait(bTmemFull + 8 * buf, t_ph); tm.stop(d0); }
      tc_fence_after();
      DbgTimer tw(dbg);
#pragma unroll 1
      for (int m = 0; m < p.NT; ++m, par ^= 1) {
        const int t = f0 + m * p.FT + fr;
        const bool row_ok = (r < RT) && (t < p.T_out);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * buf_cols + m * C);
        const long long row = ((long long)n * p.T_out + t) * p.V + w;
        const long long row_res = n * p.res_trial_rows + (long long)t * p.V + w;
        ln_epilogue_tile<C, kEpiNH, true>(p.epi, taddr, r, RT, p.V, fr, w, row_ok, row_res, row, s_part, par, h,
                                          s_patch + (warp - 3) * kPatchBytes);
      }
      // accumulator buffer drained: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bTmemEmpty + 8 * buf);
      if (++buf == TB) { buf = 0; t_ph ^= 1; }
      tw.stop(d1);
    }
    if (warp == 3 && lane == 0) { dbg_flush(dbg, 6, d0); dbg_flush(dbg, 7, d1); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// adjacency CSR ordered by (k, w): ptr[k*V + w] .. ptr[k*V + w + 1] -> (v, A[k,v,w]); and the bias
// that flows through A: bz[w][c] = sum_k bg[k*CO + c] * colsum_k[w].
__global__ void k_build_adj_csr_kw(const float *__restrict__ A, int K, int V, int *__restrict__ ptr,
                                   int2 *__restrict__ va) {
  extern __shared__ int s_cnt[];  // K*V + 1
  const int P = K * V;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const int k = i / V, w = i - k * V;
    int c = 0;
    for (int v = 0; v < V; ++v) c += (A[((long long)k * V + v) * V + w] != 0.f);
    s_cnt[i] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < P; ++i) {
      int c = s_cnt[i];
      s_cnt[i] = run;
      run += c;
    }
    s_cnt[P] = run;
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= P; i += blockDim.x) ptr[i] = s_cnt[i];
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const int k = i / V, w = i - k * V;
    int at = s_cnt[i];
    for (int v = 0; v < V; ++v) {
      const float x = A[((long long)k * V + v) * V + w];
      if (x != 0.f) {
        va[at] = make_int2(v, __float_as_int(x));
        ++at;
      }
    }
  }
}

__global__ void k_bias_through_adj(const float *__restrict__ A, const float *__restrict__ bg, int K, int V,
                                   int CO, float *__restrict__ bzT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CO * V) return;
  const int e4 = i & 3, w = (i >> 2) % V, c = ((i >> 2) / V) * 4 + e4;   // output layout [CO/4][V][4]
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    float col = 0.f;
    for (int v = 0; v < V; ++v) col += A[((long long)k * V + v) * V + w];
    s = fmaf(bg[k * CO + c], col, s);
  }
  bzT[i] = s;
}

// LayerNorm affine (C, 1, V) -> [C/4][V][4]: an epilogue thread (one joint, 4 channels per load)
// reads 16 B, and the consecutive joints of a warp share cache lines
__global__ void k_transpose_affine(const float *__restrict__ src, float *__restrict__ dst, int C, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * V) return;
  const int e4 = i & 3, w = (i >> 2) % V, c = ((i >> 2) / V) * 4 + e4;   // [C/4][V][4]
  dst[i] = src[c * V + w];
}

// 1x1 weights (K*c_out, c_in) fp32 -> bf16 planes [2][K][c_out][c_in] (same order, split only)
__global__ void k_split_bf16(const float *__restrict__ w, __nv_bfloat16 *__restrict__ wp, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  __nv_bfloat16 hi, lo;
  split_bf16(w[i], hi, lo);
  wp[i] = hi;
  wp[total + i] = lo;
}

// --------------------------------------------------------------------------- //
// Graph convolution v2: persistent + pipelined version of k_gcn_tc with dense (unpadded)
// rows.  Also serves the residual branch LN_R(conv1x1_stride(x)) of channel-changing layers
// (identity "adjacency", K = 1, frame stride folded into the tensor map, no ReLU).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2..9 transform (contraction + bf16 split into the
// swizzled A ring), 10..13 epilogue (TMEM double buffered for C <= 128).
// --------------------------------------------------------------------------- //
constexpr int kGcn2Threads = 32 * (10 + 4 * kEpiNH);   // 0 producers, 1 MMA, 2..9 transform, 10.. epilogue
constexpr int kGcn2Csr = 2048;               // ptr[<=128] + 192 entries
constexpr int kGcn2CsrMax = 192;

struct GcnTc2Params {
  int T_out, V, K, Cin, planes;
  int FT, NT, tb, groups_per_trial, items;
  int a_stage_bytes, a_ring, xs_alloc, xs_tx, xs_bufs, b_stages;
  int identity;
  const int *csr_ptr;
  const int2 *csr_va;
  EpiParams epi;
};

template <int CO, bool kRt>
__global__ void __launch_bounds__(kGcn2Threads, 1)
    k_gcn_tc2(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
              const GcnTc2Params p) {
  constexpr int kBBytes = CO * 128;
  constexpr int kTmemCols = 512;
  const int TB = p.tb;
  const int buf_cols = p.NT * CO;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int S = p.b_stages, AR = p.a_ring, XB = p.xs_bufs;
  const uint32_t sA = smem_base;
  const uint32_t sXs = sA + AR * p.a_stage_bytes;
  const uint32_t sB = sXs + XB * p.xs_alloc;
  const uint32_t sCsr = sB + S * kBBytes;
  const uint32_t sPart = sCsr + kGcn2Csr;
  const uint32_t sPatch = sPart + kPartBytes;
  const uint32_t sBar = sPatch + (kRt ? kPatchTotalRt : kPatchTotal);
  const uint32_t bXsFull = sBar, bXsEmpty = sBar + 16, bTmemFull = sBar + 32, bTmemEmpty = sBar + 48;
  const uint32_t bAFull = sBar + 64, bAEmpty = sBar + 96;
  const uint32_t bFullB = sBar + 128, bEmptyB = bFullB + 8 * S;
  const uint32_t sTmemPtr = bEmptyB + 8 * S;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  int *s_ptr = reinterpret_cast<int *>(gen_base + (sCsr - smem_base));
  int2 *s_va = reinterpret_cast<int2 *>(gen_base + (sCsr - smem_base) + 512);
  float *s_part = reinterpret_cast<float *>(gen_base + (sPart - smem_base));
  uint8_t *s_patch = gen_base + (sPatch - smem_base);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = p.Cin / 64;
  const int RT = p.FT * p.V;
  const int rows_item = p.NT * RT;
  const bool dbg = (p.epi.debug & 4) && blockIdx.x == 0;
  long long d0 = 0, d1 = 0, d2 = 0, d3 = 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bXsFull + 8 * i, 1);
      mbar_init(bXsEmpty + 8 * i, 8);
      mbar_init(bTmemFull + 8 * i, 1);
      mbar_init(bTmemEmpty + 8 * i, 4 * kEpiNH);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bAFull + 8 * i, 8);
      mbar_init(bAEmpty + 8 * i, 1);
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(bFullB + 8 * i, 1);
      mbar_init(bEmptyB + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sTmemPtr, kTmemCols);
  bool csr_smem = false;
  if (!p.identity) {
    const int nnz = __ldg(p.csr_ptr + p.K * p.V);
    csr_smem = nnz <= kGcn2CsrMax && p.K * p.V + 1 <= 128;
    if (csr_smem) {
      for (int i = threadIdx.x; i <= p.K * p.V; i += blockDim.x) s_ptr[i] = __ldg(p.csr_ptr + i);
      for (int i = threadIdx.x; i < nnz; i += blockDim.x) s_va[i] = __ldg(p.csr_va + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ---- producers: lane 0 streams the fp32 input tiles, lane 1 the weight tiles (two
    // independent threads of one warp: each waits on its own ring) ----
    if (lane == 0) {
      int xb = 0, x_ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.groups_per_trial;
        const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;
        for (int kc = 0; kc < KC; ++kc) {
          { DbgTimer tm(dbg); mbar_wait(bXsEmpty + 8 * xb, x_ph ^ 1); tm.stop(d0); }
          mbar_expect_tx(bXsFull + 8 * xb, (uint32_t)p.xs_tx);
          tma_load_4d(sXs + xb * p.xs_alloc, &tm_x, bXsFull + 8 * xb, kc * 64, 0, f0, n);
          tma_load_4d(sXs + xb * p.xs_alloc + p.xs_alloc / 2, &tm_x, bXsFull + 8 * xb, kc * 64 + 32, 0, f0, n);
          if (++xb == XB) { xb = 0; x_ph ^= 1; }
        }
      }
      dbg_flush(dbg, 4, d0);
    } else if (lane == 1) {
      int bs = 0, b_ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x)
        for (int kc = 0; kc < KC; ++kc)
          for (int k = 0; k < p.K; ++k)
            for (int ap = 0; ap < p.planes; ++ap) {
              const int nb = (ap == 0) ? p.planes : 1;
              for (int bp = 0; bp < nb; ++bp) {
                { DbgTimer tm(dbg); mbar_wait(bEmptyB + 8 * bs, b_ph ^ 1); tm.stop(d0); }
                mbar_expect_tx(bFullB + 8 * bs, kBBytes);
                tma_load_4d(sB + bs * kBBytes, &tm_w, bFullB + 8 * bs, kc * 64, 0, k, bp);
                if (++bs == S) { bs = 0; b_ph ^= 1; }
              }
            }
      dbg_flush(dbg, 5, d0);
    }
  } else if (warp == 1) {
    // ---- MMA issuer: the whole warp walks the schedule (uniform control flow, so descriptors and
    // barrier addresses live in uniform registers); one elected lane issues MMAs and commits ----
    constexpr uint32_t idesc = umma_idesc_bf16(128, CO);
    int a_s = 0, a_ph = 0, b_s = 0, b_ph = 0, buf = 0, t_ph = 0, it = 0;
    DbgTimer tall(dbg);
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      { DbgTimer tm(dbg); mbar_wait(bTmemEmpty + 8 * buf, t_ph ^ 1); tm.stop(d0); }
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * buf_cols;
      uint32_t acc = 0;
      for (int kc = 0; kc < KC; ++kc)
        for (int k = 0; k < p.K; ++k)
          for (int ap = 0; ap < p.planes; ++ap) {
            { DbgTimer tm(dbg); mbar_wait(bAFull + 8 * a_s, a_ph); tm.stop(d1); }
            tc_fence_after();
            const uint32_t a_lo0 = umma_desc_lo(sA + a_s * p.a_stage_bytes);
            const int nb = (ap == 0) ? p.planes : 1;
            for (int bp = 0; bp < nb; ++bp) {
              { DbgTimer tm(dbg); mbar_wait(bFullB + 8 * b_s, b_ph); tm.stop(d2); }
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_lo = umma_desc_lo(sB + b_s * kBBytes);
                for (int m = 0; m < p.NT; ++m) {
                  const uint32_t a_lo = a_lo0 + (uint32_t)(m * RT * 8);      // rows * 128 B >> 4
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tacc + m * CO, umma_desc_join(a_lo + 2 * kk), umma_desc_join(b_lo + 2 * kk), idesc,
                              acc | (uint32_t)kk);
                }
                umma_commit(bEmptyB + 8 * b_s);
              }
              __syncwarp();
              acc = 1;
              if (++b_s == S) { b_s = 0; b_ph ^= 1; }
            }
            if (elect_one()) umma_commit(bAEmpty + 8 * a_s);
            __syncwarp();
            if (++a_s == AR) { a_s = 0; a_ph ^= 1; }
          }
      if (elect_one()) umma_commit(bTmemFull + 8 * buf);
      __syncwarp();
      if (++buf == TB) { buf = 0; t_ph ^= 1; }
    }
    tall.stop(d3);
    if (lane == 0) {
      dbg_flush(dbg, 0, d0); dbg_flush(dbg, 1, d1); dbg_flush(dbg, 2, d2); dbg_flush(dbg, 3, d3);
      dbg_flush(dbg, 11, it);
    }
  } else if (warp < 10) {
    // ---- transform warps: the adjacency contraction on the staged fp32 tile, then the bf16
    // hi/lo split into the swizzled UMMA A ring ----
    // NT == 2: one thread per row (256 threads >= rows of an item); NT == 1: two threads per
    // row, one 32-channel sub-tile each.  The fp32 input tile is staged as two 32-channel
    // sub-tiles with the 128-B TMA swizzle, so threads of a warp (consecutive rows) read 16-B
    // chunks from distinct banks.  Per (k, sub-tile) a thread accumulates its row's 8 float4
    // groups over the CSR entries of (k, joint) -- entries outer, groups inner, so an entry's
    // address arithmetic is done once -- splits them into bf16 hi/lo, stores hi as 16-B chunks
    // into the A stage and keeps lo packed in registers for the following lo-plane stage.
    const int tid = (warp - 2) * 32 + lane;
    const bool split = rows_item <= 128;
    const int r = split ? (tid & 127) : tid;
    const int my_sub = tid >> 7;                        // used when split
    const bool r_ok = r < rows_item;
    const int f = r / p.V, w = r - f * p.V;
    const int src_base = f * p.V;                       // first source row of this row's frame
    const int sw = (r & 7) << 4;
    int as = 0, a_ph = 0, xb = 0, x_ph = 0;
    uint4 lo_stash[8];
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      if (kRt && r_ok && (!split || my_sub == 0)) {
        // RT step: start pulling this row's FIFO slot and accumulator into L2 now; the epilogue
        // reaches them only after the MMAs of this item, and would otherwise wait on HBM latency.
        const int n_ = item / p.groups_per_trial;
        const int b = (item - n_ * p.groups_per_trial) * p.NT * p.FT + f;
        if (b < p.T_out) {
          const int cnt = __ldg(p.epi.rt_counter + b);
          const long long ro = ((long long)b * p.V + w) * CO;
          const char *fp = reinterpret_cast<const char *>(p.epi.rt_fifo + (long long)(cnt % p.epi.rt_F) * p.epi.rt_slot + ro);
          const char *ap = reinterpret_cast<const char *>(p.epi.rt_acc + (long long)(cnt % p.epi.rt_S) * p.epi.rt_slot + ro);
#pragma unroll
          for (int o = 0; o < CO * 4; o += 128) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(fp + o));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + o));
          }
        }
      }
      for (int kc = 0; kc < KC; ++kc) {
        { DbgTimer tm(dbg); mbar_wait(bXsFull + 8 * xb, x_ph); tm.stop(d0); }
        const uint8_t *xs = gen_base + (sXs - smem_base) + xb * p.xs_alloc;
        const int half = p.xs_alloc / 2;
        for (int k = 0; k < p.K; ++k) {
          int e0 = 0, e1 = 0;
          if (r_ok && !p.identity) {
            e0 = csr_smem ? s_ptr[k * p.V + w] : __ldg(p.csr_ptr + k * p.V + w);
            e1 = csr_smem ? s_ptr[k * p.V + w + 1] : __ldg(p.csr_ptr + k * p.V + w + 1);
          }
          {
            { DbgTimer tm(dbg); mbar_wait(bAEmpty + 8 * as, a_ph ^ 1); tm.stop(d1); }
            DbgTimer tcomp(dbg);
            uint8_t *dst_row = gen_base + as * p.a_stage_bytes + r * 128;
            if (r_ok && !(p.epi.debug & 2)) {
#pragma unroll
              for (int sub = 0; sub < 2; ++sub) {
                if (split && sub != my_sub) continue;
                float4 acc[8];
                if (p.identity) {
                  const uint8_t *b0 = xs + sub * half + r * 128;
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[j] = *reinterpret_cast<const float4 *>(b0 + ((j << 4) ^ sw));
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                  for (int e = e0; e < e1; ++e) {
                    const int2 va = csr_smem ? s_va[e] : __ldg(p.csr_va + e);
                    const float a = __int_as_float(va.y);
                    const int rs = src_base + va.x;
                    const uint8_t *b0 = xs + sub * half + rs * 128;
                    const int ssw = (rs & 7) << 4;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                      const float4 xv = *reinterpret_cast<const float4 *>(b0 + ((j << 4) ^ ssw));
                      acc[j].x = fmaf(a, xv.x, acc[j].x);
                      acc[j].y = fmaf(a, xv.y, acc[j].y);
                      acc[j].z = fmaf(a, xv.z, acc[j].z);
                      acc[j].w = fmaf(a, xv.w, acc[j].w);
                    }
                  }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  uint32_t hi[4], lo[4];
#pragma unroll
                  for (int u = 0; u < 2; ++u) {
                    const float4 q = acc[2 * jj + u];
                    const __nv_bfloat162 h01 = __floats2bfloat162_rn(q.x, q.y);
                    const __nv_bfloat162 h23 = __floats2bfloat162_rn(q.z, q.w);
                    const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                    const __nv_bfloat162 l01 = __floats2bfloat162_rn(q.x - f01.x, q.y - f01.y);
                    const __nv_bfloat162 l23 = __floats2bfloat162_rn(q.z - f23.x, q.w - f23.y);
                    hi[2 * u] = *reinterpret_cast<const uint32_t *>(&h01);
                    hi[2 * u + 1] = *reinterpret_cast<const uint32_t *>(&h23);
                    lo[2 * u] = *reinterpret_cast<const uint32_t *>(&l01);
                    lo[2 * u + 1] = *reinterpret_cast<const uint32_t *>(&l23);
                  }
                  // channels sub*32 + jj*8 .. +8 = 16-B chunk (sub*4 + jj) of the 128-B row
                  *reinterpret_cast<uint4 *>(dst_row + ((((sub << 2) | jj) << 4) ^ sw)) =
                      make_uint4(hi[0], hi[1], hi[2], hi[3]);
                  lo_stash[sub * 4 + jj] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bAFull + 8 * as);
            if (++as == AR) { as = 0; a_ph ^= 1; }
            tcomp.stop(d2);
          }
          if (p.planes == 2) {
            mbar_wait(bAEmpty + 8 * as, a_ph ^ 1);
            uint8_t *dst_row = gen_base + as * p.a_stage_bytes + r * 128;
            if (r_ok && !(p.epi.debug & 2)) {
#pragma unroll
              for (int sub = 0; sub < 2; ++sub) {
                if (split && sub != my_sub) continue;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                  *reinterpret_cast<uint4 *>(dst_row + ((((sub << 2) | jj) << 4) ^ sw)) = lo_stash[sub * 4 + jj];
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bAFull + 8 * as);
            if (++as == AR) { as = 0; a_ph ^= 1; }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bXsEmpty + 8 * xb);
        if (++xb == XB) { xb = 0; x_ph ^= 1; }
      }
    }
    if (warp == 2 && lane == 0) { dbg_flush(dbg, 8, d0); dbg_flush(dbg, 9, d1); dbg_flush(dbg, 10, d2); }
  } else {
    // ---- epilogue warps 10.. (kEpiNH per TMEM lane quarter, splitting the channel range) ----
    const int q = warp & 3;
    const int h = (warp - 10) >> 2;
    const int r = q * 32 + lane;
    const int fr = r / p.V, w = r - fr * p.V;
    int buf = 0, t_ph = 0, par = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int n = item / p.groups_per_trial;
      const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;
      { DbgTimer tm(dbg); mbar_wait(bTmemFull + 8 * buf, t_ph); tm.stop(d0); }
      tc_fence_after();
      DbgTimer tw(dbg);
#pragma unroll 1
      for (int m = 0; m < p.NT; ++m, par ^= 1) {
        const int t = f0 + m * p.FT + fr;
        const bool row_ok = (r < RT) && (t < p.T_out);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * buf_cols + m * CO);
        const long long row = ((long long)n * p.T_out + t) * p.V + w;
        if (kRt)
          rt_epilogue_tile<CO, kEpiNH>(p.epi, taddr, r, RT, p.V, fr, w, row_ok, row, t, s_part, par, h,
                                       s_patch + (warp - 10) * 2 * kPatchBytes);
        else if (p.epi.raw)
          raw_epilogue_tile<CO, kEpiNH>(p.epi, taddr, p.V, w, row_ok, row, h, s_patch + (warp - 10) * kPatchBytes);
        else
        {
          const long long row_o =
              p.epi.out_T ? ((long long)n * p.epi.out_T + t + p.epi.out_t0) * p.V + w : row;
          ln_epilogue_tile<CO, kEpiNH>(p.epi, taddr, r, RT, p.V, fr, w, row_ok, row, row_o, s_part, par, h,
                                       s_patch + (warp - 10) * kPatchBytes);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bTmemEmpty + 8 * buf);
      if (++buf == TB) { buf = 0; t_ph ^= 1; }
      tw.stop(d1);
    }
    if (warp == 10 && lane == 0) { dbg_flush(dbg, 6, d0); dbg_flush(dbg, 7, d1); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// --------------------------------------------------------------------------- //
// host side: tensor maps
// --------------------------------------------------------------------------- //
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor, `rank` dims (innermost first), 128B swizzle, zero OOB fill
inline int make_tmap_bf16(CUtensorMap *m, const void *base, int rank, const uint64_t *dims,
                          const uint64_t *strides_bytes /* rank-1 */, const uint32_t *box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is unavailable");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base),
                  reinterpret_cast<const cuuint64_t *>(dims), reinterpret_cast<const cuuint64_t *>(strides_bytes),
                  reinterpret_cast<const cuuint32_t *>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

inline int make_tmap(CUtensorMap *m, CUtensorMapDataType dt, CUtensorMapSwizzle sw, const void *base, int rank,
                    const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is unavailable");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void *>(base), reinterpret_cast<const cuuint64_t *>(dims),
                  reinterpret_cast<const cuuint64_t *>(strides_bytes), reinterpret_cast<const cuuint32_t *>(box),
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

inline int make_tmap_f32_noswizzle(CUtensorMap *m, const void *base, int rank, const uint64_t *dims,
                                   const uint64_t *strides_bytes, const uint32_t *box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail("cuTensorMapEncodeTiled is unavailable");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void *>(base),
                  reinterpret_cast<const cuuint64_t *>(dims), reinterpret_cast<const cuuint64_t *>(strides_bytes),
                  reinterpret_cast<const cuuint32_t *>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(f32) failed with CUresult %d", (int)r);
  return 0;
}

inline bool gcn_tc_supported(int c_in, int c_out, int V, int K) {
  return (c_out == 64 || c_out == 128 || c_out == 256) && c_in % 64 == 0 && c_in >= 64 && V <= 28 &&
         V >= 2 && K >= 1 && K * V + 1 <= 256;
}

inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// x: fp32 view [N][T_view][V][c_in] with frame stride `fstride` frames (residual branch: stride s);
// wp: bf16 [2][K][c_out][c_in]
template <int CO>
int launch_gcn_tc2_c(const float *x, const __nv_bfloat16 *wp, GcnTc2Params p, int N, int T_full, int fstride,
                     cudaStream_t st) {
  const int V = p.V, kMaxSmem = 232448;
  p.FT = 128 / V;
  // tiles per item; NT * CO * tb <= 512 TMEM columns.  The RT step is HBM-bound, not weight-bound:
  // single tiles leave shared memory for its double-buffered state patches.
  p.NT = (CO >= 256 || p.epi.rt_fifo) ? 1 : 2;
  p.tb = 512 / (p.NT * CO) >= 2 ? 2 : 1;
  if (p.FT < 1) return fail("gcn tensor-core kernel: %d joints do not fit a 128-row tile", V);
  const int RT = p.FT * V;
  p.a_stage_bytes = ((p.NT * RT + 128 - RT) * 128 + 1023) & ~1023;
  p.xs_tx = p.NT * p.FT * V * 64 * 4;
  p.xs_alloc = 2 * ((p.NT * p.FT * V * 128 + 1023) & ~1023);   // two 32-channel swizzled sub-tiles
  const int fixed = kGcn2Csr + kPartBytes + (p.epi.rt_fifo ? kPatchTotalRt : kPatchTotal) + 512 + 1024;
  // prefer: double-buffered input tile, 3-deep A ring, >= 2 weight stages; back off as smem requires
  const int tries[4][2] = {{2, 3}, {2, 2}, {1, 3}, {1, 2}};
  int ok = 0;
  for (int i = 0; i < 4 && !ok; ++i) {
    p.xs_bufs = tries[i][0];
    p.a_ring = tries[i][1];
    const int left = kMaxSmem - fixed - p.xs_bufs * p.xs_alloc - p.a_ring * p.a_stage_bytes;
    p.b_stages = left / (CO * 128);
    if (p.b_stages > 6) p.b_stages = 6;
    if (p.b_stages >= 2) ok = 1;
  }
  if (!ok) return fail("gcn tensor-core kernel: shared memory does not fit V=%d", V);
  const int smem = fixed + p.xs_bufs * p.xs_alloc + p.a_ring * p.a_stage_bytes + p.b_stages * CO * 128;
  p.groups_per_trial = (p.T_out + p.NT * p.FT - 1) / (p.NT * p.FT);
  p.items = N * p.groups_per_trial;

  CUtensorMap tm_x, tm_w;
  const uint64_t xd[4] = {(uint64_t)p.Cin, (uint64_t)V, (uint64_t)p.T_out, (uint64_t)N};
  const uint64_t xst[3] = {(uint64_t)p.Cin * 4, (uint64_t)fstride * V * p.Cin * 4, (uint64_t)T_full * V * p.Cin * 4};
  const uint32_t xb[4] = {32, (uint32_t)V, (uint32_t)(p.NT * p.FT), 1};
  if (make_tmap(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, x, 4, xd, xst, xb)) return 1;
  const uint64_t wd[4] = {(uint64_t)p.Cin, (uint64_t)CO, (uint64_t)p.K, 2};
  const uint64_t wst[3] = {(uint64_t)p.Cin * 2, (uint64_t)CO * p.Cin * 2, (uint64_t)p.K * CO * p.Cin * 2};
  const uint32_t wb[4] = {64, (uint32_t)CO, 1, 1};
  if (make_tmap_bf16(&tm_w, wp, 4, wd, wst, wb)) return 1;
  const int grid = p.items < num_sms() ? p.items : num_sms();
  if (p.epi.rt_fifo) {
    STGCN_CUDA_OK(cudaFuncSetAttribute(k_gcn_tc2<CO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_gcn_tc2<CO, true><<<grid, kGcn2Threads, smem, st>>>(tm_x, tm_w, p);
  } else {
    STGCN_CUDA_OK(cudaFuncSetAttribute(k_gcn_tc2<CO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_gcn_tc2<CO, false><<<grid, kGcn2Threads, smem, st>>>(tm_x, tm_w, p);
  }
  return 0;
}

inline int launch_gcn_tc2(int CO, const float *x, const __nv_bfloat16 *wp, const GcnTc2Params &p, int N,
                          int T_full, int fstride, cudaStream_t st) {
  switch (CO) {
    case 64: return launch_gcn_tc2_c<64>(x, wp, p, N, T_full, fstride, st);
    case 128: return launch_gcn_tc2_c<128>(x, wp, p, N, T_full, fstride, st);
    case 256: return launch_gcn_tc2_c<256>(x, wp, p, N, T_full, fstride, st);
  }
  return fail("gcn tensor-core kernel: unsupported channel count %d", CO);
}

// ---- v2 launcher -----------------------------------------------------------------
inline bool tcn_tc2_supported(int C, int V, int G, int stride, int T) {
  return (C == 64 || C == 128 || C == 256) && V >= 2 && V <= 64 && G <= 15 && (G & 1) &&
         (stride == 1 || (stride == 2 && T >= 2));
}

// u planes: bf16 [planes][N][T][V][C] (T = input frames); wp: bf16 [2][G][C][C]; out/res rows over T_out
// CTA-pair variant (kernels_tc_pair.cuh); STGCN_PAIR=0 selects the single-CTA kernel
inline bool tcn_pair_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_PAIR");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}
template <int C>
int launch_tcn_pair(const CUtensorMap &tm_u0, const CUtensorMap &tm_u1, const CUtensorMap &tm_wh,
                    const TcnTc2Params &p, int grid, int smem, int sets, cudaStream_t st);
// STGCN_TCN_SETS=1 / 2 forces one / two sets of eight epilogue warps in the pair kernel (0 = by mode)
inline int tcn_sets_wanted() {
  static int n = -1;
  if (n < 0) {
    const char *e = getenv("STGCN_TCN_SETS");
    n = e ? atoi(e) : 0;
  }
  return n;
}

// `halo`: the plane buffer holds `halo` extra frames before and after the T frames of every trial
// (T-split: filled by the neighbouring ranks, or zeros at the sequence ends); output frame tau
// still reads input frames stride*tau + j - pad of the T-frame sequence.
// `trial_frames` / `buf_frames` (stride 1 only): trial n starts `trial_frames` frames after trial n-1 in a plane
// of `buf_frames` frames -- overlapping windows of one shared frame sequence (sliding-window inference); the
// residual is then addressed with the same trial pitch.  0 = dense (N, T) trials.
template <int C>
int launch_tcn_tc2_c(const __nv_bfloat16 *u, const __nv_bfloat16 *wp, TcnTc2Params p, int N, int T, int stride,
                     int halo, cudaStream_t st, int trial_frames = 0, long long buf_frames = 0) {
  const int V = p.V, pad = (p.G - 1) / 2;
  if (trial_frames && (stride != 1 || halo)) return fail("tcn tensor-core kernel: windowed trials need stride 1, no halo");
  p.res_trial_rows = trial_frames ? (long long)trial_frames * V : (long long)p.T_out * V;
  const int kMaxSmem = 232448;
  // frames per tile: as many whole frames as fit 128 rows; shrink for the big stride-2 case
  int FT = 128 / V;
  int a_rows = 0;
  for (;; --FT) {
    if (FT < 1) return fail("tcn tensor-core kernel: %d joints do not fit a 128-row tile", V);
    p.FT = FT;
    p.NT = C >= 256 ? 1 : 2;
    p.tb = 512 / (p.NT * C) >= 2 ? 2 : 1;
    const int out_f = p.NT * FT, spill = 128 - FT * V;
    if (stride == 1) {
      const int wf = out_f + 2 * pad;
      p.n_loads = 1;
      p.load_f0[0] = -pad; p.load_row[0] = 0; p.load_bytes[0] = wf * V * 128;
      for (int j = 0; j < p.G; ++j) p.tap_row[j] = j * V;
      a_rows = wf * V + spill;
    } else {
      // input frame of tap j for output frame tau: 2*tau + d, d = j - pad; parity d & 1, position
      // tau + (d >> 1) in that parity's frame sequence
      int lo[2] = {1 << 30, 1 << 30}, hi[2] = {-(1 << 30), -(1 << 30)};
      for (int j = 0; j < p.G; ++j) {
        const int d = j - pad, par = d & 1, pos = d >> 1;
        if (pos < lo[par]) lo[par] = pos;
        if (pos > hi[par]) hi[par] = pos;
      }
      if (hi[1] < lo[1]) return fail("tcn tensor-core kernel: stride 2 needs a kernel of at least 3 taps");
      int nf[2], base_row[2];
      nf[0] = out_f + hi[0] - lo[0];
      nf[1] = out_f + hi[1] - lo[1];
      base_row[0] = 0;
      base_row[1] = nf[0] * V;
      p.n_loads = 2;
      for (int q = 0; q < 2; ++q) {
        p.load_f0[q] = lo[q]; p.load_row[q] = base_row[q]; p.load_bytes[q] = nf[q] * V * 128;
      }
      for (int j = 0; j < p.G; ++j) {
        const int d = j - pad, par = d & 1, pos = d >> 1;
        p.tap_row[j] = base_row[par] + (pos - lo[par]) * V;
      }
      a_rows = (nf[0] + nf[1]) * V + spill;
    }
    p.a_stage_bytes = (a_rows * 128 + 1023) & ~1023;
    const int left = kMaxSmem - 2 * p.a_stage_bytes - kPartBytes - kPatchTotal - 512 - 1024;
    p.b_stages = left / (C * 128);
    if (p.b_stages > 8) p.b_stages = 8;
    if (p.b_stages >= 2) break;
  }
  p.groups_per_trial = (p.T_out + p.NT * p.FT - 1) / (p.NT * p.FT);
  p.items = N * p.groups_per_trial;
  const int smem = 2 * p.a_stage_bytes + p.b_stages * C * 128 + kPartBytes + kPatchTotal + 512 + 1024;

  CUtensorMap tm_u0, tm_u1, tm_w;
  if (halo) {
    if (halo % stride) return fail("tcn tensor-core kernel: halo must be a multiple of the stride");
    T += 2 * halo;                                   // frames per trial in the buffer
    for (int q = 0; q < p.n_loads; ++q) p.load_f0[q] += halo / stride;
  }
  const uint64_t plane_stride = (uint64_t)(trial_frames ? buf_frames : (long long)N * T) * V * C * 2;
  if (stride == 1) {
    const uint64_t ud[5] = {(uint64_t)C, (uint64_t)V, (uint64_t)T, (uint64_t)N, (uint64_t)p.planes};
    const uint64_t us[4] = {(uint64_t)C * 2, (uint64_t)V * C * 2,
                            (uint64_t)(trial_frames ? trial_frames : T) * V * C * 2, plane_stride};
    const uint32_t ub[5] = {64, (uint32_t)V, (uint32_t)(p.load_bytes[0] / (V * 128)), 1, 1};
    if (make_tmap_bf16(&tm_u0, u, 5, ud, us, ub)) return 1;
    tm_u1 = tm_u0;
  } else {
    for (int q = 0; q < 2; ++q) {
      const uint64_t tq = q == 0 ? (uint64_t)(T + 1) / 2 : (uint64_t)T / 2;
      const uint64_t ud[5] = {(uint64_t)C, (uint64_t)V, tq, (uint64_t)N, (uint64_t)p.planes};
      const uint64_t us[4] = {(uint64_t)C * 2, (uint64_t)2 * V * C * 2, (uint64_t)T * V * C * 2, plane_stride};
      const uint32_t ub[5] = {64, (uint32_t)V, (uint32_t)(p.load_bytes[q] / (V * 128)), 1, 1};
      if (make_tmap_bf16(q == 0 ? &tm_u0 : &tm_u1, u + (size_t)q * V * C, 5, ud, us, ub)) return 1;
    }
  }
  const uint64_t wd[4] = {(uint64_t)C, (uint64_t)C, (uint64_t)p.G, 2};
  const uint64_t wst[3] = {(uint64_t)C * 2, (uint64_t)C * C * 2, (uint64_t)p.G * C * C * 2};
  const uint32_t wb[4] = {64, (uint32_t)C, 1, 1};
  if (make_tmap_bf16(&tm_w, wp, 4, wd, wst, wb)) return 1;
  if (tcn_pair_enabled() && p.items >= 2) {
    // CTA pairs (cta_group::2): every CTA stages only its half of each weight tile
    const int kBHalf = (C / 2) * 128;
    // two sets of eight epilogue warps (one per tile of the item) when the item has two tiles and the
    // extra statistics / patch buffers still leave three weight stages.  Measured: -7.5 % on the
    // temporal class in bf16 mode (one MMA per product: the kernel is epilogue-bound), +1 % in
    // bf16x3 mode (MMA / operand-bound; fewer weight stages and 96 registers per thread cost more
    // than the second set gains) -- so only bf16 mode uses it unless STGCN_TCN_SETS forces it.
    // (also with bf16-plane residual / output: the extra split and unpack work makes the C = 64 epilogue the
    // bottleneck again, +1.5 % on the whole step)
    const int want = tcn_sets_wanted() ? tcn_sets_wanted()
                                       : ((p.planes == 1 || p.epi.res_hi || (p.epi.out_hi && !p.epi.out_f32)) ? 2 : 1);
    int sets = (p.NT == 2 && want >= 2 &&
                (kMaxSmem - 2 * p.a_stage_bytes - 2 * (kPartBytes + kPatchTotal) - 1536) / kBHalf >= 3) ? 2 : 1;
    const int fixed_p = sets * (kPartBytes + kPatchTotal) + 512 + 1024;
    int S = (kMaxSmem - 2 * p.a_stage_bytes - fixed_p) / kBHalf;
    if (S > 12) S = 12;
    if (S >= 2) {
      p.b_stages = S;
      CUtensorMap tm_wh;
      const uint32_t wbh[4] = {64, (uint32_t)(C / 2), 1, 1};
      if (make_tmap_bf16(&tm_wh, wp, 4, wd, wst, wbh)) return 1;
      const int smem_p = 2 * p.a_stage_bytes + S * kBHalf + fixed_p;
      int pairs = (p.items + 1) / 2;
      if (pairs > num_sms() / 2) pairs = num_sms() / 2;
      return launch_tcn_pair<C>(tm_u0, tm_u1, tm_wh, p, 2 * pairs, smem_p, sets, st);
    }
  }
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_tcn_tc2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  k_tcn_tc2<C><<<grid, kTcn2Threads, smem, st>>>(tm_u0, tm_u1, tm_w, p);
  return 0;
}

inline int launch_tcn_tc2(int C, const __nv_bfloat16 *u, const __nv_bfloat16 *wp, const TcnTc2Params &p, int N,
                          int T, int stride, int halo, cudaStream_t st, int trial_frames = 0,
                          long long buf_frames = 0) {
  switch (C) {
    case 64: return launch_tcn_tc2_c<64>(u, wp, p, N, T, stride, halo, st, trial_frames, buf_frames);
    case 128: return launch_tcn_tc2_c<128>(u, wp, p, N, T, stride, halo, st, trial_frames, buf_frames);
    case 256: return launch_tcn_tc2_c<256>(u, wp, p, N, T, stride, halo, st, trial_frames, buf_frames);
  }
  return fail("tcn tensor-core kernel: unsupported channel count %d", C);
}

}  // namespace tc
}  // namespace stgcn
