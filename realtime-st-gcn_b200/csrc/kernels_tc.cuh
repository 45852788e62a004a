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
template <int kSleepNs = 64>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try_wait_hint(bar, parity, 20000u)) {
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
// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

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
// Temporal (Gamma x 1, stride 1) convolution as a TMA-fed tcgen05 implicit GEMM, fused with
// bias + LayerNorm(C,V) + residual + ReLU  (stgcn.py:154-161,193).
//
//   out[n,t,v,:] = relu( LN_{C,V}( sum_j Wt[:, :, j] u[n, t+j-pad, v, :] + bt ) + res[n,t,v,:] )
//
// One CTA = 8 consecutive output frames of one trial = two 128-row UMMA tiles (4 frames x 32
// padded joint rows each), all C output channels (so the LayerNorm statistics of a frame are
// CTA-local: one epilogue warp owns one frame, one lane one joint).  The input window of 16
// frames is staged ONCE per 64-channel K chunk; every temporal tap reuses it through a
// descriptor whose start address is shifted by whole frames (32 rows = 4096 B, swizzle-atom
// aligned).  TMA zero-fills joints 25..31 and frames outside [0,T): the conv's zero padding.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM -> registers -> HBM).
// --------------------------------------------------------------------------- //
constexpr int kFrameRows = 32;
constexpr int kOutFrames = 8;
constexpr int kInFrames = 16;
constexpr int kABytes = kInFrames * kFrameRows * 128;  // 65536: one plane, one 64-channel chunk
constexpr int kTcnThreads = 192;

template <int C>
struct TcnCfg {
  static constexpr int kBBytes = C * 128;                               // [C rows][64 ch] bf16
  static constexpr int kStages = C == 64 ? 8 : (C == 128 ? 5 : 3);
  static constexpr int kBarBytes = 256;
  static constexpr int kSmem = 2 * kABytes + kStages * kBBytes + kBarBytes + 1024;  // + align slack
  static constexpr int kTmemCols = 2 * C < 32 ? 32 : 2 * C;             // 128 / 256 / 512 (powers of two)
};

struct TcnTcParams {
  int T, V, G, pad;
  int planes;           // 1: bf16, 2: bf16x3
  const float *bias;    // [C]
  const float *n_w;     // [C][V]
  const float *n_b;
  const float *res;     // fp32 [rows][C] or nullptr
  float *out;           // fp32 [rows][C]
  float eps;
};

template <int C>
__global__ void __launch_bounds__(kTcnThreads, 1)
    k_tcn_tc(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_w,
             const TcnTcParams p) {
  using Cfg = TcnCfg<C>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + 2 * kABytes;
  const uint32_t sBar = sB + S * Cfg::kBBytes;
  // barrier map (8 B each): fullA[2] emptyA[2] fullB[S] emptyB[S] tmem_full, then tmem ptr
  const uint32_t bFullA = sBar, bEmptyA = sBar + 16, bFullB = sBar + 32, bEmptyB = bFullB + 8 * S;
  const uint32_t bTmemFull = bEmptyB + 8 * S;
  const uint32_t sTmemPtr = bTmemFull + 8;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * kOutFrames;
  const int KC = C / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_u);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bFullA + 8 * i, 1);
      mbar_init(bEmptyA + 8 * i, 1);
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(bFullB + 8 * i, 1);
      mbar_init(bEmptyB + 8 * i, 1);
    }
    mbar_init(bTmemFull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sTmemPtr, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // The K loop is a flat schedule of "A stages" (one plane of one 64-channel chunk of the
  // 16-frame input window), each followed by the weight tiles multiplied against it:
  //   A = hi plane: for every tap j: B = hi(j), [B = lo(j)]      (hi*hi, hi*lo)
  //   A = lo plane: for every tap j: B = hi(j)                   (lo*hi)
  if (warp == 0) {
    if (lane == 0) {
      int a_it = 0, b_it = 0;
      for (int kc = 0; kc < KC; ++kc) {
        for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
          const int as = a_it & 1;
          mbar_wait(bEmptyA + 8 * as, ((a_it >> 1) & 1) ^ 1);
          mbar_expect_tx(bFullA + 8 * as, kABytes);
          tma_load_5d(sA + as * kABytes, &tm_u, bFullA + 8 * as, kc * 64, 0, t0 - p.pad, n, ap);
          const int nb = (ap == 0) ? p.planes : 1;
          for (int j = 0; j < p.G; ++j) {
            for (int bp = 0; bp < nb; ++bp, ++b_it) {
              const int bs = b_it % S;
              mbar_wait(bEmptyB + 8 * bs, ((b_it / S) & 1) ^ 1);
              mbar_expect_tx(bFullB + 8 * bs, Cfg::kBBytes);
              tma_load_4d(sB + bs * Cfg::kBBytes, &tm_w, bFullB + 8 * bs, kc * 64, 0, j, bp);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, C);
      int a_it = 0, b_it = 0;
      uint32_t acc = 0;
      for (int kc = 0; kc < KC; ++kc) {
        for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
          const int as = a_it & 1;
          mbar_wait(bFullA + 8 * as, (a_it >> 1) & 1);
          tc_fence_after();
          const int nb = (ap == 0) ? p.planes : 1;
          for (int j = 0; j < p.G; ++j) {
            for (int bp = 0; bp < nb; ++bp, ++b_it) {
              const int bs = b_it % S;
              mbar_wait(bFullB + 8 * bs, (b_it / S) & 1);
              tc_fence_after();
#pragma unroll
              for (int m = 0; m < 2; ++m) {
                const uint32_t a0 = sA + as * kABytes + (4 * m + j) * (kFrameRows * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(tmem_base + m * C, umma_desc_sw128(a0 + k * 32),
                            umma_desc_sw128(sB + bs * Cfg::kBBytes + k * 32), idesc, acc | (uint32_t)k);
                }
              }
              acc = 1;
              umma_commit(bEmptyB + 8 * bs);  // frees the weight stage when these MMAs retire
            }
          }
          umma_commit(bEmptyA + 8 * as);      // frees the input-window stage
        }
      }
      umma_commit(bTmemFull);                 // accumulators complete
    }
  } else {
    // ---- epilogue: warp q owns TMEM lanes [32q, 32q+32) = frame q of each tile; lane = joint ----
    const int q = warp & 3;
    mbar_wait(bTmemFull, 0);
    tc_fence_after();
    const float inv_n = 1.f / (float)(p.V * C), inv_nm1 = 1.f / (float)(p.V * C - 1);
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
      const int t = t0 + 4 * m + q;
      const bool row_ok = (t < p.T) && (lane < p.V);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * C);
      const long long row = ((long long)n * p.T + t) * p.V + lane;
      float v[32];
      float s = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < C; cb += 32) {
        tmem_ld32(taddr + cb, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) s += v[i] + __ldg(p.bias + cb + i);
      }
      const float mean = warp_sum(row_ok ? s : 0.f) * inv_n;
      float ss = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < C; cb += 32) {
        tmem_ld32(taddr + cb, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = v[i] + __ldg(p.bias + cb + i) - mean;
          ss = fmaf(d, d, ss);
        }
      }
      const float rstd = 1.f / sqrtf(warp_sum(row_ok ? ss : 0.f) * inv_nm1 + p.eps);
#pragma unroll 1
      for (int cb = 0; cb < C; cb += 32) {
        tmem_ld32(taddr + cb, v);
        if (row_ok) {
          float *dst = p.out + row * C + cb;
          const float *rs = p.res ? p.res + row * C + cb : nullptr;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = cb + i + e;
              o[e] = (v[i + e] + __ldg(p.bias + c) - mean) * rstd * __ldg(p.n_w + c * p.V + lane) +
                     __ldg(p.n_b + c * p.V + lane);
            }
            if (rs) {
              const float4 r4 = *reinterpret_cast<const float4 *>(rs + i);
              o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
            }
            *reinterpret_cast<float4 *>(dst + i) =
                make_float4(fmaxf(o[0], 0.f), fmaxf(o[1], 0.f), fmaxf(o[2], 0.f), fmaxf(o[3], 0.f));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// --------------------------------------------------------------------------- //
// Shared LayerNorm epilogue of the persistent kernels.  One thread owns one accumulator row
// r = (frame r / V, joint r % V) of a 128-row tile held in TMEM (C fp32 columns):
//   y = LN_{C,V}(acc + bias) * g + b  [+ res]  [relu]  -> fp32 rows or split-bf16 planes.
// The (C,V) statistics of a frame are reduced across its V rows through `s_part`
// (float[2][128] in shared memory) and named barrier 1 (the 128 epilogue threads).
// --------------------------------------------------------------------------- //
struct EpiParams {
  const float *bias;
  int bias_sc, bias_sw;          // bias index = c * bias_sc + w * bias_sw
  const float *n_w, *n_b;        // LayerNorm affine, reference layout (C, V)
  const float *res;              // fp32 [rows][C] added after the norm, or null
  float *out_f32;                // fp32 [rows][C] or null
  __nv_bfloat16 *out_hi, *out_lo;  // split-bf16 planes [rows][C] or null
  int relu;
  float eps;
  int debug;                     // measurement aid: 1 = skip the epilogue body, 2 = skip transform math
};

// Statistics in ONE pass over TMEM: each row accumulates sum / sum of squares of (x - shift) with
// shift = its first element (so the squares do not cancel), giving a row mean and a row M2; the V
// rows of a frame are then merged exactly (Chan et al.):  M2 = sum M2_r + C * sum (m_r - mean)^2.
// `s_part` is float[2][256]; `tile_parity` alternates the half used so one barrier per tile suffices.
template <int C>
__device__ __forceinline__ void ln_epilogue_tile(const EpiParams &e, uint32_t taddr, int r, int RT, int V, int fr,
                                                 int w, bool row_ok, long long row, float *s_part,
                                                 int tile_parity) {
  if (e.debug & 1) return;
  const float *bias = e.bias + w * e.bias_sw;
  float *sp = s_part + tile_parity * 256;
  float v[16];
  float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll 1
  for (int cb = 0; cb < C; cb += 16) {
    tmem_ld16(taddr + cb, v);
    if (cb == 0) shift = v[0] + __ldg(bias);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = v[i] + __ldg(bias + (cb + i) * e.bias_sc) - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
  }
  const float m_r = shift + s1 * (1.f / (float)C);
  const float M2_r = fmaxf(s2 - s1 * s1 * (1.f / (float)C), 0.f);
  sp[r] = row_ok ? m_r : 0.f;
  sp[128 + r] = row_ok ? M2_r : 0.f;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  float mean = 0.f, rstd = 0.f;
  if (r < RT) {
    float tm = 0.f, tq = 0.f;
    for (int j = 0; j < V; ++j) tm += sp[fr * V + j];
    mean = tm * (1.f / (float)V);
    for (int j = 0; j < V; ++j) {
      const float dm = sp[fr * V + j] - mean;
      tq += sp[128 + fr * V + j] + (float)C * dm * dm;
    }
    rstd = 1.f / sqrtf(tq * (1.f / (float)(V * C - 1)) + e.eps);
  }
#pragma unroll 1
  for (int cb = 0; cb < C; cb += 16) {
    tmem_ld16(taddr + cb, v);
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = cb + i;
        v[i] = (v[i] + __ldg(bias + c * e.bias_sc) - mean) * rstd * __ldg(e.n_w + c * V + w) +
               __ldg(e.n_b + c * V + w);
      }
      if (e.res) {
        const float4 *rs = reinterpret_cast<const float4 *>(e.res + row * C + cb);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 r4 = rs[i];
          v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
        }
      }
      if (e.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (e.out_f32) {
        float4 *dst = reinterpret_cast<float4 *>(e.out_f32 + row * C + cb);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      if (e.out_hi) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[2 * i], h0, l0);
          split_bf16(v[2 * i + 1], h1, l1);
          __nv_bfloat162 hh(h0, h1), ll(l0, l1);
          hi[i] = *reinterpret_cast<uint32_t *>(&hh);
          lo[i] = *reinterpret_cast<uint32_t *>(&ll);
        }
        uint4 *dh = reinterpret_cast<uint4 *>(e.out_hi + row * C + cb);
        dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        if (e.out_lo) {
          uint4 *dl = reinterpret_cast<uint4 *>(e.out_lo + row * C + cb);
          dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
      }
    }
  }
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
constexpr int kTcn2Threads = 224;   // warps: 0 A producer, 1 MMA, 2 B producer, 3..6 epilogue

struct TcnTc2Params {
  int T_out, V, G, planes;
  int FT, NT;                 // frames per 128-row tile, tiles per item
  int groups_per_trial, items;
  int a_stage_bytes, b_stages;
  int n_loads;                // TMA loads per A stage (1: stride 1, 2: stride 2 even/odd windows)
  int load_f0[2];             // first frame of the window relative to the item's first output frame
  int load_row[2];            // destination row in the stage
  int load_bytes[2];
  int tap_row[16];            // first stage row of tap j for tile 0
  EpiParams epi;
};

template <int C>
__global__ void __launch_bounds__(kTcn2Threads, 1)
    k_tcn_tc2(const __grid_constant__ CUtensorMap tm_u0, const __grid_constant__ CUtensorMap tm_u1,
              const __grid_constant__ CUtensorMap tm_w, const TcnTc2Params p) {
  constexpr int kBBytes = C * 128;
  constexpr int TB = (4 * C <= 512) ? 2 : 1;      // TMEM accumulator buffers (2 tiles each)
  constexpr int kTmemCols = TB * 2 * C;           // 256 / 512 / 512
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int S = p.b_stages;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + 2 * p.a_stage_bytes;
  const uint32_t sPart = sB + S * kBBytes;                    // float[2][256]
  const uint32_t sBar = sPart + 2048;
  const uint32_t bFullA = sBar, bEmptyA = sBar + 16, bTmemFull = sBar + 32, bTmemEmpty = sBar + 48;
  const uint32_t bFullB = sBar + 64, bEmptyB = bFullB + 8 * S;
  const uint32_t sTmemPtr = bEmptyB + 8 * S;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  float *s_part = reinterpret_cast<float *>(gen_base + (sPart - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = C / 64;
  const int RT = p.FT * p.V;                      // useful rows per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_u0);
    tma_prefetch_desc(&tm_u1);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bFullA + 8 * i, 1);
      mbar_init(bEmptyA + 8 * i, 1);
      mbar_init(bTmemFull + 8 * i, 1);
      mbar_init(bTmemEmpty + 8 * i, 4);           // one arrive per epilogue warp
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
      int a_it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.groups_per_trial;
        const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;   // first output frame
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
            const int as = a_it & 1;
            mbar_wait(bEmptyA + 8 * as, ((a_it >> 1) & 1) ^ 1);
            mbar_expect_tx(bFullA + 8 * as, (uint32_t)(p.load_bytes[0] + (p.n_loads > 1 ? p.load_bytes[1] : 0)));
            tma_load_5d(sA + as * p.a_stage_bytes + p.load_row[0] * 128, &tm_u0, bFullA + 8 * as, kc * 64, 0,
                        f0 + p.load_f0[0], n, ap);
            if (p.n_loads > 1)
              tma_load_5d(sA + as * p.a_stage_bytes + p.load_row[1] * 128, &tm_u1, bFullA + 8 * as, kc * 64, 0,
                          f0 + p.load_f0[1], n, ap);
          }
      }
    }
  } else if (warp == 2) {
    // ---- weight-tile producer ----
    if (lane == 0) {
      int b_it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x)
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap) {
            const int nb = (ap == 0) ? p.planes : 1;
            for (int j = 0; j < p.G; ++j)
              for (int bp = 0; bp < nb; ++bp, ++b_it) {
                const int bs = b_it % S;
                mbar_wait(bEmptyB + 8 * bs, ((b_it / S) & 1) ^ 1);
                mbar_expect_tx(bFullB + 8 * bs, kBBytes);
                tma_load_4d(sB + bs * kBBytes, &tm_w, bFullB + 8 * bs, kc * 64, 0, j, bp);
              }
          }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, C);
      int a_it = 0, b_it = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int buf = it % TB;
        mbar_wait(bTmemEmpty + 8 * buf, ((it / TB) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * 2 * C;
        uint32_t acc = 0;
        for (int kc = 0; kc < KC; ++kc)
          for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
            const int as = a_it & 1;
            mbar_wait(bFullA + 8 * as, (a_it >> 1) & 1);
            tc_fence_after();
            const int nb = (ap == 0) ? p.planes : 1;
            for (int j = 0; j < p.G; ++j)
              for (int bp = 0; bp < nb; ++bp, ++b_it) {
                const int bs = b_it % S;
                mbar_wait(bFullB + 8 * bs, (b_it / S) & 1);
                tc_fence_after();
                for (int m = 0; m < p.NT; ++m) {
                  const uint32_t a0 = sA + as * p.a_stage_bytes + (p.tap_row[j] + m * RT) * 128;
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_bf16(tacc + m * C, umma_desc_sw128(a0 + k * 32),
                              umma_desc_sw128(sB + bs * kBBytes + k * 32), idesc, acc | (uint32_t)k);
                }
                acc = 1;
                umma_commit(bEmptyB + 8 * bs);
              }
            umma_commit(bEmptyA + 8 * as);
          }
        umma_commit(bTmemFull + 8 * buf);
      }
    }
  } else {
    // ---- epilogue: thread <-> accumulator row r = 32*(warp&3) + lane = (frame r / V, joint r % V) ----
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int fr = r / p.V, w = r - fr * p.V;
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int buf = it % TB;
      const int n = item / p.groups_per_trial;
      const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;
      mbar_wait(bTmemFull + 8 * buf, (it / TB) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < p.NT; ++m) {
        const int t = f0 + m * p.FT + fr;
        const bool row_ok = (r < RT) && (t < p.T_out);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * C + m * C);
        const long long row = ((long long)n * p.T_out + t) * p.V + w;
        ln_epilogue_tile<C>(p.epi, taddr, r, RT, p.V, fr, w, row_ok, row, s_part, m & 1);
      }
      // accumulator buffer drained: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bTmemEmpty + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// --------------------------------------------------------------------------- //
// Graph-convolution stage on tensor cores, fused with LayerNorm(C,V) + ReLU
// (tgcn.py:70-79 + stgcn.py:152-153):
//
//   u[n,t,w,:] = relu( LN_{C,V}( sum_k sum_v A[k,v,w] * (Wg_k x[n,t,v,:] + bg_k) ) )
//
// The 1x1 feature transform and the V x V adjacency contraction act on different indices, so
// they commute: the contraction is applied FIRST, on the C_in-channel input tile staged in
// shared memory (A is a tree adjacency: ~1 non-zero per (k,w), kept in CSR, dense A still
// works), producing xa_k[(t,w), ci] = sum_v A[k,v,w] x[(t,v), ci]; then ONE GEMM with
// K = 3*C_in gives z directly:  z[(t,w), c] = sum_k sum_ci xa_k[(t,w),ci] Wg[k*C_out+c, ci].
// That keeps the whole z tile (8 frames x C_out) in TMEM, so the LayerNorm statistics are one
// warp-shuffle reduction per frame in the epilogue, and it needs no 3*C_out-wide intermediate.
// The bias term flows through A: bz[w,c] = sum_k bg[k*C_out+c] * colsum_k[w] (precomputed).
//
// Warp roles: 0 = TMA producer (fp32 input tile + bf16 weight tiles), 1 = MMA issuer,
// 2..9 = transform warps (contraction + bf16 hi/lo split -> swizzled UMMA A operand in smem),
// of which 2..5 then run the epilogue.
// --------------------------------------------------------------------------- //
constexpr int kGcnThreads = 320;
constexpr int kGcnXform = 8;                 // transform warps
constexpr int kGcnAStage = 2 * 128 * 128;    // 2 tiles x 128 rows x 128 B = 32768
constexpr int kGcnARing = 3;
constexpr int kGcnCsrMax = 384;              // CSR entries cached in shared memory (tree graphs: ~75)
constexpr int kGcnCsrBytes = 4096;           // ptr[K*V+1] + kGcnCsrMax (v, a) pairs
constexpr int kGcnMaxJoints = 4;             // joints per transform warp (V <= 32, 8 warps)

template <int CO>
struct GcnCfg {
  static constexpr int kBBytes = CO * 128;
  static constexpr int kStages = CO == 256 ? 2 : 4;
  static constexpr int kTmemCols = 2 * CO;
  // input tile: 8 frames x V joints x 64 fp32 channels, rounded up to 1 KB
  static int xs_bytes(int V) { return (kOutFrames * V * 64 * 4 + 1023) & ~1023; }
  static int smem(int V) {
    return kGcnARing * kGcnAStage + xs_bytes(V) + kStages * kBBytes + kGcnCsrBytes + 256 + 1024;
  }
};

struct GcnTcParams {
  int T, V, K, Cin;
  int planes;
  int xs_alloc;         // bytes reserved for the fp32 input tile (GcnCfg::xs_bytes)
  const int *csr_ptr;   // [K*V + 1], (k,w)-major
  const int2 *csr_va;   // per entry: (source joint v, bits of A[k,v,w])
  const float *bzT;     // [CO][V] bias through the adjacency
  const float *n_w, *n_b;
  __nv_bfloat16 *out_hi, *out_lo;   // [rows][CO] planes (tensor-core temporal stage) or null
  float *out_f32;                   // [rows][CO] (CUDA-core temporal stage) or null
  float eps;
};


template <int CO>
__global__ void __launch_bounds__(kGcnThreads, 1)
    k_gcn_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
             const GcnTcParams p) {
  using Cfg = GcnCfg<CO>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sXs = sA + kGcnARing * kGcnAStage;
  const uint32_t sB = sXs + p.xs_alloc;
  const uint32_t sCsr = sB + S * Cfg::kBBytes;
  const uint32_t sBar = sCsr + kGcnCsrBytes;
  const uint32_t bXsFull = sBar, bXsEmpty = sBar + 8;
  const uint32_t bAFull = sBar + 16, bAEmpty = bAFull + 8 * kGcnARing;
  const uint32_t bFullB = bAEmpty + 8 * kGcnARing, bEmptyB = bFullB + 8 * S;
  const uint32_t bTmemFull = bEmptyB + 8 * S;
  const uint32_t sTmemPtr = bTmemFull + 8;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  const float *xs = reinterpret_cast<const float *>(gen_base + (sXs - smem_base));
  uint8_t *a_gen = gen_base;  // A ring starts at smem_base
  int *s_ptr = reinterpret_cast<int *>(gen_base + (sCsr - smem_base));          // [K*V + 1]
  int2 *s_va = reinterpret_cast<int2 *>(gen_base + (sCsr - smem_base) + 1024);  // [kGcnCsrMax]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * kOutFrames;
  const int KC = p.Cin / 64;
  const uint32_t xs_bytes = (uint32_t)(kOutFrames * p.V * 64 * 4);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    mbar_init(bXsFull, 1);
    mbar_init(bXsEmpty, kGcnXform);
    for (int i = 0; i < kGcnARing; ++i) {
      mbar_init(bAFull + 8 * i, kGcnXform);
      mbar_init(bAEmpty + 8 * i, 1);
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(bFullB + 8 * i, 1);
      mbar_init(bEmptyB + 8 * i, 1);
    }
    mbar_init(bTmemFull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sTmemPtr, Cfg::kTmemCols);
  const int nnz = __ldg(p.csr_ptr + p.K * p.V);
  const bool csr_smem = nnz <= kGcnCsrMax && p.K * p.V + 1 <= 256;
  if (csr_smem) {
    for (int i = threadIdx.x; i <= p.K * p.V; i += blockDim.x) s_ptr[i] = __ldg(p.csr_ptr + i);
    for (int i = threadIdx.x; i < nnz; i += blockDim.x) s_va[i] = __ldg(p.csr_va + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    if (lane == 0) {
      int b_it = 0;
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(bXsEmpty, (kc & 1) ^ 1);
        mbar_expect_tx(bXsFull, xs_bytes);
        tma_load_4d(sXs, &tm_x, bXsFull, kc * 64, 0, t0, n);
        for (int k = 0; k < p.K; ++k)
          for (int ap = 0; ap < p.planes; ++ap) {
            const int nb = (ap == 0) ? p.planes : 1;
            for (int bp = 0; bp < nb; ++bp, ++b_it) {
              const int bs = b_it % S;
              mbar_wait(bEmptyB + 8 * bs, ((b_it / S) & 1) ^ 1);
              mbar_expect_tx(bFullB + 8 * bs, Cfg::kBBytes);
              tma_load_4d(sB + bs * Cfg::kBBytes, &tm_w, bFullB + 8 * bs, kc * 64, 0, k, bp);
            }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, CO);
      int a_it = 0, b_it = 0;
      uint32_t acc = 0;
      for (int kc = 0; kc < KC; ++kc)
        for (int k = 0; k < p.K; ++k)
          for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
            const int as = a_it % kGcnARing;
            mbar_wait(bAFull + 8 * as, (a_it / kGcnARing) & 1);
            tc_fence_after();
            const int nb = (ap == 0) ? p.planes : 1;
            for (int bp = 0; bp < nb; ++bp, ++b_it) {
              const int bs = b_it % S;
              mbar_wait(bFullB + 8 * bs, (b_it / S) & 1);
              tc_fence_after();
#pragma unroll
              for (int m = 0; m < 2; ++m) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16(tmem_base + m * CO, umma_desc_sw128(sA + as * kGcnAStage + m * 16384 + kk * 32),
                            umma_desc_sw128(sB + bs * Cfg::kBBytes + kk * 32), idesc, acc | (uint32_t)kk);
              }
              acc = 1;
              umma_commit(bEmptyB + 8 * bs);
            }
            umma_commit(bAEmpty + 8 * as);
          }
      umma_commit(bTmemFull);
    }
  } else {
    // ---- transform warps: xa_k = A_k-contraction of the staged fp32 tile -> bf16 planes ----
    // Warp tw owns joints w = tw, tw+8, ... for all 8 frames (the 8 frames share every CSR
    // entry, giving 8 independent shared-memory loads per entry); lane owns channels 2l, 2l+1.
    // The fp32 results stay in registers between the hi-plane and the lo-plane stage.
    const int tw = warp - 2;
    int a_it = 0;
    float2 xa[kGcnMaxJoints][kOutFrames];
    for (int kc = 0; kc < KC; ++kc) {
      mbar_wait(bXsFull, kc & 1);
      for (int k = 0; k < p.K; ++k)
        for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
          const int as = a_it % kGcnARing;
          if (ap == 0) {
#pragma unroll
            for (int jw = 0; jw < kGcnMaxJoints; ++jw) {
              const int w = tw + jw * kGcnXform;
#pragma unroll
              for (int f = 0; f < kOutFrames; ++f) xa[jw][f] = make_float2(0.f, 0.f);
              if (w < p.V) {
                const int e0 = csr_smem ? s_ptr[k * p.V + w] : __ldg(p.csr_ptr + k * p.V + w);
                const int e1 = csr_smem ? s_ptr[k * p.V + w + 1] : __ldg(p.csr_ptr + k * p.V + w + 1);
                for (int e = e0; e < e1; ++e) {
                  const int2 va = csr_smem ? s_va[e] : __ldg(p.csr_va + e);
                  const float a = __int_as_float(va.y);
                  const float *xr = xs + va.x * 64 + 2 * lane;
#pragma unroll
                  for (int f = 0; f < kOutFrames; ++f) {
                    const float2 xv = *reinterpret_cast<const float2 *>(xr + f * p.V * 64);
                    xa[jw][f].x = fmaf(a, xv.x, xa[jw][f].x);
                    xa[jw][f].y = fmaf(a, xv.y, xa[jw][f].y);
                  }
                }
              }
            }
          }
          mbar_wait(bAEmpty + 8 * as, ((a_it / kGcnARing) & 1) ^ 1);
          uint8_t *stage = a_gen + as * kGcnAStage;
#pragma unroll
          for (int jw = 0; jw < kGcnMaxJoints; ++jw) {
            const int w = tw + jw * kGcnXform;
            if (w < p.V) {
#pragma unroll
              for (int f = 0; f < kOutFrames; ++f) {
                __nv_bfloat16 hx, lx, hy, ly;
                split_bf16(xa[jw][f].x, hx, lx);
                split_bf16(xa[jw][f].y, hy, ly);
                const __nv_bfloat162 pk = ap == 0 ? __nv_bfloat162(hx, hy) : __nv_bfloat162(lx, ly);
                const int R = f * kFrameRows + w;               // row in the 256-row stage
                const int chunk = (lane >> 2) ^ (R & 7);        // 128B swizzle: 16B chunk ^ (row % 8)
                *reinterpret_cast<__nv_bfloat162 *>(stage + R * 128 + chunk * 16 + (lane & 3) * 4) = pk;
              }
            }
          }
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(bAFull + 8 * as);
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(bXsEmpty);
    }
    if (warp < 6) {
      // ---- epilogue (warps 2..5): one frame per warp, one joint per lane ----
      const int q = warp & 3;
      mbar_wait(bTmemFull, 0);
      tc_fence_after();
      const float inv_n = 1.f / (float)(p.V * CO), inv_nm1 = 1.f / (float)(p.V * CO - 1);
      const int lv = lane < p.V ? lane : 0;
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        const int t = t0 + 4 * m + q;
        const bool row_ok = (t < p.T) && (lane < p.V);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * CO);
        const long long row = ((long long)n * p.T + t) * p.V + lane;
        float v[32];
        float s = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < CO; cb += 32) {
          tmem_ld32(taddr + cb, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) s += v[i] + __ldg(p.bzT + (cb + i) * p.V + lv);
        }
        const float mean = warp_sum(row_ok ? s : 0.f) * inv_n;
        float ss = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < CO; cb += 32) {
          tmem_ld32(taddr + cb, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float d = v[i] + __ldg(p.bzT + (cb + i) * p.V + lv) - mean;
            ss = fmaf(d, d, ss);
          }
        }
        const float rstd = 1.f / sqrtf(warp_sum(row_ok ? ss : 0.f) * inv_nm1 + p.eps);
#pragma unroll 1
        for (int cb = 0; cb < CO; cb += 32) {
          tmem_ld32(taddr + cb, v);
          if (row_ok) {
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = cb + i;
              const float y = (v[i] + __ldg(p.bzT + c * p.V + lane) - mean) * rstd * __ldg(p.n_w + c * p.V + lane) +
                              __ldg(p.n_b + c * p.V + lane);
              o[i] = fmaxf(y, 0.f);
            }
            if (p.out_f32) {
              float *dst = p.out_f32 + row * CO + cb;
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4 *>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
            }
            if (p.out_hi) {
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(o[2 * i], h0, l0);
                split_bf16(o[2 * i + 1], h1, l1);
                __nv_bfloat162 hh(h0, h1), ll(l0, l1);
                hi[i] = *reinterpret_cast<uint32_t *>(&hh);
                lo[i] = *reinterpret_cast<uint32_t *>(&ll);
              }
              uint4 *dh = reinterpret_cast<uint4 *>(p.out_hi + row * CO + cb);
#pragma unroll
              for (int i = 0; i < 4; ++i) dh[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
              if (p.out_lo) {
                uint4 *dl = reinterpret_cast<uint4 *>(p.out_lo + row * CO + cb);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  dl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// adjacency CSR ordered by (k, w): ptr[k*V + w] .. ptr[k*V + w + 1] -> (v, A[k,v,w]); and the bias
// that flows through A, transposed for coalesced epilogue reads: bzT[c][w].
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
  const int c = i / V, w = i - c * V;
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    float col = 0.f;
    for (int v = 0; v < V; ++v) col += A[((long long)k * V + v) * V + w];
    s = fmaf(bg[k * CO + c], col, s);
  }
  bzT[i] = s;
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
constexpr int kGcn2Threads = 480;   // + warp 14: weight-tile producer
constexpr int kGcn2Csr = 2048;               // ptr[<=128] + 192 entries
constexpr int kGcn2CsrMax = 192;

struct GcnTc2Params {
  int T_out, V, K, Cin, planes;
  int FT, NT, groups_per_trial, items;
  int a_stage_bytes, a_ring, xs_alloc, xs_tx, xs_bufs, b_stages;
  int identity;
  const int *csr_ptr;
  const int2 *csr_va;
  EpiParams epi;
};

template <int CO>
__global__ void __launch_bounds__(kGcn2Threads, 1)
    k_gcn_tc2(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
              const GcnTc2Params p) {
  constexpr int kBBytes = CO * 128;
  constexpr int TB = (4 * CO <= 512) ? 2 : 1;
  constexpr int kTmemCols = TB * 2 * CO;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int S = p.b_stages, AR = p.a_ring, XB = p.xs_bufs;
  const uint32_t sA = smem_base;
  const uint32_t sXs = sA + AR * p.a_stage_bytes;
  const uint32_t sB = sXs + XB * p.xs_alloc;
  const uint32_t sCsr = sB + S * kBBytes;
  const uint32_t sPart = sCsr + kGcn2Csr;
  const uint32_t sBar = sPart + 2048;
  const uint32_t bXsFull = sBar, bXsEmpty = sBar + 16, bTmemFull = sBar + 32, bTmemEmpty = sBar + 48;
  const uint32_t bAFull = sBar + 64, bAEmpty = sBar + 96;
  const uint32_t bFullB = sBar + 128, bEmptyB = bFullB + 8 * S;
  const uint32_t sTmemPtr = bEmptyB + 8 * S;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  int *s_ptr = reinterpret_cast<int *>(gen_base + (sCsr - smem_base));
  int2 *s_va = reinterpret_cast<int2 *>(gen_base + (sCsr - smem_base) + 512);
  float *s_part = reinterpret_cast<float *>(gen_base + (sPart - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = p.Cin / 64;
  const int RT = p.FT * p.V;
  const int rows_item = p.NT * RT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bXsFull + 8 * i, 1);
      mbar_init(bXsEmpty + 8 * i, 8);
      mbar_init(bTmemFull + 8 * i, 1);
      mbar_init(bTmemEmpty + 8 * i, 4);
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
    // ---- fp32 input-tile producer ----
    if (lane == 0) {
      int x_it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int n = item / p.groups_per_trial;
        const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;
        for (int kc = 0; kc < KC; ++kc, ++x_it) {
          const int xb = x_it % XB;
          mbar_wait(bXsEmpty + 8 * xb, ((x_it / XB) & 1) ^ 1);
          mbar_expect_tx(bXsFull + 8 * xb, (uint32_t)p.xs_tx);
          tma_load_4d(sXs + xb * p.xs_alloc, &tm_x, bXsFull + 8 * xb, kc * 64, 0, f0, n);
          tma_load_4d(sXs + xb * p.xs_alloc + p.xs_alloc / 2, &tm_x, bXsFull + 8 * xb, kc * 64 + 32, 0, f0, n);
        }
      }
    }
  } else if (warp == 14) {
    // ---- weight-tile producer ----
    if (lane == 0) {
      int b_it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x)
        for (int kc = 0; kc < KC; ++kc)
          for (int k = 0; k < p.K; ++k)
            for (int ap = 0; ap < p.planes; ++ap) {
              const int nb = (ap == 0) ? p.planes : 1;
              for (int bp = 0; bp < nb; ++bp, ++b_it) {
                const int bs = b_it % S;
                mbar_wait(bEmptyB + 8 * bs, ((b_it / S) & 1) ^ 1);
                mbar_expect_tx(bFullB + 8 * bs, kBBytes);
                tma_load_4d(sB + bs * kBBytes, &tm_w, bFullB + 8 * bs, kc * 64, 0, k, bp);
              }
            }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, CO);
      int a_it = 0, b_it = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int buf = it % TB;
        mbar_wait(bTmemEmpty + 8 * buf, ((it / TB) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * 2 * CO;
        uint32_t acc = 0;
        for (int kc = 0; kc < KC; ++kc)
          for (int k = 0; k < p.K; ++k)
            for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
              const int as = a_it % AR;
              mbar_wait(bAFull + 8 * as, (a_it / AR) & 1);
              tc_fence_after();
              const int nb = (ap == 0) ? p.planes : 1;
              for (int bp = 0; bp < nb; ++bp, ++b_it) {
                const int bs = b_it % S;
                mbar_wait(bFullB + 8 * bs, (b_it / S) & 1);
                tc_fence_after();
                for (int m = 0; m < p.NT; ++m) {
                  const uint32_t a0 = sA + as * p.a_stage_bytes + m * RT * 128;
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_bf16(tacc + m * CO, umma_desc_sw128(a0 + kk * 32),
                              umma_desc_sw128(sB + bs * kBBytes + kk * 32), idesc, acc | (uint32_t)kk);
                }
                acc = 1;
                umma_commit(bEmptyB + 8 * bs);
              }
              umma_commit(bAEmpty + 8 * as);
            }
        umma_commit(bTmemFull + 8 * buf);
      }
    }
  } else if (warp < 10) {
    // ---- transform warps: ONE THREAD PER ROW (256 threads >= rows of an item) ----
    // The fp32 input tile is staged as two 32-channel sub-tiles with the 128-B TMA swizzle, so
    // threads of a warp (consecutive rows) read 16-B chunks from distinct banks.  Each thread
    // walks its row in groups of 4 channels: gather-sum over the CSR entries of (k, joint),
    // split into bf16 hi/lo, store hi as 8 B into the swizzled UMMA A stage and keep lo packed
    // in registers for the following lo-plane stage.
    const int r = (warp - 2) * 32 + lane;
    const bool r_ok = r < rows_item;
    const int f = r / p.V, w = r - f * p.V;
    const int src_base = f * p.V;                       // first source row of this row's frame
    int a_it = 0, x_it = 0;
    uint2 lo_stash[16];
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      for (int kc = 0; kc < KC; ++kc, ++x_it) {
        const int xb = x_it % XB;
        mbar_wait(bXsFull + 8 * xb, (x_it / XB) & 1);
        const uint8_t *xs = gen_base + (sXs - smem_base) + xb * p.xs_alloc;
        const int half = p.xs_alloc / 2;
        for (int k = 0; k < p.K; ++k)
          for (int ap = 0; ap < p.planes; ++ap, ++a_it) {
            const int as = a_it % AR;
            int e0 = 0, e1 = 0;
            if (ap == 0 && r_ok && !p.identity) {
              e0 = csr_smem ? s_ptr[k * p.V + w] : __ldg(p.csr_ptr + k * p.V + w);
              e1 = csr_smem ? s_ptr[k * p.V + w + 1] : __ldg(p.csr_ptr + k * p.V + w + 1);
            }
            mbar_wait(bAEmpty + 8 * as, ((a_it / AR) & 1) ^ 1);
            uint8_t *dst_row = gen_base + as * p.a_stage_bytes + r * 128;
            if (r_ok && !(p.epi.debug & 2)) {
              if (ap == 0) {
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                  const int sub = g >> 3, j = g & 7;           // 32-channel sub-tile, 16-B chunk
                  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (p.identity) {
                    acc = *reinterpret_cast<const float4 *>(xs + sub * half + r * 128 + ((j ^ (r & 7)) << 4));
                  } else {
                    for (int e = e0; e < e1; ++e) {
                      const int2 va = csr_smem ? s_va[e] : __ldg(p.csr_va + e);
                      const float a = __int_as_float(va.y);
                      const int rs = src_base + va.x;
                      const float4 xv =
                          *reinterpret_cast<const float4 *>(xs + sub * half + rs * 128 + ((j ^ (rs & 7)) << 4));
                      acc.x = fmaf(a, xv.x, acc.x);
                      acc.y = fmaf(a, xv.y, acc.y);
                      acc.z = fmaf(a, xv.z, acc.z);
                      acc.w = fmaf(a, xv.w, acc.w);
                    }
                  }
                  const __nv_bfloat162 h01 = __floats2bfloat162_rn(acc.x, acc.y);
                  const __nv_bfloat162 h23 = __floats2bfloat162_rn(acc.z, acc.w);
                  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                  const __nv_bfloat162 l01 = __floats2bfloat162_rn(acc.x - f01.x, acc.y - f01.y);
                  const __nv_bfloat162 l23 = __floats2bfloat162_rn(acc.z - f23.x, acc.w - f23.y);
                  lo_stash[g] = make_uint2(*reinterpret_cast<const uint32_t *>(&l01),
                                           *reinterpret_cast<const uint32_t *>(&l23));
                  *reinterpret_cast<uint2 *>(dst_row + ((((sub << 2) | (j >> 1)) ^ (r & 7)) << 4) + ((j & 1) << 3)) =
                      make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
                }
              } else {
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                  const int sub = g >> 3, j = g & 7;
                  *reinterpret_cast<uint2 *>(dst_row + ((((sub << 2) | (j >> 1)) ^ (r & 7)) << 4) + ((j & 1) << 3)) =
                      lo_stash[g];
                }
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bAFull + 8 * as);
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(bXsEmpty + 8 * xb);
      }
    }
  } else if (warp < 14) {
    // ---- epilogue warps 10..13 ----
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int fr = r / p.V, w = r - fr * p.V;
    int it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int buf = it % TB;
      const int n = item / p.groups_per_trial;
      const int f0 = (item - n * p.groups_per_trial) * p.NT * p.FT;
      mbar_wait(bTmemFull + 8 * buf, (it / TB) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < p.NT; ++m) {
        const int t = f0 + m * p.FT + fr;
        const bool row_ok = (r < RT) && (t < p.T_out);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * CO + m * CO);
        const long long row = ((long long)n * p.T_out + t) * p.V + w;
        ln_epilogue_tile<CO>(p.epi, taddr, r, RT, p.V, fr, w, row_ok, row, s_part, m & 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bTmemEmpty + 8 * buf);
    }
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

// x: fp32 [N][T][V][c_in]; wp: bf16 [2][K][c_out][c_in]
template <int CO>
int launch_gcn_tc_c(const float *x, const __nv_bfloat16 *wp, const GcnTcParams &p, int N, cudaStream_t st) {
  CUtensorMap tm_x, tm_w;
  const uint64_t xd[4] = {(uint64_t)p.Cin, (uint64_t)p.V, (uint64_t)p.T, (uint64_t)N};
  const uint64_t xst[3] = {(uint64_t)p.Cin * 4, (uint64_t)p.V * p.Cin * 4, (uint64_t)p.T * p.V * p.Cin * 4};
  const uint32_t xb[4] = {64, (uint32_t)p.V, kOutFrames, 1};
  if (make_tmap_f32_noswizzle(&tm_x, x, 4, xd, xst, xb)) return 1;
  const uint64_t wd[4] = {(uint64_t)p.Cin, (uint64_t)CO, (uint64_t)p.K, 2};
  const uint64_t wst[3] = {(uint64_t)p.Cin * 2, (uint64_t)CO * p.Cin * 2, (uint64_t)p.K * CO * p.Cin * 2};
  const uint32_t wb[4] = {64, (uint32_t)CO, 1, 1};
  if (make_tmap_bf16(&tm_w, wp, 4, wd, wst, wb)) return 1;
  GcnTcParams q = p;
  q.xs_alloc = GcnCfg<CO>::xs_bytes(p.V);
  const int smem = GcnCfg<CO>::smem(p.V);
  if (smem > 232448) return fail("gcn tensor-core kernel: %d joints need %d B of shared memory", p.V, smem);
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_gcn_tc<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((p.T + kOutFrames - 1) / kOutFrames, N);
  k_gcn_tc<CO><<<grid, kGcnThreads, smem, st>>>(tm_x, tm_w, q);
  return 0;
}

inline int launch_gcn_tc(int CO, const float *x, const __nv_bfloat16 *wp, const GcnTcParams &p, int N,
                         cudaStream_t st) {
  switch (CO) {
    case 64: return launch_gcn_tc_c<64>(x, wp, p, N, st);
    case 128: return launch_gcn_tc_c<128>(x, wp, p, N, st);
    case 256: return launch_gcn_tc_c<256>(x, wp, p, N, st);
  }
  return fail("gcn tensor-core kernel: unsupported channel count %d", CO);
}

inline bool tcn_tc_supported(int C, int V, int G, int stride) {
  return (C == 64 || C == 128 || C == 256) && V <= kFrameRows && G <= 9 && (G & 1) && stride == 1;
}

// u planes: bf16 [planes][N][T][V][C]; wp: bf16 [2][G][C][C]
template <int C>
int launch_tcn_tc_c(const __nv_bfloat16 *u, const __nv_bfloat16 *wp, const TcnTcParams &p, int N,
                    cudaStream_t st) {
  CUtensorMap tm_u, tm_w;
  const uint64_t ud[5] = {(uint64_t)C, (uint64_t)p.V, (uint64_t)p.T, (uint64_t)N, (uint64_t)p.planes};
  const uint64_t us[4] = {(uint64_t)C * 2, (uint64_t)p.V * C * 2, (uint64_t)p.T * p.V * C * 2,
                          (uint64_t)N * p.T * p.V * C * 2};
  const uint32_t ub[5] = {64, kFrameRows, kInFrames, 1, 1};
  if (make_tmap_bf16(&tm_u, u, 5, ud, us, ub)) return 1;
  const uint64_t wd[4] = {(uint64_t)C, (uint64_t)C, (uint64_t)p.G, 2};
  const uint64_t wst[3] = {(uint64_t)C * 2, (uint64_t)C * C * 2, (uint64_t)p.G * C * C * 2};
  const uint32_t wb[4] = {64, (uint32_t)C, 1, 1};
  if (make_tmap_bf16(&tm_w, wp, 4, wd, wst, wb)) return 1;
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_tcn_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcnCfg<C>::kSmem));
  dim3 grid((p.T + kOutFrames - 1) / kOutFrames, N);
  k_tcn_tc<C><<<grid, kTcnThreads, TcnCfg<C>::kSmem, st>>>(tm_u, tm_w, p);
  return 0;
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
  p.NT = 2;
  if (p.FT < 1) return fail("gcn tensor-core kernel: %d joints do not fit a 128-row tile", V);
  const int RT = p.FT * V;
  p.a_stage_bytes = ((p.NT * RT + 128 - RT) * 128 + 1023) & ~1023;
  p.xs_tx = p.NT * p.FT * V * 64 * 4;
  p.xs_alloc = 2 * ((p.NT * p.FT * V * 128 + 1023) & ~1023);   // two 32-channel swizzled sub-tiles
  const int fixed = kGcn2Csr + 2048 + 512 + 1024;
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
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_gcn_tc2<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  k_gcn_tc2<CO><<<grid, kGcn2Threads, smem, st>>>(tm_x, tm_w, p);
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
template <int C>
int launch_tcn_tc2_c(const __nv_bfloat16 *u, const __nv_bfloat16 *wp, TcnTc2Params p, int N, int T, int stride,
                     cudaStream_t st) {
  const int V = p.V, pad = (p.G - 1) / 2;
  const int kMaxSmem = 232448;
  // frames per tile: as many whole frames as fit 128 rows; shrink for the big stride-2 case
  int FT = 128 / V;
  int a_rows = 0;
  for (;; --FT) {
    if (FT < 1) return fail("tcn tensor-core kernel: %d joints do not fit a 128-row tile", V);
    p.FT = FT;
    p.NT = 2;
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
    const int left = kMaxSmem - 2 * p.a_stage_bytes - 2048 - 512 - 1024;
    p.b_stages = left / (C * 128);
    if (p.b_stages > 8) p.b_stages = 8;
    if (p.b_stages >= 2) break;
  }
  p.groups_per_trial = (p.T_out + p.NT * p.FT - 1) / (p.NT * p.FT);
  p.items = N * p.groups_per_trial;
  const int smem = 2 * p.a_stage_bytes + p.b_stages * C * 128 + 2048 + 512 + 1024;

  CUtensorMap tm_u0, tm_u1, tm_w;
  const uint64_t plane_stride = (uint64_t)N * T * V * C * 2;
  if (stride == 1) {
    const uint64_t ud[5] = {(uint64_t)C, (uint64_t)V, (uint64_t)T, (uint64_t)N, (uint64_t)p.planes};
    const uint64_t us[4] = {(uint64_t)C * 2, (uint64_t)V * C * 2, (uint64_t)T * V * C * 2, plane_stride};
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
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_tcn_tc2<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = p.items < num_sms() ? p.items : num_sms();
  k_tcn_tc2<C><<<grid, kTcn2Threads, smem, st>>>(tm_u0, tm_u1, tm_w, p);
  return 0;
}

inline int launch_tcn_tc2(int C, const __nv_bfloat16 *u, const __nv_bfloat16 *wp, const TcnTc2Params &p, int N,
                          int T, int stride, cudaStream_t st) {
  switch (C) {
    case 64: return launch_tcn_tc2_c<64>(u, wp, p, N, T, stride, st);
    case 128: return launch_tcn_tc2_c<128>(u, wp, p, N, T, stride, st);
    case 256: return launch_tcn_tc2_c<256>(u, wp, p, N, T, stride, st);
  }
  return fail("tcn tensor-core kernel: unsupported channel count %d", C);
}

inline int launch_tcn_tc(int C, const __nv_bfloat16 *u, const __nv_bfloat16 *wp, const TcnTcParams &p, int N,
                         cudaStream_t st) {
  switch (C) {
    case 64: return launch_tcn_tc_c<64>(u, wp, p, N, st);
    case 128: return launch_tcn_tc_c<128>(u, wp, p, N, st);
    case 256: return launch_tcn_tc_c<256>(u, wp, p, N, st);
  }
  return fail("tcn tensor-core kernel: unsupported channel count %d", C);
}

}  // namespace tc
}  // namespace stgcn
