// RT-ST-GCN continual step, state half, as a bulk-copy-staged streaming kernel (sm_100a).
//
// Per stream and layer the step moves four arrays of V*C values that nothing else touches this
// step: the oldest FIFO slot (read), the same slot (write z), the accumulator (read + write)
// (rtstgcn.py:611-625), plus z (from the GEMM, L2-resident), the residual and the output.  The first
// version of this stage (k_rt_update, kernels_simt.cuh) kept a stream's values in registers across two
// block reductions: 64-128 registers per thread, 2-4 blocks per SM, and the load -> reduce -> store
// phases of so few blocks left the memory system idle half of the time (38-53 % of the DRAM peak).
// Here persistent CTAs walk over the streams and a producer warp streams the arrays, cut into chunks of
// one float4 per compute thread, into a ring of shared-memory stages with cp.async.bulk (the TMA engine,
// completion on an mbarrier), kRtStages chunks ahead and across stream boundaries, while the compute
// threads take chunk after chunk out of shared memory: the bytes in flight are the stages (48-192 KB per
// SM), not registers, and the loads never wait for a reduction.  (A first staged version with two
// whole-stream stages per CTA ran no faster than k_rt_update: one stage in flight per CTA is too little.)
// Stores go straight from registers (fire and forget).
//
// fifo_bf16 (bf16 mode with the per-joint-weight GEMM path): the FIFO holds bf16(z) and the SAME rounded
// value is added to the fp32 accumulator now and subtracted F frames later, so the running sum cannot
// drift; state traffic per element 2 + 2 + 4 + 4 B instead of 16 B.
#pragma once
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"

namespace stgcn {

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ float4 bf16x4_to_f4(uint2 b) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&b.x));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&b.y));
  return make_float4(a.x, a.y, c.x, c.y);
}
__device__ __forceinline__ uint2 f4_to_bf16x4(float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), c = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t *>(&a), *reinterpret_cast<const uint32_t *>(&c));
}

constexpr int kRtStages = 3;     // ring of chunk stages per CTA (2-8 CTAs per SM: while one CTA reduces and
                                 // normalises a stream, the others keep the loads flowing)
constexpr int kRtMaxChunks = 4;  // chunks (of TH float4 per array) per stream

// sum over the TH compute threads (the producer warp does not take part): named barrier 1
template <int TH>
__device__ __forceinline__ float rt_block_sum(float v, float *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("bar.sync 1, %0;" ::"n"(TH) : "memory");   // protect s_red reuse
  if (lane == 0) s_red[warp] = v;
  asm volatile("bar.sync 1, %0;" ::"n"(TH) : "memory");
  float t = (lane < TH / 32) ? s_red[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// A stream's V*C values are cut into chunks of TH float4 (one per compute thread); a stage holds one chunk of
// each array: z | acc | fifo slot | residual, TH*16 B each.  The producer warp streams chunk after chunk,
// across stream boundaries, kRtStages ahead; a compute thread keeps its (at most kRtMaxChunks) updated
// accumulator values and residuals in registers until the stream's LayerNorm statistics are known.
// POOL: the instantiation that also pools over the joints (last layer).  Its 8 KB scratch is kept out of the other
// layers' kernel: with it two 512-thread CTAs no longer fit the 196 KB shared-memory configuration, and the larger
// one leaves 28 KB of L1 for 51 KB of LayerNorm tables (ncu: 145 -> 182 us per launch).
template <int TH, bool POOL>
__global__ void __launch_bounds__(TH + 32, TH == 512 ? 2 : (TH == 256 ? 4 : 6)) k_rt_stream(const RtUpdateArgs p) {
  extern __shared__ __align__(128) uint8_t rt_smem[];
  __shared__ float s_red[32];
  __shared__ float4 s_pool[POOL ? TH : 1];            // pool_out: per-thread sums over this thread's joints
  __shared__ __align__(8) unsigned long long s_bar[2 * kRtStages];
  const int tid = threadIdx.x;
  const int n = p.V * p.C, n4 = n >> 2, C4 = p.C >> 2;
  const int nch = (n4 + TH - 1) / TH;
  constexpr uint32_t kArr = TH * 16u;                 // bytes of one array chunk in a stage
  constexpr uint32_t kStage = 4u * kArr;
  const uint32_t sm0 = tc::smem_u32(rt_smem);
  const uint32_t bFull = tc::smem_u32(s_bar), bEmpty = bFull + 8 * kRtStages;
  const bool res_planes = p.res_mode == 1 && !p.res && p.res_hi;
  const bool res_rows = p.res_mode != 0 && p.res;
  const int my = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // streams of this CTA

  if (tid == 0) {
    for (int s = 0; s < kRtStages; ++s) {
      tc::mbar_init(bFull + 8 * s, 1);
      tc::mbar_init(bEmpty + 8 * s, TH / 32);
    }
    tc::fence_barrier_init();
  }
  __syncthreads();
  griddep_wait();                       // z, state and counters come from earlier kernels of the chain

  if (tid >= TH) {
    // ---- producer warp: one lane streams the chunks of this CTA's streams into the stage ring ----
    if (tid == TH) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my; ++it) {
        const int b = blockIdx.x + it * gridDim.x;
        const long long base = (long long)b * n;
        const int cnt = __ldg(p.counter + b);
        const float *ap = p.acc + (long long)(cnt % p.S) * p.slot + base;
        const long long foff = (long long)(cnt % p.F) * p.slot + base;
        for (int c = 0; c < nch; ++c) {
          const int e0 = c * TH;                                   // first float4 of the chunk
          const uint32_t len = (uint32_t)((n4 - e0 < TH) ? n4 - e0 : TH);
          const uint32_t b16 = len * 16u, b8 = len * 8u;
          const uint32_t dst = sm0 + s * kStage, bar = bFull + 8 * s;
          tc::mbar_wait(bEmpty + 8 * s, ph ^ 1u);
          tc::mbar_expect_tx(bar, 2 * b16 + (p.fifo_bf16 ? b8 : b16) +
                                      (res_rows ? b16 : (res_planes ? (p.res_lo ? b16 : b8) : 0u)));
          bulk_g2s(dst, p.z + base + 4 * e0, b16, bar);
          bulk_g2s(dst + kArr, ap + 4 * e0, b16, bar);
          if (p.fifo_bf16)
            bulk_g2s(dst + 2 * kArr, p.fifo16 + foff + 4 * e0, b8, bar);
          else
            bulk_g2s(dst + 2 * kArr, p.fifo + foff + 4 * e0, b16, bar);
          if (res_rows) {
            bulk_g2s(dst + 3 * kArr, p.res + base + 4 * e0, b16, bar);
          } else if (res_planes) {
            bulk_g2s(dst + 3 * kArr, p.res_hi + base + 4 * e0, b8, bar);
            if (p.res_lo) bulk_g2s(dst + 3 * kArr + kArr / 2, p.res_lo + base + 4 * e0, b8, bar);
          }
          if (++s == kRtStages) { s = 0; ph ^= 1u; }
        }
      }
      griddep_launch();                 // last copies issued: the next kernel may move in under this CTA's tail
    }
    return;
  }

  // ---- compute threads ----
  const bool c4_pow2 = (C4 & (C4 - 1)) == 0;
  const int c4_sh = __ffs(C4) - 1;
  const int lane = tid & 31;
  const float inv_n = 1.f / (float)n, inv_nm1 = 1.f / (float)(n - 1);
  int s = 0;
  uint32_t ph = 0;
  for (int it = 0; it < my; ++it) {
    const int b = blockIdx.x + it * gridDim.x;
    const long long base = (long long)b * n;
    const int cnt = __ldg(p.counter + b);
    float *ap = p.acc + (long long)(cnt % p.S) * p.slot + base;
    const long long foff = (long long)(cnt % p.F) * p.slot + base;
    float4 av[kRtMaxChunks], rv[kRtMaxChunks];
    float sum = 0.f, sum_r = 0.f;
    // ---- pass 1, chunk by chunk: acc <- (acc + z) + (-fifo[slot]); fifo[slot] <- z (rtstgcn.py:611-625) ----
#pragma unroll
    for (int c = 0; c < kRtMaxChunks; ++c) {
      av[c] = rv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nch) {
        const int i = c * TH + tid;
        const uint8_t *stg = rt_smem + (size_t)s * kStage;
        tc::mbar_wait(bFull + 8 * s, ph);
        if (i < n4) {
          float4 zz = reinterpret_cast<const float4 *>(stg)[tid];
          float4 a = reinterpret_cast<const float4 *>(stg + kArr)[tid], ff;
          if (p.fifo_bf16) {
            const uint2 zq = f4_to_bf16x4(zz);
            zz = bf16x4_to_f4(zq);                     // the rounded value is what enters AND later leaves the sum
            ff = bf16x4_to_f4(reinterpret_cast<const uint2 *>(stg + 2 * kArr)[tid]);
            *reinterpret_cast<uint2 *>(p.fifo16 + foff + 4 * i) = zq;
          } else {
            ff = reinterpret_cast<const float4 *>(stg + 2 * kArr)[tid];
            *reinterpret_cast<float4 *>(p.fifo + foff + 4 * i) = zz;
          }
          a.x = (a.x + zz.x) + (-ff.x);
          a.y = (a.y + zz.y) + (-ff.y);
          a.z = (a.z + zz.z) + (-ff.z);
          a.w = (a.w + zz.w) + (-ff.w);
          *reinterpret_cast<float4 *>(ap + 4 * i) = a;
          av[c] = a;
          sum += (a.x + a.y) + (a.z + a.w);
          if (res_rows) {
            rv[c] = reinterpret_cast<const float4 *>(stg + 3 * kArr)[tid];
          } else if (res_planes) {
            rv[c] = bf16x4_to_f4(reinterpret_cast<const uint2 *>(stg + 3 * kArr)[tid]);
            if (p.res_lo) {
              const float4 l = bf16x4_to_f4(reinterpret_cast<const uint2 *>(stg + 3 * kArr + kArr / 2)[tid]);
              rv[c].x += l.x; rv[c].y += l.y; rv[c].z += l.z; rv[c].w += l.w;
            }
          }
          sum_r += (rv[c].x + rv[c].y) + (rv[c].z + rv[c].w);
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bEmpty + 8 * s);          // this warp is done with the stage
        if (++s == kRtStages) { s = 0; ph ^= 1u; }
      }
    }
    if (p.debug & 8192) continue;                              // (measurement build: streaming pass only)
    const float mean = rt_block_sum<TH>(sum, s_red) * inv_n;
    float mean_r = 0.f, rstd_r = 1.f;
    if (p.res_mode == 2) mean_r = rt_block_sum<TH>(sum_r, s_red) * inv_n;
    // ---- pass 2: centred second moments from registers ----
    float q = 0.f, qr = 0.f;
#pragma unroll
    for (int c = 0; c < kRtMaxChunks; ++c)
      if (c < nch && c * TH + tid < n4) {
        const float d0 = av[c].x - mean, d1 = av[c].y - mean, d2 = av[c].z - mean, d3 = av[c].w - mean;
        q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
        const float e0 = rv[c].x - mean_r, e1 = rv[c].y - mean_r, e2 = rv[c].z - mean_r, e3 = rv[c].w - mean_r;
        qr = fmaf(e0, e0, qr); qr = fmaf(e1, e1, qr); qr = fmaf(e2, e2, qr); qr = fmaf(e3, e3, qr);
      }
    const float rstd = 1.f / sqrtf(rt_block_sum<TH>(q, s_red) * inv_nm1 + p.eps);
    if (p.res_mode == 2) rstd_r = 1.f / sqrtf(rt_block_sum<TH>(qr, s_red) * inv_nm1 + p.eps);
    // ---- pass 3: out = relu( relu(LN(acc)) + res ) ----
    float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < kRtMaxChunks; ++c)
      if (c < nch && c * TH + tid < n4) {
        const int i = c * TH + tid;
        const float4 a = av[c], r = rv[c];
        const int w = c4_pow2 ? (i >> c4_sh) : (i / C4), g = i - w * C4;
        const int ti = (g * p.V + w) * 4;                      // [C/4][V][4]
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.n_wT + ti));
        const float4 o4 = __ldg(reinterpret_cast<const float4 *>(p.n_bT + ti));
        float4 v;
        v.x = fmaxf((a.x - mean) * rstd * g4.x + o4.x, 0.f);
        v.y = fmaxf((a.y - mean) * rstd * g4.y + o4.y, 0.f);
        v.z = fmaxf((a.z - mean) * rstd * g4.z + o4.z, 0.f);
        v.w = fmaxf((a.w - mean) * rstd * g4.w + o4.w, 0.f);
        if (p.res_mode == 1) {
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        } else if (p.res_mode == 2) {
          const float4 rg = __ldg(reinterpret_cast<const float4 *>(p.r_wT + ti));
          const float4 ro = __ldg(reinterpret_cast<const float4 *>(p.r_bT + ti));
          v.x += (r.x - mean_r) * rstd_r * rg.x + ro.x;
          v.y += (r.y - mean_r) * rstd_r * rg.y + ro.y;
          v.z += (r.z - mean_r) * rstd_r * rg.z + ro.z;
          v.w += (r.w - mean_r) * rstd_r * rg.w + ro.w;
        }
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        psum.x += v.x; psum.y += v.y; psum.z += v.z; psum.w += v.w;
        if (p.out) *reinterpret_cast<float4 *>(p.out + base + 4 * i) = v;
        if (p.out_hi) {
          const uint2 h = f4_to_bf16x4(v);
          *reinterpret_cast<uint2 *>(p.out_hi + base + 4 * i) = h;
          if (p.out_lo) {
            const float4 hf = bf16x4_to_f4(h);
            *reinterpret_cast<uint2 *>(p.out_lo + base + 4 * i) =
                f4_to_bf16x4(make_float4(v.x - hf.x, v.y - hf.y, v.z - hf.z, v.w - hf.w));
          }
        }
      }
    if (POOL && p.pool_out) {
      // mean over the joints (the model's AvgPool2d((1, V)), rtstgcn.py:127): TH is a multiple of C/4, so a thread's
      // float4 always belongs to channel group tid % (C/4); the TH / (C/4) partial sums of a group are added in a
      // fixed order.  The next stream's first block reduction separates these reads from the next writes.
      s_pool[tid] = psum;
      asm volatile("bar.sync 1, %0;" ::"n"(TH) : "memory");
      if (tid < C4) {
        float4 t = s_pool[tid];
        for (int k = C4; k < TH; k += C4) {
          const float4 u = s_pool[tid + k];
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        const float inv_v = 1.f / (float)p.V;
        *reinterpret_cast<float4 *>(p.pool_out + (long long)b * p.C + 4 * tid) =
            make_float4(t.x * inv_v, t.y * inv_v, t.z * inv_v, t.w * inv_v);
      }
    }
  }
}
// pooled output needs every thread to stay on one channel group across its chunks
inline bool rt_stream_pool_supported(int V, int C);

inline int rt_stream_threads(int V, int C) {
  const int n4 = V * C / 4;
  return n4 <= 128 * kRtMaxChunks ? 128 : (n4 <= 256 * kRtMaxChunks ? 256 : 512);
}
inline bool rt_stream_supported(int V, int C) { return C % 8 == 0 && V * C / 4 <= 512 * kRtMaxChunks; }
inline bool rt_stream_pool_supported(int V, int C) {
  return rt_stream_supported(V, C) && C % 4 == 0 && rt_stream_threads(V, C) % (C / 4) == 0;
}

template <int TH, bool POOL>
int launch_rt_stream_t(const RtUpdateArgs &a, cudaStream_t st) {
  const size_t smem = (size_t)kRtStages * 4 * TH * 16;
  int per_sm = (int)((size_t)225 * 1024 / (smem + 1024 + (POOL ? TH * 16 : 0)));
  const int cap = TH == 512 ? 2 : (TH == 256 ? 4 : 6);
  if (per_sm > cap) per_sm = cap;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)tc::num_sms() * per_sm;
  if (grid > a.B) grid = a.B;
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_rt_stream<TH, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (POOL)   // two CTAs of the pooling variant need the largest carve-out
    STGCN_CUDA_OK(cudaFuncSetAttribute(k_rt_stream<TH, POOL>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
  STGCN_CUDA_OK(launch_pdl(k_rt_stream<TH, POOL>, dim3((unsigned)grid), dim3(TH + 32), (size_t)smem, st, a));
  return 0;
}

inline int launch_rt_stream(const RtUpdateArgs &a, cudaStream_t st) {
  const bool pool = a.pool_out != nullptr;
  switch (rt_stream_threads(a.V, a.C)) {
    case 128: return pool ? launch_rt_stream_t<128, true>(a, st) : launch_rt_stream_t<128, false>(a, st);
    case 256: return pool ? launch_rt_stream_t<256, true>(a, st) : launch_rt_stream_t<256, false>(a, st);
  }
  return pool ? launch_rt_stream_t<512, true>(a, st) : launch_rt_stream_t<512, false>(a, st);
}

}  // namespace stgcn
