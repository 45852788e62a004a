// RT-ST-GCN continual step, state half, as a bulk-copy-staged streaming kernel (sm_100a).
//
// Per stream and layer the step moves four arrays of V*C values that nothing else touches this
// step: the oldest FIFO slot (read), the same slot (write z), the accumulator (read + write)
// (rtstgcn.py:611-625), plus z (from the GEMM, L2-resident), the residual and the output.  The first
// version of this stage (k_rt_update, kernels_simt.cuh) kept a stream's values in registers across two
// block reductions: 64-128 registers per thread, 2-4 blocks per SM, and the load -> reduce -> store
// phases of so few blocks left the memory system idle half of the time (38-53 % of the DRAM peak).
// Here persistent CTAs walk over the streams and ONE thread per CTA streams the next stream's arrays
// into a shared-memory stage with cp.async.bulk (the TMA engine, completion on an mbarrier) while the
// CTA's threads work on the current stage out of shared memory: the bytes in flight are the stages
// (2 per CTA, 26-205 KB per SM), not registers, the compute passes are short, and the loads never wait
// for a reduction.  Stores go straight from registers (fire and forget).
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

constexpr int kRtStages = 2;

// stage layout (bytes, n = V*C): z [4n] | acc [4n] | fifo slot [4n, or 2n as bf16] | residual [4n: fp32 rows, or
// hi plane 2n + lo plane 2n]
template <int TH>
__global__ void __launch_bounds__(TH, TH == 512 ? 1 : (TH == 256 ? 4 : 8)) k_rt_stream(const RtUpdateArgs p) {
  extern __shared__ __align__(128) uint8_t rt_smem[];
  __shared__ float s_red[32];
  __shared__ __align__(8) unsigned long long s_bar[kRtStages];
  const int tid = threadIdx.x;
  const int n = p.V * p.C, n4 = n >> 2, C4 = p.C >> 2;
  const uint32_t arr = (uint32_t)n * 4u;
  const uint32_t stage_bytes = 4u * arr;
  const uint32_t sm0 = tc::smem_u32(rt_smem);
  const uint32_t bar0 = tc::smem_u32(s_bar);
  const bool c4_pow2 = (C4 & (C4 - 1)) == 0;
  const int c4_sh = __ffs(C4) - 1;
  const bool res_planes = p.res_mode == 1 && !p.res && p.res_hi;
  const bool res_rows = p.res_mode != 0 && p.res;
  const int my = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // streams of this CTA

  if (tid == 0) {
    for (int s = 0; s < kRtStages; ++s) tc::mbar_init(bar0 + 8 * s, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int it) {
    const int b = blockIdx.x + it * gridDim.x;
    const int s = it % kRtStages;
    const long long base = (long long)b * n;
    const int cnt = __ldg(p.counter + b);
    const uint32_t dst = sm0 + s * stage_bytes, bar = bar0 + 8 * s;
    const uint32_t fbytes = p.fifo_bf16 ? arr / 2 : arr;
    const uint32_t rbytes = res_rows ? arr : (res_planes ? (p.res_lo ? arr : arr / 2) : 0u);
    tc::mbar_expect_tx(bar, 2 * arr + fbytes + rbytes);
    bulk_g2s(dst, p.z + base, arr, bar);
    bulk_g2s(dst + arr, p.acc + (long long)(cnt % p.S) * p.slot + base, arr, bar);
    if (p.fifo_bf16)
      bulk_g2s(dst + 2 * arr, p.fifo16 + (long long)(cnt % p.F) * p.slot + base, arr / 2, bar);
    else
      bulk_g2s(dst + 2 * arr, p.fifo + (long long)(cnt % p.F) * p.slot + base, arr, bar);
    if (res_rows) {
      bulk_g2s(dst + 3 * arr, p.res + base, arr, bar);
    } else if (res_planes) {
      bulk_g2s(dst + 3 * arr, p.res_hi + base, arr / 2, bar);
      if (p.res_lo) bulk_g2s(dst + 3 * arr + arr / 2, p.res_lo + base, arr / 2, bar);
    }
  };
  if (tid == 0)
    for (int it = 0; it < kRtStages && it < my; ++it) issue(it);

  const float inv_n = 1.f / (float)n, inv_nm1 = 1.f / (float)(n - 1);
  for (int it = 0; it < my; ++it) {
    const int s = it % kRtStages;
    const uint32_t ph = (uint32_t)(it / kRtStages) & 1u;
    const int b = blockIdx.x + it * gridDim.x;
    const long long base = (long long)b * n;
    const int cnt = __ldg(p.counter + b);
    float *ap = p.acc + (long long)(cnt % p.S) * p.slot + base;
    float *fp = p.fifo_bf16 ? nullptr : p.fifo + (long long)(cnt % p.F) * p.slot + base;
    __nv_bfloat16 *fp16 = p.fifo_bf16 ? p.fifo16 + (long long)(cnt % p.F) * p.slot + base : nullptr;
    uint8_t *stg = rt_smem + (size_t)s * stage_bytes;
    const float4 *sz = reinterpret_cast<const float4 *>(stg);
    float4 *sa = reinterpret_cast<float4 *>(stg + arr);
    const float4 *sf = reinterpret_cast<const float4 *>(stg + 2 * arr);
    const uint2 *sf16 = reinterpret_cast<const uint2 *>(stg + 2 * arr);
    const float4 *sr = reinterpret_cast<const float4 *>(stg + 3 * arr);
    const uint2 *srh = reinterpret_cast<const uint2 *>(stg + 3 * arr);
    const uint2 *srl = reinterpret_cast<const uint2 *>(stg + 3 * arr + arr / 2);
    tc::mbar_wait(bar0 + 8 * s, ph);

    // ---- pass 1: acc <- (acc + z) + (-fifo[slot]); fifo[slot] <- z (reference order, rtstgcn.py:611-625) ----
    float sum = 0.f, sum_r = 0.f;
    for (int i = tid; i < n4; i += TH) {
      float4 zz = sz[i], a = sa[i], ff;
      if (p.fifo_bf16) {
        const uint2 zq = f4_to_bf16x4(zz);
        zz = bf16x4_to_f4(zq);                       // the rounded value is what enters AND later leaves the sum
        ff = bf16x4_to_f4(sf16[i]);
        *reinterpret_cast<uint2 *>(fp16 + 4 * i) = zq;
      } else {
        ff = sf[i];
        *reinterpret_cast<float4 *>(fp + 4 * i) = zz;
      }
      a.x = (a.x + zz.x) + (-ff.x);
      a.y = (a.y + zz.y) + (-ff.y);
      a.z = (a.z + zz.z) + (-ff.z);
      a.w = (a.w + zz.w) + (-ff.w);
      sa[i] = a;
      *reinterpret_cast<float4 *>(ap + 4 * i) = a;
      sum += (a.x + a.y) + (a.z + a.w);
      if (p.res_mode == 2) {
        const float4 r = sr[i];
        sum_r += (r.x + r.y) + (r.z + r.w);
      }
    }
    const float mean = block_sum(sum, s_red) * inv_n;
    float mean_r = 0.f, rstd_r = 1.f;
    if (p.res_mode == 2) mean_r = block_sum(sum_r, s_red) * inv_n;
    // ---- pass 2: centred second moments (each thread re-reads the elements it wrote) ----
    float q = 0.f, qr = 0.f;
    for (int i = tid; i < n4; i += TH) {
      const float4 a = sa[i];
      const float d0 = a.x - mean, d1 = a.y - mean, d2 = a.z - mean, d3 = a.w - mean;
      q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
      if (p.res_mode == 2) {
        const float4 r = sr[i];
        const float e0 = r.x - mean_r, e1 = r.y - mean_r, e2 = r.z - mean_r, e3 = r.w - mean_r;
        qr = fmaf(e0, e0, qr); qr = fmaf(e1, e1, qr); qr = fmaf(e2, e2, qr); qr = fmaf(e3, e3, qr);
      }
    }
    const float rstd = 1.f / sqrtf(block_sum(q, s_red) * inv_nm1 + p.eps);
    if (p.res_mode == 2) rstd_r = 1.f / sqrtf(block_sum(qr, s_red) * inv_nm1 + p.eps);
    // ---- pass 3: out = relu( relu(LN(acc)) + res ) ----
    for (int i = tid; i < n4; i += TH) {
      const float4 a = sa[i];
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
        float4 r;
        if (res_rows) {
          r = sr[i];
        } else {
          r = bf16x4_to_f4(srh[i]);
          if (p.res_lo) {
            const float4 l = bf16x4_to_f4(srl[i]);
            r.x += l.x; r.y += l.y; r.z += l.z; r.w += l.w;
          }
        }
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      } else if (p.res_mode == 2) {
        const float4 r = sr[i];
        const float4 rg = __ldg(reinterpret_cast<const float4 *>(p.r_wT + ti));
        const float4 ro = __ldg(reinterpret_cast<const float4 *>(p.r_bT + ti));
        v.x += (r.x - mean_r) * rstd_r * rg.x + ro.x;
        v.y += (r.y - mean_r) * rstd_r * rg.y + ro.y;
        v.z += (r.z - mean_r) * rstd_r * rg.z + ro.z;
        v.w += (r.w - mean_r) * rstd_r * rg.w + ro.w;
      }
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
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
    // the stage is free once every thread has finished reading it; the refill is an async-proxy write
    // after generic-proxy accesses of the same bytes -> proxy fence
    __syncthreads();
    if (tid == 0 && it + kRtStages < my) {
      tc::fence_proxy_async();
      issue(it + kRtStages);
    }
  }
}

inline bool rt_stream_supported(int V, int C) { return C % 8 == 0 && (size_t)kRtStages * 16 * V * C <= 220 * 1024; }

template <int TH>
int launch_rt_stream_t(const RtUpdateArgs &a, cudaStream_t st) {
  const size_t smem = (size_t)kRtStages * 16 * a.V * a.C;
  int per_sm = (int)((size_t)225 * 1024 / (smem + 1024));
  if (per_sm > 2048 / TH) per_sm = 2048 / TH;
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)tc::num_sms() * per_sm;
  if (grid > a.B) grid = a.B;
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_rt_stream<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_rt_stream<TH><<<(unsigned)grid, TH, smem, st>>>(a);
  return 0;
}

inline int launch_rt_stream(const RtUpdateArgs &a, cudaStream_t st) {
  const int n4 = a.V * a.C / 4;
  if (n4 <= 192) return launch_rt_stream_t<128>(a, st);
  if (n4 <= 1024) return launch_rt_stream_t<256>(a, st);
  return launch_rt_stream_t<512>(a, st);
}

}  // namespace stgcn
