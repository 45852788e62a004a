// fp32 CUDA-core kernels of the ST-GCN / RT-ST-GCN forward path (sm_100a).
//
// Internal activation layout is channels-last: rows r = (n*T + t)*V + v, each row
// holding C contiguous fp32 channels ("NTVC").  The reference layout (N,C,T,V) is
// converted only at API entry/exit.
//
// These kernels are the STGCN_MATH_FP32 arithmetic (exact reference arithmetic,
// also the path for shapes the tcgen05 kernels do not cover: C % 64 != 0, tiny
// batches).  The tensor-core kernels live in kernels_tc.cuh.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace stgcn {

// --------------------------------------------------------------------------- //
// layout conversion  x[n][c][p] <-> y[n][p][c]   (p = t*V + v)
// --------------------------------------------------------------------------- //
// The channels-last side may be padded to `ld >= C` channels (zero filled / ignored).
__global__ void k_cp_to_pc(const float *__restrict__ x, float *__restrict__ y, int C, long long P, int ld) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float *xs = x + n * C * P;
  float *ys = y + n * ld * P;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < P) ? xs[(long long)c * P + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i;
    int c = c0 + threadIdx.x;
    if (c < ld && p < P) ys[p * ld + c] = tile[threadIdx.x][i];
  }
}

__global__ void k_pc_to_cp(const float *__restrict__ x, float *__restrict__ y, int C, long long P, int ld) {
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float *xs = x + n * ld * P;
  float *ys = y + n * C * P;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i;
    int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < P) ? xs[p * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long p = p0 + threadIdx.x;
    if (c < C && p < P) ys[(long long)c * P + p] = tile[threadIdx.x][i];
  }
}

// --------------------------------------------------------------------------- //
// weight repack: temporal conv weight (c_out, c_in, G) -> (c_out, G, c_in) so the
// implicit-GEMM K index is (tap, channel) with channels contiguous.
// --------------------------------------------------------------------------- //
__global__ void k_pack_tcn_w(const float *__restrict__ w, float *__restrict__ wp, int c_out, int c_in,
                             int G, int c_out_p, int c_in_p) {
  // destination (c_out_p, G, c_in_p), zero padded
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)c_out_p * c_in_p * G;
  if (i >= total) return;
  int ci = i % c_in_p;
  long long r = i / c_in_p;
  int j = r % G;
  int co = r / G;
  wp[i] = (co < c_out && ci < c_in) ? w[((long long)co * c_in + ci) * G + j] : 0.f;
}

__global__ void k_pad_vec(const float *__restrict__ src, float *__restrict__ dst, int n, int n_p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_p) dst[i] = (i < n) ? src[i] : 0.f;
}

// --------------------------------------------------------------------------- //
// implicit GEMM over time taps:
//   Y[m, co] = bias[co] + sum_{j<G} sum_{c<C_in} X[src(m,j), c] * Wp[co, j*C_in + c]
//   m = (n, tau, w),  src = (n, stride*tau + j - pad, w), zero outside [0, T_in)
// G = 1 gives the 1x1 convs (gcn feature transform, strided residual conv).
// --------------------------------------------------------------------------- //
struct ConvGeom {
  int M;  // N * T_out * V
  int T_in, T_out, V, C_in, C_out, G, stride, pad, Ktot;
};

template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
    k_gemm_conv(const float *__restrict__ X, const float *__restrict__ Wp, const float *__restrict__ bias,
                float *__restrict__ Y, ConvGeom g) {
  static_assert(BM == 128 && BN == 64 && BK == 16, "tile shape is tied to the thread mapping");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // A-load assignment: two float4 per thread (row = idx>>2, kq = idx&3)
  int a_row[2], a_n[2], a_tau[2], a_w[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = tid + i * 256;
    a_row[i] = idx >> 2;
    int m = m0 + a_row[i];
    a_ok[i] = m < g.M;
    int mm = a_ok[i] ? m : 0;
    a_w[i] = mm % g.V;
    int tmp = mm / g.V;
    a_tau[i] = tmp % g.T_out;
    a_n[i] = tmp / g.T_out;
  }
  const int kq = tid & 3;
  const int b_col = tid >> 2;
  const bool b_ok = (n0 + b_col) < g.C_out;

  float4 ra[2], rb;
  auto load_tile = [&](int k0) {
    int kk = k0 + kq * 4;
    int j = kk / g.C_in;
    int c = kk - j * g.C_in;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int ts = g.stride * a_tau[i] + j - g.pad;
      bool ok = a_ok[i] && kk < g.Ktot && ts >= 0 && ts < g.T_in;
      if (ok) {
        long long src = (((long long)a_n[i] * g.T_in + ts) * g.V + a_w[i]) * g.C_in + c;
        ra[i] = *reinterpret_cast<const float4 *>(X + src);
      } else {
        ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (b_ok && kk < g.Ktot)
      rb = *reinterpret_cast<const float4 *>(Wp + (long long)(n0 + b_col) * g.Ktot + kk);
    else
      rb = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][kq * 4 + 0][a_row[i]] = ra[i].x;
      As[buf][kq * 4 + 1][a_row[i]] = ra[i].y;
      As[buf][kq * 4 + 2][a_row[i]] = ra[i].z;
      As[buf][kq * 4 + 3][a_row[i]] = ra[i].w;
    }
    Bs[buf][kq * 4 + 0][b_col] = rb.x;
    Bs[buf][kq * 4 + 1][b_col] = rb.y;
    Bs[buf][kq * 4 + 2][b_col] = rb.z;
    Bs[buf][kq * 4 + 3][b_col] = rb.w;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (g.Ktot + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
      float4 b = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  const int co = n0 + tx * 4;
  if (co < g.C_out) {
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = *reinterpret_cast<const float4 *>(bias + co);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + ty * 8 + i;
      if (m < g.M) {
        float4 o = make_float4(acc[i][0] + bv.x, acc[i][1] + bv.y, acc[i][2] + bv.z, acc[i][3] + bv.w);
        *reinterpret_cast<float4 *>(Y + (long long)m * g.C_out + co) = o;
      }
    }
  }
}

// --------------------------------------------------------------------------- //
// adjacency in CSR-by-output-joint form.  A[k][v][w] (already A*importance):
//   ptr[w] .. ptr[w+1]  ->  (yoff = v*K*C + k*C, val) in (k,v) order.
// Dense A (e.g. AA-GCN's A+B+C) simply yields K*V entries per joint.
// One block per sample (blockIdx.x) when A is per-sample.
// --------------------------------------------------------------------------- //
__global__ void k_build_adj_csr(const float *__restrict__ A, int K, int V, int C, int *__restrict__ ptr,
                                int *__restrict__ yoff, float *__restrict__ val) {
  extern __shared__ int s_cnt[];  // V + 1
  const int n = blockIdx.x;
  const float *a = A + (long long)n * K * V * V;
  int *p = ptr + (long long)n * (V + 1);
  int *yo = yoff + (long long)n * K * V * V;
  float *va = val + (long long)n * K * V * V;
  for (int w = threadIdx.x; w < V; w += blockDim.x) {
    int c = 0;
    for (int kv = 0; kv < K * V; ++kv) c += (a[(long long)kv * V + w] != 0.f);
    s_cnt[w] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int w = 0; w < V; ++w) {
      int c = s_cnt[w];
      s_cnt[w] = run;
      run += c;
    }
    s_cnt[V] = run;
  }
  __syncthreads();
  for (int w = threadIdx.x; w <= V; w += blockDim.x) p[w] = s_cnt[w];
  for (int w = threadIdx.x; w < V; w += blockDim.x) {
    int at = s_cnt[w];
    for (int kv = 0; kv < K * V; ++kv) {
      float x = a[(long long)kv * V + w];
      if (x != 0.f) {
        int k = kv / V, v = kv - k * V;
        yo[at] = v * K * C + k * C;
        va[at] = x;
        ++at;
      }
    }
  }
}

// --------------------------------------------------------------------------- //
// block reduction helper (sum), all threads get the result
// --------------------------------------------------------------------------- //
__device__ __forceinline__ float block_sum(float v, float *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // protect s_red reuse
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (lane < nw) ? s_red[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// --------------------------------------------------------------------------- //
// per-frame fused kernel.  One frame = V rows x C channels (all statistics of the
// reference LayerNorm live inside one frame).
//   producer: a = load | adjacency(y) | RT aggregate (adjacency + FIFO/accumulator)
//   a' = LN(a) (optional), relu_mid (optional)
//   out = relu_out( a' + [ b | LN(b) ] )
// --------------------------------------------------------------------------- //
enum { FRAME_LOAD = 0, FRAME_ADJ = 1, FRAME_RT = 2 };
enum { B_NONE = 0, B_RAW = 1, B_LN = 2 };

struct FrameArgs {
  int producer;
  long long frames;       // N*T (RT: B)
  int frames_per_sample;  // T, to pick the per-sample adjacency
  int K, V, C;
  const float *a;  // FRAME_LOAD: [frames*V, C]
  const float *y;  // FRAME_ADJ/RT: [frames*V, K*C]
  const int *adj_ptr;
  const int *adj_yoff;
  const float *adj_val;
  int adj_per_sample;
  int norm_a;  // 1: LayerNorm(C,V) on a
  const float *na_w, *na_b;
  int relu_mid;
  int b_mode;
  const float *b;
  const float *nb_w, *nb_b;
  int relu_out;
  float eps;
  float *out;
  // optional split-bf16 output (hi = bf16(v), lo = bf16(v - hi)) feeding the tcgen05 kernels;
  // when out_hi is set, `out` is not written.  out_lo may be null (single-plane bf16 mode).
  __nv_bfloat16 *out_hi, *out_lo;
  // RT state
  float *fifo, *acc;
  const int *counter;
  int F, S;
};

__global__ void __launch_bounds__(256) k_frame(FrameArgs p) {
  extern __shared__ __align__(16) float s_buf[];
  __shared__ float s_red[32];
  const int VC = p.V * p.C;
  float *za = s_buf;
  float *zb = s_buf + VC;
  const float inv_n = 1.f / (float)VC;
  const float inv_nm1 = 1.f / (float)(VC - 1);

  for (long long f = blockIdx.x; f < p.frames; f += gridDim.x) {
    const long long row0 = f * p.V;
    // ---- producer ----
    if (p.producer == FRAME_LOAD) {
      const float *src = p.a + row0 * p.C;
      for (int i = threadIdx.x; i < VC; i += blockDim.x) za[i] = src[i];
    } else {
      const long long smp = p.adj_per_sample ? (f / p.frames_per_sample) : 0;
      const int *ptr = p.adj_ptr + smp * (p.V + 1);
      const int *yo = p.adj_yoff + smp * (long long)p.K * p.V * p.V;
      const float *va = p.adj_val + smp * (long long)p.K * p.V * p.V;
      const float *ysrc = p.y + row0 * (long long)p.K * p.C;
      int fi = 0, ai = 0;
      if (p.producer == FRAME_RT) {
        int t = p.counter[f];
        fi = t % p.F;
        ai = t % p.S;
      }
      for (int i = threadIdx.x; i < VC; i += blockDim.x) {
        const int w = i / p.C, c = i - w * p.C;
        float z = 0.f;
        const int e1 = ptr[w + 1];
        for (int e = ptr[w]; e < e1; ++e) z = fmaf(ysrc[yo[e] + c], va[e], z);
        if (p.producer == FRAME_RT) {
          // acc <- (acc + z) + (-fifo[fi]);  fifo[fi] <- z   (rtstgcn.py:611-621)
          const long long so = (f * p.V + w) * (long long)p.C + c;
          const long long per = p.frames * (long long)VC;
          float *fp = p.fifo + (long long)fi * per + so;
          float *ap = p.acc + (long long)ai * per + so;
          float a = *ap + z;
          a = a + (-*fp);
          *ap = a;
          *fp = z;
          z = a;
        }
        za[i] = z;
      }
    }
    if (p.b_mode == B_LN) {
      const float *src = p.b + row0 * p.C;
      for (int i = threadIdx.x; i < VC; i += blockDim.x) zb[i] = src[i];
    }
    __syncthreads();

    // ---- statistics ----
    float mean_a = 0.f, rstd_a = 1.f, mean_b = 0.f, rstd_b = 1.f;
    if (p.norm_a) {
      float s = 0.f;
      for (int i = threadIdx.x; i < VC; i += blockDim.x) s += za[i];
      mean_a = block_sum(s, s_red) * inv_n;
      float q = 0.f;
      for (int i = threadIdx.x; i < VC; i += blockDim.x) {
        float d = za[i] - mean_a;
        q = fmaf(d, d, q);
      }
      rstd_a = 1.f / sqrtf(block_sum(q, s_red) * inv_nm1 + p.eps);
    }
    if (p.b_mode == B_LN) {
      float s = 0.f;
      for (int i = threadIdx.x; i < VC; i += blockDim.x) s += zb[i];
      mean_b = block_sum(s, s_red) * inv_n;
      float q = 0.f;
      for (int i = threadIdx.x; i < VC; i += blockDim.x) {
        float d = zb[i] - mean_b;
        q = fmaf(d, d, q);
      }
      rstd_b = 1.f / sqrtf(block_sum(q, s_red) * inv_nm1 + p.eps);
    }

    // ---- epilogue ----
    float *dst = p.out + row0 * p.C;
    for (int i = threadIdx.x; i < VC; i += blockDim.x) {
      const int w = i / p.C, c = i - w * p.C;
      const int pi = c * p.V + w;  // reference affine layout (C, 1, V)
      float v = za[i];
      if (p.norm_a) v = (v - mean_a) * rstd_a * p.na_w[pi] + p.na_b[pi];
      if (p.relu_mid) v = fmaxf(v, 0.f);
      if (p.b_mode == B_RAW) {
        v += p.b[row0 * p.C + i];
      } else if (p.b_mode == B_LN) {
        v += (zb[i] - mean_b) * rstd_b * p.nb_w[pi] + p.nb_b[pi];
      }
      if (p.relu_out) v = fmaxf(v, 0.f);
      if (p.out_hi) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        p.out_hi[row0 * p.C + i] = hi;
        if (p.out_lo) p.out_lo[row0 * p.C + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
      } else {
        dst[i] = v;
      }
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------------- //
// stand-alone module kernels working directly on the reference layout (N,C,T,V)
// --------------------------------------------------------------------------- //
// LayerNorm over (C,V) per (n,t), unbiased variance (layernorm.py:22-28).
__global__ void __launch_bounds__(256)
    k_layernorm_nctv(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                     float *__restrict__ y, int N, int C, int T, int V, float eps) {
  extern __shared__ __align__(16) float s_buf[];
  __shared__ float s_red[32];
  const int CV = C * V;
  const float inv_n = 1.f / (float)CV, inv_nm1 = 1.f / (float)(CV - 1);
  const long long frames = (long long)N * T;
  for (long long f = blockIdx.x; f < frames; f += gridDim.x) {
    const long long n = f / T, t = f - n * T;
    for (int i = threadIdx.x; i < CV; i += blockDim.x) {
      int c = i / V, v = i - c * V;
      s_buf[i] = x[((n * C + c) * T + t) * V + v];
    }
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < CV; i += blockDim.x) s += s_buf[i];
    const float mean = block_sum(s, s_red) * inv_n;
    float q = 0.f;
    for (int i = threadIdx.x; i < CV; i += blockDim.x) {
      float d = s_buf[i] - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.f / sqrtf(block_sum(q, s_red) * inv_nm1 + eps);
    for (int i = threadIdx.x; i < CV; i += blockDim.x) {
      int c = i / V, v = i - c * V;
      y[((n * C + c) * T + t) * V + v] = (s_buf[i] - mean) * rstd * w[i] + b[i];
    }
    __syncthreads();
  }
}

// Batch statistics on (N,C,T,V).  mode 0: per channel; mode 1: per (v,c) feature v*C+c.
__global__ void __launch_bounds__(256)
    k_bn_stats_nctv(const float *__restrict__ x, int C, int T, int V, int mode, int chunk,
                    double *__restrict__ sum, double *__restrict__ sumsq) {
  extern __shared__ double s_d[];  // mode 1: 2*V
  __shared__ double s_w[2][8];
  const int c = blockIdx.y;
  const long long n = blockIdx.z;
  const long long plane = (long long)T * V;
  const float *xs = x + (n * C + c) * plane;
  long long i0 = (long long)blockIdx.x * chunk, i1 = i0 + chunk;
  if (i1 > plane) i1 = plane;
  if (mode == 1) {
    for (int v = threadIdx.x; v < 2 * V; v += blockDim.x) s_d[v] = 0.0;
    __syncthreads();
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      float val = xs[i];
      int v = (int)(i % V);
      atomicAdd(&s_d[v], (double)val);
      atomicAdd(&s_d[V + v], (double)val * val);
    }
    __syncthreads();
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      atomicAdd(&sum[v * C + c], s_d[v]);
      atomicAdd(&sumsq[v * C + c], s_d[V + v]);
    }
  } else {
    double s = 0.0, q = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      float val = xs[i];
      s += val;
      q += (double)val * val;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_w[0][threadIdx.x >> 5] = s;
      s_w[1][threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double ts = 0, tq = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
        ts += s_w[0][k];
        tq += s_w[1][k];
      }
      atomicAdd(&sum[c], ts);
      atomicAdd(&sumsq[c], tq);
    }
  }
}

__global__ void __launch_bounds__(256)
    k_bn_apply_nctv(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ b,
                    const double *__restrict__ sum, const double *__restrict__ sumsq, float *__restrict__ y,
                    int N, int C, int T, int V, int mode, double inv_count, float eps) {
  const long long plane = (long long)T * V;
  const long long total = (long long)N * C * plane;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % V);
    const int c = (int)((i / plane) % C);
    const int f = mode == 1 ? v * C + c : c;
    double mean = sum[f] * inv_count;
    double var = sumsq[f] * inv_count - mean * mean;
    if (var < 0) var = 0;
    y[i] = (float)(((double)x[i] - mean) / sqrt(var + (double)eps)) * w[f] + b[f];
  }
}

// --------------------------------------------------------------------------- //
// RT-ST-GCN continual step, state half (rtstgcn.py:611-625, 548-553), as a streaming kernel:
// one block per stream; z is the graph-convolved frame (bias included) written by the tensor-core
// GEMM kernel and still L2-resident.  Per element, in the reference's order:
//   acc <- (acc + z) + (-fifo[slot]);  fifo[slot] <- z;  o = acc
//   out = relu( relu(LN_{C,V}(o)) + res ),  res = x | LN_R(q) | nothing
// Everything is float4 and row-contiguous, every thread issues all its loads before the first use,
// and the 6.9 GB of state move at HBM speed instead of through a latency-bound GEMM epilogue.
// --------------------------------------------------------------------------- //
struct RtUpdateArgs {
  int B, V, C;
  const float *z;                 // [B*V, C]
  float *fifo, *acc;              // [F][B*V, C], [S][B*V, C]
  __nv_bfloat16 *fifo16;          // fifo_bf16: the FIFO as bf16 [F][B*V, C] (k_rt_stream only; fifo == null)
  int fifo_bf16;
  const int *counter;             // [B]
  int F, S;
  long long slot;                 // B*V*C
  const float *n_wT, *n_bT;       // LayerNorm affine as [C/4][V][4]
  int res_mode;                   // 0 none, 1 raw rows, 2 LayerNorm_R(rows)
  const float *res;               // [B*V, C]
  const __nv_bfloat16 *res_hi, *res_lo;   // res_mode 1 with the rows as bf16 hi/lo planes (res == null)
  __nv_bfloat16 *out_hi, *out_lo; // also / instead write the output as bf16 planes (next layer's GEMM operand)
  const float *r_wT, *r_bT;       // affine of the residual norm, [C/4][V][4]
  float eps;
  float *out;                     // [B*V, C]
  float *pool_out;                // k_rt_stream only: instead of `out`, the mean over the joints [B, C] (last layer)
  int debug;                      // measurement build only (STGCN_DEBUG bits)
};

__device__ __forceinline__ float4 ld_nc_stream4(const float *p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ uint2 ld_nc_stream2(const __nv_bfloat16 *p) {
  uint2 v;
  asm volatile("ld.global.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

template <int NV>
__global__ void __launch_bounds__(256, 2) k_rt_update(RtUpdateArgs p) {
  __shared__ float s_red[32];
  const int b = blockIdx.x;
  const int VC4 = (p.V * p.C) >> 2, C4 = p.C >> 2;
  const bool c4_pow2 = (C4 & (C4 - 1)) == 0;
  const int c4_sh = __ffs(C4) - 1;
  const int cnt = __ldg(p.counter + b);
  const long long base = (long long)b * p.V * p.C;
  float *fp = p.fifo + (long long)(cnt % p.F) * p.slot + base;
  float *ap = p.acc + (long long)(cnt % p.S) * p.slot + base;
  const float *zp = p.z + base;
  const float *rp = p.res ? p.res + base : nullptr;
  float4 a[NV], r[NV], zz[NV], ff[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + 256 * j;
    a[j] = zz[j] = ff[j] = r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < VC4) {
      zz[j] = ld_nc_stream4(zp + 4 * i);
      ff[j] = ld_nc_stream4(fp + 4 * i);
      a[j] = ld_nc_stream4(ap + 4 * i);
      if (rp) {
        r[j] = ld_nc_stream4(rp + 4 * i);
      } else if (p.res_mode == 1 && p.res_hi) {
        // raw bits only: converting here would make every iteration wait for its own load and
        // serialise the loads of the following iterations
        const uint2 rh = ld_nc_stream2(p.res_hi + base + 4 * i);
        const uint2 rl = p.res_lo ? ld_nc_stream2(p.res_lo + base + 4 * i) : make_uint2(0u, 0u);
        r[j] = make_float4(__uint_as_float(rh.x), __uint_as_float(rh.y), __uint_as_float(rl.x), __uint_as_float(rl.y));
      }
    }
  }
  float s = 0.f, sr = 0.f;
  const bool res_planes = p.res_mode == 1 && !rp && p.res_hi;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < VC4) {
      if (res_planes) {
        const uint32_t h0 = __float_as_uint(r[j].x), h1 = __float_as_uint(r[j].y);
        const uint32_t l0 = __float_as_uint(r[j].z), l1 = __float_as_uint(r[j].w);
        const float2 a01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&h0));
        const float2 a23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&h1));
        const float2 b01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&l0));
        const float2 b23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&l1));
        r[j] = make_float4(a01.x + b01.x, a01.y + b01.y, a23.x + b23.x, a23.y + b23.y);
      }
      a[j].x = (a[j].x + zz[j].x) + (-ff[j].x);
      a[j].y = (a[j].y + zz[j].y) + (-ff[j].y);
      a[j].z = (a[j].z + zz[j].z) + (-ff[j].z);
      a[j].w = (a[j].w + zz[j].w) + (-ff[j].w);
      *reinterpret_cast<float4 *>(fp + 4 * i) = zz[j];
      *reinterpret_cast<float4 *>(ap + 4 * i) = a[j];
      s += (a[j].x + a[j].y) + (a[j].z + a[j].w);
      sr += (r[j].x + r[j].y) + (r[j].z + r[j].w);
    }
  }
  const float inv_n = 1.f / (float)(p.V * p.C), inv_nm1 = 1.f / (float)(p.V * p.C - 1);
  const float mean = block_sum(s, s_red) * inv_n;
  float mean_r = 0.f, rstd_r = 1.f;
  if (p.res_mode == 2) mean_r = block_sum(sr, s_red) * inv_n;
  float q = 0.f, qr = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < VC4) {
      const float d0 = a[j].x - mean, d1 = a[j].y - mean, d2 = a[j].z - mean, d3 = a[j].w - mean;
      q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
      const float e0 = r[j].x - mean_r, e1 = r[j].y - mean_r, e2 = r[j].z - mean_r, e3 = r[j].w - mean_r;
      qr = fmaf(e0, e0, qr); qr = fmaf(e1, e1, qr); qr = fmaf(e2, e2, qr); qr = fmaf(e3, e3, qr);
    }
  }
  const float rstd = 1.f / sqrtf(block_sum(q, s_red) * inv_nm1 + p.eps);
  if (p.res_mode == 2) rstd_r = 1.f / sqrtf(block_sum(qr, s_red) * inv_nm1 + p.eps);
  float *op = p.out ? p.out + base : nullptr;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < VC4) {
      const int w = c4_pow2 ? (i >> c4_sh) : (i / C4), g = i - w * C4;   // no integer division for C = 64/128/256                 // joint, channel group
      const int ti = (g * p.V + w) * 4;                     // [C/4][V][4]
      const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.n_wT + ti));
      const float4 o4 = __ldg(reinterpret_cast<const float4 *>(p.n_bT + ti));
      float4 v;
      v.x = fmaxf((a[j].x - mean) * rstd * g4.x + o4.x, 0.f);
      v.y = fmaxf((a[j].y - mean) * rstd * g4.y + o4.y, 0.f);
      v.z = fmaxf((a[j].z - mean) * rstd * g4.z + o4.z, 0.f);
      v.w = fmaxf((a[j].w - mean) * rstd * g4.w + o4.w, 0.f);
      if (p.res_mode == 1) {
        v.x += r[j].x; v.y += r[j].y; v.z += r[j].z; v.w += r[j].w;
      } else if (p.res_mode == 2) {
        const float4 rg = __ldg(reinterpret_cast<const float4 *>(p.r_wT + ti));
        const float4 ro = __ldg(reinterpret_cast<const float4 *>(p.r_bT + ti));
        v.x += (r[j].x - mean_r) * rstd_r * rg.x + ro.x;
        v.y += (r[j].y - mean_r) * rstd_r * rg.y + ro.y;
        v.z += (r[j].z - mean_r) * rstd_r * rg.z + ro.z;
        v.w += (r[j].w - mean_r) * rstd_r * rg.w + ro.w;
      }
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      if (p.out) *reinterpret_cast<float4 *>(op + 4 * i) = v;
      if (p.out_hi) {
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2 *>(p.out_hi + base + 4 * i) =
            make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
        if (p.out_lo) {
          const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
          const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
          *reinterpret_cast<uint2 *>(p.out_lo + base + 4 * i) =
              make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
        }
      }
    }
  }
}

inline int launch_rt_update(const RtUpdateArgs &a, cudaStream_t st) {
  const int nv = (a.V * a.C / 4 + 255) / 256;
  if (nv <= 2) k_rt_update<2><<<a.B, 256, 0, st>>>(a);
  else if (nv <= 4) k_rt_update<4><<<a.B, 256, 0, st>>>(a);
  else if (nv <= 7) k_rt_update<7><<<a.B, 256, 0, st>>>(a);
  else return fail("rt update: V*C = %d too large", a.V * a.C);
  return 0;
}
inline bool rt_update_supported(int V, int C) { return C % 4 == 0 && (V * C / 4 + 255) / 256 <= 7; }

// --------------------------------------------------------------------------- //
// RT head: mean over the V joints of every stream + fcn_out (rtstgcn.py:153-157), kRtHeadStreams streams
// per block so that the classifier matrix is read once per block instead of once per stream (the
// two-kernel pool + fc cost 57 us of a 1.5 ms step at 4096 streams, mostly re-reading W from L2).
// x [B*V, C] fp32 rows -> logits [B, classes]
// --------------------------------------------------------------------------- //
constexpr int kRtHeadStreams = 8;
// streams per block: enough blocks to fill the GPU at small batches, W reuse at large ones
inline int rt_head_streams(int B) { return B >= 2048 ? kRtHeadStreams : (B >= 512 ? 4 : (B >= 128 ? 2 : 1)); }
// top5 (optional, [B][5] int32): indices of the five largest logits of every stream, best first -- the
// reference's Statistics (utils/statistics.py:4-16: torch.topk(predictions, 5, dim=1)) computed where the
// logits are produced, so an evaluation loop needs only 20 bytes per stream-frame from the device
__global__ void __launch_bounds__(256)
    k_rt_head(const float *__restrict__ x, int B, int V, int C, const float *__restrict__ W,
              const float *__restrict__ bias, int classes, float *__restrict__ logits, int *__restrict__ top5,
              int spb) {
  extern __shared__ float s_pool[];                       // [spb][C] + [spb][classes]; spb streams per block
  float *s_log = s_pool + spb * C;
  griddep_wait();
  griddep_launch();
  const int b0 = blockIdx.x * spb;
  const int nb = B - b0 < spb ? B - b0 : spb;
  const float inv_v = 1.f / (float)V;
  // pooling: float4 lanes over the channels, 256 / (C/4) streams in parallel, all V loads of a stream in flight
  const int C4 = C >> 2;
  if ((C & 3) == 0 && C4 <= 256) {
    const int c4 = threadIdx.x % C4, sg = threadIdx.x / C4, SG = 256 / C4;
    if (sg < SG)
      for (int s = sg; s < nb; s += SG) {
        const float4 *xp = reinterpret_cast<const float4 *>(x + ((long long)(b0 + s) * V) * C) + c4;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        int v = 0;
        for (; v + 6 <= V; v += 6) {                       // six independent 16-byte loads in flight per thread
          float4 u[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) u[i] = __ldg(xp + (long long)(v + i) * C4);
#pragma unroll
          for (int i = 0; i < 6; i += 2) {
            a0.x += u[i].x; a0.y += u[i].y; a0.z += u[i].z; a0.w += u[i].w;
            a1.x += u[i + 1].x; a1.y += u[i + 1].y; a1.z += u[i + 1].z; a1.w += u[i + 1].w;
          }
        }
        for (; v + 1 < V; v += 2) {
          const float4 u0 = __ldg(xp + (long long)v * C4), u1 = __ldg(xp + (long long)(v + 1) * C4);
          a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
          a1.x += u1.x; a1.y += u1.y; a1.z += u1.z; a1.w += u1.w;
        }
        if (v < V) {
          const float4 u0 = __ldg(xp + (long long)v * C4);
          a0.x += u0.x; a0.y += u0.y; a0.z += u0.z; a0.w += u0.w;
        }
        reinterpret_cast<float4 *>(s_pool + s * C)[c4] =
            make_float4((a0.x + a1.x) * inv_v, (a0.y + a1.y) * inv_v, (a0.z + a1.z) * inv_v, (a0.w + a1.w) * inv_v);
      }
  } else {
    for (int idx = threadIdx.x; idx < nb * C; idx += 256) {
      const int s = idx / C, c = idx - s * C;
      const float *xp = x + ((long long)(b0 + s) * V) * C + c;
      float a = 0.f;
      for (int v = 0; v < V; ++v) a += __ldg(xp + (long long)v * C);
      s_pool[s * C + c] = a * inv_v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // a warp takes a class: its W row is read once (C/32 values per lane, C <= 512 in registers) and applied to
  // every stream of the block
  for (int m = warp; m < classes; m += 8) {
    float wr[16];
    const bool in_regs = C <= 512;
#pragma unroll
    for (int j = 0; j < 16; ++j) wr[j] = (in_regs && lane + 32 * j < C) ? __ldg(W + (long long)m * C + lane + 32 * j) : 0.f;
    const float bm = __ldg(bias + m);
    for (int s = 0; s < nb; ++s) {
      float a = 0.f;
      if (in_regs) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (lane + 32 * j < C) a = fmaf(wr[j], s_pool[s * C + lane + 32 * j], a);
      } else {
        for (int c = lane; c < C; c += 32) a = fmaf(__ldg(W + (long long)m * C + c), s_pool[s * C + c], a);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) {
        a += bm;
        logits[(long long)(b0 + s) * classes + m] = a;
        if (top5) s_log[s * classes + m] = a;
      }
    }
  }
  if (!top5) return;
  __syncthreads();
  // one warp per stream: five rounds of warp arg-max over the classes
  for (int s = warp; s < nb; s += 8) {
    float *row = s_log + s * classes;
    for (int r = 0; r < 5; ++r) {
      float best = -3.402823466e38f;
      int bi = 0x7fffffff;
      for (int m = lane; m < classes; m += 32) {
        const float v = row[m];
        if (v > best || (v == best && m < bi)) { best = v; bi = m; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) {
        top5[(long long)(b0 + s) * 5 + r] = r < classes ? bi : -1;
        if (bi < classes) row[bi] = -3.402823466e38f;
      }
      __syncwarp();
    }
  }
}

// indices of the five largest of `classes` logits per row (one warp per row; best first, ties to the lower index)
__global__ void __launch_bounds__(256) k_topk5(const float *__restrict__ logits, int rows, int classes, int *__restrict__ top5) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float *lp = logits + (long long)row * classes;
  int taken[5] = {-1, -1, -1, -1, -1};
  for (int r = 0; r < 5; ++r) {
    float best = -3.402823466e38f;
    int bi = 0x7fffffff;
    for (int m = lane; m < classes; m += 32) {
      bool used = false;
#pragma unroll
      for (int j = 0; j < 5; ++j) used |= (taken[j] == m);
      const float v = lp[m];
      if (!used && (v > best || (v == best && m < bi))) { best = v; bi = m; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    taken[r] = bi;
    if (lane == 0) top5[(long long)row * 5 + r] = bi < classes ? bi : -1;
  }
}

// per-stream frame counters advance modulo `period` = lcm of every layer's ring and accumulator sizes
// (the kernels only use cnt % F and cnt % S), so they never overflow on a long-running stream;
// period == 0: plain increment
__global__ void k_advance_counters(int *counter, int first, int count, int period) {
  griddep_wait();
  griddep_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    int c = counter[first + i] + 1;
    if (period > 0 && c >= period) c = 0;
    counter[first + i] = c;
  }
}

// --------------------------------------------------------------------------- //
// batch-statistics BatchNorm (two phase).  x is [rows, C]; sums are double.
// --------------------------------------------------------------------------- //
__global__ void __launch_bounds__(256)
    k_channel_stats(const float *__restrict__ x, long long rows, int C, long long rows_per_block,
                    double *__restrict__ sum, double *__restrict__ sumsq) {
  __shared__ double s_s[256], s_q[256];
  const int cpt = C < 256 ? C : 256;
  const int rsplit = 256 / cpt;
  const int c0 = threadIdx.x % cpt, rs = threadIdx.x / cpt;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  for (int cg = 0; cg < C; cg += cpt) {
    const int c = cg + c0;
    double s = 0.0, q = 0.0;
    if (rs < rsplit && c < C) {
      for (long long r = r0 + rs; r < r1; r += rsplit) {
        float v = x[r * C + c];
        s += v;
        q += (double)v * v;
      }
    }
    s_s[threadIdx.x] = s;
    s_q[threadIdx.x] = q;
    __syncthreads();
    if (rs == 0 && c < C) {
      for (int k = 1; k < rsplit; ++k) {
        s += s_s[k * cpt + c0];
        q += s_q[k * cpt + c0];
      }
      atomicAdd(&sum[c], s);
      atomicAdd(&sumsq[c], q);
    }
    __syncthreads();
  }
}

// out = relu_out( bn_a(a) [relu_mid] + [ b | bn_b(b) ] ), per-channel batch statistics.
struct BnApplyArgs {
  const float *a;
  const double *a_sum, *a_sumsq;
  const float *a_w, *a_b;
  int relu_mid;
  int b_mode;  // B_NONE / B_RAW / B_LN (here: BN of b)
  const float *b;
  const double *b_sum, *b_sumsq;
  const float *b_w, *b_b;
  int relu_out;
  long long rows;
  int C;
  double inv_count;
  float eps;
  float *out;
};

__global__ void __launch_bounds__(256) k_bn_apply(BnApplyArgs p) {
  const long long total = p.rows * p.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % p.C);
    double mean = p.a_sum[c] * p.inv_count;
    double var = p.a_sumsq[c] * p.inv_count - mean * mean;
    if (var < 0) var = 0;
    float v = (float)(((double)p.a[i] - mean) / sqrt(var + (double)p.eps)) * p.a_w[c] + p.a_b[c];
    if (p.relu_mid) v = fmaxf(v, 0.f);
    if (p.b_mode == B_RAW) {
      v += p.b[i];
    } else if (p.b_mode == B_LN) {
      double mb = p.b_sum[c] * p.inv_count;
      double vb = p.b_sumsq[c] * p.inv_count - mb * mb;
      if (vb < 0) vb = 0;
      v += (float)(((double)p.b[i] - mb) / sqrt(vb + (double)p.eps)) * p.b_w[c] + p.b_b[c];
    }
    if (p.relu_out) v = fmaxf(v, 0.f);
    p.out[i] = v;
  }
}

// Batch-statistics BatchNorm as a per-channel affine: coef[c] = (scale, shift) with scale = w / sqrt(var + eps),
// shift = b - mean * scale (moments in double), then one float4 streaming pass (tensor-core BatchNorm path).
__global__ void k_bn_coeffs(const double *__restrict__ sum, const double *__restrict__ sumsq,
                            const float *__restrict__ w, const float *__restrict__ b, double inv_count, float eps,
                            int C, float2 *__restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sum[c] * inv_count;
  double var = sumsq[c] * inv_count - mean * mean;
  if (var < 0) var = 0;
  const double sc = (double)w[c] / sqrt(var + (double)eps);
  coef[c] = make_float2((float)sc, (float)((double)b[c] - mean * sc));
}

struct BnApply4Args {
  const float *a;
  const float2 *a_coef;
  int b_mode;                 // B_NONE / B_RAW / B_LN (here: BatchNorm of b with b_coef)
  const float *b;
  const float2 *b_coef;
  int relu_out;
  long long n4;               // rows * C / 4
  int C;
  float *out;                 // fp32 rows, and / or
  __nv_bfloat16 *out_hi, *out_lo;   // bf16 hi / lo planes
};

__global__ void __launch_bounds__(256) k_bn_apply4(BnApply4Args p) {
  const int C4 = p.C >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 a = *reinterpret_cast<const float4 *>(p.a + 4 * i);
    const float4 k01 = *reinterpret_cast<const float4 *>(p.a_coef + c), k23 = *reinterpret_cast<const float4 *>(p.a_coef + c + 2);
    float4 v = make_float4(fmaf(a.x, k01.x, k01.y), fmaf(a.y, k01.z, k01.w), fmaf(a.z, k23.x, k23.y), fmaf(a.w, k23.z, k23.w));
    if (p.b_mode == B_RAW) {
      const float4 r = *reinterpret_cast<const float4 *>(p.b + 4 * i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    } else if (p.b_mode == B_LN) {
      const float4 r = *reinterpret_cast<const float4 *>(p.b + 4 * i);
      const float4 q01 = *reinterpret_cast<const float4 *>(p.b_coef + c), q23 = *reinterpret_cast<const float4 *>(p.b_coef + c + 2);
      v.x += fmaf(r.x, q01.x, q01.y); v.y += fmaf(r.y, q01.z, q01.w);
      v.z += fmaf(r.z, q23.x, q23.y); v.w += fmaf(r.w, q23.z, q23.w);
    }
    if (p.relu_out) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if (p.out) *reinterpret_cast<float4 *>(p.out + 4 * i) = v;
    if (p.out_hi) {
      const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
      *reinterpret_cast<uint2 *>(p.out_hi + 4 * i) =
          make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
      if (p.out_lo) {
        const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
        *reinterpret_cast<uint2 *>(p.out_lo + 4 * i) =
            make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
      }
    }
  }
}

// --------------------------------------------------------------------------- //
// input stage: norm_in + fcn_in (stgcn.py:82-85).  x is [frames, V, C_in] (NTVC).
//   LN mode: per-frame LayerNorm over (C_in, V), affine (C_in, V)
//   BN mode: per-feature (v*C_in + c) batch statistics, precomputed sums
// out[frame, v, co] = b[co] + sum_c W[co][c] * xn[v][c]
// --------------------------------------------------------------------------- //
struct EmbedArgs {
  const float *x;
  long long frames;
  int V, C_in, C0;
  int norm;  // STGCN_NORM_*
  const float *n_w, *n_b;
  const double *bn_sum, *bn_sumsq;
  double bn_inv_count;
  float eps;
  const float *W, *bias;
  float *out;
};

__global__ void __launch_bounds__(256) k_embed(EmbedArgs p) {
  extern __shared__ __align__(16) float s_buf[];
  __shared__ float s_red[32];
  const int VCi = p.V * p.C_in;
  float *xn = s_buf;                  // V*C_in
  float *sw = s_buf + VCi;            // C0*C_in
  float *sb = sw + p.C0 * p.C_in;     // C0
  float *sc = sb + p.C0;              // BN: scale[V*C_in], shift[V*C_in]
  for (int i = threadIdx.x; i < p.C0 * p.C_in; i += blockDim.x) sw[i] = p.W[i];
  for (int i = threadIdx.x; i < p.C0; i += blockDim.x) sb[i] = p.bias[i];
  if (p.norm == 1) {
    for (int i = threadIdx.x; i < VCi; i += blockDim.x) {
      double mean = p.bn_sum[i] * p.bn_inv_count;
      double var = p.bn_sumsq[i] * p.bn_inv_count - mean * mean;
      if (var < 0) var = 0;
      double rstd = 1.0 / sqrt(var + (double)p.eps);
      sc[i] = (float)rstd * p.n_w[i];
      sc[VCi + i] = p.n_b[i] - (float)(mean * rstd) * p.n_w[i];
    }
  }
  __syncthreads();
  const float inv_n = 1.f / (float)VCi, inv_nm1 = 1.f / (float)(VCi - 1);
  for (long long f = blockIdx.x; f < p.frames; f += gridDim.x) {
    const float *src = p.x + f * VCi;
    for (int i = threadIdx.x; i < VCi; i += blockDim.x) xn[i] = src[i];
    __syncthreads();
    if (p.norm == 0) {
      float s = 0.f;
      for (int i = threadIdx.x; i < VCi; i += blockDim.x) s += xn[i];
      float mean = block_sum(s, s_red) * inv_n;
      float q = 0.f;
      for (int i = threadIdx.x; i < VCi; i += blockDim.x) {
        float d = xn[i] - mean;
        q = fmaf(d, d, q);
      }
      float rstd = 1.f / sqrtf(block_sum(q, s_red) * inv_nm1 + p.eps);
      for (int i = threadIdx.x; i < VCi; i += blockDim.x) {
        int v = i / p.C_in, c = i - v * p.C_in;
        int pi = c * p.V + v;
        xn[i] = (xn[i] - mean) * rstd * p.n_w[pi] + p.n_b[pi];
      }
    } else {
      for (int i = threadIdx.x; i < VCi; i += blockDim.x) xn[i] = xn[i] * sc[i] + sc[VCi + i];
    }
    __syncthreads();
    float *dst = p.out + f * (long long)p.V * p.C0;
    for (int i = threadIdx.x; i < p.V * p.C0; i += blockDim.x) {
      int v = i / p.C0, co = i - v * p.C0;
      float acc = sb[co];
      for (int c = 0; c < p.C_in; ++c) acc = fmaf(sw[co * p.C_in + c], xn[v * p.C_in + c], acc);
      dst[i] = acc;
    }
    __syncthreads();
  }
}

// Input stage, LayerNorm mode, one WARP per frame, reading the reference layout (N, C_in, T, V)
// directly (no separate layout pass): V*C_in <= 128 values per frame live in registers, the
// statistics are warp shuffles, and the C0 outputs of a joint are written as coalesced rows.
struct EmbedWarpArgs {
  const float *x;            // (N, C_in, T, V), or any view with element strides (sn, sc, st, 1)
  long long sn, sc, st;      // strides of trial / channel / frame (sliding windows: sn = st = V)
  int N, T, V, C_in, C0;
  const float *n_w, *n_b;    // (C_in, V)
  float eps;
  const float *W, *bias;     // (C0, C_in), (C0)
  float *out;                // [N*T*V, C0], or (out == null) bf16 hi/lo planes of the same rows
  __nv_bfloat16 *out_hi, *out_lo;
};

__global__ void __launch_bounds__(256) k_embed_warp(EmbedWarpArgs p) {
  extern __shared__ __align__(16) float s_buf[];
  const int VCi = p.V * p.C_in;
  float *sw = s_buf;                              // C0*C_in
  float *sb = sw + p.C0 * p.C_in;                 // C0
  float *sxn = sb + p.C0;                         // [8 warps][VCi]
  for (int i = threadIdx.x; i < p.C0 * p.C_in; i += blockDim.x) sw[i] = p.W[i];
  for (int i = threadIdx.x; i < p.C0; i += blockDim.x) sb[i] = p.bias[i];
  __syncthreads();
  griddep_wait();                                 // the input frames (and the output buffer's previous readers)
  griddep_launch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // C_in == 3 and 32 % (C0/4) == 0 (the shipped trunks: 3 -> 64): a lane always produces the same four
  // output channels, so its weights live in registers and the inner product needs no shared-memory reads
  const bool reg_w = p.C_in == 3 && (p.C0 >> 2) <= 32 && 32 % (p.C0 >> 2) == 0;
  const int co_l = (lane % (p.C0 >> 2)) * 4;
  float wr[12];
  float4 br = make_float4(0.f, 0.f, 0.f, 0.f);
  if (reg_w) {
#pragma unroll
    for (int k = 0; k < 12; ++k) wr[k] = sw[(co_l + k / 3) * 3 + (k % 3)];
    br = make_float4(sb[co_l], sb[co_l + 1], sb[co_l + 2], sb[co_l + 3]);
  }
  float *xn = sxn + warp * VCi;
  const float inv_n = 1.f / (float)VCi, inv_nm1 = 1.f / (float)(VCi - 1);
  const long long frames = (long long)p.N * p.T;
  for (long long f = (long long)blockIdx.x * 8 + warp; f < frames; f += (long long)gridDim.x * 8) {
    const long long n = f / p.T;
    const int t = (int)(f - n * p.T);
    float v[4];
    int cidx[4], vidx[4];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = lane + 32 * q;                // i = c*V + vj (the reference layout of one frame)
      v[q] = 0.f;
      cidx[q] = i / p.V;
      vidx[q] = i - cidx[q] * p.V;
      if (i < VCi) {
        v[q] = p.x[n * p.sn + cidx[q] * p.sc + t * p.st + vidx[q]];
        s += v[q];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_n;
    float qd = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (lane + 32 * q < VCi) {
        const float d = v[q] - mean;
        qd = fmaf(d, d, qd);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qd += __shfl_xor_sync(0xffffffffu, qd, o);
    const float rstd = 1.f / sqrtf(qd * inv_nm1 + p.eps);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = lane + 32 * q;
      if (i < VCi) xn[vidx[q] * p.C_in + cidx[q]] = (v[q] - mean) * rstd * __ldg(p.n_w + i) + __ldg(p.n_b + i);
    }
    __syncwarp();
    const long long dst_o = f * (long long)p.V * p.C0;
    float *dst = p.out + dst_o;
    const int C04 = p.C0 >> 2;                                 // C0 % 4 == 0 (checked by the caller)
    if (reg_w) {
      // lane -> fixed channel group co (32 % (C0/4) == 0): its 4 x C_in weights and biases sit in registers
      for (int i = lane; i < p.V * C04; i += 32) {
        const int vj = i / C04;
        const float x0 = xn[vj * 3], x1 = xn[vj * 3 + 1], x2 = xn[vj * 3 + 2];
        float4 acc;
        acc.x = fmaf(wr[2], x2, fmaf(wr[1], x1, fmaf(wr[0], x0, br.x)));
        acc.y = fmaf(wr[5], x2, fmaf(wr[4], x1, fmaf(wr[3], x0, br.y)));
        acc.z = fmaf(wr[8], x2, fmaf(wr[7], x1, fmaf(wr[6], x0, br.z)));
        acc.w = fmaf(wr[11], x2, fmaf(wr[10], x1, fmaf(wr[9], x0, br.w)));
        if (p.out) {
          *reinterpret_cast<float4 *>(dst + vj * p.C0 + co_l) = acc;
        } else {
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(acc.x, acc.y), h23 = __floats2bfloat162_rn(acc.z, acc.w);
          const long long o = dst_o + vj * p.C0 + co_l;
          *reinterpret_cast<uint2 *>(p.out_hi + o) =
              make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
          if (p.out_lo) {
            const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
            const __nv_bfloat162 l01 = __floats2bfloat162_rn(acc.x - f01.x, acc.y - f01.y);
            const __nv_bfloat162 l23 = __floats2bfloat162_rn(acc.z - f23.x, acc.w - f23.y);
            *reinterpret_cast<uint2 *>(p.out_lo + o) =
                make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
          }
        }
      }
      __syncwarp();
      continue;
    }
    for (int i = lane; i < p.V * C04; i += 32) {
      const int vj = i / C04, co = (i - vj * C04) * 4;
      float4 acc = make_float4(sb[co], sb[co + 1], sb[co + 2], sb[co + 3]);
      for (int c = 0; c < p.C_in; ++c) {
        const float xv = xn[vj * p.C_in + c];
        acc.x = fmaf(sw[co * p.C_in + c], xv, acc.x);
        acc.y = fmaf(sw[(co + 1) * p.C_in + c], xv, acc.y);
        acc.z = fmaf(sw[(co + 2) * p.C_in + c], xv, acc.z);
        acc.w = fmaf(sw[(co + 3) * p.C_in + c], xv, acc.w);
      }
      if (p.out) {
        *reinterpret_cast<float4 *>(dst + vj * p.C0 + co) = acc;
      } else {
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(acc.x, acc.y), h23 = __floats2bfloat162_rn(acc.z, acc.w);
        const long long o = dst_o + vj * p.C0 + co;
        *reinterpret_cast<uint2 *>(p.out_hi + o) =
            make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
        if (p.out_lo) {
          const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(acc.x - f01.x, acc.y - f01.y);
          const __nv_bfloat162 l23 = __floats2bfloat162_rn(acc.z - f23.x, acc.w - f23.y);
          *reinterpret_cast<uint2 *>(p.out_lo + o) =
              make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
        }
      }
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------------------- //
// pooling + classifier (stgcn.py:92-95, rtstgcn.py:149-152)
//   part[n][chunk][c] = sum over the chunk's rows;  logits = Wo * mean + bo
// --------------------------------------------------------------------------- //
__global__ void __launch_bounds__(256)
    k_pool_partial(const float *__restrict__ x, long long R, int C, int rows_per_chunk, int nchunk,
                   float *__restrict__ part) {
  const long long n = blockIdx.y;
  const int ch = blockIdx.x;
  long long r0 = (long long)ch * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > R) r1 = R;
  const float *xs = x + n * R * C;
  const int C4 = C >> 2;
  if ((C & 3) == 0 && C4 <= 256 && 256 % C4 == 0) {
    // float4 lanes over the channels, 256 / (C/4) row groups in parallel, four independent
    // accumulators per thread; the row groups are merged in a fixed order (deterministic)
    __shared__ float4 s_acc[256];
    const int c4 = threadIdx.x % C4, rg = threadIdx.x / C4, RG = 256 / C4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    const float4 *xp = reinterpret_cast<const float4 *>(xs) + c4;
    long long r = r0 + rg;
    for (; r + 3 * RG < r1; r += 4 * RG) {
      const float4 v0 = __ldg(xp + r * C4), v1 = __ldg(xp + (r + RG) * C4), v2 = __ldg(xp + (r + 2 * RG) * C4),
                   v3 = __ldg(xp + (r + 3 * RG) * C4);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
      a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; r < r1; r += RG) {
      const float4 v0 = __ldg(xp + r * C4);
      a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    a0.x = (a0.x + a1.x) + (a2.x + a3.x);
    a0.y = (a0.y + a1.y) + (a2.y + a3.y);
    a0.z = (a0.z + a1.z) + (a2.z + a3.z);
    a0.w = (a0.w + a1.w) + (a2.w + a3.w);
    s_acc[threadIdx.x] = a0;
    __syncthreads();
    if (rg == 0) {
      float4 t = s_acc[c4];
      for (int g = 1; g < RG; ++g) {
        const float4 u = s_acc[g * C4 + c4];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      reinterpret_cast<float4 *>(part + (n * nchunk + ch) * C)[c4] = t;
    }
    return;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += xs[r * C + c];
    part[(n * nchunk + ch) * C + c] = s;
  }
}

// RT-ST-GCN training-time temporal stage (rtstgcn.py:366-381): causal sum of `taps` frames spaced
// `stride` apart, o[n,t,:] = sum_{j<taps} z[n, t - j*stride, :] (zero before the sequence start).
// z, o are [N*T, VC] channels-last rows of one frame each.
__global__ void k_causal_tap_sum(const float *__restrict__ z, float *__restrict__ o, long long N, int T, int VC,
                                 int taps, int stride) {
  const long long total = N * T * (long long)(VC / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % (VC / 4));
    const long long f = i / (VC / 4);
    const int t = (int)(f % T);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < taps && t - j * stride >= 0; ++j) {
      const float4 v = *reinterpret_cast<const float4 *>(z + (f - (long long)j * stride) * VC + q * 4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4 *>(o + f * VC + q * 4) = s;
  }
}

// mean over joints: x (N, C, L, V) -> y (N, C, L)   (nn.AvgPool2d((1, V)), rtstgcn.py:127,149)
__global__ void k_mean_joints(const float *__restrict__ x, float *__restrict__ y, long long rows, int V) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float s = 0.f;
  for (int v = 0; v < V; ++v) s += x[i * V + v];
  y[i] = s / (float)V;
}

__global__ void k_pool_sum(const float *__restrict__ part, int nchunk, int C, int N, float *__restrict__ sums) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * C) return;
  const long long n = i / C;
  const int c = (int)(i - n * C);
  float s = 0.f;
  for (int k = 0; k < nchunk; ++k) s += part[(n * nchunk + k) * C + c];
  sums[i] = s;
}

__global__ void __launch_bounds__(256)
    k_pool_fc(const float *__restrict__ part, int nchunk, int C, float inv_R, const float *__restrict__ W,
              const float *__restrict__ bias, int classes, float *__restrict__ logits) {
  extern __shared__ float s_pool[];  // C
  const long long n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nchunk; ++k) s += part[(n * nchunk + k) * C + c];
    s_pool[c] = s * inv_R;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int m = warp; m < classes; m += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(W[(long long)m * C + c], s_pool[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) logits[n * classes + m] = s + bias[m];
  }
}

}  // namespace stgcn
