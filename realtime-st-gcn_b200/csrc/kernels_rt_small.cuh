// RT-ST-GCN continual step for a FEW streams (latency path; rtstgcn.py:137-157, 528-553, 591-627).
//
// At batch 1 the step is 25 rows x 19.8 M MACs: far too small for a 128-row tensor-core tile and,
// run as one kernel per layer, dominated by launch and pipeline-fill latency (0.38 ms measured).
// This kernel runs the WHOLE step -- input norm + fcn_in, every online layer with its FIFO /
// accumulator update, pooling and fcn_out -- in ONE launch: one thread-block cluster of kNC CTAs
// per stream.  Inside a layer the CTAs split the output channels (each streams 1/kNC of the
// weights from L2, coalesced row reads broadcast to the warp); the exchange step is one cluster
// weights from L2 through a cp.async ring); a layer has two exchange steps, both over distributed
// shared memory with a cluster barrier: (1) the partial LayerNorm statistics of every CTA's slice
// of the updated accumulator go to ALL CTAs; (2) every CTA normalises its own slice (still in
// registers) and writes it into ALL CTAs' copy of the next layer's input.  The 1x1 feature
// transforms run as 3xTF32 warp-level MMAs (x = hi + lo, hi*hi + hi*lo + lo*hi: fp32 parity to
// ~1e-6) in EVERY math mode, including STGCN_MATH_FP32 and STGCN_MATH_BF16 (the latency path is
// not MMA-bound); adjacency, state update and LayerNorm are fp32 FMAs in the reference's order.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace stgcn {
namespace rts {

namespace cg = cooperative_groups;

constexpr int kNC = 8;            // CTAs per cluster (portable maximum)
constexpr int kThreads = 256;
constexpr int kMaxLayers = 16;
constexpr int kRowsPerWarp = 4;   // output rows a warp accumulates at once (register tile)
constexpr int kChunk = 32;        // input channels per staged weight chunk
constexpr int kWPitch = kChunk + 4;  // floats per staged weight row (+4: conflict-free fragment loads)
#ifndef STGCN_RTS_STAGES
#define STGCN_RTS_STAGES 3   // 3 / 5 / 8 measured: 0.150 / 0.150 / 0.148 ms per 1-stream step (not the bound)
#endif
constexpr int kStages = STGCN_RTS_STAGES;   // weight-chunk ring depth (cp.async, kStages - 1 chunks in flight)
constexpr int kCsrPtrMax = 128;   // K*V + 1 row pointers staged in shared memory
constexpr int kCsrNnzMax = 256;   // adjacency non-zeros staged in shared memory (tree graphs: ~3V)

struct Layer {
  int c_in, c_out, F, S, residual;            // F ring slots, S accumulators; residual: STGCN_RES_*
  const float *gcn_w, *gcn_b;                 // (K*c_out, c_in), (K*c_out)
  const float *n1_w, *n1_b;                   // (c_out, V)
  const float *res_w, *nr_w, *nr_b;           // (c_out, c_in) no bias; (c_out, V)
  const int *kw_ptr;                          // adjacency CSR ordered by (k, w)
  const int2 *kw_va;
  float *fifo, *acc;                          // [F][B][V][c_out], [S][B][V][c_out]
};

// Measurement aid (STGCN_DEBUG & 4): cycles of CTA (0,0) per phase: 0 input 1 gemm 2 adj+state 3 stats+barrier 4 normalise 5 tail
__device__ unsigned long long g_dbg[8];

struct Params {
  int num_layers, V, K, in_feat, num_classes, B, c_max, debug;
  int period;                                 // frame counters wrap at this value (0: never)
  int single_pass;                            // math = bf16: one TF32 product instead of the 3xTF32 split
  float eps;
  const float *x;                             // (B, in_feat, 1, V)
  float *logits;                              // (B, num_classes)
  int *counter;                               // [B]
  const float *norm_in_w, *norm_in_b;         // (in_feat, V)
  const float *fcn_in_w, *fcn_in_b;           // (c0, in_feat)
  const float *fcn_out_w, *fcn_out_b;         // (classes, c_last)
  Layer layer[kMaxLayers];
};

// fp32 -> (hi, lo) TF32 pair: x = hi + lo to ~2^-21; hi*hi + hi*lo + lo*hi on the tensor core keeps
// the 1x1 feature transform inside the fp32 parity bound (3xTF32).
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ float block_sum(float v, float *red) {
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) s += red[i];
  return s;
}

// shared-memory floats needed for `c_max` channels, `V` joints, `K` partitions
inline size_t smem_floats(int c_max, int V, int K) {
  const size_t cv = (size_t)c_max * V;
  const size_t ys = (size_t)(K + 1) * (c_max / kNC) * V;
  const size_t wr = (size_t)kStages * (K + 1) * (c_max / kNC) * kWPitch;  // weight-chunk ring
  const size_t stage = 2 * (size_t)(c_max / kNC) * V;      // this CTA's normalised slice, per layer parity
  return 2 * cv /* x ping-pong */ + ys + wr + 2 * kNC * 4 /* stats */ + 64 +
         kCsrPtrMax + 2 * kCsrNnzMax + stage + 4 /* mbarrier */;
}

__global__ void __cluster_dims__(kNC, 1, 1) __launch_bounds__(kThreads, 1) k_rt_small(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  const int V = p.V, K = p.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t cv = (size_t)p.c_max * V;
  float *xbuf[2] = {sm, sm + cv};
  float *ybuf = sm + 2 * cv;
  float *wring = ybuf + (size_t)(K + 1) * (p.c_max / kNC) * V;   // [kStages][rows_max][kChunk]
  const int stage_floats = (K + 1) * (p.c_max / kNC) * kWPitch;
  float *stats = wring + (size_t)kStages * stage_floats;         // [2][kNC][4]
  float *red = stats + 2 * kNC * 4;
  int *s_ptr = reinterpret_cast<int *>(red + 64);                // [kCsrPtrMax]
  int2 *s_va = reinterpret_cast<int2 *>(s_ptr + kCsrPtrMax);      // [kCsrNnzMax]
  float *xstage = reinterpret_cast<float *>(s_va + kCsrNnzMax);   // [2][c_max/kNC][V]: slice on its way to the peers
  const uint32_t xbar = tc::smem_u32(xstage + 2 * (size_t)(p.c_max / kNC) * V);   // "next layer's input complete"
  if (tid == 0) {
    tc::mbar_init(xbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  const int cnt = p.counter[b];
  const bool dbg = (p.debug & 4) && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = dbg ? clock64() : 0;
  auto lap = [&](int i) {
    if (dbg) {
      const long long now = clock64();
      tph[i] += now - tlast;
      tlast = now;
    }
  };

  // ---- weight streaming: this CTA's rows of every layer's 1x1 weights (and residual conv) pass
  // through a ring of kChunk-channel chunks, copied with cp.async two chunks ahead of their use --
  // across layer boundaries too, so the copies also overlap the cluster barriers.
  int total_chunks = 0;
  for (int l = 0; l < p.num_layers; ++l) total_chunks += p.layer[l].c_in / kChunk;
  int il = 0, ic = 0;                                  // layer / chunk of the next chunk to issue
  int issued = 0;
  auto issue_next = [&]() {
    if (issued < total_chunks) {
      const Layer &L = p.layer[il];
      const int Cs = L.c_out / kNC, c0 = rank * Cs;
      const int rows = K * Cs + (L.residual == 2 ? Cs : 0);
      float *dst = wring + (size_t)(issued % kStages) * stage_floats;
      for (int pc = tid; pc < rows * (kChunk / 4); pc += kThreads) {
        const int j = pc / (kChunk / 4), q = pc - j * (kChunk / 4);
        const float *src;
        if (j < K * Cs) {
          const int k = j / Cs, cl = j - k * Cs;
          src = L.gcn_w + (size_t)(k * L.c_out + c0 + cl) * L.c_in;
        } else {
          src = L.res_w + (size_t)(c0 + j - K * Cs) * L.c_in;
        }
        src += ic * kChunk + q * 4;
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + j * kWPitch + q * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
      }
      if (++ic == L.c_in / kChunk) { ic = 0; ++il; }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty) group keeps the count uniform
    ++issued;
  };
#pragma unroll
  for (int i = 0; i < kStages - 1; ++i) issue_next();
  int consumed = 0;

  // ---- input stage (every CTA, redundantly): LayerNorm over (in_feat, V), then fcn_in ----
  {
    const int n_in = p.in_feat * V;
    float *xin = ybuf;                                   // scratch
    float v = 0.f;
    if (tid < n_in) v = p.x[(size_t)b * n_in + tid];
    const float mean = block_sum(tid < n_in ? v : 0.f, red) / (float)n_in;
    const float d = tid < n_in ? v - mean : 0.f;
    const float var = block_sum(d * d, red) / (float)(n_in - 1);
    const float rstd = 1.f / sqrtf(var + p.eps);
    if (tid < n_in) xin[tid] = d * rstd * p.norm_in_w[tid] + p.norm_in_b[tid];
    __syncthreads();
    const int c0n = p.layer[0].c_in;
    for (int i = tid; i < c0n * V; i += kThreads) {
      const int c = i / V, w = i - c * V;
      float s = p.fcn_in_b[c];
      for (int ci = 0; ci < p.in_feat; ++ci) s = fmaf(p.fcn_in_w[c * p.in_feat + ci], xin[ci * V + w], s);
      xbuf[0][i] = s;
    }
    __syncthreads();
  }

  lap(0);
  int cur = 0;
  for (int l = 0; l < p.num_layers; ++l) {
    const Layer &L = p.layer[l];
    const int par = l & 1;
    const int Cs = L.c_out / kNC, c0 = rank * Cs;
    const bool res_conv = L.residual == 2;
    const float *x = xbuf[cur];
    const int rows = K * Cs + (res_conv ? Cs : 0);
    // this layer's adjacency (A * importance differs per layer) -> shared memory; lands during (a)
    const int nnz = __ldg(L.kw_ptr + K * V);
    const bool csr_s = nnz <= kCsrNnzMax;
    for (int i = tid; i <= K * V; i += kThreads) s_ptr[i] = __ldg(L.kw_ptr + i);
    if (csr_s)
      for (int i = tid; i < nnz; i += kThreads) s_va[i] = __ldg(L.kw_va + i);

    const int fi = cnt % L.F, ai = cnt % L.S;
    const size_t slot = (size_t)p.B * V * L.c_out;
    float oreg[4], rreg[4], aold[4], fold[4], g1[4], b1[4], gr[4], br[4];
    float so = 0.f, sr = 0.f;
    const int n_el = V * Cs;
    // the state loads do not depend on this step's arithmetic: issue them before the GEMM
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * kThreads;
      aold[it] = 0.f;
      fold[it] = 0.f;
      g1[it] = b1[it] = gr[it] = br[it] = 0.f;
      if (idx < n_el) {
        const int w = idx / Cs, cl = idx - w * Cs;
        const size_t so_ = ((size_t)b * V + w) * L.c_out + c0 + cl;
        fold[it] = L.fifo[(size_t)fi * slot + so_];
        aold[it] = L.acc[(size_t)ai * slot + so_];
        const int at = (c0 + cl) * V + w;                 // reference affine layout (C, 1, V)
        g1[it] = __ldg(L.n1_w + at);
        b1[it] = __ldg(L.n1_b + at);
        if (res_conv) {
          gr[it] = __ldg(L.nr_w + at);
          br[it] = __ldg(L.nr_b + at);
        }
      }
    }

    // ---- (a) this CTA's rows of the 1x1 feature transform (+ residual conv): y[j][v] ----
    // A 25-row problem cannot fill a 128-row tcgen05 tile; the shape that fits is the transposed
    // product on warp-level MMAs: D[row j][joint v] = W[j][ci] * x[ci][v], m16n8k8 TF32 with the
    // 3xTF32 split (fp32 parity).  Warp w owns rows 16w..16w+15 (<= 128 rows per CTA) and all four
    // 8-joint column tiles; A fragments come from the staged weight chunk, B fragments from the
    // fp32 activations in shared memory, both split into (hi, lo) on the fly.
    const int g = lane >> 2, tq = lane & 3;
    const int jr = warp * 16;
    const bool active = jr < rows;
    const int NT = (V + 7) >> 3;
    float dacc[4][4];
    {
      float bias0 = 0.f, bias1 = 0.f;
      const int ja = jr + g, jb = jr + g + 8;
      if (ja < K * Cs) bias0 = __ldg(L.gcn_b + (ja / Cs) * L.c_out + c0 + ja % Cs);
      if (jb < K * Cs) bias1 = __ldg(L.gcn_b + (jb / Cs) * L.c_out + c0 + jb % Cs);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        dacc[nt][0] = bias0; dacc[nt][1] = bias0; dacc[nt][2] = bias1; dacc[nt][3] = bias1;
      }
    }
    lap(1);
    for (int ch = 0; ch < L.c_in / kChunk; ++ch) {
      issue_next();                                             // chunk consumed + kStages - 1
      lap(6);
      asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");   // chunk `consumed` has landed (this thread's part)
      __syncthreads();                                          // ... and everybody else's
      lap(7);
      if (active) {
        const float *wst = wring + (size_t)(consumed % kStages) * stage_floats + (jr + g) * kWPitch + tq;
        const float *xc = x + (size_t)(ch * kChunk + tq) * V + g;
        if (p.single_pass) {
          // math = bf16: one TF32 product (2^-11 operands, inside the mode's stated 2e-2 tolerance)
#pragma unroll
          for (int k0 = 0; k0 < kChunk; k0 += 8) {
            uint32_t ah[4];
            ah[0] = to_tf32(wst[k0]);
            ah[1] = to_tf32(wst[8 * kWPitch + k0]);
            ah[2] = to_tf32(wst[k0 + 4]);
            ah[3] = to_tf32(wst[8 * kWPitch + k0 + 4]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              if (nt < NT) {
                uint32_t bh[2];
                bh[0] = to_tf32(xc[k0 * V + nt * 8]);
                bh[1] = to_tf32(xc[(k0 + 4) * V + nt * 8]);
                mma_tf32(dacc[nt], ah, bh);
              }
            }
          }
        } else {
#pragma unroll
          for (int k0 = 0; k0 < kChunk; k0 += 8) {
            uint32_t ah[4], al[4];
            split_tf32(wst[k0], ah[0], al[0]);
            split_tf32(wst[8 * kWPitch + k0], ah[1], al[1]);
            split_tf32(wst[k0 + 4], ah[2], al[2]);
            split_tf32(wst[8 * kWPitch + k0 + 4], ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              if (nt < NT) {                                          // 8-joint column tiles that hold joints
                uint32_t bh[2], bl[2];
                split_tf32(xc[k0 * V + nt * 8], bh[0], bl[0]);
                split_tf32(xc[(k0 + 4) * V + nt * 8], bh[1], bl[1]);
                mma_tf32(dacc[nt], al, bh);
                mma_tf32(dacc[nt], ah, bl);
                mma_tf32(dacc[nt], ah, bh);
              }
            }
          }
        }
      }
      ++consumed;
      __syncthreads();                                          // the stage may be overwritten by a later issue
    }
    if (active) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nt * 8 + 2 * tq;
        const int ja = jr + g, jb = jr + g + 8;
        if (col < V) {
          if (ja < rows) ybuf[ja * V + col] = dacc[nt][0];
          if (jb < rows) ybuf[jb * V + col] = dacc[nt][2];
        }
        if (col + 1 < V) {
          if (ja < rows) ybuf[ja * V + col + 1] = dacc[nt][1];
          if (jb < rows) ybuf[jb * V + col + 1] = dacc[nt][3];
        }
      }
    }
    __syncthreads();

    lap(1);
    // ---- (b,c) adjacency contraction + FIFO / accumulator update for (w, c_local) ----
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * kThreads;
      oreg[it] = 0.f;
      rreg[it] = 0.f;
      if (idx < n_el) {
        const int w = idx / Cs, cl = idx - w * Cs;
        float z = 0.f;
        for (int k = 0; k < K; ++k) {
          const int e1 = s_ptr[k * V + w + 1];
          for (int e = s_ptr[k * V + w]; e < e1; ++e) {
            const int2 va = csr_s ? s_va[e] : __ldg(L.kw_va + e);
            z = fmaf(ybuf[(k * Cs + cl) * V + va.x], __int_as_float(va.y), z);
          }
        }
        const size_t so_ = ((size_t)b * V + w) * L.c_out + c0 + cl;
        float a = aold[it] + z;
        a = a + (-fold[it]);                              // rtstgcn.py:611-612
        L.acc[(size_t)ai * slot + so_] = a;
        L.fifo[(size_t)fi * slot + so_] = z;              // rtstgcn.py:621
        oreg[it] = a;
        so += a;
        if (res_conv) {
          rreg[it] = ybuf[(K * Cs + cl) * V + w];
          sr += rreg[it];
        }
      }
    }
    if (n_el > 4 * kThreads) __trap();
    lap(2);
    // partial statistics over this CTA's n_el elements (mean, then M2 from registers) -> every CTA
    const float mo = block_sum(so, red) / (float)n_el;
    const float mr = res_conv ? block_sum(sr, red) / (float)n_el : 0.f;
    float qo = 0.f, qr = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it)
      if (tid + it * kThreads < n_el) {
        const float d = oreg[it] - mo;
        qo = fmaf(d, d, qo);
        const float e = rreg[it] - mr;
        qr = fmaf(e, e, qr);
      }
    qo = block_sum(qo, red);
    if (res_conv) qr = block_sum(qr, red);
    if (tid < kNC) {
      float *st = cluster.map_shared_rank(stats, tid) + (par * kNC + rank) * 4;
      st[0] = mo; st[1] = qo; st[2] = mr; st[3] = qr;
    }
    cluster.sync();                                       // exchange 1: LayerNorm statistics
    lap(3);

    // ---- (e) merge the statistics; normalise THIS CTA's slice (still in registers) and write it
    // into every CTA's copy of the next layer's input (distributed shared memory) ----
    float mean_o = 0.f, mean_r = 0.f;
#pragma unroll
    for (int rk = 0; rk < kNC; ++rk) {
      mean_o += stats[(par * kNC + rk) * 4];
      mean_r += stats[(par * kNC + rk) * 4 + 2];
    }
    mean_o *= 1.f / kNC;
    mean_r *= 1.f / kNC;
    float M2o = 0.f, M2r = 0.f;
#pragma unroll
    for (int rk = 0; rk < kNC; ++rk) {
      const float dm = stats[(par * kNC + rk) * 4] - mean_o, dr = stats[(par * kNC + rk) * 4 + 2] - mean_r;
      M2o += stats[(par * kNC + rk) * 4 + 1] + (float)n_el * dm * dm;
      M2r += stats[(par * kNC + rk) * 4 + 3] + (float)n_el * dr * dr;
    }
    const float inv_nm1 = 1.f / (float)(L.c_out * V - 1);
    const float rstd_o = 1.f / sqrtf(M2o * inv_nm1 + p.eps);
    const float rstd_r = 1.f / sqrtf(M2r * inv_nm1 + p.eps);
    // The slice [Cs][V] is contiguous in every CTA's copy of the next input: it is staged locally and sent to
    // each CTA of the cluster with ONE bulk shared::cta -> shared::cluster copy that completes on the
    // receiver's mbarrier -- no per-element remote stores (they were uncoalesced 4-byte DSMEM writes, 8 per
    // value) and no second cluster barrier.  Re-use is safe without further synchronisation: a peer sends
    // layer l+1 only after exchange 1 of layer l+1, i.e. after every CTA has finished layer l, and the staging
    // slot of a parity is rewritten two layers later, when all receivers have seen this layer's bytes.
    float *xn = xbuf[cur ^ 1];
    float *stg = xstage + (size_t)par * (p.c_max / kNC) * V;
    const uint32_t slice_bytes = (uint32_t)(Cs * V * sizeof(float));
    if (tid == 0) tc::mbar_expect_tx(xbar, kNC * slice_bytes);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * kThreads;
      if (idx < n_el) {
        const int w = idx / Cs, cl = idx - w * Cs;
        const int at = (c0 + cl) * V + w;
        float y = fmaxf((oreg[it] - mean_o) * rstd_o * g1[it] + b1[it], 0.f);                 // bn_relu
        if (L.residual == 1) y = fmaxf(y + x[at], 0.f);
        else if (res_conv) y = fmaxf(y + (rreg[it] - mean_r) * rstd_r * gr[it] + br[it], 0.f);
        stg[cl * V + w] = y;
      }
    }
    tc::fence_proxy_async();                              // generic writes of the slice -> async-proxy reads
    __syncthreads();
    if (tid < kNC) {
      const uint32_t src = tc::smem_u32(stg);
      const uint32_t dst_local = tc::smem_u32(xn + (size_t)c0 * V);
      uint32_t dst, bar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(dst_local), "r"(tid));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"(xbar), "r"(tid));
      asm volatile(
          "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
          "r"(src), "r"(slice_bytes), "r"(bar)
          : "memory");
    }
    tc::mbar_wait(xbar, (uint32_t)(l & 1));               // all eight slices of the next input have landed here
    cur ^= 1;
    lap(4);
  }

  // ---- pooling over joints + classifier (rank 0 of the cluster) ----
  if (rank == 0) {
    const int C = p.layer[p.num_layers - 1].c_out;
    const float *x = xbuf[cur];
    float *pooled = ybuf;
    for (int c = tid; c < C; c += kThreads) {
      float s = 0.f;
      for (int w = 0; w < V; ++w) s += x[c * V + w];
      pooled[c] = s / (float)V;
    }
    __syncthreads();
    for (int m = warp; m < p.num_classes; m += kThreads / 32) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s = fmaf(__ldg(p.fcn_out_w + (size_t)m * C + c), pooled[c], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) p.logits[(size_t)b * p.num_classes + m] = s + __ldg(p.fcn_out_b + m);
    }
    if (tid == 0) p.counter[b] = (p.period > 0 && cnt + 1 >= p.period) ? 0 : cnt + 1;
  }
  cluster.sync();   // no CTA may exit while others can still write into its shared memory
  lap(5);
  if (dbg)
    for (int i = 0; i < 8; ++i) atomicAdd(&g_dbg[i], (unsigned long long)tph[i]);
}

}  // namespace rts
}  // namespace stgcn
