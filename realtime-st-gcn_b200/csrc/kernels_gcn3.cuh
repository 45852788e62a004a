// Graph convolution v3 (sm_100a): reference operation order on CTA pairs.
//
//   y[(t,v), k*C+c] = sum_ci W[k*C+c, ci] x[(t,v), ci]          tcgen05 GEMM, N = K*64 per block
//   z[(t,w), c]     = bz[w,c] + sum_{k,v} A[k,v,w] y[(t,v), k*C+c]   epilogue, from shared memory
//   out             = relu(LN_{C,V}(z))  (or LN_R(y + b) for the residual 1x1 branch, K = 1)
//
// Why this order (tgcn.py:70-79 does the same): the activations arrive as bf16 hi/lo planes written
// by the producing kernel's epilogue, so the A operand of the GEMM is a plain TMA load -- the
// CUDA-core adjacency transform + re-split of kernel v2 (which bound it at 10-40 % tensor-pipe)
// disappears, the MMA gets N = 192 instead of N = C (no longer shared-memory bound), and the
// V x V contraction touches each accumulator once with ~3 FMAs (tree adjacency).
//
// One CTA pair (cta_group::2, M = 256) per two 128-row tiles (FT whole frames each); per CTA:
//   warp 0   lane 0: TMA producer of the activation chunks ([128 rows][64 ch] bf16 per plane, ring of 8)
//            lane 1: TMA producer of this CTA's half of every weight tile
//   warp 1   MMA issuer (leader CTA only); accumulators double buffered in TMEM (2 x K*64 columns)
//   warps 2..21 epilogue (20 warps: the CUDA-core work is latency-bound, so it is spread over as
//            many warps as the register file allows).  Per 32-channel pass: (1) warps 2..17 copy the
//            pass's K*32 accumulator columns TMEM -> shared memory (thread = row); (2) all twenty
//            warps contract over the adjacency: four warps own one frame, a quarter-warp one output
//            row (8 lanes x float4 = 32 channels, conflict-free 128-B reads), results stay in
//            registers for all passes.  After the last pass the frame statistics are merged by
//            shuffles (+ one 128-thread named barrier), and the normalised rows are written as
//            64/128-B row segments.
#pragma once
#include "kernels_tc_pair.cuh"

namespace stgcn {
namespace tc {

constexpr int kG3EpiWarps = 20;
constexpr int kG3LdWarps = 16;          // epilogue warps that read TMEM (four per lane quarter)
constexpr int kG3EpiThreads = 32 * kG3EpiWarps;
constexpr int kG3Threads = 32 * (2 + kG3EpiWarps);
constexpr int kG3R = 8;                 // activation chunk ring slots
constexpr int kG3Chunk = 16384;         // [128 rows][64 ch] bf16
constexpr int kG3Steps = 8;             // gather steps held in shared memory per row (more: slow global path)
constexpr int kG3MaxEnt = 2 * 4 * kG3Steps * 4;   // padded entry table [warp-in-pair][slot][step][quarter-warp]
constexpr int kG3MaxV = 32;
constexpr int kG3EntCap = 3 * kG3MaxV * kG3MaxV;

// Per-layer gather tables, built on the device from A_eff (K, V, V):
//   entries of output joint w: ent[row_ptr[w] .. row_ptr[w+1]) = (v | k << 8, A[k,v,w])
//   rowmap[(warp-in-frame*2 + slot)*4 + quarter-warp] = output joint handled there (or -1); joints
//   are dealt in order of decreasing entry count so the four rows of one warp instruction have
//   similar trip counts, heavy and light groups paired per warp.
//   ent2[((warp-in-frame*2 + slot)*kG3Steps + step)*4 + quarter-warp]: the step-th entry of the row at
//   that position, padded with (0, 0.f); cm[warp-in-frame*2 + slot] = steps needed there (<= kG3Steps).
struct Gcn3Tables {
  int row_ptr[kG3MaxV + 1];
  int rowmap[32];
  int nent;
  int cm[8];
  int pad[1];
  int2 ent[kG3EntCap];
  int2 ent2[kG3MaxEnt];
};

__global__ void k_gcn3_tables(const float *__restrict__ A, int K, int V, int identity, Gcn3Tables *tab) {
  __shared__ int cnt[kG3MaxV + 1];
  const int w = threadIdx.x;
  if (w < V) {
    int c = 0;
    if (identity) c = 1;
    else
      for (int k = 0; k < K; ++k)
        for (int v = 0; v < V; ++v) c += (A[((long long)k * V + v) * V + w] != 0.f);
    cnt[w] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < V; ++i) {
      tab->row_ptr[i] = run;
      run += cnt[i];
    }
    tab->row_ptr[V] = run;
    tab->nent = run;
    // joints by decreasing entry count (stable insertion sort), dealt four at a time
    int order[kG3MaxV];
    for (int i = 0; i < V; ++i) {
      int j = i;
      while (j > 0 && cnt[order[j - 1]] < cnt[i]) {
        order[j] = order[j - 1];
        --j;
      }
      order[j] = i;
    }
    for (int i = 0; i < 32; ++i) tab->rowmap[i] = -1;
    for (int i = 0; i < V; ++i) {
      const int g = i >> 2, qw = i & 3;       // group g (four joints) -> warp g / 7-g of the frame, slot 0 / 1
      const int pos = g < 4 ? g * 2 : (7 - g) * 2 + 1;
      tab->rowmap[pos * 4 + qw] = order[i];
    }
  }
  __syncthreads();
  if (w < V) {
    int at = tab->row_ptr[w];
    if (identity) tab->ent[at] = make_int2(w, __float_as_int(1.f));
    else
      for (int k = 0; k < K; ++k)
        for (int v = 0; v < V; ++v) {
          const float a = A[((long long)k * V + v) * V + w];
          if (a != 0.f) tab->ent[at++] = make_int2(v | (k << 8), __float_as_int(a));
        }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kG3MaxEnt; i += blockDim.x) {
    const int qw = i & 3, step = (i >> 2) % kG3Steps, pos = (i >> 2) / kG3Steps;   // pos = warp-in-frame*2 + slot
    const int j = tab->rowmap[pos * 4 + qw];
    int2 en = make_int2(0, 0);
    if (j >= 0 && step < tab->row_ptr[j + 1] - tab->row_ptr[j]) en = tab->ent[tab->row_ptr[j] + step];
    tab->ent2[i] = en;
  }
  if (threadIdx.x < 8) {
    int m = 0;
    for (int qw = 0; qw < 4; ++qw) {
      const int j = tab->rowmap[threadIdx.x * 4 + qw];
      if (j >= 0) m = max(m, tab->row_ptr[j + 1] - tab->row_ptr[j]);
    }
    tab->cm[threadIdx.x] = min(m, kG3Steps);
  }
}

// 1x1 weights (K*CO, Cin) fp32 -> bf16 tiles [plane][block b][kc][n = k*64 + c][64 ci]: one tile is the
// B operand (N = K*64 rows, K-major) of block b / input chunk kc; a CTA of the pair loads rows
// rank*N/2 .. of it.
__global__ void k_pack_gcn3_w(const float *__restrict__ w, __nv_bfloat16 *__restrict__ wb, int CO, int Cin, int K) {
  const long long total = (long long)K * CO * Cin;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int KC = Cin / 64, BN = K * 64;
  const int ci = (int)(i & 63);
  long long r = i >> 6;
  const int n = (int)(r % BN);
  r /= BN;
  const int kc = (int)(r % KC);
  const int b = (int)(r / KC);
  const int k = n >> 6, c = n & 63;
  __nv_bfloat16 hi, lo;
  split_bf16(w[((long long)k * CO + b * 64 + c) * Cin + kc * 64 + ci], hi, lo);
  wb[i] = hi;
  wb[total + i] = lo;
}

// (C, 1, V) parameter -> [V][C] (an epilogue quarter-warp reads 128 contiguous bytes of one joint)
__global__ void k_transpose_cv(const float *__restrict__ src, float *__restrict__ dst, int C, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * V) return;
  const int w = i / C, c = i - w * C;
  dst[i] = src[c * V + w];
}
// bias through the adjacency as [V][C]: bz[w][c] = sum_k b[k*C + c] * sum_v A[k,v,w]
__global__ void k_bias_through_adj_vc(const float *__restrict__ A, const float *__restrict__ bg, int K, int V, int CO,
                                      float *__restrict__ bz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= CO * V) return;
  const int w = i / CO, c = i - w * CO;
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    float col = 0.f;
    for (int v = 0; v < V; ++v) col += A[((long long)k * V + v) * V + w];
    s = fmaf(bg[k * CO + c], col, s);
  }
  bz[i] = s;
}
// fp32 rows -> bf16 hi/lo planes (layer-level entry points; the model path writes planes directly)
__global__ void k_rows_to_planes(const float *__restrict__ x, __nv_bfloat16 *__restrict__ hi,
                                 __nv_bfloat16 *__restrict__ lo, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  __nv_bfloat16 h, l;
  split_bf16(x[i], h, l);
  hi[i] = h;
  if (lo) lo[i] = l;
}

__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap *m, uint32_t bar_cluster, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 32 lanes x 16 columns, no wait (several loads in flight, then tmem_ld_wait)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// remote arrive without release semantics (the arriving thread publishes no generic-proxy data;
// with .release the compiler emits MEMBAR + ERRBAR in front of every arrive)
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// global stores that do not allocate in L1 (the small L1 left next to ~223 KB of shared memory
// holds the parameter tables)
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream(uint2 *p, const uint2 &v) {
  asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t *r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

struct Gcn3Params {
  int T, V, Cin, planes;          // frames per trial of this view, joints, input channels, bf16 planes (1 or 2)
  int FT, tiles_per_trial, items; // whole frames per 128-row tile; items = trials * tiles_per_trial
  int S;                          // weight stages
  const Gcn3Tables *tab;
  const float *bias;              // bias_v = 1: [V][CO] table, 0: [CO] vector, null: none
  int bias_v;
  const float *n_w, *n_b;         // LayerNorm affine as [V][CO]
  float *out_f32;                 // fp32 rows [trial][T][V][CO], or
  __nv_bfloat16 *out_hi, *out_lo; // bf16 planes [trial][out_T][V][CO] (frame t stored at t + out_t0)
  int out_T, out_t0;
  int relu;
  float eps;
  int debug;
};

// exact merge of two (count, mean, M2) partial statistics (Chan et al.)
__device__ __forceinline__ void stat_merge(float &n, float &m, float &M2, float nb, float mb, float Mb) {
  const float nn = n + nb;
  const float f = nn > 0.f ? nb / nn : 0.f;
  const float d = mb - m;
  m = fmaf(d, f, m);
  M2 = M2 + Mb + d * d * n * f;
  n = nn;
}

template <int CO, int K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG3Threads, 1)
    k_gcn3(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const Gcn3Params p) {
  constexpr int NB = CO / 64;               // 64-channel blocks
  constexpr int NP = 2 * NB;                // 32-channel epilogue passes
  constexpr int BN = K * 64;                // accumulator columns per block = MMA N
  constexpr int kWHalf = (BN / 2) * 128;    // this CTA's half of a weight tile
  constexpr int YP = K * 32 + 4;            // pitch (floats) of the staged accumulator rows
  constexpr int kYBytes = 128 * YP * 4;
  constexpr int R = kG3R;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int S = p.S;
  const uint32_t sX = smem_base;
  const uint32_t sW = sX + R * kG3Chunk;
  const uint32_t sY = sW + S * kWHalf;
  const uint32_t sEnt = sY + kYBytes;
  const uint32_t sMisc = sEnt + kG3MaxEnt * 8;
  const uint32_t sBar = sMisc + 768;
  const uint32_t bXFull = sBar, bXEmpty = sBar + 8 * R, bTFull = sBar + 16 * R, bTEmpty = bTFull + 16;
  const uint32_t bWFull = bTEmpty + 16, bWEmpty = bWFull + 8 * S;
  const uint32_t sTmemPtr = bWEmpty + 8 * S;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  float *s_y = reinterpret_cast<float *>(gen_base + (sY - smem_base));
  int2 *s_ent = reinterpret_cast<int2 *>(gen_base + (sEnt - smem_base));
  int *s_rowptr = reinterpret_cast<int *>(gen_base + (sMisc - smem_base));          // [33]
  int *s_rowmap = s_rowptr + 36;                                                     // [32]
  float *s_stat = reinterpret_cast<float *>(s_rowmap + 32);                          // [5 frames][4 warps][4]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int KC = p.Cin / 64;
  const int CH = KC * p.planes;             // activation chunks per tile
  const int iters = (p.items + 2 * npairs - 1 - 2 * pair) / (2 * npairs);   // same for both CTAs of the pair
#ifdef STGCN_G3_DEBUG   // per-role cycle counters (STGCN_DEBUG=4); compiled out by default: they cost ~20 registers
  const bool dbg = (p.debug & 4) && blockIdx.x == 0;
#else
  constexpr bool dbg = false;
#endif
  long long d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0, d5 = 0, d6 = 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < R; ++i) {
      mbar_init(bXFull + 8 * i, 2);                   // one arrive.expect_tx per CTA of the pair (leader's is used)
      mbar_init(bXEmpty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bTFull + 8 * i, 1);
      mbar_init(bTEmpty + 8 * i, 2 * kG3LdWarps);     // the TMEM-reading warps of BOTH CTAs
    }
    for (int i = 0; i < S; ++i) {
      mbar_init(bWFull + 8 * i, 2);
      mbar_init(bWEmpty + 8 * i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2(sTmemPtr, 512);
  // gather tables -> shared memory (entries as byte offsets into a frame's staged rows)
  for (int i = threadIdx.x; i <= p.V; i += blockDim.x) s_rowptr[i] = __ldg(&p.tab->row_ptr[i]);
  for (int i = threadIdx.x; i < 32; i += blockDim.x) s_rowmap[i] = __ldg(&p.tab->rowmap[i]);
  for (int i = threadIdx.x; i < kG3MaxEnt; i += blockDim.x) {
    const int2 en = p.tab->ent2[i];
    s_ent[i] = make_int2((en.x & 0xff) * (YP * 4) + (en.x >> 8) * 128, en.y);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // barriers of both CTAs initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    if (lane == 0) {
      // ---- activation chunks of this CTA's own tiles ----
      const uint32_t tx = (uint32_t)(p.FT * p.V * 128);
      int idx = 0;
      for (int it = 0; it < iters; ++it) {
        const int item = 2 * (pair + it * npairs) + (int)rank;
        const bool valid = item < p.items;
        const int n = valid ? item / p.tiles_per_trial : 0;
        const int f0 = valid ? (item - n * p.tiles_per_trial) * p.FT : 0;
        const int n_ld = valid ? n : 0x3fffff;        // out-of-range trial: TMA zero-fills (dummy tile of an odd tail)
        for (int j = 0; j < CH; ++j, ++idx) {
          const int slot = idx % R;
          { DbgTimer tm(dbg); mbar_wait(bXEmpty + 8 * slot, (((uint32_t)idx / R) & 1) ^ 1); tm.stop(d0); }
          const uint32_t lbar = mapa_cta(bXFull + 8 * slot, 0);
          mbar_expect_tx_cluster(lbar, tx);
          tma_load_5d_2sm(sX + slot * kG3Chunk, &tm_x, lbar, (j / p.planes) * 64, 0, f0, n_ld, j % p.planes);
        }
      }
      dbg_flush(dbg, 4, d0);
    } else if (lane == 1) {
      // ---- this CTA's half (rows rank*BN/2 ..) of every weight tile ----
      int ws = 0, w_ph = 0;
      for (int it = 0; it < iters; ++it)
        for (int b = 0; b < NB; ++b)
          for (int kc = 0; kc < KC; ++kc)
            for (int pl = 0; pl < p.planes; ++pl) {
              { DbgTimer tm(dbg); mbar_wait(bWEmpty + 8 * ws, w_ph ^ 1); tm.stop(d0); }
              const uint32_t lbar = mapa_cta(bWFull + 8 * ws, 0);
              mbar_expect_tx_cluster(lbar, kWHalf);
              tma_load_3d_2sm(sW + ws * kWHalf, &tm_w, lbar, 0, (int)rank * (BN / 2), (pl * NB + b) * KC + kc);
              if (++ws == S) { ws = 0; w_ph ^= 1; }
            }
      dbg_flush(dbg, 5, d0);
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---- leader: MMA issuer for the pair (whole warp walks the schedule, one elected lane issues) ----
      constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
      int ws = 0, w_ph = 0, buf = 0, t_ph = 0, idx0 = 0;
      DbgTimer tall(dbg);
      for (int it = 0; it < iters; ++it, idx0 += CH) {
#pragma unroll 1
        for (int b = 0; b < NB; ++b) {
          { DbgTimer tm(dbg); mbar_wait_cluster(bTEmpty + 8 * buf, t_ph ^ 1); tm.stop(d0); }
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * BN;
          uint32_t acc = 0;
          for (int kc = 0; kc < KC; ++kc) {
            const int jh = idx0 + kc * p.planes, jl = jh + 1;
            const int sh = jh % R, sl = jl % R;
            if (b == 0) {
              DbgTimer tm(dbg);
              mbar_wait_cluster(bXFull + 8 * sh, ((uint32_t)jh / R) & 1);
              if (p.planes == 2) mbar_wait_cluster(bXFull + 8 * sl, ((uint32_t)jl / R) & 1);
              tm.stop(d1);
            }
            { DbgTimer tm(dbg); mbar_wait_cluster(bWFull + 8 * ws, w_ph); tm.stop(d2); }
            tc_fence_after();
            const uint32_t a_hi = umma_desc_lo(sX + sh * kG3Chunk), a_lo = umma_desc_lo(sX + sl * kG3Chunk);
            if (elect_one()) {
              const uint32_t b_d = umma_desc_lo(sW + ws * kWHalf);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_bf16_2(tacc, umma_desc_join(a_hi + 2 * kk), umma_desc_join(b_d + 2 * kk), idesc, acc | (uint32_t)kk);
              if (p.planes == 2) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16_2(tacc, umma_desc_join(a_lo + 2 * kk), umma_desc_join(b_d + 2 * kk), idesc, 1u);
              }
              umma_commit_2(bWEmpty + 8 * ws);
            }
            __syncwarp();
            if (++ws == S) { ws = 0; w_ph ^= 1; }
            if (p.planes == 2) {
              { DbgTimer tm(dbg); mbar_wait_cluster(bWFull + 8 * ws, w_ph); tm.stop(d2); }
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_d = umma_desc_lo(sW + ws * kWHalf);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16_2(tacc, umma_desc_join(a_hi + 2 * kk), umma_desc_join(b_d + 2 * kk), idesc, 1u);
                umma_commit_2(bWEmpty + 8 * ws);
              }
              __syncwarp();
              if (++ws == S) { ws = 0; w_ph ^= 1; }
            }
            acc = 1;
            if (b == NB - 1) {
              // last use of these chunks: hand them back to the producers of both CTAs
              if (elect_one()) {
                umma_commit_2(bXEmpty + 8 * sh);
                if (p.planes == 2) umma_commit_2(bXEmpty + 8 * sl);
              }
              __syncwarp();
            }
          }
          if (elect_one()) umma_commit_2(bTFull + 8 * buf);
          __syncwarp();
          buf ^= 1;
          if (buf == 0) t_ph ^= 1;
        }
      }
      tall.stop(d3);
      if (lane == 0) {
        dbg_flush(dbg, 0, d0); dbg_flush(dbg, 1, d1); dbg_flush(dbg, 2, d2); dbg_flush(dbg, 3, d3);
        dbg_flush(dbg, 11, iters);
      }
    }
  } else {
    // ---- epilogue ----
    const int e = warp - 2;                          // 0..19
    const int f = e >> 2, wq = e & 3;                // frame of the tile / warp within the frame's four
    const int qw = lane >> 3, l8 = lane & 7;
    const int q = warp & 3, hh = e >> 2;             // TMEM lane quarter / column share (warps 2..17)
    int jw[2], cm[2];
    int nvalid = 0;
    bool overflow = false;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      jw[i] = s_rowmap[(wq * 2 + i) * 4 + qw];
      cm[i] = __ldg(&p.tab->cm[wq * 2 + i]);         // warp-uniform step count of this slot
      if (jw[i] >= 0) {
        ++nvalid;
        overflow |= s_rowptr[jw[i] + 1] - s_rowptr[jw[i]] > kG3Steps;
      }
    }
    const int cmax = max(cm[0], cm[1]);
    overflow = __any_sync(0xffffffffu, overflow);    // rows with more entries than the table holds (dense A)
    // elements of the frame this warp accumulates (for the statistics merge)
    float cnt_w = (float)(nvalid * 4 * NP);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) cnt_w += __shfl_xor_sync(0xffffffffu, cnt_w, o);
    const float inv_cnt_w = cnt_w > 0.f ? 1.f / cnt_w : 0.f;
    const bool f_in_tile = f < p.FT;
    const char *yf = reinterpret_cast<const char *>(s_y) + (size_t)f * p.V * (YP * 4) + l8 * 16;
    const int2 *ep0 = s_ent + (wq * 2) * kG3Steps * 4 + qw;
    const float *bias_l = p.bias ? p.bias + l8 * 4 : nullptr;
    const float *nw_l = p.n_w + l8 * 4, *nb_l = p.n_b + l8 * 4;
    const int jo0 = max(jw[0], 0) * CO, jo1 = max(jw[1], 0) * CO;
    int buf = 0, t_ph = 0;
    for (int it = 0; it < iters; ++it) {
      const int item = 2 * (pair + it * npairs) + (int)rank;
      const bool valid = item < p.items;
      const int n = valid ? item / p.tiles_per_trial : 0;
      const int t = (valid ? (item - n * p.tiles_per_trial) * p.FT : 0) + f;
      const bool frame_ok = valid && f_in_tile && t < p.T;
      float4 z[NP][2];
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        { DbgTimer tm(dbg); mbar_wait(bTFull + 8 * buf, t_ph); tm.stop(d0); }
        tc_fence_after();
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const int ps = b * 2 + hp;
          // the bias rows of this pass: issued here, in flight across the staging below
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            z[ps][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (frame_ok && bias_l && jw[i] >= 0)
              z[ps][i] = __ldg(reinterpret_cast<const float4 *>(bias_l + (p.bias_v ? (i ? jo1 : jo0) : 0) + ps * 32));
          }
          { DbgTimer tm(dbg); named_bar_sync(1, kG3EpiThreads); tm.stop(d1); }   // staged rows of the previous pass are consumed
          DbgTimer tp1(dbg);
          if (e < kG3LdWarps) {
            // this warp's share of the pass: K segments of 8 accumulator columns of its 32 rows
            uint32_t rr[K][8];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + 32 * hp);
#pragma unroll
            for (int s = 0; s < K; ++s) {
              const int seg = hh * K + s;               // 0 .. 4K-1: partition seg / 4, columns (seg % 4) * 8 ..
              tmem_ld8_issue(taddr + (uint32_t)((seg >> 2) * 64 + (seg & 3) * 8), rr[s]);
            }
            tmem_ld_wait();
            if (hp == 1) {
              // accumulator buffer drained (the loads above have completed): hand it back to the
              // MMA issuer on the leader's barrier.  Relaxed: no generic-proxy data is published.
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_remote_relaxed(bTEmpty + 8 * buf, 0);
            }
            float *yrow = s_y + (q * 32 + lane) * YP;
#pragma unroll
            for (int s = 0; s < K; ++s) {
              const int seg = hh * K + s;
#pragma unroll
              for (int j = 0; j < 2; ++j)
                *reinterpret_cast<uint4 *>(yrow + seg * 8 + j * 4) =
                    make_uint4(rr[s][4 * j], rr[s][4 * j + 1], rr[s][4 * j + 2], rr[s][4 * j + 3]);
            }
          }
          tp1.stop(d2);
          { DbgTimer tm(dbg); named_bar_sync(1, kG3EpiThreads); tm.stop(d3); }   // staged rows complete
          DbgTimer tp2(dbg);
          if (frame_ok) {
            // adjacency contraction: step s of slot i reads the staged row of its s-th entry; the step
            // counts are warp-uniform (rows of one instruction have similar entry counts), padding
            // entries carry a = 0
            const int2 *ep = ep0;
            for (int s = 0; s < cmax; ++s, ep += 4) {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                if (s < cm[i]) {
                  const int2 en = ep[i * kG3Steps * 4];
                  const float a = __int_as_float(en.y);
                  const float4 y = *reinterpret_cast<const float4 *>(yf + en.x);
                  z[ps][i].x = fmaf(a, y.x, z[ps][i].x);
                  z[ps][i].y = fmaf(a, y.y, z[ps][i].y);
                  z[ps][i].z = fmaf(a, y.z, z[ps][i].z);
                  z[ps][i].w = fmaf(a, y.w, z[ps][i].w);
                }
            }
            if (overflow) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                if (jw[i] < 0) continue;
                const int e0 = s_rowptr[jw[i]], cn = s_rowptr[jw[i] + 1] - e0;
                for (int s = kG3Steps; s < cn; ++s) {
                  const int2 en = p.tab->ent[e0 + s];
                  const float a = __int_as_float(en.y);
                  const float4 y =
                      *reinterpret_cast<const float4 *>(yf + (en.x & 0xff) * (YP * 4) + (en.x >> 8) * 128);
                  z[ps][i].x = fmaf(a, y.x, z[ps][i].x);
                  z[ps][i].y = fmaf(a, y.y, z[ps][i].y);
                  z[ps][i].z = fmaf(a, y.z, z[ps][i].z);
                  z[ps][i].w = fmaf(a, y.w, z[ps][i].w);
                }
              }
            }
            // statistics about a warp-common shift (lane 0's first element), so that the warp merge
            // is a plain sum and the squares do not cancel
            if (ps == 0) shift = __shfl_sync(0xffffffffu, z[0][0].x, 0);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if (jw[i] >= 0) {
                const float d0 = z[ps][i].x - shift, d1 = z[ps][i].y - shift, d2 = z[ps][i].z - shift,
                            d3 = z[ps][i].w - shift;
                s1 += (d0 + d1) + (d2 + d3);
                s2 = fmaf(d0, d0, s2);
                s2 = fmaf(d1, d1, s2);
                s2 = fmaf(d2, d2, s2);
                s2 = fmaf(d3, d3, s2);
              }
          }
          tp2.stop(d4);
        }
        buf ^= 1;
        if (buf == 0) t_ph ^= 1;
      }
      if (!frame_ok) continue;
      DbgTimer tfs(dbg);
      // ---- frame statistics: warp (shuffled sums) -> the frame's four warps (shared memory, exact merge) ----
      // the first pass of the normalisation needs the LayerNorm affine of this thread's rows: load it now
      float4 g4[2], o4[2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (jw[i] >= 0) {
          g4[i] = __ldg(reinterpret_cast<const float4 *>(nw_l + (i ? jo1 : jo0)));
          o4[i] = __ldg(reinterpret_cast<const float4 *>(nb_l + (i ? jo1 : jo0)));
        }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      float *st = s_stat + (f * 4) * 4;
      if (lane == 0) {
        st[wq * 4] = cnt_w;
        st[wq * 4 + 1] = fmaf(s1, inv_cnt_w, shift);
        st[wq * 4 + 2] = fmaxf(s2 - s1 * s1 * inv_cnt_w, 0.f);
      }
      named_bar_sync(2 + f, 128);
      // merge the four partials in a fixed order so that all four warps get bit-identical statistics
      float cnt_t = st[0], mean = st[1], M2 = st[2];
#pragma unroll
      for (int j = 1; j < 4; ++j) stat_merge(cnt_t, mean, M2, st[j * 4], st[j * 4 + 1], st[j * 4 + 2]);
      const float rstd = 1.f / sqrtf(M2 * (1.f / (float)(p.V * CO - 1)) + p.eps);
      const float nmr = -mean * rstd;
      tfs.stop(d5);
      DbgTimer tfn(dbg);
      const long long frow = (long long)n * p.T + t;
      const long long frow_o = p.out_T ? (long long)n * p.out_T + t + p.out_t0 : frow;
      // row bases of this thread's two rows (element offsets of channel l8*4)
      const long long ob0 = ((p.out_f32 ? frow : frow_o) * p.V) * CO + jo0 + l8 * 4;
      const long long ob1 = ((p.out_f32 ? frow : frow_o) * p.V) * CO + jo1 + l8 * 4;
#pragma unroll
      for (int ps = 0; ps < NP; ++ps) {
        if (ps > 0) {
#pragma unroll
          for (int i = 0; i < 2; ++i)
            if (jw[i] >= 0) {
              g4[i] = __ldg(reinterpret_cast<const float4 *>(nw_l + (i ? jo1 : jo0) + ps * 32));
              o4[i] = __ldg(reinterpret_cast<const float4 *>(nb_l + (i ? jo1 : jo0) + ps * 32));
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (jw[i] < 0) continue;
          float4 v = z[ps][i];
          v.x = fmaf(fmaf(v.x, rstd, nmr), g4[i].x, o4[i].x);
          v.y = fmaf(fmaf(v.y, rstd, nmr), g4[i].y, o4[i].y);
          v.z = fmaf(fmaf(v.z, rstd, nmr), g4[i].z, o4[i].z);
          v.w = fmaf(fmaf(v.w, rstd, nmr), g4[i].w, o4[i].w);
          if (p.relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          }
          const long long o = (i ? ob1 : ob0) + ps * 32;
          if (p.out_f32) {
            st_stream(reinterpret_cast<float4 *>(p.out_f32 + o), v);
          } else {
            const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
            st_stream(reinterpret_cast<uint2 *>(p.out_hi + o),
                      make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23)));
            if (p.out_lo) {
              const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
              const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
              const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
              st_stream(reinterpret_cast<uint2 *>(p.out_lo + o),
                        make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23)));
            }
          }
        }
      }
      tfn.stop(d6);
    }
    if (e == 0 && lane == 0) {
      dbg_flush(dbg, 6, d0); dbg_flush(dbg, 7, d1); dbg_flush(dbg, 8, d2); dbg_flush(dbg, 9, d3);
      dbg_flush(dbg, 10, d4); dbg_flush(dbg, 12, d5); dbg_flush(dbg, 13, d6);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                 // neither CTA may leave while the pair's MMAs / arrives are in flight
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

inline bool gcn3_supported(int c_in, int c_out, int V, int K) {
  return (c_out == 64 || c_out == 128 || c_out == 256) && c_in % 64 == 0 && c_in >= 64 && c_in <= 256 &&
         V >= 22 && V <= kG3MaxV && (K == 1 || K == 3);
}
// Opt-in (STGCN_GCN3=1).  Measured on B200 (profiles/r01_gcn3_*): 851 us vs 1090 us (v2) per 3.2 M
// rows at C = 64 but no gain at C = 128 and a loss at C = 256 (register spills), and the temporal
// kernel pays ~13 % for writing / reading bf16 planes instead of fp32 rows -- a net loss over the
// whole trunk, so v2 stays the default.  DESIGN.md section 4 has the analysis.
inline bool gcn3_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_GCN3");
    on = e ? atoi(e) != 0 : 0;
  }
  return on != 0;
}

// x planes: bf16 [planes][N][T_full][V][Cin]; the view takes every `fstride`-th frame (T frames);
// wb: k_pack_gcn3_w tiles.
template <int CO, int K>
int launch_gcn3_ck(const __nv_bfloat16 *x, const __nv_bfloat16 *wb, Gcn3Params p, int N, int T_full, int fstride,
                   cudaStream_t st) {
  const int V = p.V, kMaxSmem = 232448;
  constexpr int BN = K * 64, kWHalf = (BN / 2) * 128, YP = K * 32 + 4, NB = CO / 64;
  p.FT = 128 / V;
  if (p.FT > 5) p.FT = 5;
  p.tiles_per_trial = (p.T + p.FT - 1) / p.FT;
  p.items = N * p.tiles_per_trial;
  const int fixed = kG3R * kG3Chunk + 128 * YP * 4 + kG3MaxEnt * 8 + 768 + 16 * kG3R + 32 + 64 + 1024;
  int S = (kMaxSmem - fixed) / (kWHalf + 16);
  if (S > 8) S = 8;
  if (S < 2) return fail("gcn3: shared memory does not fit");
  p.S = S;
  const int smem = fixed + S * (kWHalf + 16);
  const int KC = p.Cin / 64;
  CUtensorMap tm_x, tm_w;
  const uint64_t xd[5] = {(uint64_t)p.Cin, (uint64_t)V, (uint64_t)p.T, (uint64_t)N, (uint64_t)p.planes};
  const uint64_t xs[4] = {(uint64_t)p.Cin * 2, (uint64_t)fstride * V * p.Cin * 2, (uint64_t)T_full * V * p.Cin * 2,
                          (uint64_t)N * T_full * V * p.Cin * 2};
  const uint32_t xb[5] = {64, (uint32_t)V, (uint32_t)p.FT, 1, 1};
  if (make_tmap_bf16(&tm_x, x, 5, xd, xs, xb)) return 1;
  const uint64_t wd[3] = {64, (uint64_t)BN, (uint64_t)(2 * NB * KC)};
  const uint64_t wst[2] = {128, (uint64_t)BN * 128};
  const uint32_t wbx[3] = {64, (uint32_t)(BN / 2), 1};
  if (make_tmap_bf16(&tm_w, wb, 3, wd, wst, wbx)) return 1;
  int pairs = (p.items + 1) / 2;
  if (pairs > num_sms() / 2) pairs = num_sms() / 2;
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_gcn3<CO, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_gcn3<CO, K><<<2 * pairs, kG3Threads, smem, st>>>(tm_x, tm_w, p);
  return 0;
}

inline int launch_gcn3(int CO, int K, const __nv_bfloat16 *x, const __nv_bfloat16 *wb, const Gcn3Params &p, int N,
                       int T_full, int fstride, cudaStream_t st) {
  if (K == 3) {
    switch (CO) {
      case 64: return launch_gcn3_ck<64, 3>(x, wb, p, N, T_full, fstride, st);
      case 128: return launch_gcn3_ck<128, 3>(x, wb, p, N, T_full, fstride, st);
      case 256: return launch_gcn3_ck<256, 3>(x, wb, p, N, T_full, fstride, st);
    }
  } else if (K == 1) {
    switch (CO) {
      case 64: return launch_gcn3_ck<64, 1>(x, wb, p, N, T_full, fstride, st);
      case 128: return launch_gcn3_ck<128, 1>(x, wb, p, N, T_full, fstride, st);
      case 256: return launch_gcn3_ck<256, 1>(x, wb, p, N, T_full, fstride, st);
    }
  }
  return fail("gcn3: unsupported shape (C_out %d, K %d)", CO, K);
}

}  // namespace tc
}  // namespace stgcn
