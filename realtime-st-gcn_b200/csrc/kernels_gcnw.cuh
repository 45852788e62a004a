// Graph convolution as a GEMM with per-joint, pre-scaled weights (default for model-level forwards whose
// adjacency is known to be sparse, stgcn_model_desc.reserved bit 1; STGCN_GCNW=0 disables it).
//
//   z[(n,t,w), c] = bz[w,c] + sum_{edges e = (k,v) into w} x[(n,t,v), :] . (A[k,v,w] * W_k[c, :])
//
// The adjacency weight of an edge is folded into the WEIGHTS (one pre-scaled, bf16-split copy of W_k
// per edge, built once per parameter set), and a tile is 128 consecutive frames of ONE output joint
// w: its A operands are plain strided TMA boxes of the bf16-plane input (joint v_e, 128 frames), its
// B operands the edge's weight tiles.  No CUDA-core arithmetic before the MMA, no gather after it; the
// epilogue only adds the bias and stores z (fp32).
//
// LayerNorm(C,V) needs all V joints of a frame, i.e. V different tiles.  Default: z goes to HBM and the
// streaming kernel k_ln_stream (one block per frame) normalises it.  Opt-in (STGCN_GCNW_FUSE=1): the stage as
// ONE persistent kernel with two halves connected through L2.  The GEMM epilogue writes its z tile into
// a ring of frame-group slots (a group = the V tiles of 128 frames; a few tens of MB, so the lines stay
// dirty in the 126 MB L2 and are overwritten there) together with each row's partial statistics
// (mean, M2 over its channels), and counts the tile in ready[group].  Eight more warps per CTA (the
// "LN warps") take the groups round-robin: wait for ready[group] == all tiles, merge the row partials
// of every frame (Chan et al., exact), then stream the slot once -- LayerNorm affine, ReLU, bf16 split
// -- into the temporal kernel's operand planes and release the slot (done[group]).  z never travels
// to HBM and back, and the HBM-bound normalisation overlaps the MMA-bound GEMM on the same SMs.
// Producers only ever wait for consumers of EARLIER groups and consumers never wait for producers of
// later ones, so the schedule cannot deadlock as long as all CTAs are resident (cooperative launch).
// Measured on B200 (32 trials x T = 4000, parity mode): 12.3 ms for the stage (64 MB ring, 4-step LN items;
// 17-48 ms with smaller rings / larger items) against 9.9 ms for the two kernels -- z does stay in L2 (with
// the L2 eviction-priority hints below: 1.64 GB of DRAM traffic per C = 64 launch instead of 3.2 GB), but
// the ring couples the progress of all CTAs (throughput = ring slots / loop latency, ~55 us measured with
// statically assigned tiles) -- so it is not the default.
// Tree-structured adjacency only: the weight buffer holds 6*V edges.
#pragma once
#include "kernels_tc.cuh"

namespace stgcn {
namespace tc {

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

// LayerNorm affine (C, 1, V) -> [V][C]
__global__ void k_transpose_cv(const float *__restrict__ in, float *__restrict__ out, int C, int V) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * V) return;
  const int v = i / C, c = i - v * C;
  out[i] = in[c * V + v];
}

constexpr int kGwMaxV = 32;
constexpr int kGwEdgeCap = 6 * kGwMaxV;
inline int gcnw_edge_cap(int V) { return 6 * V; }

struct GcnwTables {
  int ptr[kGwMaxV + 1];       // edges into joint w: [ptr[w], ptr[w+1])
  int nedge;                  // -1: the adjacency has more edges than the weight buffer holds
  int pad[2];
  int src[kGwEdgeCap];        // source joint v
  int kk[kGwEdgeCap];         // partition k
  float val[kGwEdgeCap];      // A[k,v,w]
};

__global__ void k_gcnw_tables(const float *__restrict__ A, int K, int V, int identity, int cap, GcnwTables *tab) {
  __shared__ int cnt[kGwMaxV + 1];
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
      tab->ptr[i] = run;
      run += cnt[i];
    }
    tab->ptr[V] = run;
    tab->nedge = run <= cap ? run : -1;
  }
  __syncthreads();
  if (w < V && tab->nedge >= 0) {
    int at = tab->ptr[w];
    if (identity) {
      tab->src[at] = w; tab->kk[at] = 0; tab->val[at] = 1.f;
    } else {
      for (int k = 0; k < K; ++k)
        for (int v = 0; v < V; ++v) {
          const float a = A[((long long)k * V + v) * V + w];
          if (a != 0.f) {
            tab->src[at] = v; tab->kk[at] = k; tab->val[at] = a;
            ++at;
          }
        }
    }
  }
}

// wsc[plane][edge][c_out][c_in] = split_bf16(A_e * W[k_e*CO + c][ci]); plane stride = cap*CO*Cin
__global__ void k_gcnw_pack(const float *__restrict__ w, const GcnwTables *__restrict__ tab, int CO, int Cin, int cap,
                            __nv_bfloat16 *__restrict__ wsc) {
  const long long per = (long long)CO * Cin;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ne = tab->nedge;
  if (ne < 0 || i >= per * ne) return;
  const int e = (int)(i / per);
  const long long r = i - (long long)e * per;
  __nv_bfloat16 hi, lo;
  split_bf16(tab->val[e] * w[(long long)tab->kk[e] * per + r], hi, lo);
  wsc[i] = hi;
  wsc[(long long)cap * per + i] = lo;
}

struct GcnwParams {
  int T, V, Cin, planes, N;     // frames per trial of this view, joints, input channels, bf16 planes, trials
  int tblocks, items;           // 128-frame blocks per trial; items = N * tblocks * V (w fastest)
  int a_stages, b_stages;
  const GcnwTables *tab;
  const float *bias;            // bias_sw = 1: [C/4][V][4] table, 0: [C] vector, null: none
  int bias_sw;
  float *out;                   // z fp32 rows [(n*T + t)*V + w][CO]
  int tma_out;                  // rows leave through TMA stores (tm_z) instead of st.global from the patch
  int debug;
  // tap mode (CoST-GCN temporal convolution over a ring of frames, ntaps > 0): V == 1, the "edges" of the
  // single joint are the taps, "source joint" of tap j = ring slot tap_src[j]; tab is not used
  int ntaps;
  int tap_src[16];
  // temporal mode (tmode == 2: the Gamma x 1 convolution of an ST-GCN layer for ANY Gamma): the "edges" of every
  // joint are the ntaps taps, source joint = the joint itself, tap j reads input frame stride*tau + j - tpad
  // (frames outside the trial are zero-filled by TMA = the convolution's zero padding); stride 2 reads the
  // even / odd frame sequences through two tensor maps
  int tmode, tpad, tstride;
  int smem_cap;               // 0 = all shared memory; else the launcher sizes its stage rings to this many bytes
                              // (a CTA that leaves room for another kernel's CTAs on the same SM)
  // fused LayerNorm stage (k_gcnw<.., FUSE = true>)
  float *zring;                 // [R][128 frames][V][CO] fp32
  float2 *sring;                // [R][128][V][kEpiNH] (mean, M2) of every (frame, joint, column group)
  unsigned *ready, *done;       // [N * tblocks] tile-arrival count / slot-released flag per frame group
  int R;                        // ring slots
  int npc;                      // position chunks per frame block (LN items per group = kGwLnBlocks * npc)
  const float *n_wV, *n_bV;     // LayerNorm affine as [V][C] (the order in which the LN warps stream a frame)
  int relu;
  float eps;
  float *out_f32;               // normalised output: fp32 rows, or
  __nv_bfloat16 *out_hi, *out_lo;   // bf16 hi / lo planes (frame t of trial n at frame t + out_t0 of out_T)
  int out_T, out_t0;
};

// ---- inter-CTA flags of the fused stage -------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// L2 eviction priorities of the fused stage: the z ring must stay resident (written, read once ~10 us later,
// overwritten in place a few hundred us later) while ~4x as many bytes stream through L2 once (x in, u out)
// TMA store of a [32 rows][128 B] SWIZZLE_128B shared-memory box into a 4-D fp32 tensor (bulk async-group)
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_hint16(void *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_hint8(void *p, uint2 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v2.u32 [%0], {%1, %2}, %3;" ::"l"(p), "r"(v.x), "r"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_hint16(const float4 *p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
// bounded spin (a protocol bug must trap, never hang the GPU)
template <bool ACQ>
__device__ __forceinline__ void wait_flag(const unsigned *p, unsigned target, unsigned have) {
  long long t0 = 0;
  unsigned spins = 0;
  while (have < target) {
    __nanosleep(200);
    have = ACQ ? ld_acquire_gpu(p) : ld_relaxed_gpu(p);
    if ((++spins & 255u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ll) {
        printf("stgcn_b200: gcnw flag timeout (block %d thread %d have %u want %u)\n", blockIdx.x, threadIdx.x, have,
               target);
        __trap();
      }
    }
  }
}
constexpr int kGwLnThreads = 256;     // LN warps of the fused stage
constexpr int kGwLnFrames = 8;        // frames per LN work item (one warp)
constexpr int kGwLnBlocks = 128 / kGwLnFrames;  // frame blocks per group; x npc position chunks = LN items per group
constexpr int kGwLnIters = 4;         // at most this many 32-position steps per item (bounds an item's latency, and
                                      // with it the number of groups the ring must hold)
constexpr int kGwPubRing = 4;         // tiles the epilogue may run ahead of the publisher warp

constexpr int kGwThreads = 32 * (3 + 4 * kEpiNH);   // 0 A producer, 1 MMA, 2 B producer, 3.. epilogue

// MERGE (C_out <= 128): one activation stage holds both bf16 planes of an (edge, 64-channel chunk)
// and one weight stage both weight planes, so the MMA issuer waits on two barriers per twelve MMAs
// instead of five per twelve, and the weight hi plane is loaded once instead of twice.
template <int CO, bool MERGE, bool FUSE>
__global__ void __launch_bounds__(kGwThreads + (FUSE ? 32 + kGwLnThreads : 0), FUSE ? 1 : 2)
    k_gcnw(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_x1,
           const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_z, const GcnwParams p) {
  constexpr int kAPlane = 128 * 128;            // [128 frames][64 ch] bf16
  constexpr int kBPlane = CO * 128;             // [CO][64 ch] bf16
  constexpr int kABytes = MERGE ? 2 * kAPlane : kAPlane;
  constexpr int kBBytes = MERGE ? 2 * kBPlane : kBPlane;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const int SA = p.a_stages, SB = p.b_stages;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + SA * kABytes;
  const uint32_t sPatch = sB + SB * kBBytes;
  const uint32_t sTab = sPatch + kPatchTotal;                    // ptr[33] + src[192]
  const uint32_t sPub = sTab + 1024;                             // FUSE: publisher barriers [kGwPubRing] + published-tile count
  const uint32_t sBar = sPub + 1024;
  const uint32_t bTmemFull = sBar, bTmemEmpty = sBar + 16;
  const uint32_t bFullA = sBar + 32, bEmptyA = bFullA + 8 * SA;
  const uint32_t bFullB = bEmptyA + 8 * SA, bEmptyB = bFullB + 8 * SB;
  const uint32_t sTmemPtr = bEmptyB + 8 * SB;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(gen_base + (sTmemPtr - smem_base));
  uint8_t *s_patch = gen_base + (sPatch - smem_base);
  int *s_ptr = reinterpret_cast<int *>(gen_base + (sTab - smem_base));
  int *s_src = s_ptr + 40;
  volatile unsigned *s_pubcount = reinterpret_cast<volatile unsigned *>(gen_base + (sPub - smem_base) + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = p.Cin / 64;
  if (p.tmode == 0 && __ldg(&p.tab->nedge) < 0) {
    // the caller vouched for a sparse adjacency (stgcn_model_desc.reserved bit 1) that is not sparse:
    // fail loudly instead of computing with a truncated edge list
    if (threadIdx.x == 0 && blockIdx.x == 0)
      printf("stgcn_b200: adjacency has more than 6*V non-zeros but the model descriptor says it is sparse\n");
    __trap();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_x1);
    tma_prefetch_desc(&tm_w);
    if (p.tma_out) tma_prefetch_desc(&tm_z);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bTmemFull + 8 * i, 1);
      mbar_init(bTmemEmpty + 8 * i, 4 * kEpiNH);
    }
    for (int i = 0; i < SA; ++i) {
      mbar_init(bFullA + 8 * i, 1);
      mbar_init(bEmptyA + 8 * i, 1);
    }
    for (int i = 0; i < SB; ++i) {
      mbar_init(bFullB + 8 * i, 1);
      mbar_init(bEmptyB + 8 * i, 1);
    }
    if (FUSE) {
      for (int i = 0; i < kGwPubRing; ++i) mbar_init(sPub + 8 * i, 4 * kEpiNH);
      *s_pubcount = 0u;
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(sTmemPtr, 512);
  if (p.tmode == 1) {
    if (threadIdx.x == 0) { s_ptr[0] = 0; s_ptr[1] = p.ntaps; }
    if (threadIdx.x < 16) s_src[threadIdx.x] = p.tap_src[threadIdx.x];
  } else if (p.tmode == 0) {
    for (int i = threadIdx.x; i <= p.V; i += blockDim.x) s_ptr[i] = __ldg(&p.tab->ptr[i]);
    for (int i = threadIdx.x; i < kGwEdgeCap; i += blockDim.x) s_src[i] = __ldg(&p.tab->src[i]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ---- A producer: for every edge of the item's joint, the source joint's 128 frames, per 64-channel
    // chunk and plane ----
    if (lane == 0) {
      // the activations are the only operand another kernel of the chain writes (weights, tables and bias are
      // prepared once): everything else of this CTA -- weight stages included -- is already in flight
      griddep_wait();
      int as = 0, a_ph = 0;
      const bool tap = p.tmode == 1;     // ring-tap mode: tensor-map dimension 1 = rows, 2 = ring slots
      const bool tmp = p.tmode == 2;     // temporal mode: same joint, frame offset per tap, even / odd maps
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int w = item % p.V;
        const int tb = (item / p.V) % p.tblocks, n = item / (p.V * p.tblocks);
        const int e0 = tmp ? 0 : s_ptr[w], e1 = tmp ? p.ntaps : s_ptr[w + 1];
        for (int e = e0; e < e1; ++e) {
          // coordinates of the edge's / tap's activation box (dimension 1, dimension 2) and its tensor map
          int c1 = tap ? tb * 128 : s_src[e], c2 = tap ? s_src[e] : tb * 128;
          const CUtensorMap *tm = &tm_x;
          if (tmp) {
            const int d = e - p.tpad;
            c1 = w;
            if (p.tstride == 2) {
              c2 = tb * 128 + (d >> 1);               // position in the even / odd frame sequence
              tm = (d & 1) ? &tm_x1 : &tm_x;
            } else {
              c2 = tb * 128 + d;
            }
          }
          for (int kc = 0; kc < KC; ++kc) {
            if (MERGE) {
              mbar_wait(bEmptyA + 8 * as, a_ph ^ 1);
              mbar_expect_tx(bFullA + 8 * as, (uint32_t)(p.planes * kAPlane));
              for (int ap = 0; ap < p.planes; ++ap)
                tma_load_5d(sA + as * kABytes + ap * kAPlane, tm, bFullA + 8 * as, kc * 64, c1, c2, n, ap);
              if (++as == SA) { as = 0; a_ph ^= 1; }
            } else {
              for (int ap = 0; ap < p.planes; ++ap) {
                mbar_wait(bEmptyA + 8 * as, a_ph ^ 1);
                mbar_expect_tx(bFullA + 8 * as, kAPlane);
                tma_load_5d(sA + as * kABytes, tm, bFullA + 8 * as, kc * 64, c1, c2, n, ap);
                if (++as == SA) { as = 0; a_ph ^= 1; }
              }
            }
          }
        }
      }
      // this CTA's last activation loads are issued: what remains is the tail of the pipeline, under which the
      // next kernel's CTAs may become resident (triggering at the start keeps them resident -- and, measured,
      // slowing this kernel -- for its whole duration)
      griddep_launch();
    }
  } else if (warp == 2) {
    // ---- B producer: the edge's pre-scaled weight tiles (hi, lo for an A hi stage; hi for an A lo stage) ----
    if (lane == 0) {
      int bs = 0, b_ph = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int w = item % p.V;
        const int e0 = p.tmode == 2 ? 0 : s_ptr[w], e1 = p.tmode == 2 ? p.ntaps : s_ptr[w + 1];
        for (int e = e0; e < e1; ++e)
          for (int kc = 0; kc < KC; ++kc) {
            if (MERGE) {
              mbar_wait(bEmptyB + 8 * bs, b_ph ^ 1);
              mbar_expect_tx(bFullB + 8 * bs, (uint32_t)(p.planes * kBPlane));
              for (int bp = 0; bp < p.planes; ++bp)
                tma_load_4d(sB + bs * kBBytes + bp * kBPlane, &tm_w, bFullB + 8 * bs, kc * 64, 0, e, bp);
              if (++bs == SB) { bs = 0; b_ph ^= 1; }
            } else {
              for (int ap = 0; ap < p.planes; ++ap) {
                const int nb = (ap == 0) ? p.planes : 1;
                for (int bp = 0; bp < nb; ++bp) {
                  mbar_wait(bEmptyB + 8 * bs, b_ph ^ 1);
                  mbar_expect_tx(bFullB + 8 * bs, kBPlane);
                  tma_load_4d(sB + bs * kBBytes, &tm_w, bFullB + 8 * bs, kc * 64, 0, e, bp);
                  if (++bs == SB) { bs = 0; b_ph ^= 1; }
                }
              }
            }
          }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ----
    constexpr uint32_t idesc = umma_idesc_bf16(128, CO);
    int a_s = 0, a_ph = 0, b_s = 0, b_ph = 0, buf = 0, t_ph = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int w = item % p.V;
      mbar_wait(bTmemEmpty + 8 * buf, t_ph ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * CO;
      uint32_t acc = 0;
      const int e0 = p.tmode == 2 ? 0 : s_ptr[w], e1 = p.tmode == 2 ? p.ntaps : s_ptr[w + 1];
      for (int e = e0; e < e1; ++e)
        for (int kc = 0; kc < KC; ++kc) {
          if (MERGE) {
            mbar_wait(bFullA + 8 * a_s, a_ph);
            mbar_wait(bFullB + 8 * b_s, b_ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_hi = umma_desc_lo(sA + a_s * kABytes), a_lo = umma_desc_lo(sA + a_s * kABytes + kAPlane);
              const uint32_t b_hi = umma_desc_lo(sB + b_s * kBBytes), b_lo = umma_desc_lo(sB + b_s * kBBytes + kBPlane);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tacc, umma_desc_join(a_hi + 2 * k), umma_desc_join(b_hi + 2 * k), idesc, acc | (uint32_t)k);
              if (p.planes == 2) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tacc, umma_desc_join(a_hi + 2 * k), umma_desc_join(b_lo + 2 * k), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tacc, umma_desc_join(a_lo + 2 * k), umma_desc_join(b_hi + 2 * k), idesc, 1u);
              }
              umma_commit(bEmptyB + 8 * b_s);
              umma_commit(bEmptyA + 8 * a_s);
            }
            __syncwarp();
            acc = 1;
            if (++b_s == SB) { b_s = 0; b_ph ^= 1; }
            if (++a_s == SA) { a_s = 0; a_ph ^= 1; }
            continue;
          }
          for (int ap = 0; ap < p.planes; ++ap) {
            mbar_wait(bFullA + 8 * a_s, a_ph);
            tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(sA + a_s * kABytes);
            const int nb = (ap == 0) ? p.planes : 1;
            for (int bp = 0; bp < nb; ++bp) {
              mbar_wait(bFullB + 8 * b_s, b_ph);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_lo = umma_desc_lo(sB + b_s * kBBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tacc, umma_desc_join(a_lo + 2 * k), umma_desc_join(b_lo + 2 * k), idesc, acc | (uint32_t)k);
                umma_commit(bEmptyB + 8 * b_s);
              }
              __syncwarp();
              acc = 1;
              if (++b_s == SB) { b_s = 0; b_ph ^= 1; }
            }
            if (elect_one()) umma_commit(bEmptyA + 8 * a_s);
            __syncwarp();
            if (++a_s == SA) { a_s = 0; a_ph ^= 1; }
          }
        }
      if (elect_one()) umma_commit(bTmemFull + 8 * buf);
      __syncwarp();
      buf ^= 1;
      if (buf == 0) t_ph ^= 1;
    }
  } else if (warp < 3 + 4 * kEpiNH) {
    // ---- epilogue: z = acc + bias, rows (frames) leave through the warp's patch as whole 128-B lines ----
    const int q = warp & 3;
    const int h = (warp - 3) >> 2;
    constexpr int CH = CO / kEpiNH;
    const int c0 = h * CH;
    uint8_t *patch = s_patch + (warp - 3) * kPatchBytes;
    uint8_t *mine = patch + lane * kPatchPitch;
    uint8_t *tpatch = s_patch + (warp - 3) * 4096;           // TMA-store form: [32 rows][128 B], 1024-B aligned
    int buf = 0, t_ph = 0, tile_k = 0;
    const uint64_t pol_keep = FUSE ? l2_policy_evict_last() : 0ull;
    float v[16];
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int w = item % p.V;
      const int grp = item / p.V;                                  // frame group (n, tb)
      const int tb = grp % p.tblocks, n = grp / p.tblocks;
      const bool has_edges = p.tmode == 2 || s_ptr[w + 1] > s_ptr[w];
      const int t = tb * 128 + q * 32 + lane;
      const bool row_ok = t < p.T;
      const uint32_t okmask = __ballot_sync(0xffffffffu, row_ok);
      // FUSE: rows go to the group's ring slot [128 frames][V][CO]; otherwise to the z tensor
      const int slot = FUSE ? grp % p.R : 0;
      float *zout = FUSE ? p.zring + (size_t)slot * (128 * p.V) * CO : p.out;
      const long long row0 = FUSE ? (long long)(q * 32) * p.V + w
                                  : ((long long)n * p.T + tb * 128 + q * 32) * p.V + w;   // row of lane 0; lane rr: + rr*V
      const int pstep = p.bias_sw ? p.V : 1;
      const float4 *bias4 =
          p.bias ? reinterpret_cast<const float4 *>(p.bias) + (c0 >> 2) * pstep + (p.bias_sw ? w : 0) : nullptr;
      // the slot must have been released by the LN warps that read its previous group; the flag load is
      // issued before the accumulator wait so that its L2 latency is hidden.  Relaxed on purpose: only
      // stores follow (they cannot be speculated), and an acquire would invalidate the L1 that holds
      // the bias tables once per tile.
      const unsigned ln_per_group = (unsigned)(kGwLnBlocks * p.npc);
      unsigned freed = ln_per_group;
      if (FUSE && grp >= p.R && lane == 0) freed = ld_relaxed_gpu(p.done + (grp - p.R));
      mbar_wait(bTmemFull + 8 * buf, t_ph);
      tc_fence_after();
      if (FUSE) {
        if (grp >= p.R && lane == 0) wait_flag<false>(p.done + (grp - p.R), ln_per_group, freed);
        __syncwarp();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * CO);
      float shift = 0.f, s1 = 0.f, s2 = 0.f;          // FUSE: statistics of this row's CH columns about their first value
#pragma unroll 1
      for (int sb = 0; sb < CH; sb += 32) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int cb = sb + half * 16;
          tmem_ld16(taddr + c0 + cb, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias4) b4 = __ldg(bias4 + ((cb >> 2) + i) * pstep);
            float4 o = has_edges ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (FUSE) {
              if (cb == 0 && i == 0) shift = o.x;
              const float d0 = o.x - shift, d1 = o.y - shift, d2 = o.z - shift, d3 = o.w - shift;
              s1 += (d0 + d1) + (d2 + d3);
              s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
            }
            if (!FUSE && p.tma_out)      // SWIZZLE_128B box row: 16-B chunk j of row r sits at chunk j ^ (r & 7)
              *reinterpret_cast<float4 *>(tpatch + lane * 128 + (((half * 4 + i) ^ (lane & 7)) << 4)) = o;
            else
              *reinterpret_cast<float4 *>(mine + half * 64 + i * 16) = o;
          }
        }
        if (!FUSE && p.tma_out) {
          // the box [32 frames][32 channels] of this joint goes out as ONE TMA store (rows past the end of the
          // trial are clipped by the tensor map); the patch is reused once the store has read it
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tm_z, smem_u32(tpatch), c0 + sb, w, tb * 128 + q * 32, n);
            bulk_commit();
            bulk_wait_read0();
          }
          __syncwarp();
          continue;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int pc = i * 32 + lane, rr = pc >> 3, qq = pc & 7;
          if ((okmask >> rr) & 1) {
            float *dst = zout + (row0 + (long long)rr * p.V) * CO + c0 + sb + qq * 4;
            const float4 val = *reinterpret_cast<const float4 *>(patch + rr * kPatchPitch + qq * 16);
            if (FUSE) st_hint16(dst, val, pol_keep);
            else *reinterpret_cast<float4 *>(dst) = val;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bTmemEmpty + 8 * buf);
      if (FUSE) {
        // row partial (mean, M2); then hand the tile to the publisher warp through a CTA-scope mbarrier: the
        // gpu-scope release (a fence that waits for this SM's outstanding stores) must not sit in the
        // epilogue's critical path
        if (row_ok) {
          const float m_r = shift + s1 * (1.f / (float)CH);
          const float M2_r = fmaxf(s2 - s1 * s1 * (1.f / (float)CH), 0.f);
          p.sring[(((size_t)slot * 128 + q * 32 + lane) * p.V + w) * kEpiNH + h] = make_float2(m_r, M2_r);
        }
        __syncwarp();
        if (lane == 0) {
          while (*s_pubcount + (unsigned)kGwPubRing <= (unsigned)tile_k) __nanosleep(64);
          mbar_arrive(sPub + 8 * (tile_k % kGwPubRing));
        }
        ++tile_k;
      }
      buf ^= 1;
      if (buf == 0) t_ph ^= 1;
    }
    if (!FUSE && p.tma_out && lane == 0) bulk_wait0();        // all stores of this warp have left shared memory and landed
  } else if (FUSE && warp == 3 + 4 * kEpiNH) {
    // ---- publisher: tiles whose eight epilogue warps have arrived become visible to the other CTAs ----
    // (bar.sync-style cumulativity: the epilogue's stores happen-before its CTA-scope arrive, the fence
    // below publishes them at gpu scope; several finished tiles share one fence)
    if (lane == 0) {
      const int ntiles = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      int k = 0;
      while (k < ntiles) {
        mbar_wait<32>(sPub + 8 * (k % kGwPubRing), (uint32_t)(k / kGwPubRing) & 1u);
        int k2 = k + 1;
        while (k2 < ntiles && k2 < k + kGwPubRing - 1 &&
               mbar_try_wait(sPub + 8 * (k2 % kGwPubRing), (uint32_t)(k2 / kGwPubRing) & 1u))
          ++k2;
        if (!(p.debug & 1024)) __threadfence();             // (debug bit: timing without the release fence)
        for (int j = k; j < k2; ++j)
          atomicAdd(p.ready + (blockIdx.x + j * gridDim.x) / p.V, (unsigned)(4 * kEpiNH));
        *s_pubcount = (unsigned)k2;
        k = k2;
      }
    }
  } else if (FUSE && warp > 3 + 4 * kEpiNH) {
    // ---- LN warps: normalise complete frame groups out of their ring slots ----
    // Work item = kGwLnFrames consecutive frames of a group, one WARP per item, items handed out in order by
    // a global ticket counter (work conserving: only the few groups in flight have work).  No block barrier
    // and no fence: a warp polls ready[group] with relaxed loads and acquires once; it has consumed all its
    // loads of the slot before it counts the item in done[group].
    constexpr int C4 = CO / 4, kSh = (CO == 64) ? 4 : (CO == 128 ? 5 : 6);
    const int V = p.V, NP = V * kEpiNH;                     // row partials per frame
    const int groups = p.N * p.tblocks;
    const unsigned per_group = (unsigned)(kGwLnBlocks * p.npc);
    const unsigned lnitems = (unsigned)groups * per_group;
    const int pc_len = ((V * C4 + 31) / 32 + p.npc - 1) / p.npc * 32;   // positions per chunk (multiple of 32)
    unsigned *ticket = p.ready + 2 * (size_t)groups;        // [ready | done | ticket]
    const unsigned target = (unsigned)(V * 4 * kEpiNH);     // epilogue warps per group
    const uint32_t vmagic = (uint32_t)((0x100000000ull + (unsigned)V - 1) / (unsigned)V);   // x / V for x < 65536
    const float inv_np = 1.f / (float)NP, inv_cv = 1.f / (float)(V * CO - 1);
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    unsigned li = 0;
    if (lane == 0) li = atomicAdd(ticket, 1u);
    li = __shfl_sync(0xffffffffu, li, 0);
    while (li < lnitems) {
      const int grp = (int)(li / per_group), sub = (int)(li % per_group);
      const int fb = sub / p.npc, pc = sub - fb * p.npc;          // frame block, position chunk
      const int slot = grp % p.R;
      const int tb = grp % p.tblocks, n = grp / p.tblocks;
      const int f0 = fb * kGwLnFrames;
      int nf = p.T - tb * 128 - f0;                          // valid frames of this item
      nf = nf < 0 ? 0 : (nf > kGwLnFrames ? kGwLnFrames : nf);
      unsigned nli = 0;
      if (lane == 0) {
        nli = atomicAdd(ticket, 1u);                         // next ticket: its latency hides behind this item
        wait_flag<false>(p.ready + grp, target, ld_relaxed_gpu(p.ready + grp));
        __threadfence();                                     // relaxed load + fence = acquire
      }
      __syncwarp();
      if (p.debug & 4096) nf = 0;                            // (debug bit: LN warps only hand the slots back)
      // statistics of the item's frames (merge of the V * kEpiNH row partials, Chan et al.)
      float fm[kGwLnFrames], fr[kGwLnFrames];
      const float2 *sp = p.sring + ((size_t)slot * 128 + f0) * NP;
      float2 a0[kGwLnFrames], a1[kGwLnFrames];
#pragma unroll
      for (int f = 0; f < kGwLnFrames; ++f) {
        a0[f] = (f < nf && lane < NP) ? __ldcg(sp + (size_t)f * NP + lane) : make_float2(0.f, 0.f);
        a1[f] = (f < nf && lane + 32 < NP) ? __ldcg(sp + (size_t)f * NP + lane + 32) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int f = 0; f < kGwLnFrames; ++f) {
        const float ref = __shfl_sync(0xffffffffu, a0[f].x, 0);
        const float e0 = lane < NP ? a0[f].x - ref : 0.f, e1 = lane + 32 < NP ? a1[f].x - ref : 0.f;
        float sd = e0 + e1, sdd = fmaf(e0, e0, e1 * e1), sm2 = a0[f].y + a1[f].y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sd += __shfl_xor_sync(0xffffffffu, sd, o);
          sdd += __shfl_xor_sync(0xffffffffu, sdd, o);
          sm2 += __shfl_xor_sync(0xffffffffu, sm2, o);
        }
        const float mean = ref + sd * inv_np;
        const float tq = sm2 + (float)(CO / kEpiNH) * fmaxf(sdd - sd * sd * inv_np, 0.f);
        fr[f] = 1.f / sqrtf(tq * inv_cv + p.eps);
        fm[f] = -mean * fr[f];
      }
      // streaming pass, position-major: a lane owns the positions p = lane, lane + 32, ... of a frame ((joint,
      // channel quad), contiguous in the slot, in the output and in the [V][C] affine tables) and applies each
      // position's affine to all frames of the item -- the two table loads are amortised over kGwLnFrames
      // elements and every load instruction of the warp covers 512 contiguous bytes (this kernel leaves
      // ~8 KB of L1: per-element table loads in the [C/4][V][4] epilogue layout ran at L2 latency each)
      const int VC4 = V * C4;
      const float4 *zs = reinterpret_cast<const float4 *>(p.zring + ((size_t)slot * 128 + f0) * V * CO);
      const float4 *gw = reinterpret_cast<const float4 *>(p.n_wV), *gb = reinterpret_cast<const float4 *>(p.n_bV);
      const long long fo = (p.out_T ? (long long)n * p.out_T + p.out_t0 : (long long)n * p.T) + tb * 128 + f0;
      const size_t ob = (size_t)fo * V * CO;
      (void)kSh; (void)vmagic;
#pragma unroll 1
      const int pos_end = (pc + 1) * pc_len < VC4 ? (pc + 1) * pc_len : VC4;
      for (int pos = pc * pc_len + lane; pos < pos_end && nf > 0; pos += 32) {
        float4 a[kGwLnFrames];
        const float4 gg = __ldg(gw + pos), oo = __ldg(gb + pos);
#pragma unroll
        for (int f = 0; f < kGwLnFrames; ++f)
          a[f] = f < nf ? ld_hint16(zs + (size_t)f * VC4 + pos, pol_keep) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int f = 0; f < kGwLnFrames; ++f) {
          if (f < nf) {
            const float rs = fr[f], nmr = fm[f];
            const size_t i = (size_t)f * VC4 + pos;
            float4 r;
            r.x = fmaf(fmaf(a[f].x, rs, nmr), gg.x, oo.x);
            r.y = fmaf(fmaf(a[f].y, rs, nmr), gg.y, oo.y);
            r.z = fmaf(fmaf(a[f].z, rs, nmr), gg.z, oo.z);
            r.w = fmaf(fmaf(a[f].w, rs, nmr), gg.w, oo.w);
            if (p.relu) {
              r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
            }
            if ((p.debug & 2048) && r.x != 12345.678f) continue;   // (debug bit: no LN stores)
            if (p.out_f32) {
              st_hint16(p.out_f32 + ob + 4 * i, r, pol_stream);
            } else {
              const __nv_bfloat162 h01 = __floats2bfloat162_rn(r.x, r.y), h23 = __floats2bfloat162_rn(r.z, r.w);
              st_hint8(p.out_hi + ob + 4 * i,
                       make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23)), pol_stream);
              if (p.out_lo) {
                const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                const __nv_bfloat162 l01 = __floats2bfloat162_rn(r.x - f01.x, r.y - f01.y);
                const __nv_bfloat162 l23 = __floats2bfloat162_rn(r.z - f23.x, r.w - f23.y);
                st_hint8(p.out_lo + ob + 4 * i,
                         make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23)), pol_stream);
              }
            }
          }
        }
      }
      // every lane has consumed its loads of the slot (their values were used above): count the item; the slot
      // is released to the epilogue of group grp + R when all kGwLnItems items are in
      __syncwarp();
      if (lane == 0) atomicAdd(p.done + grp, 1u);
      li = __shfl_sync(0xffffffffu, nli, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// LayerNorm(C,V) (+ ReLU) of z fp32 frames as a streaming kernel: one block per frame; output fp32 rows
// or bf16 hi/lo planes (frame t of trial n stored at frame t + out_t0 of out_T frames).
struct LnStreamArgs {
  long long frames;
  int T, V, C;
  const float *z;
  const float *n_wT, *n_bT;       // [C/4][V][4]
  int relu;
  float eps;
  float *out_f32;
  __nv_bfloat16 *out_hi, *out_lo;
  int out_T, out_t0;
  // optional residual added after the norm (and after relu_mid), before the final relu (CoST-GCN:
  // relu(LN2(q) + res)): rows of frame f at res + f*V*C, as fp32 or as bf16 hi (+ lo) planes
  const float *res_f32;
  const __nv_bfloat16 *res_hi, *res_lo;
  int relu_mid;
};

template <int NV>
__global__ void __launch_bounds__(256, 2) k_ln_stream(LnStreamArgs p) {
  __shared__ float s_red[32];
  griddep_wait();                                            // z comes from the GEMM launched just before
  griddep_launch();
  const long long f = blockIdx.x;
  const int VC4 = (p.V * p.C) >> 2, C4 = p.C >> 2;
  const bool c4_pow2 = (C4 & (C4 - 1)) == 0;
  const int c4_sh = __ffs(C4) - 1;
  const float *zp = p.z + f * (long long)p.V * p.C;
  float4 a[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + 256 * j;
    a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < VC4) {
      a[j] = ld_stream(reinterpret_cast<const float4 *>(zp) + i);
      s += (a[j].x + a[j].y) + (a[j].z + a[j].w);
    }
  }
  const float inv_n = 1.f / (float)(p.V * p.C), inv_nm1 = 1.f / (float)(p.V * p.C - 1);
  const float mean = block_sum(s, s_red) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < VC4) {
      const float d0 = a[j].x - mean, d1 = a[j].y - mean, d2 = a[j].z - mean, d3 = a[j].w - mean;
      q = fmaf(d0, d0, q); q = fmaf(d1, d1, q); q = fmaf(d2, d2, q); q = fmaf(d3, d3, q);
    }
  }
  const float rstd = 1.f / sqrtf(block_sum(q, s_red) * inv_nm1 + p.eps);
  const float nmr = -mean * rstd;
  long long fo = f;
  if (p.out_T) {                                             // halo layout only: keeps the division off the main path
    const int n = (int)f / p.T;                              // frames < 2^31 (checked by the launcher)
    fo = (long long)n * p.out_T + ((int)f - n * p.T) + p.out_t0;
  }
  const long long ob = fo * (long long)p.V * p.C;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + 256 * j;
    if (i < VC4) {
      const int w = c4_pow2 ? (i >> c4_sh) : (i / C4), g = i - w * C4;   // no integer division for C = 64/128/256
      const int ti = (g * p.V + w) * 4;
      const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.n_wT + ti));
      const float4 o4 = __ldg(reinterpret_cast<const float4 *>(p.n_bT + ti));
      float4 v;
      v.x = fmaf(fmaf(a[j].x, rstd, nmr), g4.x, o4.x);
      v.y = fmaf(fmaf(a[j].y, rstd, nmr), g4.y, o4.y);
      v.z = fmaf(fmaf(a[j].z, rstd, nmr), g4.z, o4.z);
      v.w = fmaf(fmaf(a[j].w, rstd, nmr), g4.w, o4.w);
      if (p.relu_mid) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      const long long ib = f * (long long)p.V * p.C + 4 * i;
      if (p.res_f32) {
        const float4 r4 = ld_stream(reinterpret_cast<const float4 *>(p.res_f32 + ib));
        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
      } else if (p.res_hi) {
        const uint2 rh = *reinterpret_cast<const uint2 *>(p.res_hi + ib);
        const float2 h01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rh.x));
        const float2 h23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rh.y));
        v.x += h01.x; v.y += h01.y; v.z += h23.x; v.w += h23.y;
        if (p.res_lo) {
          const uint2 rl = *reinterpret_cast<const uint2 *>(p.res_lo + ib);
          const float2 l01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rl.x));
          const float2 l23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rl.y));
          v.x += l01.x; v.y += l01.y; v.z += l23.x; v.w += l23.y;
        }
      }
      if (p.relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      if (p.out_f32) {
        *reinterpret_cast<float4 *>(p.out_f32 + ob + 4 * i) = v;
      } else {
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2 *>(p.out_hi + ob + 4 * i) =
            make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
        if (p.out_lo) {
          const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
          const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
          *reinterpret_cast<uint2 *>(p.out_lo + ob + 4 * i) =
              make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
        }
      }
    }
  }
}

// The same LayerNorm with one WARP per frame (V*C/4 <= 32*NV float4 held in registers): statistics by warp
// shuffles only -- no shared memory, no block barriers, a third of the instructions per element of the block
// form, which is issue-bound for C = 64 (ncu: 79 % issue, 50 % DRAM).  Statistics in one pass about the frame's
// first value (shifted sums; merged exactly as in frame_stats).
template <int NV>
__global__ void __launch_bounds__(256) k_ln_warp(LnStreamArgs p) {
  const int lane = threadIdx.x & 31;
  griddep_wait();
  griddep_launch();
  const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (f >= p.frames) return;
  const int VC4 = (p.V * p.C) >> 2, C4 = p.C >> 2;
  const int c4_sh = __ffs(C4) - 1;                          // C4 is a power of two here (C = 64 / 128 / 256)
  const float *zp = p.z + f * (long long)p.V * p.C;
  float4 a[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    a[j] = i < VC4 ? ld_stream(reinterpret_cast<const float4 *>(zp) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float shift = __shfl_sync(0xffffffffu, a[0].x, 0);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (lane + 32 * j < VC4) {
      const float d0 = a[j].x - shift, d1 = a[j].y - shift, d2 = a[j].z - shift, d3 = a[j].w - shift;
      s1 += (d0 + d1) + (d2 + d3);
      s2 = fmaf(d0, d0, s2); s2 = fmaf(d1, d1, s2); s2 = fmaf(d2, d2, s2); s2 = fmaf(d3, d3, s2);
    }
  }
  // per-lane partial (mean, M2) merged across the warp (Chan et al.): exact for unequal counts as well
  float cnt = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) cnt += (lane + 32 * j < VC4) ? 4.f : 0.f;
  float mean = cnt > 0.f ? shift + s1 / cnt : 0.f;
  float m2 = cnt > 0.f ? fmaxf(s2 - s1 * s1 / cnt, 0.f) : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float cb = __shfl_xor_sync(0xffffffffu, cnt, o);
    const float mb = __shfl_xor_sync(0xffffffffu, mean, o);
    const float qb = __shfl_xor_sync(0xffffffffu, m2, o);
    const float tot = cnt + cb;
    if (tot > 0.f) {
      const float dm = mb - mean;
      m2 = m2 + qb + dm * dm * (cnt * cb / tot);
      mean = mean + dm * (cb / tot);
    }
    cnt = tot;
  }
  const float rstd = 1.f / sqrtf(m2 / (float)(p.V * p.C - 1) + p.eps);
  const float nmr = -mean * rstd;
  long long fo = f;
  if (p.out_T) {                                             // halo layout only: keeps the division off the main path
    const int n = (int)f / p.T;                              // frames < 2^31 (checked by the launcher)
    fo = (long long)n * p.out_T + ((int)f - n * p.T) + p.out_t0;
  }
  const long long ob = fo * (long long)p.V * p.C;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    if (i < VC4) {
      const int w = i >> c4_sh, g = i - (w << c4_sh);
      const int ti = (g * p.V + w) * 4;
      const float4 g4 = __ldg(reinterpret_cast<const float4 *>(p.n_wT + ti));
      const float4 o4 = __ldg(reinterpret_cast<const float4 *>(p.n_bT + ti));
      float4 v;
      v.x = fmaf(fmaf(a[j].x, rstd, nmr), g4.x, o4.x);
      v.y = fmaf(fmaf(a[j].y, rstd, nmr), g4.y, o4.y);
      v.z = fmaf(fmaf(a[j].z, rstd, nmr), g4.z, o4.z);
      v.w = fmaf(fmaf(a[j].w, rstd, nmr), g4.w, o4.w);
      if (p.relu_mid) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      const long long ib = f * (long long)p.V * p.C + 4 * i;
      if (p.res_f32) {
        const float4 r4 = ld_stream(reinterpret_cast<const float4 *>(p.res_f32 + ib));
        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
      } else if (p.res_hi) {
        const uint2 rh = *reinterpret_cast<const uint2 *>(p.res_hi + ib);
        const float2 h01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rh.x));
        const float2 h23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rh.y));
        v.x += h01.x; v.y += h01.y; v.z += h23.x; v.w += h23.y;
        if (p.res_lo) {
          const uint2 rl = *reinterpret_cast<const uint2 *>(p.res_lo + ib);
          const float2 l01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rl.x));
          const float2 l23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&rl.y));
          v.x += l01.x; v.y += l01.y; v.z += l23.x; v.w += l23.y;
        }
      }
      if (p.relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      if (p.out_f32) {
        *reinterpret_cast<float4 *>(p.out_f32 + ob + 4 * i) = v;
      } else {
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2 *>(p.out_hi + ob + 4 * i) =
            make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
        if (p.out_lo) {
          const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
          const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
          *reinterpret_cast<uint2 *>(p.out_lo + ob + 4 * i) =
              make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
        }
      }
    }
  }
}

// STGCN_LN_WARP: largest V*C/4 handled by the warp-per-frame form (default 416 = C 64 x V 26; 0 = never)
inline int ln_warp_limit() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("STGCN_LN_WARP");
    v = e ? atoi(e) : 416;
  }
  return v;
}

inline int launch_ln_stream(const LnStreamArgs &a, cudaStream_t st) {
  const int nv = (a.V * a.C / 4 + 255) / 256;
  if (a.frames <= 0) return 0;
  if (a.frames > 0x7fffffffLL) return fail("ln stream: too many frames");
  const int vc4 = a.V * a.C / 4, c4 = a.C / 4;
  if ((c4 & (c4 - 1)) == 0 && vc4 <= ln_warp_limit()) {
    const unsigned blocks = (unsigned)((a.frames + 7) / 8);
    if (vc4 <= 32 * 13) { STGCN_CUDA_OK(launch_pdl(k_ln_warp<13>, dim3(blocks), dim3(256), (size_t)0, st, a)); return 0; }
    if (vc4 <= 32 * 25) { STGCN_CUDA_OK(launch_pdl(k_ln_warp<25>, dim3(blocks), dim3(256), (size_t)0, st, a)); return 0; }
  }
  const dim3 grid((unsigned)a.frames), block(256);
  if (nv <= 2) STGCN_CUDA_OK(launch_pdl(k_ln_stream<2>, grid, block, (size_t)0, st, a));
  else if (nv <= 4) STGCN_CUDA_OK(launch_pdl(k_ln_stream<4>, grid, block, (size_t)0, st, a));
  else if (nv <= 7) STGCN_CUDA_OK(launch_pdl(k_ln_stream<7>, grid, block, (size_t)0, st, a));
  else return fail("ln stream: V*C = %d too large", a.V * a.C);
  return 0;
}

inline bool gcnw_supported(int c_in, int c_out, int V, int K) {
  return (c_out == 64 || c_out == 128 || c_out == 256) && c_in % 64 == 0 && c_in >= 64 && V >= 2 && V <= kGwMaxV &&
         K >= 1 && (V * c_out / 4 + 255) / 256 <= 7;
}
inline bool gcnw_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_GCNW");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}

// The one-kernel form of the stage is opt-in (STGCN_GCNW_FUSE=1): it removes z's HBM round trip (measured:
// 1.64 GB instead of 3.2 GB of DRAM traffic per C = 64 launch pair) but the cross-CTA ring couples the
// progress of all CTAs, and on B200 it runs 1.7x slower than GEMM + k_ln_stream
// STGCN_GCNW_TMA_OUT=0: the GEMM's rows leave with st.global from the warp's patch instead of TMA stores
inline bool gcnw_tma_out_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_GCNW_TMA_OUT");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}
inline bool gcnw_fuse_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_GCNW_FUSE");
    on = e ? atoi(e) != 0 : 0;
  }
  return on != 0;
}
// ring geometry of the fused stage: slots of [128 frames][V][CO] fp32.  The ring must hold the groups
// the GEMM side produces during one trip round the produce -> normalise -> release loop and still stay
// L2-resident (eviction-priority hints): 64 MB / 6..64 slots measured best (sweep in profiles/)
inline int gcnw_env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e && atoi(e) > 0 ? atoi(e) : dflt;
}
inline int gcnw_ring_slots(int V, int CO) {
  static const int mb = gcnw_env_int("STGCN_GCNW_RING_MB", 64);       // tuning aid
  const size_t slot = (size_t)128 * V * CO * sizeof(float);
  int R = (int)((size_t)mb * 1024 * 1024 / slot);
  return R > 64 ? 64 : (R < 6 ? 6 : R);
}
inline int gcnw_ln_chunks(int V, int CO) {
  static const int iters = gcnw_env_int("STGCN_GCNW_LN_ITERS", kGwLnIters);   // tuning aid
  const int steps = (V * CO / 4 + 31) / 32;
  return (steps + iters - 1) / iters;
}
inline size_t gcnw_ring_floats(int V, int CO) { return (size_t)gcnw_ring_slots(V, CO) * 128 * V * CO; }
inline size_t gcnw_sring_float2(int V, int CO) { return (size_t)gcnw_ring_slots(V, CO) * 128 * V * kEpiNH; }

// x planes: bf16 [planes][N_full][T_full][V][Cin] (the view takes every fstride-th frame, T frames, of the
// first p.N trials from `x`); wsc: k_gcnw_pack tiles [2][cap][CO][Cin].  FUSE: p.zring .. p.out_t0 set and
// p.ready / p.done zeroed by the caller (N * tblocks counters each).
// tap mode view of the activations: a ring [planes][slots][rows][C]; tensor-map dimensions (C, rows, slots, 1,
// planes) with a {64, 128, 1, 1, 1} box (strides stay increasing)
struct GcnwXView {
  long long slots = 0;       // 0: ordinary (C, V, T, N, planes) view
  long long slot_stride = 0; // elements between ring slots
};
template <int CO, bool FUSE>
int launch_gcnw_c(const __nv_bfloat16 *x, const __nv_bfloat16 *wsc, GcnwParams p, int T_full, int fstride, int cap,
                  long long plane_stride, cudaStream_t st, const GcnwXView &xv) {
  const int V = p.V;
  p.tblocks = (p.T + 127) / 128;
  p.items = p.N * p.tblocks * V;
  const int fixed = kPatchTotal + 1024 + 1024 + 512 + 1024;
  constexpr bool MERGE = CO <= 128;
  constexpr int kA = MERGE ? 2 * 128 * 128 : 128 * 128, kB = MERGE ? 2 * CO * 128 : CO * 128;
  // weight stages first (they are the long-latency stream for C >= 128), then activation stages
  int SB = MERGE ? (CO >= 128 ? 2 : 3) : 3;
  int kMaxSmem = 232448;
  if (p.smem_cap > 0 && p.smem_cap < kMaxSmem) {           // shared-SM mode: smallest rings that still pipeline
    if (!MERGE) SB = 2;
    const int need = fixed + SB * kB + 2 * kA;
    kMaxSmem = p.smem_cap > need ? p.smem_cap : need;
  }
  int SA = (kMaxSmem - fixed - SB * kB) / kA;
  if (SA > (MERGE ? 4 : 8)) SA = MERGE ? 4 : 8;
  if (SA < 2) return fail("gcnw: shared memory does not fit");
  if (MERGE && SA >= 4 && (kMaxSmem - fixed - SB * kB - SA * kA) >= kB) ++SB;
  p.a_stages = SA; p.b_stages = SB;
  const int smem = fixed + SA * kA + SB * kB;
  CUtensorMap tm_x, tm_w;
  uint64_t xd[5] = {(uint64_t)p.Cin, (uint64_t)V, (uint64_t)p.T, (uint64_t)p.N, (uint64_t)p.planes};
  uint64_t xs[4] = {(uint64_t)p.Cin * 2, (uint64_t)fstride * V * p.Cin * 2, (uint64_t)T_full * V * p.Cin * 2,
                    (uint64_t)plane_stride * 2};
  uint32_t xb[5] = {64, 1, 128, 1, 1};
  if (xv.slots > 0) {
    xd[1] = (uint64_t)p.T; xd[2] = (uint64_t)xv.slots;
    xs[0] = (uint64_t)p.Cin * 2;
    xs[1] = (uint64_t)xv.slot_stride * 2;
    xs[2] = (uint64_t)xv.slot_stride * xv.slots * 2;     // N == 1
    xb[1] = 128; xb[2] = 1;
  }
  if (make_tmap_bf16(&tm_x, x, 5, xd, xs, xb)) return 1;
  CUtensorMap tm_x1 = tm_x;
  if (p.tmode == 2) {
    // the view is the convolution INPUT: T_full frames per trial, read at stride p.tstride
    if (p.tstride == 2) {
      for (int q = 0; q < 2; ++q) {
        xd[2] = q == 0 ? (uint64_t)(T_full + 1) / 2 : (uint64_t)T_full / 2;
        xs[1] = (uint64_t)2 * V * p.Cin * 2;
        if (make_tmap_bf16(q == 0 ? &tm_x : &tm_x1, x + (size_t)q * V * p.Cin, 5, xd, xs, xb)) return 1;
      }
    } else {
      xd[2] = (uint64_t)T_full;
      xs[1] = (uint64_t)V * p.Cin * 2;
      if (make_tmap_bf16(&tm_x, x, 5, xd, xs, xb)) return 1;
      tm_x1 = tm_x;
    }
  }
  const uint64_t wd[4] = {(uint64_t)p.Cin, (uint64_t)CO, (uint64_t)cap, 2};
  const uint64_t wst[3] = {(uint64_t)p.Cin * 2, (uint64_t)CO * p.Cin * 2, (uint64_t)cap * CO * p.Cin * 2};
  const uint32_t wb[4] = {64, (uint32_t)CO, 1, 1};
  if (make_tmap_bf16(&tm_w, wsc, 4, wd, wst, wb)) return 1;
  CUtensorMap tm_z = tm_w;
  p.tma_out = (!FUSE && gcnw_tma_out_enabled()) ? 1 : 0;
  if (p.tma_out) {
    const uint64_t zd[4] = {(uint64_t)CO, (uint64_t)V, (uint64_t)p.T, (uint64_t)p.N};
    const uint64_t zs[3] = {(uint64_t)CO * 4, (uint64_t)V * CO * 4, (uint64_t)p.T * V * CO * 4};
    const uint32_t zb[4] = {32, 1, 32, 1};
    if (make_tmap(&tm_z, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, p.out, 4, zd, zs, zb)) return 1;
  }
  const int grid = p.items < num_sms() ? p.items : num_sms();
  STGCN_CUDA_OK(cudaFuncSetAttribute(k_gcnw<CO, MERGE, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (FUSE) {
    // producers wait for consumers in other CTAs: every CTA must be resident -> cooperative launch
    void *args[5] = {&tm_x, &tm_x1, &tm_w, &tm_z, &p};
    STGCN_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(&k_gcnw<CO, MERGE, FUSE>), dim3(grid),
                                              dim3(kGwThreads + 32 + kGwLnThreads), args, (size_t)smem, st));
  } else {
    STGCN_CUDA_OK(launch_pdl(k_gcnw<CO, MERGE, FUSE>, dim3(grid), dim3(kGwThreads), (size_t)smem, st, tm_x, tm_x1, tm_w,
                             tm_z, p));
  }
  return 0;
}

// plane_stride: elements between the hi and lo planes of x (= rows of the whole buffer * Cin)
inline int launch_gcnw(int CO, const __nv_bfloat16 *x, const __nv_bfloat16 *wsc, const GcnwParams &p, int T_full,
                       int fstride, int cap, long long plane_stride, cudaStream_t st,
                       const GcnwXView &xv = GcnwXView()) {
  const bool fuse = p.zring != nullptr;
  switch (CO) {
    case 64: return fuse ? launch_gcnw_c<64, true>(x, wsc, p, T_full, fstride, cap, plane_stride, st, xv)
                         : launch_gcnw_c<64, false>(x, wsc, p, T_full, fstride, cap, plane_stride, st, xv);
    case 128: return fuse ? launch_gcnw_c<128, true>(x, wsc, p, T_full, fstride, cap, plane_stride, st, xv)
                          : launch_gcnw_c<128, false>(x, wsc, p, T_full, fstride, cap, plane_stride, st, xv);
    case 256: return fuse ? launch_gcnw_c<256, true>(x, wsc, p, T_full, fstride, cap, plane_stride, st, xv)
                          : launch_gcnw_c<256, false>(x, wsc, p, T_full, fstride, cap, plane_stride, st, xv);
  }
  return fail("gcnw: unsupported channel count %d", CO);
}

}  // namespace tc
}  // namespace stgcn
