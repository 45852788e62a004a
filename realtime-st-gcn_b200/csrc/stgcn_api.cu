// C ABI of the B200-native ST-GCN / RT-ST-GCN forward path (see include/stgcn_b200.h).
#include "../../include/stgcn_b200.h"

#include <cstdlib>

#include <vector>
#include "common.cuh"
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_tc_pair.cuh"
#include "kernels_gcnw.cuh"
#include "kernels_rt_small.cuh"
#include "kernels_rt_stream.cuh"

using namespace stgcn;

namespace {

constexpr float kEps = 1e-5f;  // reference LayerNorm / BatchNorm eps (layernorm.py:8)

// Measurement aid, compiled in only with -DSTGCN_DEBUG_BUILD (STGCN_NVCC_EXTRA): STGCN_DEBUG bits skip
// parts of the kernels (1: tensor-core epilogues, 2: transform math, ...) to attribute time between
// pipeline roles -- results are then wrong on purpose.  The shipped library ignores the variable.
inline int debug_mode() {
#ifdef STGCN_DEBUG_BUILD
  static int m = -1;
  if (m < 0) {
    const char *e = getenv("STGCN_DEBUG");
    m = e ? atoi(e) : 0;
  }
  return m;
#else
  return 0;
#endif
}

// STGCN_DEBUG & 4: after a tensor-core launch, synchronise and print CTA 0's role/wait cycle counters
int debug_dump(const char *what, int c, cudaStream_t st) {
  if (!(debug_mode() & 4)) return 0;
  unsigned long long h[16];
  STGCN_CUDA_OK(cudaStreamSynchronize(st));
  STGCN_CUDA_OK(cudaMemcpyFromSymbol(h, tc::g_dbg, sizeof(h)));
  static const char *names[12] = {"mma_wait_tmem", "mma_wait_A", "mma_wait_B", "mma_total", "xA_prod_wait",
                                  "B_prod_wait", "epi_wait_tmem", "epi_work", "xf_wait_x", "xf_wait_Aempty",
                                  "xf_compute", "items"};
  fprintf(stderr, "[dbg] %s<%d>:", what, c);
  for (int i = 0; i < 12; ++i) fprintf(stderr, " %s=%llu", names[i], h[i]);
  if (!strcmp(what, "tcn"))
    fprintf(stderr, " [tcn pass 2 of CTA 0 / row 0: 8 residual load 9 tmem loads 10 tables+math+patch 15 cooperative store]");
  fprintf(stderr, " ph12(pass1|rt_wait)=%llu ph13(bar_stats|rt_update)=%llu ph14(pass2|rt_store)=%llu ph15(rt_finish)=%llu", h[12], h[13], h[14], h[15]);
  fprintf(stderr, "\n");
  memset(h, 0, sizeof(h));
  STGCN_CUDA_OK(cudaMemcpyToSymbol(tc::g_dbg, h, sizeof(h)));
  return 0;
}

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline int round4(int x) { return (x + 3) & ~3; }

// ---- launch helpers ---------------------------------------------------------
int to_ntvc(const float *x, float *y, int N, int C, long long P, int ld, cudaStream_t st) {
  dim3 grid(cdiv(P, 32), cdiv(ld, 32), N), block(32, 8);
  ProfScope ps(KC_LAYOUT, st);
  k_cp_to_pc<<<grid, block, 0, st>>>(x, y, C, P, ld);
  STGCN_LAUNCH_OK();
  return 0;
}
int to_nctv(const float *x, float *y, int N, int C, long long P, int ld, cudaStream_t st) {
  dim3 grid(cdiv(P, 32), cdiv(C, 32), N), block(32, 8);
  ProfScope ps(KC_LAYOUT, st);
  k_pc_to_cp<<<grid, block, 0, st>>>(x, y, C, P, ld);
  STGCN_LAUNCH_OK();
  return 0;
}

// Y[M, C_out] = conv over taps of X (NTVC), weights packed (C_out, G, C_in)
int launch_gemm(const float *X, const float *Wp, const float *bias, float *Y, int N, int T_in, int V,
                int C_in, int C_out, int G, int stride, cudaStream_t st) {
  STGCN_REQUIRE(C_in % 4 == 0 && C_out % 4 == 0, "gemm: channels must be multiples of 4 (got %d -> %d)",
                C_in, C_out);
  ConvGeom g;
  g.T_in = T_in;
  g.T_out = (T_in - 1) / stride + 1;
  g.V = V;
  g.C_in = C_in;
  g.C_out = C_out;
  g.G = G;
  g.stride = stride;
  g.pad = (G - 1) / 2;
  g.Ktot = G * C_in;
  long long M = (long long)N * g.T_out * V;
  STGCN_REQUIRE(M < (1ll << 31) - 256, "gemm: too many rows (%lld)", M);
  g.M = (int)M;
  dim3 grid(cdiv(M, 128), cdiv(C_out, 64));
  ProfScope ps(G > 1 ? KC_GEMM_TCN : KC_GEMM_1X1, st);
  k_gemm_conv<128, 64, 16><<<grid, 256, 0, st>>>(X, Wp, bias, Y, g);
  STGCN_LAUNCH_OK();
  return 0;
}

int launch_frame(const FrameArgs &a, cudaStream_t st) {
  size_t smem = (size_t)a.V * a.C * sizeof(float) * (a.b_mode == B_LN ? 2 : 1);
  STGCN_REQUIRE(smem <= 200 * 1024, "frame kernel: V*C too large for shared memory (%zu B)", smem);
  if (smem > 48 * 1024)
    STGCN_CUDA_OK(cudaFuncSetAttribute(k_frame, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = a.frames < 148 * 8 ? a.frames : 148 * 8;
  ProfScope ps(KC_FRAME, st);
  k_frame<<<(unsigned)blocks, 256, smem, st>>>(a);
  STGCN_LAUNCH_OK();
  return 0;
}

// STGCN_TAPS=0: layers the frame-tile temporal kernel does not cover (Gamma > 15) and BatchNorm layers use
// the fp32 CUDA-core kernels instead of the all-taps tensor-core path
inline bool taps_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_TAPS");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}

struct AdjCsr {
  int *ptr;
  int *yoff;
  float *val;
};

int build_csr(const float *A, int a_per_sample, int N, int K, int V, int C, Bump &ws, AdjCsr &csr,
              cudaStream_t st) {
  const int nA = a_per_sample ? N : 1;
  csr.ptr = ws.take<int>((size_t)nA * (V + 1));
  csr.yoff = ws.take<int>((size_t)nA * K * V * V);
  csr.val = ws.take<float>((size_t)nA * K * V * V);
  if (ws.measuring()) return 0;
  STGCN_REQUIRE(!ws.overflow, "workspace too small (adjacency)");
  STGCN_REQUIRE(V <= 1024, "too many joints (%d)", V);
  ProfScope ps(KC_MISC, st);
  k_build_adj_csr<<<nA, 64, (V + 1) * sizeof(int), st>>>(A, K, V, C, csr.ptr, csr.yoff, csr.val);
  STGCN_LAUNCH_OK();
  return 0;
}

int channel_stats(const float *x, long long rows, int C, double *sums /* [2*C] */, cudaStream_t st) {
  STGCN_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  long long rpb = 1024;
  long long blocks = (rows + rpb - 1) / rpb;
  if (blocks > 148 * 16) {
    blocks = 148 * 16;
    rpb = (rows + blocks - 1) / blocks;
    blocks = (rows + rpb - 1) / rpb;
  }
  ProfScope ps(KC_BN, st);
  k_channel_stats<<<(unsigned)blocks, 256, 0, st>>>(x, rows, C, rpb, sums, sums + C);
  STGCN_LAUNCH_OK();
  return 0;
}

int check_layer(const stgcn_layer_desc &d) {
  STGCN_REQUIRE(d.kernel % 2 == 1, "temporal kernel must be odd (got %d)", d.kernel);
  STGCN_REQUIRE(d.stride >= 1, "stride must be >= 1");
  STGCN_REQUIRE(d.c_in % 4 == 0 && d.c_out % 4 == 0,
                "layer channels must be multiples of 4 (got %d -> %d)", d.c_in, d.c_out);
  STGCN_REQUIRE(d.residual != STGCN_RES_IDENTITY || (d.c_in == d.c_out && (d.stride == 1 || d.rt)),
                "identity residual needs c_in == c_out and stride 1");
  return 0;
}

// ---- per-layer prepared operands of the tensor-core kernels ----------------------
// bf16 hi/lo weight planes, the (k,w)-ordered adjacency CSR and the bias that flows through the
// adjacency.  They depend only on the parameters, so a model prepares them once
// (stgcn_model_prepare) and every forward reuses them; the layer-level entry points build them
// per call in the workspace.
struct LayerPrep {
  bool gcn = false, tcn = false, res = false, csr = false;
  int *kw_ptr = nullptr;
  int2 *kw_va = nullptr;
  float *bzT = nullptr;
  __nv_bfloat16 *wg16 = nullptr, *wp16 = nullptr, *wr16 = nullptr;
  float *zero = nullptr;  // c_out zeros: bias of the bias-free RT residual conv (rtstgcn.py:503)
  float *n1T = nullptr, *n2T = nullptr, *nrT = nullptr;  // LayerNorm affine (V, C): weight then bias
  // graph conv with per-joint pre-scaled weights (kernels_gcnw.cuh): edge tables + weight tiles
  bool gw = false, gwr = false;
  tc::GcnwTables *gwtab = nullptr, *gwtabr = nullptr;
  __nv_bfloat16 *wsc = nullptr, *wscr = nullptr;
  float *n1V = nullptr, *nrV = nullptr;   // LayerNorm affine as [V][C] (LN warps of the fused stage): weight then bias
  // "all taps" layer path (temporal convolution as a per-joint-tile GEMM over the taps, kernels_gcnw.cuh tmode 2):
  // LayerNorm layers whose temporal kernel does not fit the frame-tile kernel (> 15 taps), and BatchNorm layers
  bool taps = false;
};

LayerPrep prep_take(const stgcn_layer_desc &d, int K, int V, Bump &ws, bool sparse_adj = false) {
  LayerPrep P;
  const bool ln = d.norm == STGCN_NORM_LAYERNORM;
  P.gcn = ln && !d.a_per_sample && tc::gcn_tc_supported(d.c_in, d.c_out, V, K);
  P.csr = !d.a_per_sample && K * V + 1 <= 1024;
  P.tcn = ln && d.rt != 1 && tc::tcn_tc2_supported(d.c_out, V, d.kernel, d.stride, 2);
  P.res = ln && d.residual == STGCN_RES_CONV && tc::gcn_tc_supported(d.c_in, d.c_out, V, 1);
  if (P.csr) {
    P.kw_ptr = ws.take<int>((size_t)K * V + 1);
    P.kw_va = ws.take<int2>((size_t)K * V * V);
  }
  if (P.gcn) {
    P.bzT = ws.take<float>((size_t)d.c_out * V);
    P.wg16 = ws.take<__nv_bfloat16>((size_t)2 * K * d.c_out * d.c_in);
  }
  if (P.tcn) P.wp16 = ws.take<__nv_bfloat16>((size_t)2 * d.c_out * d.c_out * d.kernel);
  if (P.res) {
    P.wr16 = ws.take<__nv_bfloat16>((size_t)2 * d.c_out * d.c_in);
    P.zero = ws.take<float>((size_t)d.c_out);
    P.nrT = ws.take<float>((size_t)2 * d.c_out * V);
  }
  if (P.gcn) P.n1T = ws.take<float>((size_t)2 * d.c_out * V);
  if (P.tcn) P.n2T = ws.take<float>((size_t)2 * d.c_out * V);
  P.gw = sparse_adj && P.gcn && tc::gcnw_enabled() && tc::gcnw_supported(d.c_in, d.c_out, V, K) &&
         (d.residual != STGCN_RES_CONV || P.res);
  const bool bn = d.norm == STGCN_NORM_BATCHNORM;
  const bool shapes_ok = !d.a_per_sample && d.rt == 0 && sparse_adj && tc::gcnw_enabled() &&
                         tc::gcnw_supported(d.c_in, d.c_out, V, K) && tc::gcnw_supported(d.c_out, d.c_out, V, 1) &&
                         d.kernel <= tc::kGwEdgeCap && taps_enabled();
  P.taps = shapes_ok && ((ln && P.gw && !P.tcn) || bn);
  if (P.taps) {
    if (!P.wp16) P.wp16 = ws.take<__nv_bfloat16>((size_t)2 * d.c_out * d.c_out * d.kernel);
    if (ln && !P.n2T) P.n2T = ws.take<float>((size_t)2 * d.c_out * V);
    if (bn) {
      P.bzT = ws.take<float>((size_t)d.c_out * V);
      P.gw = true;                                   // edge tables + pre-scaled weights below
    }
  }
  P.gwr = P.gw && d.residual == STGCN_RES_CONV;
  if (P.gw) {
    const size_t cap = (size_t)tc::gcnw_edge_cap(V);
    P.gwtab = ws.take<tc::GcnwTables>(1);
    P.wsc = ws.take<__nv_bfloat16>(2 * cap * d.c_out * d.c_in);
    P.n1V = ws.take<float>((size_t)2 * d.c_out * V);
    if (P.gwr) {
      P.nrV = ws.take<float>((size_t)2 * d.c_out * V);
      P.gwtabr = ws.take<tc::GcnwTables>(1);
      P.wscr = ws.take<__nv_bfloat16>(2 * cap * d.c_out * d.c_in);
    }
  }
  return P;
}

int prep_run(const stgcn_layer_desc &d, int K, int V, const LayerPrep &P, cudaStream_t st) {
  ProfScope ps(KC_MISC, st);
  if (P.csr) {
    tc::k_build_adj_csr_kw<<<1, 128, (K * V + 1) * sizeof(int), st>>>(d.a_eff, K, V, P.kw_ptr, P.kw_va);
    STGCN_LAUNCH_OK();
  }
  if (P.gcn) {
    const long long nw = (long long)K * d.c_out * d.c_in;
    tc::k_bias_through_adj<<<cdiv((long long)d.c_out * V, 256), 256, 0, st>>>(d.a_eff, d.gcn_b, K, V, d.c_out,
                                                                              P.bzT);
    STGCN_LAUNCH_OK();
    tc::k_split_bf16<<<cdiv(nw, 256), 256, 0, st>>>(d.gcn_w, P.wg16, nw);
    STGCN_LAUNCH_OK();
  }
  if (P.tcn || P.taps) {
    const long long nw = (long long)d.c_out * d.c_out * d.kernel;
    tc::k_pack_tcn_w_bf16<<<cdiv(nw, 256), 256, 0, st>>>(d.tcn_w, P.wp16, d.c_out, d.c_out, d.kernel);
    STGCN_LAUNCH_OK();
  }
  if (P.taps && !P.gcn) {                            // BatchNorm layers: the bias that flows through the adjacency
    tc::k_bias_through_adj<<<cdiv((long long)d.c_out * V, 256), 256, 0, st>>>(d.a_eff, d.gcn_b, K, V, d.c_out,
                                                                              P.bzT);
    STGCN_LAUNCH_OK();
  }
  if (P.res) {
    const long long nwr = (long long)d.c_out * d.c_in;
    tc::k_split_bf16<<<cdiv(nwr, 256), 256, 0, st>>>(d.res_w, P.wr16, nwr);
    STGCN_LAUNCH_OK();
    STGCN_CUDA_OK(cudaMemsetAsync(P.zero, 0, sizeof(float) * d.c_out, st));
  }
  if (P.gw) {
    const int cap = tc::gcnw_edge_cap(V);
    const long long per = (long long)d.c_out * d.c_in;
    tc::k_gcnw_tables<<<1, 32, 0, st>>>(d.a_eff, K, V, 0, cap, P.gwtab);
    STGCN_LAUNCH_OK();
    tc::k_gcnw_pack<<<cdiv(per * cap, 256), 256, 0, st>>>(d.gcn_w, P.gwtab, d.c_out, d.c_in, cap, P.wsc);
    STGCN_LAUNCH_OK();
    const int cvv = d.c_out * V;
    if (d.norm == STGCN_NORM_LAYERNORM) {
    tc::k_transpose_cv<<<cdiv(cvv, 256), 256, 0, st>>>(d.n1_w, P.n1V, d.c_out, V);
    STGCN_LAUNCH_OK();
    tc::k_transpose_cv<<<cdiv(cvv, 256), 256, 0, st>>>(d.n1_b, P.n1V + cvv, d.c_out, V);
    STGCN_LAUNCH_OK();
    }
    if (P.gwr) {
      if (d.norm == STGCN_NORM_LAYERNORM) {
      tc::k_transpose_cv<<<cdiv(cvv, 256), 256, 0, st>>>(d.nr_w, P.nrV, d.c_out, V);
      STGCN_LAUNCH_OK();
      tc::k_transpose_cv<<<cdiv(cvv, 256), 256, 0, st>>>(d.nr_b, P.nrV + cvv, d.c_out, V);
      STGCN_LAUNCH_OK();
      }
      tc::k_gcnw_tables<<<1, 32, 0, st>>>(nullptr, 1, V, 1, cap, P.gwtabr);
      STGCN_LAUNCH_OK();
      tc::k_gcnw_pack<<<cdiv(per * cap, 256), 256, 0, st>>>(d.res_w, P.gwtabr, d.c_out, d.c_in, cap, P.wscr);
      STGCN_LAUNCH_OK();
    }
  }
  const int cv = d.c_out * V;
  const float *src[3][2] = {{d.n1_w, d.n1_b}, {d.n2_w, d.n2_b}, {d.nr_w, d.nr_b}};
  float *dst[3] = {P.n1T, P.n2T, P.nrT};
  for (int i = 0; i < 3; ++i)
    if (dst[i])
      for (int j = 0; j < 2; ++j) {
        tc::k_transpose_affine<<<cdiv(cv, 256), 256, 0, st>>>(src[i][j], dst[i] + (size_t)j * cv, d.c_out, V);
        STGCN_LAUNCH_OK();
      }
  return 0;
}

// ---- graph-conv stage with per-joint pre-scaled weights (kernels_gcnw.cuh) --------------------
// z = GEMM(x planes) + bias, then LayerNorm(C,V) (+ ReLU) into l.out_*.  Fused form: one cooperative
// persistent kernel, z through an L2-resident ring; two-kernel form: z in HBM + k_ln_stream.
// Scratch comes from `ws` and is released on return.  In measuring mode only the sizes are taken.
int gcnw_stage(int c_out, const __nv_bfloat16 *xh, const __nv_bfloat16 *wsc, tc::GcnwParams g, tc::LnStreamArgs l,
               const float *affine_vc, int T_full, int fstride, long long plane_stride, Bump &ws, cudaStream_t st) {
  const size_t mark = ws.mark();
  const int V = g.V, cap = tc::gcnw_edge_cap(V);
  const long long rows = (long long)g.N * g.T * V;
  // The stage is a per-frame function: when the trials are dense in memory (no halo frames, frame stride 1, or
  // stride 2 over an even number of frames) they form ONE frame sequence, so a 128-frame tile may span trials --
  // short trials (sliding windows, T = 50 .. 300; the deep layers at T/4) no longer leave tiles part empty.
  if (g.N > 1 && l.out_T == 0 && g.tmode == 0 && (long long)g.N * T_full < (1ll << 30) &&
      ((fstride == 1 && T_full == g.T) || (fstride == 2 && T_full == 2 * g.T))) {
    T_full *= g.N;
    g.T *= g.N;
    l.T = g.T;
    g.N = 1;
  }
  if (tc::gcnw_fuse_enabled()) {
    const int groups = g.N * ((g.T + 127) / 128);
    g.zring = ws.take<float>(tc::gcnw_ring_floats(V, c_out));
    g.sring = ws.take<float2>(tc::gcnw_sring_float2(V, c_out));
    g.ready = ws.take<unsigned>((size_t)2 * groups + 1);           // [ready | done | LN ticket counter]
    g.done = g.ready ? g.ready + groups : nullptr;
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (graph-conv stage ring)");
      g.R = tc::gcnw_ring_slots(V, c_out);
      g.npc = tc::gcnw_ln_chunks(V, c_out);
      g.n_wV = affine_vc; g.n_bV = affine_vc + (size_t)c_out * V; g.relu = l.relu; g.eps = l.eps;
      g.out_f32 = l.out_f32; g.out_hi = l.out_hi; g.out_lo = l.out_lo;
      g.out_T = l.out_T; g.out_t0 = l.out_t0;
      STGCN_CUDA_OK(cudaMemsetAsync(g.ready, 0, sizeof(unsigned) * (2 * groups + 1), st));
      ProfScope ps(KC_GEMM_1X1, st);
      if (tc::launch_gcnw(c_out, xh, wsc, g, T_full, fstride, cap, plane_stride, st)) return 1;
      STGCN_LAUNCH_OK();
    }
  } else {
    float *zb = ws.take<float>((size_t)rows * c_out);
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (graph-conv stage)");
      g.out = zb;
      {
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcnw(c_out, xh, wsc, g, T_full, fstride, cap, plane_stride, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      l.z = zb;
      ProfScope ps(KC_FRAME, st);
      if (tc::launch_ln_stream(l, st)) return 1;
      STGCN_LAUNCH_OK();
    }
  }
  ws.release(mark);
  return 0;
}

// ---- "all taps" ST-GCN layer: both stages as per-joint-tile GEMMs (kernels_gcnw.cuh) --------------------
// The temporal convolution runs as the same GEMM as the graph convolution with the taps as edges (tmode 2),
// which has no limit on Gamma (the frame-tile kernel k_tcn_tc2p stages a window of FT + Gamma - 1 frames and
// stops at 15 taps) and writes its accumulators raw, which is what batch-statistics BatchNorm needs (moments
// over the whole call first, then normalise).  LayerNorm: x planes -> [GEMM -> z -> LN1+ReLU] -> u planes ->
// [tap GEMM -> q] -> LN2 + residual + ReLU (k_ln_stream).  BatchNorm: the same GEMMs, k_channel_stats +
// k_bn_apply in between (fp32 rows; rows -> planes conversions before each GEMM).
int gcnw_taps(int c, const __nv_bfloat16 *u, const __nv_bfloat16 *wp16, const float *bias, float *q, int N, int T,
              int T_out, int V, int kernel, int stride, int planes, long long plane_stride, cudaStream_t st) {
  STGCN_REQUIRE(stride == 1 || (stride == 2 && T >= 2), "temporal GEMM: stride must be 1 or 2 (got %d)", stride);
  tc::GcnwParams g{};
  g.T = T_out; g.V = V; g.Cin = c; g.planes = planes; g.N = N;
  g.bias = bias; g.bias_sw = 0;
  g.out = q;
  g.tmode = 2; g.ntaps = kernel; g.tpad = (kernel - 1) / 2; g.tstride = stride;
  ProfScope ps(KC_GEMM_TCN, st);
  if (tc::launch_gcnw(c, u, wp16, g, T, stride, kernel, plane_stride, st)) return 1;
  STGCN_LAUNCH_OK();
  return 0;
}

int layer_forward_taps(const stgcn_layer_desc &d, int K, int V, int math, const float *x, float *out, int N, int T,
                       Bump &ws, cudaStream_t st, const LayerPrep &P, bool x_planes, bool out_planes) {
  const size_t mark = ws.mark();
  const bool bn = d.norm == STGCN_NORM_BATCHNORM;
  const int planes = math == STGCN_MATH_BF16X3 ? 2 : 1;
  const int T_out = (T - 1) / d.stride + 1;
  const long long rows = (long long)N * T * V, rows_out = (long long)N * T_out * V;
  const int co = d.c_out, cap = tc::gcnw_edge_cap(V);
  STGCN_REQUIRE(!bn || (!x_planes && !out_planes), "BatchNorm layers exchange fp32 rows");
  // layer input as bf16 planes
  const __nv_bfloat16 *xh = reinterpret_cast<const __nv_bfloat16 *>(x);
  if (!x_planes) {
    __nv_bfloat16 *sc = ws.take<__nv_bfloat16>((size_t)planes * rows * d.c_in);
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (layer input planes)");
      const long long tot = rows * d.c_in;
      ProfScope ps(KC_LAYOUT, st);
      tc::k_rows_to_planes<<<cdiv(tot, 256), 256, 0, st>>>(x, sc, planes == 2 ? sc + tot : nullptr, tot);
      STGCN_LAUNCH_OK();
    }
    xh = sc;
  }
  const __nv_bfloat16 *xl = planes == 2 ? xh + (size_t)rows * d.c_in : nullptr;
  __nv_bfloat16 *u16 = ws.take<__nv_bfloat16>((size_t)planes * rows * co);
  __nv_bfloat16 *u16_lo = (planes == 2 && u16) ? u16 + (size_t)rows * co : nullptr;
  float *q = ws.take<float>((size_t)rows_out * co);
  const bool res_conv = d.residual == STGCN_RES_CONV;
  float *resb = res_conv ? ws.take<float>((size_t)rows_out * co) : nullptr;
  if (!bn) {
    // ---- LayerNorm ----
    tc::LnStreamArgs l{};
    l.frames = (long long)N * T; l.T = T; l.V = V; l.C = co;
    l.n_wT = P.n1T; l.n_bT = P.n1T + (size_t)co * V;
    l.relu = 1; l.eps = kEps;
    l.out_hi = u16; l.out_lo = u16_lo;
    tc::GcnwParams g{};
    g.T = T; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = N;
    g.tab = P.gwtab;
    g.bias = P.bzT; g.bias_sw = 1;
    if (gcnw_stage(co, xh, P.wsc, g, l, P.n1V, T, 1, rows * d.c_in, ws, st)) return 1;
    if (res_conv) {
      tc::LnStreamArgs lr{};
      lr.frames = (long long)N * T_out; lr.T = T_out; lr.V = V; lr.C = co;
      lr.n_wT = P.nrT; lr.n_bT = P.nrT + (size_t)co * V;
      lr.relu = 0; lr.eps = kEps;
      lr.out_f32 = resb;
      tc::GcnwParams gr{};
      gr.T = T_out; gr.V = V; gr.Cin = d.c_in; gr.planes = planes; gr.N = N;
      gr.tab = P.gwtabr;
      gr.bias = d.res_b; gr.bias_sw = 0;
      if (gcnw_stage(co, xh, P.wscr, gr, lr, P.nrV, T, d.stride, rows * d.c_in, ws, st)) return 1;
    }
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (layer, all-taps path)");
      if (gcnw_taps(co, u16, P.wp16, d.tcn_b, q, N, T, T_out, V, d.kernel, d.stride, planes, rows * co, st)) return 1;
      tc::LnStreamArgs f{};
      f.frames = (long long)N * T_out; f.T = T_out; f.V = V; f.C = co;
      f.z = q;
      f.n_wT = P.n2T; f.n_bT = P.n2T + (size_t)co * V;
      f.relu = 1; f.eps = kEps;
      if (d.residual == STGCN_RES_IDENTITY) { f.res_hi = xh; f.res_lo = xl; }
      else if (res_conv) f.res_f32 = resb;
      if (out_planes) {
        f.out_hi = reinterpret_cast<__nv_bfloat16 *>(out);
        f.out_lo = planes == 2 ? f.out_hi + (size_t)rows_out * co : nullptr;
      } else {
        f.out_f32 = out;
      }
      ProfScope ps(KC_FRAME, st);
      if (tc::launch_ln_stream(f, st)) return 1;
      STGCN_LAUNCH_OK();
    }
    ws.release(mark);
    return 0;
  }
  // ---- BatchNorm (batch statistics over the whole call, stgcn.py:152,160,171) ----
  float *z = ws.take<float>((size_t)rows * co);
  double *sums = ws.take<double>((size_t)4 * co);
  float2 *coef = ws.take<float2>((size_t)2 * co);
  if (!ws.measuring()) {
    STGCN_REQUIRE(!ws.overflow, "workspace too small (layer, all-taps BatchNorm path)");
    {
      tc::GcnwParams g{};
      g.T = T; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = N;
      g.tab = P.gwtab;
      g.bias = P.bzT; g.bias_sw = 1;
      g.out = z;
      ProfScope ps(KC_GEMM_1X1, st);
      if (tc::launch_gcnw(co, xh, P.wsc, g, T, 1, cap, rows * d.c_in, st)) return 1;
      STGCN_LAUNCH_OK();
    }
    if (channel_stats(z, rows, co, sums, st)) return 1;
    {
      ProfScope ps(KC_BN, st);
      k_bn_coeffs<<<cdiv(co, 128), 128, 0, st>>>(sums, sums + co, d.n1_w, d.n1_b, 1.0 / (double)rows, kEps, co, coef);
      STGCN_LAUNCH_OK();
      BnApply4Args b{};
      b.a = z; b.a_coef = coef; b.relu_out = 1; b.n4 = rows * co / 4; b.C = co;
      b.out_hi = u16; b.out_lo = u16_lo;                     // relu(BN1(z)) straight into the temporal GEMM's operand planes
      k_bn_apply4<<<148 * 8, 256, 0, st>>>(b);
      STGCN_LAUNCH_OK();
    }
    if (res_conv) {
      tc::GcnwParams g{};
      g.T = T_out; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = N;
      g.tab = P.gwtabr;
      g.bias = d.res_b; g.bias_sw = 0;
      g.out = resb;
      ProfScope ps(KC_GEMM_1X1, st);
      if (tc::launch_gcnw(co, xh, P.wscr, g, T, d.stride, cap, rows * d.c_in, st)) return 1;
      STGCN_LAUNCH_OK();
    }
    if (gcnw_taps(co, u16, P.wp16, d.tcn_b, q, N, T, T_out, V, d.kernel, d.stride, planes, rows * co, st)) return 1;
    if (channel_stats(q, rows_out, co, sums, st)) return 1;
    if (res_conv && channel_stats(resb, rows_out, co, sums + 2 * co, st)) return 1;
    ProfScope ps(KC_BN, st);
    k_bn_coeffs<<<cdiv(co, 128), 128, 0, st>>>(sums, sums + co, d.n2_w, d.n2_b, 1.0 / (double)rows_out, kEps, co, coef);
    STGCN_LAUNCH_OK();
    BnApply4Args b{};
    b.a = q; b.a_coef = coef;
    if (d.residual == STGCN_RES_IDENTITY) { b.b_mode = B_RAW; b.b = x; }
    else if (res_conv) {
      k_bn_coeffs<<<cdiv(co, 128), 128, 0, st>>>(sums + 2 * co, sums + 3 * co, d.nr_w, d.nr_b, 1.0 / (double)rows_out,
                                                kEps, co, coef + co);
      STGCN_LAUNCH_OK();
      b.b_mode = B_LN; b.b = resb; b.b_coef = coef + co;
    }
    b.relu_out = 1; b.n4 = rows_out * co / 4; b.C = co;
    b.out = out;
    k_bn_apply4<<<148 * 8, 256, 0, st>>>(b);
    STGCN_LAUNCH_OK();
  }
  ws.release(mark);
  return 0;
}

// ---- ST-GCN layer on channels-last activations --------------------------------
// x [N*T*V, c_in] -> out [N*T_out*V, c_out].  Scratch comes from `ws` (released on return).
// `pp`: prepared operands (model path) or null (built here, per call).
constexpr int kHalo = 4;   // halo frames carried on each side of the temporal-conv input in T-split mode

// x_planes / out_planes: the buffer holds bf16 hi/lo planes [plane][rows][C] (one plane in bf16
// mode) instead of fp32 rows -- the inter-layer format of the per-joint-weight graph-conv path.
int layer_forward_ntvc(const stgcn_layer_desc &d, int K, int V, int math, const float *x, float *out,
                       int N, int T, Bump &ws, cudaStream_t st, const LayerPrep *pp = nullptr,
                       const stgcn_halo_desc *halo = nullptr, int layer_index = 0, bool x_planes = false,
                       bool out_planes = false, bool sparse_adj = false) {
  if (check_layer(d)) return 1;
  const size_t mark = ws.mark();
  const int T_out = (T - 1) / d.stride + 1;
  const long long rows = (long long)N * T * V, rows_out = (long long)N * T_out * V;
  const bool bn = d.norm == STGCN_NORM_BATCHNORM;
  LayerPrep local;
  if (!pp && math != STGCN_MATH_FP32) {
    local = prep_take(d, K, V, ws, sparse_adj);
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (layer operands)");
      if (prep_run(d, K, V, local, st)) return 1;
    }
    pp = &local;
  }
  if (math != STGCN_MATH_FP32 && pp && pp->taps && !halo) {
    const int r = layer_forward_taps(d, K, V, math, x, out, N, T, ws, st, *pp, x_planes, out_planes);
    ws.release(mark);
    return r;
  }
  // tensor-core temporal stage: LayerNorm, stride 1/2, C in {64,128,256}
  const bool tc_tcn = math != STGCN_MATH_FP32 && pp && pp->tcn &&
                      tc::tcn_tc2_supported(d.c_out, V, d.kernel, d.stride, T);
  const int planes = math == STGCN_MATH_BF16X3 ? 2 : 1;
  // tensor-core graph-convolution stage: LayerNorm, shared adjacency, C_in % 64 == 0
  const bool tc_gcn = math != STGCN_MATH_FP32 && pp && pp->gcn;
  // per-joint-weight GEMM path: the layer input (and the identity residual) travel as bf16 planes
  const bool use_gw = tc_gcn && tc_tcn && pp->gw;
  STGCN_REQUIRE(use_gw || (!x_planes && !out_planes), "bf16-plane activations need the per-joint-weight graph-conv path");
  const __nv_bfloat16 *xh = nullptr, *xl = nullptr;
  if (use_gw) {
    if (x_planes) {
      xh = reinterpret_cast<const __nv_bfloat16 *>(x);
    } else {
      __nv_bfloat16 *sc = ws.take<__nv_bfloat16>((size_t)planes * rows * d.c_in);
      if (!ws.measuring()) {
        STGCN_REQUIRE(!ws.overflow, "workspace too small (layer input planes)");
        const long long tot = rows * d.c_in;
        ProfScope ps(KC_LAYOUT, st);
        tc::k_rows_to_planes<<<cdiv(tot, 256), 256, 0, st>>>(x, sc, planes == 2 ? sc + tot : nullptr, tot);
        STGCN_LAUNCH_OK();
      }
      xh = sc;
    }
    xl = (planes == 2 && xh) ? xh + (size_t)rows * d.c_in : nullptr;
  }

  const int hf = halo ? kHalo : 0;                          // halo frames per side
  if (halo) {
    STGCN_REQUIRE(tc_gcn && tc_tcn, "T-split needs the tensor-core path (LayerNorm, math != fp32, C in {64,128,256})");
    STGCN_REQUIRE((d.kernel - 1) / 2 <= kHalo && T >= kHalo, "T-split: kernel %d / chunk of %d frames unsupported",
                  d.kernel, T);
  }
  const long long rows_u = (long long)N * (T + 2 * hf) * V;  // rows of the temporal-conv input buffer
  float *u = tc_tcn ? nullptr : ws.take<float>((size_t)rows * d.c_out);
  __nv_bfloat16 *u16 = tc_tcn ? ws.take<__nv_bfloat16>((size_t)planes * rows_u * d.c_out) : nullptr;
  __nv_bfloat16 *u16_lo = (tc_tcn && planes == 2 && u16) ? u16 + (size_t)rows_u * d.c_out : nullptr;
  double *sums = bn ? ws.take<double>((size_t)4 * d.c_out) : nullptr;
  if (use_gw) {
    // GEMM with per-joint pre-scaled weights; LayerNorm + ReLU + split either inside the same persistent
    // kernel (z travels through an L2-resident ring) or, STGCN_GCNW_FUSE=0, as the streaming kernel
    tc::LnStreamArgs l{};
    l.frames = (long long)N * T; l.T = T; l.V = V; l.C = d.c_out;
    l.n_wT = pp->n1T; l.n_bT = pp->n1T + (size_t)d.c_out * V;
    l.relu = 1; l.eps = kEps;
    l.out_hi = u16; l.out_lo = u16_lo;
    if (hf) { l.out_T = T + 2 * hf; l.out_t0 = hf; }
    tc::GcnwParams g{};
    g.T = T; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = N;
    g.tab = pp->gwtab;
    g.bias = pp->bzT; g.bias_sw = 1;
    g.debug = debug_mode();
    if (gcnw_stage(d.c_out, xh, pp->wsc, g, l, pp->n1V, T, 1, rows * d.c_in, ws, st)) return 1;
  } else if (tc_gcn) {
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (layer gcn stage)");
      tc::GcnTc2Params g{};
      g.T_out = T; g.V = V; g.K = K; g.Cin = d.c_in; g.planes = planes;
      g.csr_ptr = pp->kw_ptr; g.csr_va = pp->kw_va;
      g.epi.bias = pp->bzT; g.epi.bias_sw = 1;
      g.epi.n_wT = pp->n1T; g.epi.n_bT = pp->n1T + (size_t)d.c_out * V;
      g.epi.out_hi = u16; g.epi.out_lo = u16_lo; g.epi.out_f32 = u;
      if (hf) { g.epi.out_T = T + 2 * hf; g.epi.out_t0 = hf; }
      g.epi.relu = 1;
      g.epi.eps = kEps;
      g.epi.debug = debug_mode();
      ProfScope ps(KC_GEMM_1X1, st);
      if (tc::launch_gcn_tc2(d.c_out, x, pp->wg16, g, N, T, 1, st)) return 1;
      STGCN_LAUNCH_OK();
      if (debug_dump("gcn", d.c_out, st)) return 1;
    }
  } else {
    AdjCsr csr;
    if (build_csr(d.a_eff, d.a_per_sample, N, K, V, d.c_out, ws, csr, st)) return 1;
    const size_t m2 = ws.mark();
    float *y = ws.take<float>((size_t)rows * K * d.c_out);
    float *z = bn ? ws.take<float>((size_t)rows * d.c_out) : nullptr;
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (layer gcn stage)");
      if (launch_gemm(x, d.gcn_w, d.gcn_b, y, N, T, V, d.c_in, K * d.c_out, 1, 1, st)) return 1;
      FrameArgs a{};
      a.producer = FRAME_ADJ;
      a.frames = (long long)N * T;
      a.frames_per_sample = T;
      a.K = K; a.V = V; a.C = d.c_out;
      a.y = y;
      a.adj_ptr = csr.ptr; a.adj_yoff = csr.yoff; a.adj_val = csr.val;
      a.adj_per_sample = d.a_per_sample;
      a.eps = kEps;
      if (!bn) {
        a.norm_a = 1; a.na_w = d.n1_w; a.na_b = d.n1_b; a.relu_out = 1; a.out = u;
        a.out_hi = u16;
        a.out_lo = u16_lo;
        if (launch_frame(a, st)) return 1;
      } else {
        a.out = z;
        if (launch_frame(a, st)) return 1;
        if (channel_stats(z, rows, d.c_out, sums, st)) return 1;
        BnApplyArgs b{};
        b.a = z; b.a_sum = sums; b.a_sumsq = sums + d.c_out; b.a_w = d.n1_w; b.a_b = d.n1_b;
        b.relu_out = 1; b.rows = rows; b.C = d.c_out; b.inv_count = 1.0 / (double)rows; b.eps = kEps;
        b.out = u;
        ProfScope ps(KC_BN, st);
        k_bn_apply<<<148 * 8, 256, 0, st>>>(b);
        STGCN_LAUNCH_OK();
      }
    }
    ws.release(m2);
  }
  if (tc_tcn) {
    // channel-changing / strided residual: LN_R(conv1x1_stride(x)) precomputed into `resb`
    const bool res_conv = d.residual == STGCN_RES_CONV;
    const bool res_tc = res_conv && pp->res;
    float *resb = res_conv ? ws.take<float>((size_t)rows_out * d.c_out) : nullptr;
    float *qr = (res_conv && !res_tc) ? ws.take<float>((size_t)rows_out * d.c_out) : nullptr;
    if (res_tc && use_gw) {
      // residual 1x1 conv + LayerNorm_R with the same stage (identity edges, strided frames)
      tc::LnStreamArgs l{};
      l.frames = (long long)N * T_out; l.T = T_out; l.V = V; l.C = d.c_out;
      l.n_wT = pp->nrT; l.n_bT = pp->nrT + (size_t)d.c_out * V;
      l.relu = 0; l.eps = kEps;
      l.out_f32 = resb;
      tc::GcnwParams g{};
      g.T = T_out; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = N;
      g.tab = pp->gwtabr;
      g.bias = d.res_b; g.bias_sw = 0;
      g.debug = debug_mode();
      if (gcnw_stage(d.c_out, xh, pp->wscr, g, l, pp->nrV, T, d.stride, rows * d.c_in, ws, st)) return 1;
    }
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (layer tcn stage)");
      if (res_tc && use_gw) {
      } else if (res_tc) {
        tc::GcnTc2Params g{};
        g.T_out = T_out; g.V = V; g.K = 1; g.Cin = d.c_in; g.planes = planes;
        g.identity = 1;
        g.epi.bias = d.res_b; g.epi.bias_sw = 0;
        g.epi.n_wT = pp->nrT; g.epi.n_bT = pp->nrT + (size_t)d.c_out * V;
        g.epi.out_f32 = resb;
        g.epi.relu = 0;
        g.epi.eps = kEps;
        g.epi.debug = debug_mode();
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcn_tc2(d.c_out, x, pp->wr16, g, N, T, d.stride, st)) return 1;
        STGCN_LAUNCH_OK();
      } else if (res_conv) {
        if (launch_gemm(x, d.res_w, d.res_b, qr, N, T, V, d.c_in, d.c_out, 1, d.stride, st)) return 1;
        FrameArgs a{};
        a.producer = FRAME_LOAD;
        a.frames = (long long)N * T_out;
        a.frames_per_sample = T_out;
        a.K = K; a.V = V; a.C = d.c_out;
        a.a = qr;
        a.norm_a = 1; a.na_w = d.nr_w; a.na_b = d.nr_b;
        a.eps = kEps;
        a.out = resb;
        if (launch_frame(a, st)) return 1;
      }
      if (halo) {
        // boundary frames of u: pack -> exchange with the neighbouring ranks -> unpack (or zero padding)
        const size_t frame_b = (size_t)V * d.c_out * sizeof(__nv_bfloat16);
        const size_t pitch = (size_t)(T + 2 * hf) * frame_b, width = (size_t)hf * frame_b;
        const size_t segs = (size_t)planes * N, bytes = width * segs;
        STGCN_REQUIRE(bytes <= halo->capacity, "T-split: halo staging buffers too small (%zu B needed)", bytes);
        char *ub = reinterpret_cast<char *>(u16);
        if (halo->has_left)
          STGCN_CUDA_OK(cudaMemcpy2DAsync(halo->send_left, width, ub + (size_t)hf * frame_b, pitch, width, segs,
                                          cudaMemcpyDeviceToDevice, st));
        if (halo->has_right)
          STGCN_CUDA_OK(cudaMemcpy2DAsync(halo->send_right, width, ub + (size_t)T * frame_b, pitch, width, segs,
                                          cudaMemcpyDeviceToDevice, st));
        if (halo->has_left || halo->has_right) {
          STGCN_REQUIRE(halo->exchange, "T-split: exchange callback missing");
          STGCN_REQUIRE(halo->exchange(halo->ctx, layer_index, bytes) == 0, "T-split: halo exchange failed (layer %d)",
                        layer_index);
        }
        if (halo->has_left)
          STGCN_CUDA_OK(cudaMemcpy2DAsync(ub, pitch, halo->recv_left, width, width, segs, cudaMemcpyDeviceToDevice, st));
        else
          STGCN_CUDA_OK(cudaMemset2DAsync(ub, pitch, 0, width, segs, st));
        char *right = ub + (size_t)(hf + T) * frame_b;
        if (halo->has_right)
          STGCN_CUDA_OK(cudaMemcpy2DAsync(right, pitch, halo->recv_right, width, width, segs, cudaMemcpyDeviceToDevice, st));
        else
          STGCN_CUDA_OK(cudaMemset2DAsync(right, pitch, 0, width, segs, st));
      }
      tc::TcnTc2Params p{};
      p.T_out = T_out; p.V = V; p.G = d.kernel;
      p.planes = planes;
      p.epi.bias = d.tcn_b; p.epi.bias_sw = 0;
      p.epi.n_wT = pp->n2T; p.epi.n_bT = pp->n2T + (size_t)d.c_out * V;
      if (d.residual == STGCN_RES_IDENTITY && use_gw) {
        p.epi.res_hi = xh; p.epi.res_lo = xl;            // the layer input as planes (hi + lo = x to 2^-17)
      } else {
        p.epi.res = d.residual == STGCN_RES_IDENTITY ? x : resb;
      }
      if (out_planes) {
        p.epi.out_hi = reinterpret_cast<__nv_bfloat16 *>(out);
        p.epi.out_lo = planes == 2 ? p.epi.out_hi + (size_t)rows_out * d.c_out : nullptr;
      } else {
        p.epi.out_f32 = out;
      }
      p.epi.relu = 1;
      p.epi.eps = kEps;
      p.epi.debug = debug_mode();
      ProfScope ps(KC_GEMM_TCN, st);
      if (tc::launch_tcn_tc2(d.c_out, u16, pp->wp16, p, N, T, d.stride, hf, st)) return 1;
      STGCN_LAUNCH_OK();
      if (debug_dump("tcn", d.c_out, st)) return 1;
    }
    ws.release(mark);
    return 0;
  }
  float *wp = ws.take<float>((size_t)d.c_out * d.c_out * d.kernel);
  float *q = ws.take<float>((size_t)rows_out * d.c_out);
  float *qr = d.residual == STGCN_RES_CONV ? ws.take<float>((size_t)rows_out * d.c_out) : nullptr;
  if (!ws.measuring()) {
    STGCN_REQUIRE(!ws.overflow, "workspace too small (layer tcn stage)");
    long long nw = (long long)d.c_out * d.c_out * d.kernel;
    {
      ProfScope ps(KC_MISC, st);
      k_pack_tcn_w<<<cdiv(nw, 256), 256, 0, st>>>(d.tcn_w, wp, d.c_out, d.c_out, d.kernel, d.c_out, d.c_out);
      STGCN_LAUNCH_OK();
    }
    if (launch_gemm(u, wp, d.tcn_b, q, N, T, V, d.c_out, d.c_out, d.kernel, d.stride, st)) return 1;
    if (qr && launch_gemm(x, d.res_w, d.res_b, qr, N, T, V, d.c_in, d.c_out, 1, d.stride, st)) return 1;
    if (!bn) {
      FrameArgs a{};
      a.producer = FRAME_LOAD;
      a.frames = (long long)N * T_out;
      a.frames_per_sample = T_out;
      a.K = K; a.V = V; a.C = d.c_out;
      a.a = q;
      a.norm_a = 1; a.na_w = d.n2_w; a.na_b = d.n2_b;
      a.eps = kEps;
      if (d.residual == STGCN_RES_IDENTITY) { a.b_mode = B_RAW; a.b = x; }
      else if (d.residual == STGCN_RES_CONV) { a.b_mode = B_LN; a.b = qr; a.nb_w = d.nr_w; a.nb_b = d.nr_b; }
      a.relu_out = 1;
      a.out = out;
      if (launch_frame(a, st)) return 1;
    } else {
      if (channel_stats(q, rows_out, d.c_out, sums, st)) return 1;
      BnApplyArgs b{};
      b.a = q; b.a_sum = sums; b.a_sumsq = sums + d.c_out; b.a_w = d.n2_w; b.a_b = d.n2_b;
      if (d.residual == STGCN_RES_IDENTITY) { b.b_mode = B_RAW; b.b = x; }
      else if (d.residual == STGCN_RES_CONV) {
        if (channel_stats(qr, rows_out, d.c_out, sums + 2 * d.c_out, st)) return 1;
        b.b_mode = B_LN; b.b = qr; b.b_sum = sums + 2 * d.c_out; b.b_sumsq = sums + 3 * d.c_out;
        b.b_w = d.nr_w; b.b_b = d.nr_b;
      }
      b.relu_out = 1; b.rows = rows_out; b.C = d.c_out; b.inv_count = 1.0 / (double)rows_out; b.eps = kEps;
      b.out = out;
      ProfScope ps(KC_BN, st);
      k_bn_apply<<<148 * 8, 256, 0, st>>>(b);
      STGCN_LAUNCH_OK();
    }
  }
  ws.release(mark);
  return 0;
}

// STGCN_RT_SPLIT=0 selects the fused step (state update inside the GEMM kernel's epilogue)
inline bool rt_split_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_RT_SPLIT");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}

// STGCN_RT_STREAM=0 selects the register-resident state kernel (k_rt_update) instead of the bulk-copy-staged
// one (k_rt_stream)
inline bool rt_stream_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_RT_STREAM");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}
int launch_rt_state(const RtUpdateArgs &u, cudaStream_t st) {
  if (rt_stream_enabled() && rt_stream_supported(u.V, u.C)) return launch_rt_stream(u, st);
  if (u.fifo_bf16) return fail("rt step: the bf16 FIFO layout needs the staged state kernel (STGCN_RT_STREAM)");
  return launch_rt_update(u, st);
}

// ---- RT online layer on channels-last frames -----------------------------------
// x [B*V, c_in] -> out [B*V, c_out]; fifo [F][B][V][C], acc [S][B][V][C]; counter[B].
int rt_layer_step_ntvc(const stgcn_layer_desc &d, int K, int V, int math, const float *x, float *out,
                       float *fifo, float *acc, const int *counter, int B, Bump &ws, cudaStream_t st,
                       const LayerPrep *pp = nullptr, bool use_gw = false, bool x_planes = false,
                       bool out_planes = false, bool fifo_bf16 = false, long long slot_rows = 0,
                       int gw_smem_cap = 0, cudaEvent_t gemm_wait = nullptr, cudaEvent_t gemm_done = nullptr,
                       bool pool_out = false) {
  // pool_out (per-joint-weight path, last layer): `out` receives the mean over the joints [B, c_out] instead of rows
  // slot_rows: rows (streams * V) between consecutive FIFO / accumulator slots when the call covers only a range
  // of the streams the state was laid out for (0 = the B of this call); gw_smem_cap: see GcnwParams::smem_cap
  if (check_layer(d)) return 1;
  STGCN_REQUIRE(!slot_rows || use_gw, "rt layer: stream ranges need the per-joint-weight GEMM path");
  STGCN_REQUIRE(!fifo_bf16 || use_gw, "rt layer: the bf16 FIFO layout needs the per-joint-weight GEMM path");
  STGCN_REQUIRE(d.norm == STGCN_NORM_LAYERNORM,
                "continual inference needs LayerNorm: batch statistics of a single frame are undefined "
                "(reference raises at models/utils/batchnorm.py:20)");
  LayerPrep local;
  const size_t mark0 = ws.mark();
  if (!pp && math != STGCN_MATH_FP32) {
    local = prep_take(d, K, V, ws);
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (rt layer operands)");
      if (prep_run(d, K, V, local, st)) return 1;
    }
    pp = &local;
  }
  if (math != STGCN_MATH_FP32 && pp && pp->gcn && (d.residual != STGCN_RES_CONV || pp->res) && rt_split_enabled() &&
      rt_update_supported(V, d.c_out)) {
    // ---- split step: tensor-core GEMM (+ adjacency) writes z = gcn(x) raw (it stays in L2), then the
    // streaming kernel k_rt_update does the FIFO / accumulator update, LayerNorm and residual ----
    const int planes = math == STGCN_MATH_BF16X3 ? 2 : 1;
    const long long rows = (long long)B * V;
    float *zb = ws.take<float>((size_t)rows * d.c_out);
    float *qr = d.residual == STGCN_RES_CONV ? ws.take<float>((size_t)rows * d.c_out) : nullptr;
    STGCN_REQUIRE(use_gw || (!x_planes && !out_planes), "rt layer: bf16-plane activations need the per-joint-weight GEMM");
    if (use_gw && !ws.measuring()) {
      // ---- GEMM with per-joint pre-scaled weights (kernels_gcnw.cuh): streams are the frames of one
      // trial, the input arrives as bf16 planes ----
      STGCN_REQUIRE(!ws.overflow, "workspace too small (rt layer)");
      STGCN_REQUIRE(pp->gw && x_planes, "rt layer: per-joint-weight GEMM needs prepared operands and plane input");
      const __nv_bfloat16 *xh = reinterpret_cast<const __nv_bfloat16 *>(x);
      const __nv_bfloat16 *xl = planes == 2 ? xh + (size_t)rows * d.c_in : nullptr;
      const int cap = tc::gcnw_edge_cap(V);
      // two-half step: this half's GEMMs start when the other half's GEMMs of the same phase are done, so that a
      // GEMM always shares the SMs with the other half's state kernel
      if (gemm_wait) STGCN_CUDA_OK(cudaStreamWaitEvent(st, gemm_wait, 0));
      if (qr) {
        tc::GcnwParams g{};
        g.T = B; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = 1;
        g.tab = pp->gwtabr;
        g.out = qr;                            // no bias (rtstgcn.py:503)
        g.smem_cap = gw_smem_cap;
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcnw(d.c_out, xh, pp->wscr, g, B, 1, cap, rows * d.c_in, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      {
        tc::GcnwParams g{};
        g.T = B; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = 1;
        g.tab = pp->gwtab;
        g.bias = pp->bzT; g.bias_sw = 1;
        g.out = zb;
        g.smem_cap = gw_smem_cap;
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcnw(d.c_out, xh, pp->wsc, g, B, 1, cap, rows * d.c_in, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      if (gemm_done) STGCN_CUDA_OK(cudaEventRecord(gemm_done, st));
      RtUpdateArgs u{};
      u.B = B; u.V = V; u.C = d.c_out;
      u.z = zb;
      u.debug = debug_mode();
      u.fifo = fifo_bf16 ? nullptr : fifo; u.acc = acc; u.counter = counter;
      u.fifo16 = fifo_bf16 ? reinterpret_cast<__nv_bfloat16 *>(fifo) : nullptr;
      u.fifo_bf16 = fifo_bf16 ? 1 : 0;
      u.F = d.stride * (d.kernel - 1) + 1;
      u.S = d.stride;
      u.slot = (slot_rows ? slot_rows : rows) * d.c_out;
      u.n_wT = pp->n1T; u.n_bT = pp->n1T + (size_t)d.c_out * V;
      if (d.residual == STGCN_RES_IDENTITY) { u.res_mode = 1; u.res_hi = xh; u.res_lo = xl; }
      else if (d.residual == STGCN_RES_CONV) {
        u.res_mode = 2; u.res = qr;
        u.r_wT = pp->nrT; u.r_bT = pp->nrT + (size_t)d.c_out * V;
      }
      u.eps = kEps;
      if (debug_mode() & 256) { u.res_mode = 0; }                     // timing experiments only
      if (pool_out) {
        STGCN_REQUIRE(!out_planes && rt_stream_enabled() && rt_stream_pool_supported(V, d.c_out),
                      "rt layer: pooled output needs the staged state kernel");
        u.pool_out = out;
      } else if (out_planes && !(debug_mode() & 512)) {
        u.out_hi = reinterpret_cast<__nv_bfloat16 *>(out);
        u.out_lo = planes == 2 ? u.out_hi + (size_t)rows * d.c_out : nullptr;
      } else {
        u.out = out;
      }
      ProfScope ps(KC_FRAME, st);
      if (launch_rt_state(u, st)) return 1;
      STGCN_LAUNCH_OK();
      ws.release(mark0);
      return 0;
    }
    if (use_gw) {          // measuring pass: same workspace as above
      ws.release(mark0);
      return 0;
    }
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (rt layer)");
      if (qr) {
        // residual branch conv1x1(x): no bias, no stride (rtstgcn.py:503); its LayerNorm runs in k_rt_update
        tc::GcnTc2Params g{};
        g.T_out = B; g.V = V; g.K = 1; g.Cin = d.c_in; g.planes = planes;
        g.identity = 1;
        g.epi.bias = pp->zero; g.epi.bias_sw = 0;
        g.epi.out_f32 = qr;
        g.epi.raw = 1;
        g.epi.debug = debug_mode();
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcn_tc2(d.c_out, x, pp->wr16, g, 1, B, 1, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      {
        tc::GcnTc2Params g{};
        g.T_out = B; g.V = V; g.K = K; g.Cin = d.c_in; g.planes = planes;
        g.csr_ptr = pp->kw_ptr; g.csr_va = pp->kw_va;
        g.epi.bias = pp->bzT; g.epi.bias_sw = 1;
        g.epi.out_f32 = zb;
        g.epi.raw = 1;
        g.epi.debug = debug_mode();
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcn_tc2(d.c_out, x, pp->wg16, g, 1, B, 1, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      RtUpdateArgs u{};
      u.B = B; u.V = V; u.C = d.c_out;
      u.z = zb;
      u.fifo = fifo; u.acc = acc; u.counter = counter;
      u.F = d.stride * (d.kernel - 1) + 1;
      u.S = d.stride;
      u.slot = rows * d.c_out;
      u.n_wT = pp->n1T; u.n_bT = pp->n1T + (size_t)d.c_out * V;
      if (d.residual == STGCN_RES_IDENTITY) { u.res_mode = 1; u.res = x; }
      else if (d.residual == STGCN_RES_CONV) {
        u.res_mode = 2; u.res = qr;
        u.r_wT = pp->nrT; u.r_bT = pp->nrT + (size_t)d.c_out * V;
      }
      u.eps = kEps;
      u.out = out;
      ProfScope ps(KC_FRAME, st);
      if (launch_rt_state(u, st)) return 1;
      STGCN_LAUNCH_OK();
    }
    ws.release(mark0);
    return 0;
  }
  if (math != STGCN_MATH_FP32 && pp && pp->gcn && (d.residual != STGCN_RES_CONV || pp->res)) {
    // ---- tensor-core step: the B streams form one "trial" of B frames (rows (b, w)) ----
    const int planes = math == STGCN_MATH_BF16X3 ? 2 : 1;
    const long long rows = (long long)B * V;
    float *resb = d.residual == STGCN_RES_CONV ? ws.take<float>((size_t)rows * d.c_out) : nullptr;
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "workspace too small (rt layer)");
      if (resb) {
        // residual branch LN_R(conv1x1(x)): no bias, no stride (rtstgcn.py:503)
        tc::GcnTc2Params g{};
        g.T_out = B; g.V = V; g.K = 1; g.Cin = d.c_in; g.planes = planes;
        g.identity = 1;
        g.epi.bias = pp->zero; g.epi.bias_sw = 0;
        g.epi.n_wT = pp->nrT; g.epi.n_bT = pp->nrT + (size_t)d.c_out * V;
        g.epi.out_f32 = resb;
        g.epi.relu = 0;
        g.epi.eps = kEps;
        g.epi.debug = debug_mode();
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcn_tc2(d.c_out, x, pp->wr16, g, 1, B, 1, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      tc::GcnTc2Params g{};
      g.T_out = B; g.V = V; g.K = K; g.Cin = d.c_in; g.planes = planes;
      g.csr_ptr = pp->kw_ptr; g.csr_va = pp->kw_va;
      g.epi.bias = pp->bzT; g.epi.bias_sw = 1;
      g.epi.n_wT = pp->n1T; g.epi.n_bT = pp->n1T + (size_t)d.c_out * V;
      g.epi.res = d.residual == STGCN_RES_IDENTITY ? x : resb;
      g.epi.out_f32 = out;
      g.epi.relu = 1;                       // relu(relu(LN(o)) + res); idempotent without a residual
      g.epi.eps = kEps;
      g.epi.debug = debug_mode();
      g.epi.rt_fifo = fifo; g.epi.rt_acc = acc; g.epi.rt_counter = counter;
      g.epi.rt_F = d.stride * (d.kernel - 1) + 1;
      g.epi.rt_S = d.stride;
      g.epi.rt_slot = rows * d.c_out;
      ProfScope ps(KC_FRAME, st);
      if (tc::launch_gcn_tc2(d.c_out, x, pp->wg16, g, 1, B, 1, st)) return 1;
      STGCN_LAUNCH_OK();
      if (debug_dump("rt", d.c_out, st)) return 1;
    }
    ws.release(mark0);
    return 0;
  }
  const size_t mark = ws.mark();
  AdjCsr csr;
  if (build_csr(d.a_eff, 0, B, K, V, d.c_out, ws, csr, st)) return 1;
  const long long rows = (long long)B * V;
  float *y = ws.take<float>((size_t)rows * K * d.c_out);
  float *qr = d.residual == STGCN_RES_CONV ? ws.take<float>((size_t)rows * d.c_out) : nullptr;
  if (!ws.measuring()) {
    STGCN_REQUIRE(!ws.overflow, "workspace too small (rt layer)");
    if (launch_gemm(x, d.gcn_w, d.gcn_b, y, B, 1, V, d.c_in, K * d.c_out, 1, 1, st)) return 1;
    // residual conv of the online layer has no bias and no stride (rtstgcn.py:503)
    if (qr && launch_gemm(x, d.res_w, nullptr, qr, B, 1, V, d.c_in, d.c_out, 1, 1, st)) return 1;
    FrameArgs a{};
    a.producer = FRAME_RT;
    a.frames = B;
    a.frames_per_sample = 1;
    a.K = K; a.V = V; a.C = d.c_out;
    a.y = y;
    a.adj_ptr = csr.ptr; a.adj_yoff = csr.yoff; a.adj_val = csr.val;
    a.norm_a = 1; a.na_w = d.n1_w; a.na_b = d.n1_b;
    a.relu_mid = 1;
    if (d.residual == STGCN_RES_IDENTITY) { a.b_mode = B_RAW; a.b = x; }
    else if (d.residual == STGCN_RES_CONV) { a.b_mode = B_LN; a.b = qr; a.nb_w = d.nr_w; a.nb_b = d.nr_b; }
    a.relu_out = 1;
    a.eps = kEps;
    a.out = out;
    a.fifo = fifo; a.acc = acc; a.counter = counter;
    a.F = d.stride * (d.kernel - 1) + 1;
    a.S = d.stride;
    if (launch_frame(a, st)) return 1;
  }
  ws.release(mark0);
  return 0;
}

// pooled_sums != null: write the per-trial channel sums (N, C) instead of the classifier output
int pool_fc(const float *x, int N, long long R, int C, const float *W, const float *bias, int classes,
            float *logits, Bump &ws, cudaStream_t st, float *pooled_sums = nullptr) {
  const int rpc = 512;
  const int nchunk = cdiv(R, rpc);
  float *part = ws.take<float>((size_t)N * nchunk * C);
  if (ws.measuring()) return 0;
  STGCN_REQUIRE(!ws.overflow, "workspace too small (pool)");
  ProfScope ps(KC_POOL, st);
  k_pool_partial<<<dim3(nchunk, N), 256, 0, st>>>(x, R, C, rpc, nchunk, part);
  STGCN_LAUNCH_OK();
  if (pooled_sums) {
    k_pool_sum<<<cdiv((long long)N * C, 256), 256, 0, st>>>(part, nchunk, C, N, pooled_sums);
    STGCN_LAUNCH_OK();
    return 0;
  }
  k_pool_fc<<<N, 256, C * sizeof(float), st>>>(part, nchunk, C, 1.f / (float)R, W, bias, classes, logits);
  STGCN_LAUNCH_OK();
  return 0;
}

// input stage: x (N,C_in,T,V) -> h0 [N*T*V, C0]
// xs: optional element strides (trial, channel, frame) of `x`; null = contiguous (N, C_in, T, V)
inline bool embed_warp_path(const stgcn_model_desc &m) {
  return m.norm == STGCN_NORM_LAYERNORM && m.num_joints * m.in_feat <= 128 && m.layers[0].c_in % 4 == 0;
}
// out_planes: 0 = fp32 rows, 1 / 2 = bf16 hi (/ lo) planes in the same buffer (warp path only)
int embed(const stgcn_model_desc &m, const float *x, float *h0, int N, int T, Bump &ws, cudaStream_t st,
          const long long *xs = nullptr, int out_planes = 0) {
  const int V = m.num_joints, Ci = m.in_feat, C0 = m.layers[0].c_in;
  const long long frames = (long long)N * T;
  STGCN_REQUIRE(!out_planes || embed_warp_path(m), "embed: bf16-plane output needs the LayerNorm input stage");
  if (embed_warp_path(m)) {
    // one warp per frame, straight from the reference layout (no layout pass, no staging buffer)
    if (ws.measuring()) return 0;
    EmbedWarpArgs e{};
    e.x = x; e.N = N; e.T = T; e.V = V; e.C_in = Ci; e.C0 = C0;
    e.sn = xs ? xs[0] : (long long)Ci * T * V;
    e.sc = xs ? xs[1] : (long long)T * V;
    e.st = xs ? xs[2] : (long long)V;
    e.n_w = m.norm_in_w; e.n_b = m.norm_in_b; e.eps = kEps;
    e.W = m.fcn_in_w; e.bias = m.fcn_in_b; e.out = h0;
    if (out_planes) {
      e.out = nullptr;
      e.out_hi = reinterpret_cast<__nv_bfloat16 *>(h0);
      e.out_lo = out_planes == 2 ? e.out_hi + (size_t)frames * V * C0 : nullptr;
    }
    const size_t smem = sizeof(float) * ((size_t)C0 * Ci + C0 + (size_t)8 * V * Ci);
    STGCN_REQUIRE(smem <= 48 * 1024, "embed: input feature map too large");
    long long blocks = (frames + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    ProfScope ps(KC_EMBED, st);
    STGCN_CUDA_OK(launch_pdl(k_embed_warp, dim3((unsigned)blocks), dim3(256), smem, st, e));
    STGCN_LAUNCH_OK();
    return 0;
  }
  STGCN_REQUIRE(!xs, "strided (sliding-window) input needs the LayerNorm input stage");
  const size_t mark = ws.mark();
  float *xin = ws.take<float>((size_t)frames * V * Ci);
  double *sums = m.norm == STGCN_NORM_BATCHNORM ? ws.take<double>((size_t)2 * V * Ci) : nullptr;
  if (!ws.measuring()) {
    STGCN_REQUIRE(!ws.overflow, "workspace too small (embed)");
    if (to_ntvc(x, xin, N, Ci, (long long)T * V, Ci, st)) return 1;
    EmbedArgs e{};
    e.x = xin; e.frames = frames; e.V = V; e.C_in = Ci; e.C0 = C0; e.norm = m.norm;
    e.n_w = m.norm_in_w; e.n_b = m.norm_in_b; e.eps = kEps;
    e.W = m.fcn_in_w; e.bias = m.fcn_in_b; e.out = h0;
    if (m.norm == STGCN_NORM_BATCHNORM) {
      if (channel_stats(xin, frames, V * Ci, sums, st)) return 1;
      e.bn_sum = sums; e.bn_sumsq = sums + V * Ci; e.bn_inv_count = 1.0 / (double)frames;
    }
    size_t smem = sizeof(float) * ((size_t)3 * V * Ci + (size_t)C0 * Ci + C0);
    STGCN_REQUIRE(smem <= 160 * 1024, "embed: input feature map too large");
    if (smem > 48 * 1024)
      STGCN_CUDA_OK(cudaFuncSetAttribute(k_embed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = frames < 148 * 8 ? frames : 148 * 8;
    ProfScope ps(KC_EMBED, st);
    k_embed<<<(unsigned)blocks, 256, smem, st>>>(e);
    STGCN_LAUNCH_OK();
  }
  ws.release(mark);
  return 0;
}

int check_model(const stgcn_model_desc *m) {
  STGCN_REQUIRE(m && m->layers && m->num_layers > 0, "model descriptor is empty");
  STGCN_REQUIRE(m->in_feat > 0 && m->num_joints > 1 && m->partitions > 0 && m->num_classes > 0,
                "bad model dimensions");
  for (int i = 0; i + 1 < m->num_layers; ++i)
    STGCN_REQUIRE(m->layers[i].c_out == m->layers[i + 1].c_in, "layer %d/%d channel mismatch", i, i + 1);
  return 0;
}

size_t model_prepare_layout(const stgcn_model_desc &m, void *base, size_t cap, LayerPrep *out);
inline bool use_prepared(const stgcn_model_desc &m) {
  return m.prepared != nullptr && m.math != STGCN_MATH_FP32 &&
         m.prepared_bytes >= model_prepare_layout(m, nullptr, 0, nullptr);
}

size_t model_prepare_layout(const stgcn_model_desc &m, void *base, size_t cap, LayerPrep *out) {
  Bump pb(base, cap);
  for (int i = 0; i < m.num_layers; ++i) {
    LayerPrep P = prep_take(m.layers[i], m.partitions, m.num_joints, pb, (m.reserved & 2) != 0);
    if (out) out[i] = P;
  }
  return pb.peak;
}

// ---- sliding windows: the per-frame work of the first layer is done once per FRAME, not once per window ------
// Window n of a chunk covers frames [n, n + W) of the chunk's Tc = n_win + W - 1 frames.  Everything before the
// first temporal convolution is a per-frame function (input norm, fcn_in, graph convolution, LayerNorm, ReLU), so
// it is evaluated on the Tc shared frames; the temporal kernel then reads its windows out of that one sequence
// through a tensor map whose trial pitch is one frame (the rows outside a window are zero-filled by TMA exactly as
// the convolution's padding), and an identity residual is addressed the same way.  STGCN_WINDOWS_SHARE=0 keeps the
// per-window evaluation (A/B and the parity test).
inline bool windows_share_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_WINDOWS_SHARE");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}
inline bool windows_share_supported(const stgcn_model_desc &m, const LayerPrep &P, int W) {
  const stgcn_layer_desc &d = m.layers[0];
  return windows_share_enabled() && m.math != STGCN_MATH_FP32 && d.norm == STGCN_NORM_LAYERNORM && embed_warp_path(m) &&
         P.gw && P.tcn && !P.taps && d.stride == 1 && d.residual != STGCN_RES_CONV &&
         tc::tcn_tc2_supported(d.c_out, m.num_joints, d.kernel, 1, W);
}
// x: strided view of the padded capture at the chunk's first frame; out: [n_win][W][V][c_out] (fp32 rows or planes)
int layer0_windows_shared(const stgcn_model_desc &m, const LayerPrep &P, const float *x, float *out, int n_win, int W,
                          const long long *xs, bool out_planes, Bump &ws, cudaStream_t st) {
  const stgcn_layer_desc &d = m.layers[0];
  const int V = m.num_joints, planes = m.math == STGCN_MATH_BF16X3 ? 2 : 1;
  const int Tc = n_win + W - 1;
  const long long rows_c = (long long)Tc * V, rows_out = (long long)n_win * W * V;
  const size_t mark = ws.mark();
  __nv_bfloat16 *h0 = ws.take<__nv_bfloat16>((size_t)planes * rows_c * d.c_in);
  __nv_bfloat16 *u16 = ws.take<__nv_bfloat16>((size_t)planes * rows_c * d.c_out);
  const long long one_trial[3] = {0, xs[1], xs[2]};
  if (embed(m, x, reinterpret_cast<float *>(h0), 1, Tc, ws, st, one_trial, planes)) return 1;
  tc::LnStreamArgs l{};
  l.frames = Tc; l.T = Tc; l.V = V; l.C = d.c_out;
  l.n_wT = P.n1T; l.n_bT = P.n1T + (size_t)d.c_out * V;
  l.relu = 1; l.eps = kEps;
  l.out_hi = u16; l.out_lo = planes == 2 && u16 ? u16 + (size_t)rows_c * d.c_out : nullptr;
  tc::GcnwParams g{};
  g.T = Tc; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = 1;
  g.tab = P.gwtab;
  g.bias = P.bzT; g.bias_sw = 1;
  g.debug = debug_mode();
  if (gcnw_stage(d.c_out, h0, P.wsc, g, l, P.n1V, Tc, 1, rows_c * d.c_in, ws, st)) return 1;
  if (!ws.measuring()) {
    STGCN_REQUIRE(!ws.overflow, "workspace too small (shared window frames)");
    tc::TcnTc2Params p{};
    p.T_out = W; p.V = V; p.G = d.kernel;
    p.planes = planes;
    p.epi.bias = d.tcn_b; p.epi.bias_sw = 0;
    p.epi.n_wT = P.n2T; p.epi.n_bT = P.n2T + (size_t)d.c_out * V;
    if (d.residual == STGCN_RES_IDENTITY) {
      p.epi.res_hi = h0;
      p.epi.res_lo = planes == 2 ? h0 + (size_t)rows_c * d.c_in : nullptr;
    }
    if (out_planes) {
      p.epi.out_hi = reinterpret_cast<__nv_bfloat16 *>(out);
      p.epi.out_lo = planes == 2 ? p.epi.out_hi + (size_t)rows_out * d.c_out : nullptr;
    } else {
      p.epi.out_f32 = out;
    }
    p.epi.relu = 1;
    p.epi.eps = kEps;
    p.epi.debug = debug_mode();
    ProfScope ps(KC_GEMM_TCN, st);
    if (tc::launch_tcn_tc2(d.c_out, u16, P.wp16, p, n_win, W, 1, 0, st, 1, Tc)) return 1;
    STGCN_LAUNCH_OK();
  }
  ws.release(mark);
  return 0;
}

// ST-GCN model on `n` trials (one chunk).  logits [n, classes]; features optional (NCTV).
int model_chunk(const stgcn_model_desc &m, const float *x, float *logits, float *features, int n, int T,
                Bump &ws, cudaStream_t st, const stgcn_halo_desc *halo = nullptr, float *pooled_sums = nullptr,
                const long long *xs = nullptr) {
  const int V = m.num_joints, K = m.partitions;
  // ping-pong activation buffers sized for the largest layer interface
  size_t max_act = (size_t)n * T * V * m.layers[0].c_in;
  int t = T;
  for (int i = 0; i < m.num_layers; ++i) {
    t = (t - 1) / m.layers[i].stride + 1;
    size_t a = (size_t)n * t * V * m.layers[i].c_out;
    if (a > max_act) max_act = a;
  }
  float *buf[2] = {ws.take<float>(max_act), ws.take<float>(max_act)};
  // which layers take bf16-plane input (per-joint-weight graph conv); a producer writes planes when
  // its consumer wants them and it can (warp input stage / tensor-core temporal epilogue)
  const bool have = use_prepared(m);
  const int planes = m.math == STGCN_MATH_BF16X3 ? 2 : 1;
  bool pl[65] = {false};
  STGCN_REQUIRE(m.num_layers <= 64, "too many layers");
  {
    Bump pm(nullptr, 0);
    int tt = T;
    for (int i = 0; i < m.num_layers; ++i) {
      const stgcn_layer_desc &d = m.layers[i];
      const LayerPrep P = prep_take(d, K, V, pm, (m.reserved & 2) != 0);
      pl[i] = m.math != STGCN_MATH_FP32 && d.norm == STGCN_NORM_LAYERNORM && P.gw &&
              ((P.tcn && tc::tcn_tc2_supported(d.c_out, V, d.kernel, d.stride, tt)) || P.taps);
      tt = (tt - 1) / d.stride + 1;
    }
  }
  const bool in0_planes = pl[0] && embed_warp_path(m);
  // sliding windows (xs with a one-frame trial pitch): first layer on the shared frames
  bool shared0 = false;
  if (xs && xs[0] == xs[2] && have && !halo && pl[0]) {
    Bump p0(const_cast<void *>(m.prepared), m.prepared_bytes);
    const LayerPrep P0 = prep_take(m.layers[0], K, V, p0, (m.reserved & 2) != 0);
    shared0 = windows_share_supported(m, P0, T);
    if (shared0 && layer0_windows_shared(m, P0, x, buf[1], n, T, xs, pl[0] && pl[1], ws, st)) return 1;
  }
  if (!shared0 && embed(m, x, buf[0], n, T, ws, st, xs, in0_planes ? planes : 0)) return 1;
  int cur = 0;
  t = T;
  Bump pb(const_cast<void *>(m.prepared), m.prepared_bytes);
  bool x_planes = in0_planes;
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    STGCN_REQUIRE(!d.rt, "stgcn_model_forward needs ST-GCN layers (rt == 0)");
    STGCN_REQUIRE(!d.a_per_sample, "per-sample adjacency is only supported by the layer-level API");
    LayerPrep P;
    if (have) P = prep_take(d, K, V, pb, (m.reserved & 2) != 0);
    const bool out_planes = pl[i] && pl[i + 1];
    if (i == 0 && shared0) {
      x_planes = out_planes;
      cur ^= 1;
      continue;
    }
    if (layer_forward_ntvc(d, K, V, m.math, buf[cur], buf[cur ^ 1], n, t, ws, st, have ? &P : nullptr, halo, i,
                           x_planes, out_planes, (m.reserved & 2) != 0))
      return 1;
    x_planes = out_planes;
    t = (t - 1) / d.stride + 1;
    cur ^= 1;
  }
  const int c_last = m.layers[m.num_layers - 1].c_out;
  if (features && !ws.measuring())
    if (to_nctv(buf[cur], features, n, c_last, (long long)t * V, c_last, st)) return 1;
  if (pool_fc(buf[cur], n, (long long)t * V, c_last, m.fcn_out_w, m.fcn_out_b, m.num_classes, logits, ws, st,
              pooled_sums))
    return 1;
  return 0;
}

size_t model_chunk_bytes(const stgcn_model_desc &m, int n, int T) {
  Bump ws(nullptr, 0);
  model_chunk(m, nullptr, nullptr, nullptr, n, T, ws, nullptr);
  return ws.peak;
}

// rows (trial-frames * V) one chunk of trials should carry when chunking is allowed
constexpr long long kChunkRows = 3200000;

// STGCN_CHUNK_ROWS overrides the chunk size (tuning aid: smaller chunks keep more of a layer's
// intermediates in the 126 MB L2, larger ones amortise launches and tails)
inline long long chunk_rows() {
  static long long r = -1;
  if (r < 0) {
    const char *e = getenv("STGCN_CHUNK_ROWS");
    r = e ? atoll(e) : kChunkRows;
    if (r <= 0) r = kChunkRows;
  }
  return r;
}

int default_chunk(const stgcn_model_desc &m, int N, int T) {
  if (m.norm == STGCN_NORM_BATCHNORM) return N;  // batch statistics span the whole call
  long long per_trial = (long long)T * m.num_joints;
  long long n = chunk_rows() / (per_trial > 0 ? per_trial : 1);
  if (n < 1) n = 1;
  if (n > N) n = N;
  return (int)n;
}

// ---- RT state layout ------------------------------------------------------------
struct RtLayout {
  size_t counters;  // offset of int32[B]
  size_t fifo[64], acc[64];
  size_t total;
};
bool rt_fifo_bf16(const stgcn_model_desc &m, int B);
int rt_layout(const stgcn_model_desc &m, int B, RtLayout &L) {
  STGCN_REQUIRE(m.num_layers <= 64, "too many layers");
  const size_t fifo_elt = rt_fifo_bf16(m, B) ? sizeof(__nv_bfloat16) : sizeof(float);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off += (bytes + 255) & ~size_t(255);
    return at;
  };
  L.counters = take(sizeof(int32_t) * (size_t)B);
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    const size_t slot_n = (size_t)B * m.num_joints * d.c_out;
    L.fifo[i] = take(slot_n * fifo_elt * (size_t)(d.stride * (d.kernel - 1) + 1));
    L.acc[i] = take(slot_n * sizeof(float) * (size_t)d.stride);
  }
  L.total = off;
  return 0;
}

// frame counters wrap at the least common multiple of every layer's FIFO / accumulator ring sizes
inline long long gcd_ll(long long a, long long b) { return b ? gcd_ll(b, a % b) : a; }
inline int rt_counter_period_of(int F, int S, long long run) {
  run = run / gcd_ll(run, F) * F;
  run = run / gcd_ll(run, S) * S;
  return run > (1ll << 30) ? 0 : (int)run;
}
int rt_counter_period(const stgcn_model_desc &m) {
  long long run = 1;
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    const int p = rt_counter_period_of(d.stride * (d.kernel - 1) + 1, d.stride, run);
    if (p == 0) return 0;
    run = p;
  }
  return (int)run;
}

// few streams: the whole step as one cluster kernel (kernels_rt_small.cuh)
// measured (back-to-back replays): 0.123-0.125 ms for 1..14 streams, 0.244 ms at 16 (the clusters no longer all
// run at once); the batched path takes 0.19-0.20 ms from 2 to 17 streams
constexpr int kSmallBatchMax = 14;
bool rt_small_supported(const stgcn_model_desc &m, int B) {
  if (m.reserved & 1) return false;                      // caller opted out (tests of the batched path)
  if (B > kSmallBatchMax || m.num_layers > rts::kMaxLayers || m.num_joints > 32) return false;
  // K*V + 1 CSR row pointers and the non-zeros are staged in fixed shared arrays (kernels_rt_small.cuh:31-32)
  if (m.in_feat * m.num_joints > rts::kThreads || m.partitions * m.num_joints + 1 > rts::kCsrPtrMax) return false;
  int c_max = m.layers[0].c_in;
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    if (d.rt != 1 || d.norm != STGCN_NORM_LAYERNORM || d.a_per_sample) return false;
    if (d.c_out % rts::kNC || d.c_in % rts::kChunk) return false;
    if (((d.c_out / rts::kNC) * m.num_joints) % 4) return false;        // a CTA's slice travels as one 16-byte-granular bulk copy
    if (m.num_joints * (d.c_out / rts::kNC) > 4 * rts::kThreads) return false;
    if ((m.partitions + 1) * (d.c_out / rts::kNC) > 128) return false;   // rows per CTA (register tile bound)
    c_max = d.c_out > c_max ? d.c_out : c_max;
    c_max = d.c_in > c_max ? d.c_in : c_max;
  }
  // the y buffer doubles as scratch for the input frame (in_feat * V) and the pooled features (c_last)
  const long long ybuf = (long long)(m.partitions + 1) * (c_max / rts::kNC) * m.num_joints;
  if ((long long)m.in_feat * m.num_joints > ybuf || m.layers[m.num_layers - 1].c_out > ybuf) return false;
  return rts::smem_floats(c_max, m.num_joints, m.partitions) * sizeof(float) <= 220 * 1024;
}

// per-joint-weight GEMM + streaming state kernel for every layer (static properties of the model only,
// so that the state layout never depends on whether operands were prepared)
bool rt_all_gw_static(const stgcn_model_desc &m) {
  if ((m.reserved & 2) == 0 || m.math == STGCN_MATH_FP32 || !rt_split_enabled() || !embed_warp_path(m) ||
      m.norm != STGCN_NORM_LAYERNORM || !tc::gcnw_enabled())
    return false;
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    if (d.norm != STGCN_NORM_LAYERNORM || d.a_per_sample || d.rt != 1) return false;
    if (!tc::gcn_tc_supported(d.c_in, d.c_out, m.num_joints, m.partitions) ||
        !tc::gcnw_supported(d.c_in, d.c_out, m.num_joints, m.partitions) || !rt_update_supported(m.num_joints, d.c_out))
      return false;
    if (d.residual == STGCN_RES_CONV && !tc::gcn_tc_supported(d.c_in, d.c_out, m.num_joints, 1)) return false;
  }
  return true;
}
// bf16 mode, many streams: the FIFO is stored as bf16 (the accumulators stay fp32): 2 + 2 + 4 + 4 bytes of
// state traffic per element and step instead of 16
bool rt_fifo_bf16(const stgcn_model_desc &m, int B) {
  if (m.math != STGCN_MATH_BF16 || rt_small_supported(m, B) || !rt_stream_enabled() || !rt_all_gw_static(m)) return false;
  for (int i = 0; i < m.num_layers; ++i)
    if (!rt_stream_supported(m.num_joints, m.layers[i].c_out)) return false;
  return true;
}

// STGCN_RT_OVERLAP: smallest stream count that is stepped as two overlapping halves (0 = never)
inline int rt_overlap_min_streams() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("STGCN_RT_OVERLAP");
    v = e ? atoi(e) : 1024;
    if (v <= 0) v = 0x7fffffff;
  }
  return v;
}
inline int rt_overlap_smem_cap() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("STGCN_RT_OVERLAP_SMEM");      // 0 = the GEMM CTAs keep all shared memory (measured best)
    v = e ? atoi(e) : 0;
  }
  return v;
}
// per-device side stream + fork / join events of the two-half step (created once; usable under stream capture)
// The caller holds the returned per-device lock while it enqueues fork .. join: the events are shared by every
// model instance of the device, and a record / wait pair of another host thread must not land between them.
constexpr int kRtSideEvents = 2 + 2 * 64;      // fork, join, and one "GEMMs done" event per layer and half
int rt_side_stream(cudaStream_t *side, cudaEvent_t **ev, std::unique_lock<std::mutex> &hold) {
  static std::mutex mu[kMaxDevices];
  static cudaStream_t streams[kMaxDevices] = {nullptr};
  static cudaEvent_t events[kMaxDevices][kRtSideEvents] = {{nullptr}};
  int dev = 0;
  STGCN_CUDA_OK(cudaGetDevice(&dev));
  STGCN_REQUIRE(dev >= 0 && dev < kMaxDevices, "rt step: device index %d out of range", dev);
  hold = std::unique_lock<std::mutex>(mu[dev]);
  if (!streams[dev]) {
    STGCN_CUDA_OK(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
    for (int i = 0; i < kRtSideEvents; ++i)
      STGCN_CUDA_OK(cudaEventCreateWithFlags(&events[dev][i], cudaEventDisableTiming));
  }
  *side = streams[dev];
  *ev = events[dev];
  return 0;
}
// STGCN_RT_OVERLAP_MODE: 0 (default) = the halves run free, 1 = the second half starts one GEMM phase late,
// 2 = GEMM phases of the two halves alternate strictly.  Measured at 2048 / 4096 streams (PKU trunk, ms):
// one batch 0.838 / 1.489; mode 0 0.786 / 1.473; mode 1 0.830 / 1.508; mode 2 0.964 / 1.571 -- a GEMM and a
// state kernel that share the SMs slow each other by what the overlap gains, so only the free-running form
// (which mostly fills the tails of the half-sized launches) is kept.
inline int rt_overlap_mode() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("STGCN_RT_OVERLAP_MODE");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// STGCN_RT_POOL=0: the last layer writes rows and the head pools them (A/B of the fused pooling)
inline bool rt_pool_fused_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("STGCN_RT_POOL");
    on = e ? atoi(e) != 0 : 1;
  }
  return on != 0;
}
// One continual step for streams [b0, b0 + nb) of a state laid out for B streams (the many-streams path), in
// three parts so that two ranges can be enqueued layer by layer on two CUDA streams.
struct RtRangeCursor {
  float *buf[2] = {nullptr, nullptr};
  int cur = 0, b0 = 0, nb = 0;
  size_t prep_off = 0;            // walk through the prepared operands
  bool pooled = false;            // the last layer wrote the mean over the joints [nb, c_last] instead of rows
};
int rt_step_range_begin(const stgcn_model_desc &m, const RtLayout &L, const float *x, int B, int b0, int nb,
                        bool all_gw, Bump &ws, cudaStream_t st, RtRangeCursor &c) {
  const int V = m.num_joints;
  STGCN_REQUIRE(all_gw || (b0 == 0 && nb == B), "rt step: stream ranges need the per-joint-weight GEMM path");
  size_t max_act = (size_t)nb * V * m.layers[0].c_in;
  for (int i = 0; i < m.num_layers; ++i) {
    size_t a = (size_t)nb * V * m.layers[i].c_out;
    if (a > max_act) max_act = a;
  }
  c.buf[0] = ws.take<float>(max_act);
  c.buf[1] = ws.take<float>(max_act);
  c.cur = 0; c.b0 = b0; c.nb = nb; c.prep_off = 0;
  const int planes = m.math == STGCN_MATH_BF16X3 ? 2 : 1;
  STGCN_REQUIRE(ws.measuring() || !ws.overflow, "rtstgcn_step: workspace too small (%zu B given)", ws.cap);
  return embed(m, x ? x + (size_t)b0 * m.in_feat * V : nullptr, c.buf[0], nb, 1, ws, st, nullptr, all_gw ? planes : 0);
}
int rt_step_range_layer(const stgcn_model_desc &m, const RtLayout &L, void *state, int B, bool all_gw, bool fifo16,
                        Bump &ws, cudaStream_t st, int gw_smem_cap, RtRangeCursor &c, int i,
                        cudaEvent_t gemm_wait, cudaEvent_t gemm_done) {
  const int V = m.num_joints, K = m.partitions;
  const bool have = use_prepared(m), sparse = (m.reserved & 2) != 0;
  const stgcn_layer_desc &d = m.layers[i];
  STGCN_REQUIRE(d.rt == 1, "rtstgcn_step needs online layers (rt == 1)");
  char *sb = static_cast<char *>(state);
  int *counter = ws.measuring() ? nullptr : reinterpret_cast<int *>(sb + L.counters) + c.b0;
  const size_t first = (size_t)c.b0 * V * d.c_out;           // elements before this range inside one slot
  float *fifo = ws.measuring() ? nullptr
                               : reinterpret_cast<float *>(sb + L.fifo[i] + first * (fifo16 && all_gw ? 2 : 4));
  float *acc = ws.measuring() ? nullptr : reinterpret_cast<float *>(sb + L.acc[i]) + first;
  LayerPrep P;
  if (have) {
    Bump pb(const_cast<void *>(m.prepared), m.prepared_bytes);
    pb.off = c.prep_off;
    P = prep_take(d, K, V, pb, sparse);
    c.prep_off = pb.off;
  }
  const bool last = i + 1 == m.num_layers;
  // last layer: the state kernel pools over the joints itself (the head then reads [nb, C] instead of [nb*V, C])
  const bool pool = last && all_gw && rt_pool_fused_enabled() && rt_stream_enabled() &&
                    rt_stream_pool_supported(V, d.c_out) && !(debug_mode() & 8192);
  if (rt_layer_step_ntvc(d, K, V, m.math, c.buf[c.cur], c.buf[c.cur ^ 1], fifo, acc, counter, c.nb, ws, st,
                         have ? &P : nullptr, all_gw, all_gw, all_gw && !last, fifo16 && all_gw,
                         c.nb == B ? 0 : (long long)B * V, gw_smem_cap, gemm_wait, gemm_done, pool))
    return 1;
  c.pooled = pool;
  c.cur ^= 1;
  return 0;
}
int rt_step_range_end(const stgcn_model_desc &m, float *logits, int *top5, cudaStream_t st, RtRangeCursor &c) {
  const int V = m.num_joints, c_last = m.layers[m.num_layers - 1].c_out;
  ProfScope ps(KC_POOL, st);
  const int spb = rt_head_streams(c.nb);
  STGCN_CUDA_OK(launch_pdl(k_rt_head, dim3(cdiv(c.nb, spb)), dim3(256),
                           sizeof(float) * spb * (c_last + m.num_classes), st, (const float *)c.buf[c.cur], c.nb,
                           c.pooled ? 1 : V, c_last, m.fcn_out_w, m.fcn_out_b, m.num_classes, logits + (size_t)c.b0 * m.num_classes,
                           top5 ? top5 + (size_t)c.b0 * 5 : (int *)nullptr, spb));
  STGCN_LAUNCH_OK();
  return 0;
}
int rt_step_range(const stgcn_model_desc &m, const RtLayout &L, const float *x, void *state, float *logits, int *top5,
                  int B, int b0, int nb, bool all_gw, bool fifo16, Bump &ws, cudaStream_t st, int gw_smem_cap) {
  RtRangeCursor c;
  if (rt_step_range_begin(m, L, x, B, b0, nb, all_gw, ws, st, c)) return 1;
  for (int i = 0; i < m.num_layers; ++i)
    if (rt_step_range_layer(m, L, state, B, all_gw, fifo16, ws, st, gw_smem_cap, c, i, nullptr, nullptr)) return 1;
  if (!ws.measuring() && rt_step_range_end(m, logits, top5, st, c)) return 1;
  return 0;
}

int rt_step(const stgcn_model_desc &m, const float *x, void *state, float *logits, int B, Bump &ws,
            cudaStream_t st, int *top5 = nullptr) {
  const int V = m.num_joints, K = m.partitions;
  STGCN_REQUIRE(m.norm == STGCN_NORM_LAYERNORM,
                "continual inference needs LayerNorm (reference raises at models/utils/batchnorm.py:20)");
  RtLayout L;
  if (rt_layout(m, B, L)) return 1;
  if (rt_small_supported(m, B)) {
    const size_t mark = ws.mark();
    rts::Params P{};
    P.num_layers = m.num_layers; P.V = V; P.K = K; P.in_feat = m.in_feat; P.num_classes = m.num_classes; P.B = B;
    P.eps = kEps;
    P.period = rt_counter_period(m);
    P.single_pass = m.math == STGCN_MATH_BF16 ? 1 : 0;
    P.x = x; P.logits = logits;
    P.norm_in_w = m.norm_in_w; P.norm_in_b = m.norm_in_b;
    P.fcn_in_w = m.fcn_in_w; P.fcn_in_b = m.fcn_in_b;
    P.fcn_out_w = m.fcn_out_w; P.fcn_out_b = m.fcn_out_b;
    char *sb = static_cast<char *>(state);
    P.counter = ws.measuring() ? nullptr : reinterpret_cast<int *>(sb + L.counters);
    Bump pb(const_cast<void *>(m.prepared), m.prepared_bytes);
    const bool have = use_prepared(m);
    int c_max = m.layers[0].c_in;
    for (int i = 0; i < m.num_layers; ++i) {
      const stgcn_layer_desc &d = m.layers[i];
      LayerPrep lp;
      if (have) {
        lp = prep_take(d, K, V, pb, (m.reserved & 2) != 0);
      } else {
        // no prepared operands (math = fp32, or the caller did not prepare): the cluster kernel reads
        // only the (k, w)-ordered adjacency CSR, so build just that -- not the tensor-core operand set
        lp.kw_ptr = ws.take<int>((size_t)K * V + 1);
        lp.kw_va = ws.take<int2>((size_t)K * V * V);
        if (!ws.measuring()) {
          STGCN_REQUIRE(!ws.overflow, "workspace too small (rt step adjacency)");
          ProfScope ps(KC_MISC, st);
          tc::k_build_adj_csr_kw<<<1, 128, (K * V + 1) * sizeof(int), st>>>(d.a_eff, K, V, lp.kw_ptr, lp.kw_va);
          STGCN_LAUNCH_OK();
        }
      }
      rts::Layer &R = P.layer[i];
      R.c_in = d.c_in; R.c_out = d.c_out; R.F = d.stride * (d.kernel - 1) + 1; R.S = d.stride;
      R.residual = d.residual;
      R.gcn_w = d.gcn_w; R.gcn_b = d.gcn_b; R.n1_w = d.n1_w; R.n1_b = d.n1_b;
      R.res_w = d.res_w; R.nr_w = d.nr_w; R.nr_b = d.nr_b;
      R.kw_ptr = lp.kw_ptr; R.kw_va = lp.kw_va;
      R.fifo = ws.measuring() ? nullptr : reinterpret_cast<float *>(sb + L.fifo[i]);
      R.acc = ws.measuring() ? nullptr : reinterpret_cast<float *>(sb + L.acc[i]);
      c_max = d.c_out > c_max ? d.c_out : c_max;
      c_max = d.c_in > c_max ? d.c_in : c_max;
    }
    P.c_max = c_max;
    if (!ws.measuring()) {
      const size_t smem = rts::smem_floats(c_max, V, K) * sizeof(float);
      STGCN_CUDA_OK(cudaFuncSetAttribute(rts::k_rt_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ProfScope ps(KC_FRAME, st);
      P.debug = debug_mode();
      rts::k_rt_small<<<dim3(rts::kNC, B), rts::kThreads, smem, st>>>(P);
      STGCN_LAUNCH_OK();
      if (debug_mode() & 4) {
        unsigned long long h[8];
        STGCN_CUDA_OK(cudaStreamSynchronize(st));
        STGCN_CUDA_OK(cudaMemcpyFromSymbol(h, rts::g_dbg, sizeof(h)));
        fprintf(stderr, "[dbg] rt_small: input=%llu gemm(+preload)=%llu adj_state=%llu stats_barrier=%llu normalise=%llu tail=%llu issue=%llu wait=%llu\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
        memset(h, 0, sizeof(h));
        STGCN_CUDA_OK(cudaMemcpyToSymbol(rts::g_dbg, h, sizeof(h)));
      }
    }
    ws.release(mark);
    return 0;
  }
  const bool have = use_prepared(m);
  const bool sparse = (m.reserved & 2) != 0;
  // per-joint-weight GEMM path: all layers or none (the activations then travel as bf16 planes)
  bool all_gw = have && sparse && m.math != STGCN_MATH_FP32 && rt_split_enabled() && embed_warp_path(m);
  {
    Bump pm(nullptr, 0);
    for (int i = 0; i < m.num_layers && all_gw; ++i) {
      const stgcn_layer_desc &d = m.layers[i];
      const LayerPrep Q = prep_take(d, K, V, pm, sparse);
      all_gw = Q.gw && (d.residual != STGCN_RES_CONV || Q.res) && rt_update_supported(V, d.c_out);
    }
  }
  const bool fifo16 = rt_fifo_bf16(m, B);
  STGCN_REQUIRE(!fifo16 || all_gw || ws.measuring(),
                "rtstgcn_step: bf16 continual mode stores the FIFO as bf16 and needs prepared operands "
                "(stgcn_model_prepare)");
  int *counter = ws.measuring() ? nullptr : reinterpret_cast<int *>(static_cast<char *>(state) + L.counters);
  const int halves = (all_gw && B >= rt_overlap_min_streams()) ? 2 : 1;
  if (halves == 1) {
    if (rt_step_range(m, L, x, state, logits, top5, B, 0, B, all_gw, fifo16, ws, st, 0)) return 1;
  } else {
    // Two halves of the streams as two independent pipelines on two CUDA streams: the launches of one half fill
    // the tails of the other's (-5 % at 1024 streams, -6 % at 2048, -1 % at 4096).  Pairing one half's GEMM with
    // the other half's state kernel on purpose (STGCN_RT_OVERLAP_MODE, STGCN_RT_OVERLAP_SMEM) does not pay.
    const int nb0 = ((B / 2 + 127) / 128) * 128 < B ? ((B / 2 + 127) / 128) * 128 : B / 2;
    cudaStream_t side = nullptr;
    cudaEvent_t *ev = nullptr;
    std::unique_lock<std::mutex> hold;
    if (!ws.measuring() && rt_side_stream(&side, &ev, hold)) return 1;
    size_t part[2];
    for (int h = 0; h < 2; ++h) {
      Bump wm(nullptr, 0);
      if (rt_step_range(m, L, nullptr, nullptr, nullptr, nullptr, B, h ? nb0 : 0, h ? B - nb0 : nb0, all_gw, fifo16, wm,
                        nullptr, 0))
        return 1;
      part[h] = (wm.peak + 255) & ~size_t(255);
    }
    char *w0 = ws.take<char>(part[0]), *w1 = ws.take<char>(part[1]);
    if (!ws.measuring()) {
      STGCN_REQUIRE(!ws.overflow, "rtstgcn_step: workspace too small (%zu B given)", ws.cap);
      cudaEvent_t fork = ev[0], join = ev[1];
      STGCN_CUDA_OK(cudaEventRecord(fork, st));
      STGCN_CUDA_OK(cudaStreamWaitEvent(side, fork, 0));
      Bump b0(w0, part[0]), b1(w1, part[1]);
      const int cap = rt_overlap_smem_cap(), mode = rt_overlap_mode(), NL = m.num_layers;
      // half 0, layer i: waits for half 1's GEMMs of layer i-1 (mode 2), records doneA[i];
      // half 1, layer i: waits for doneA[i] (mode 2; mode 1: layer 0 only), records doneB[i] (mode 2).
      // The two halves are ENQUEUED one after the other, so half 0 can only wait on events of half 1 that were
      // recorded earlier in host order: enqueue layer by layer, alternating the halves.
      cudaEvent_t *doneA = ev + 2, *doneB = ev + 2 + 64;
      int rc = 0;
      if (mode == 0) {
        rc = rt_step_range(m, L, x, state, logits, top5, B, 0, nb0, all_gw, fifo16, b0, st, cap);
        rc = rc || rt_step_range(m, L, x, state, logits, top5, B, nb0, B - nb0, all_gw, fifo16, b1, side, cap);
      } else {
        cudaEvent_t waitA[64] = {nullptr}, recA[64] = {nullptr}, waitB[64] = {nullptr}, recB[64] = {nullptr};
        for (int i = 0; i < NL; ++i) {
          if (mode == 2 || i == 0) { recA[i] = doneA[i]; waitB[i] = doneA[i]; }
          if (mode == 2 && i + 1 < NL) { recB[i] = doneB[i]; waitA[i + 1] = doneB[i]; }
        }
        RtRangeCursor ca, cb;
        rc = rt_step_range_begin(m, L, x, B, 0, nb0, all_gw, b0, st, ca) ||
             rt_step_range_begin(m, L, x, B, nb0, B - nb0, all_gw, b1, side, cb);
        for (int i = 0; i < NL && !rc; ++i) {
          rc = rt_step_range_layer(m, L, state, B, all_gw, fifo16, b0, st, cap, ca, i, waitA[i], recA[i]) ||
               rt_step_range_layer(m, L, state, B, all_gw, fifo16, b1, side, cap, cb, i, waitB[i], recB[i]);
        }
        rc = rc || rt_step_range_end(m, logits, top5, st, ca) || rt_step_range_end(m, logits, top5, side, cb);
      }
      cudaEventRecord(join, side);                           // always rejoin (also after an error, also under capture)
      cudaStreamWaitEvent(st, join, 0);
      if (rc) return 1;
    }
  }
  if (!ws.measuring()) {
    STGCN_CUDA_OK(launch_pdl(k_advance_counters, dim3(cdiv(B, 256)), dim3(256), (size_t)0, st, counter, 0, B,
                             rt_counter_period(m)));
    STGCN_LAUNCH_OK();
  }
  return 0;
}


// ---- CoST-GCN continual step (models/costgcn/costgcn.py:81-99, 190-211) ---------------------------
// Per layer and stream: a ring of the last F = stride*(kernel-1)+1 frames u = relu(LN1(gcn(x))) and a ring
// of the last kernel/2 + 1 residuals, both as bf16 hi/lo planes [plane][slot][B*V][C] (the reference
// keeps newest-first FIFOs of z and re-normalises the whole FIFO every step; LayerNorm is per frame, so
// storing u once per frame is the same thing -- including the never-written slots, which the reference
// reads as LN(0) = bias: the ring is initialised with relu(tcn.0.bias)).  The ring position is the
// caller's frame index t, common to all streams; resetting a stream re-initialises its ring entries.
struct CostLayout {
  size_t u[64], r[64];
  size_t total;
};
inline int cost_res_slots(const stgcn_layer_desc &d) { return d.residual == STGCN_RES_NONE ? 0 : d.kernel / 2 + 1; }
int cost_layout(const stgcn_model_desc &m, int B, CostLayout &L) {
  STGCN_REQUIRE(m.num_layers <= 64, "too many layers");
  const int planes = m.math == STGCN_MATH_BF16X3 ? 2 : 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t at = off;
    off += (bytes + 255) & ~size_t(255);
    return at;
  };
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    const size_t slot = (size_t)B * m.num_joints * d.c_out * sizeof(__nv_bfloat16);
    L.u[i] = take(slot * planes * (size_t)(d.stride * (d.kernel - 1) + 1));
    L.r[i] = take(slot * planes * (size_t)cost_res_slots(d));
  }
  L.total = off;
  return 0;
}

int cost_check(const stgcn_model_desc &m) {
  STGCN_REQUIRE(m.norm == STGCN_NORM_LAYERNORM && m.math != STGCN_MATH_FP32 && (m.reserved & 2) && embed_warp_path(m),
                "costgcn: the B200 path needs LayerNorm, math in {bf16x3, bf16} and a sparse (tree) adjacency");
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    STGCN_REQUIRE(d.rt == 2 && d.norm == STGCN_NORM_LAYERNORM && !d.a_per_sample, "costgcn: layer %d is not a CoST-GCN layer", i);
    STGCN_REQUIRE(d.kernel % 2 == 1 && d.kernel <= 15, "costgcn: temporal kernel must be odd and <= 15 (got %d)", d.kernel);
    STGCN_REQUIRE(tc::gcnw_supported(d.c_in, d.c_out, m.num_joints, m.partitions) &&
                      tc::gcnw_supported(d.c_out, d.c_out, m.num_joints, 1),
                  "costgcn: layer %d: channels must be 64, 128 or 256 (got %d -> %d)", i, d.c_in, d.c_out);
  }
  return 0;
}

// u ring entries of streams [first, first + count) <- relu(tcn.0.bias) (hi / lo planes); n1_b is (C, V)
__global__ void k_cost_ring_init(const float *__restrict__ n1_b, __nv_bfloat16 *__restrict__ ring, int planes, int slots,
                                 int B, int V, int C, int first, int count) {
  const long long per = (long long)count * V * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per * slots) return;
  const int sl = (int)(i / per);
  const long long r = i - (long long)sl * per;
  const int c = (int)(r % C);
  const int v = (int)((r / C) % V);
  const long long b = first + r / ((long long)V * C);
  const float val = fmaxf(n1_b[c * V + v], 0.f);
  __nv_bfloat16 hi, lo;
  tc::split_bf16(val, hi, lo);
  const long long slot_n = (long long)B * V * C;
  const long long at = (long long)sl * slot_n + (b * V + v) * C + c;
  ring[at] = hi;
  if (planes == 2) ring[(long long)slots * slot_n + at] = lo;
}

int cost_step(const stgcn_model_desc &m, const float *x, void *state, long long t, float *logits, int B, Bump &ws,
              cudaStream_t st) {
  const int V = m.num_joints, K = m.partitions;
  if (cost_check(m)) return 1;
  STGCN_REQUIRE(ws.measuring() || use_prepared(m), "costgcn_step needs prepared operands (stgcn_model_prepare)");
  CostLayout L;
  if (cost_layout(m, B, L)) return 1;
  const int planes = m.math == STGCN_MATH_BF16X3 ? 2 : 1;
  const long long rows = (long long)B * V;
  size_t max_act = (size_t)rows * m.layers[0].c_in;
  for (int i = 0; i < m.num_layers; ++i) {
    const size_t a = (size_t)rows * m.layers[i].c_out;
    if (a > max_act) max_act = a;
  }
  // activations between layers travel as bf16 planes; the last layer writes fp32 rows for the head
  __nv_bfloat16 *buf[2] = {ws.take<__nv_bfloat16>(planes * max_act), ws.take<__nv_bfloat16>(planes * max_act)};
  float *zq = ws.take<float>(max_act);
  float *qres = ws.take<float>(max_act);
  float *last = ws.take<float>(max_act);
  if (ws.measuring()) return 0;
  STGCN_REQUIRE(!ws.overflow, "costgcn_step: workspace too small (%zu B given)", ws.cap);
  if (embed(m, x, reinterpret_cast<float *>(buf[0]), B, 1, ws, st, nullptr, planes)) return 1;
  char *sb = static_cast<char *>(state);
  Bump pb(const_cast<void *>(m.prepared), m.prepared_bytes);
  const int cap = tc::gcnw_edge_cap(V);
  int cur = 0;
  for (int i = 0; i < m.num_layers; ++i) {
    const stgcn_layer_desc &d = m.layers[i];
    const LayerPrep P = prep_take(d, K, V, pb, true);
    STGCN_REQUIRE(P.gw && P.tcn && (d.residual != STGCN_RES_CONV || P.gwr), "costgcn: layer %d has no tensor-core operands", i);
    const int F = d.stride * (d.kernel - 1) + 1, R2 = cost_res_slots(d);
    const long long slot_in = rows * d.c_in, slot_n = rows * d.c_out;
    const __nv_bfloat16 *xh = buf[cur];
    __nv_bfloat16 *ur = reinterpret_cast<__nv_bfloat16 *>(sb + L.u[i]);
    __nv_bfloat16 *rr = reinterpret_cast<__nv_bfloat16 *>(sb + L.r[i]);
    const int su = (int)(t % F);
    const bool is_last = i + 1 == m.num_layers;
    // ---- residual of this frame -> ring slot t mod R2 (costgcn.py:193-198) ----
    if (d.residual == STGCN_RES_IDENTITY) {
      const int sr = (int)(t % R2);
      for (int pl = 0; pl < planes; ++pl)
        STGCN_CUDA_OK(cudaMemcpyAsync(rr + ((size_t)pl * R2 + sr) * slot_n, xh + (size_t)pl * slot_in,
                                      (size_t)slot_n * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, st));
    } else if (d.residual == STGCN_RES_CONV) {
      const int sr = (int)(t % R2);
      tc::GcnwParams g{};
      g.T = B; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = 1;
      g.tab = P.gwtabr;
      g.bias = d.res_b; g.bias_sw = 0;
      g.out = qres;
      {
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcnw(d.c_out, xh, P.wscr, g, B, 1, cap, slot_in, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      tc::LnStreamArgs l{};
      l.frames = B; l.T = B; l.V = V; l.C = d.c_out;
      l.z = qres;
      l.n_wT = P.nrT; l.n_bT = P.nrT + (size_t)d.c_out * V;
      l.relu = 0; l.eps = kEps;
      l.out_hi = rr + (size_t)sr * slot_n;
      l.out_lo = planes == 2 ? rr + ((size_t)R2 + sr) * slot_n : nullptr;
      ProfScope ps(KC_FRAME, st);
      if (tc::launch_ln_stream(l, st)) return 1;
      STGCN_LAUNCH_OK();
    }
    // ---- graph convolution, LayerNorm 1, ReLU -> u ring slot t mod F (costgcn.py:200-207) ----
    {
      tc::GcnwParams g{};
      g.T = B; g.V = V; g.Cin = d.c_in; g.planes = planes; g.N = 1;
      g.tab = P.gwtab;
      g.bias = P.bzT; g.bias_sw = 1;
      g.out = zq;
      {
        ProfScope ps(KC_GEMM_1X1, st);
        if (tc::launch_gcnw(d.c_out, xh, P.wsc, g, B, 1, cap, slot_in, st)) return 1;
        STGCN_LAUNCH_OK();
      }
      tc::LnStreamArgs l{};
      l.frames = B; l.T = B; l.V = V; l.C = d.c_out;
      l.z = zq;
      l.n_wT = P.n1T; l.n_bT = P.n1T + (size_t)d.c_out * V;
      l.relu = 1; l.eps = kEps;
      l.out_hi = ur + (size_t)su * slot_n;
      l.out_lo = planes == 2 ? ur + ((size_t)F + su) * slot_n : nullptr;
      ProfScope ps(KC_FRAME, st);
      if (tc::launch_ln_stream(l, st)) return 1;
      STGCN_LAUNCH_OK();
    }
    // ---- Gamma x 1 convolution over the ring: tap j reads the frame j*stride steps in the past ----
    {
      tc::GcnwParams g{};
      g.T = (int)rows; g.V = 1; g.Cin = d.c_out; g.planes = planes; g.N = 1;
      g.bias = d.tcn_b; g.bias_sw = 0;
      g.out = zq;
      g.ntaps = d.kernel; g.tmode = 1;
      for (int j = 0; j < d.kernel; ++j) g.tap_src[j] = (int)((((t - (long long)j * d.stride) % F) + F) % F);
      tc::GcnwXView xv;
      xv.slots = F; xv.slot_stride = slot_n;
      ProfScope ps(KC_GEMM_TCN, st);
      if (tc::launch_gcnw(d.c_out, ur, P.wp16, g, (int)rows, 1, d.kernel, (long long)F * slot_n, st, xv)) return 1;
      STGCN_LAUNCH_OK();
    }
    // ---- LayerNorm 2, + residual of kernel/2 frames ago, ReLU (costgcn.py:209-211) ----
    {
      tc::LnStreamArgs l{};
      l.frames = B; l.T = B; l.V = V; l.C = d.c_out;
      l.z = zq;
      l.n_wT = P.n2T; l.n_bT = P.n2T + (size_t)d.c_out * V;
      l.relu = 1; l.eps = kEps;
      if (R2 > 0) {
        const int rd = (int)((((t - d.kernel / 2) % R2) + R2) % R2);
        l.res_hi = rr + (size_t)rd * slot_n;
        l.res_lo = planes == 2 ? rr + ((size_t)R2 + rd) * slot_n : nullptr;
      }
      if (is_last) {
        l.out_f32 = last;
      } else {
        l.out_hi = buf[cur ^ 1];
        l.out_lo = planes == 2 ? buf[cur ^ 1] + slot_n : nullptr;
      }
      ProfScope ps(KC_FRAME, st);
      if (tc::launch_ln_stream(l, st)) return 1;
      STGCN_LAUNCH_OK();
    }
    cur ^= 1;
  }
  const int c_last = m.layers[m.num_layers - 1].c_out;
  ProfScope ps(KC_POOL, st);
  const int spb = rt_head_streams(B);
  k_rt_head<<<cdiv(B, spb), 256, sizeof(float) * spb * (c_last + m.num_classes), st>>>(
      last, B, V, c_last, m.fcn_out_w, m.fcn_out_b, m.num_classes, logits, nullptr, spb);
  STGCN_LAUNCH_OK();
  return 0;
}
}  // namespace

// =============================================================================
// exported C ABI
// =============================================================================
extern "C" {

int stgcn_abi_version(void) { return STGCN_ABI_VERSION; }
const char *stgcn_last_error(void) { return err_buf(); }

long long stgcn_launch_count(void) { return prof_global().launches.load(); }

int stgcn_profile_begin(void) {
  Profiler &p = prof();
  std::lock_guard<std::mutex> g(p.mu);
  p.n = 0;
  for (int i = 0; i < KC_COUNT; ++i) {
    p.ms[i] = 0.f;
    p.count[i] = 0;
  }
  if (!p.on.exchange(true)) prof_global().active.fetch_add(1);
  return 0;
}

int stgcn_profile_end(float *ms_per_class, long long *launches_per_class, int n_classes) {
  Profiler &p = prof();
  if (p.on.exchange(false)) prof_global().active.fetch_sub(1);
  STGCN_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> g(p.mu);
  for (int i = 0; i < p.n; ++i) {
    float ms = 0.f;
    STGCN_CUDA_OK(cudaEventElapsedTime(&ms, p.ev[i][0], p.ev[i][1]));
    p.ms[p.cls[i]] += ms;
    p.count[p.cls[i]] += 1;
  }
  for (int i = 0; i < n_classes && i < KC_COUNT; ++i) {
    if (ms_per_class) ms_per_class[i] = p.ms[i];
    if (launches_per_class) launches_per_class[i] = p.count[i];
  }
  p.n = 0;
  return 0;
}

int stgcn_device_check(int device) {
  cudaDeviceProp prop;
  STGCN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  STGCN_REQUIRE(prop.major == 10, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                prop.major, prop.minor);
  return 0;
}

// ---- LayerNorm (layernorm.py:22-28) ----------------------------------------
int stgcn_layernorm_forward(const float *x, const float *w, const float *b, float *y, int N, int C, int T,
                            int V, float eps, void *stream) {
  // Stand-alone module call in the reference layout: frames are gathered with stride T*V.
  STGCN_REQUIRE(N > 0 && C > 0 && T > 0 && V > 0 && (long long)C * V > 1, "layernorm: bad shape");
  size_t smem = (size_t)C * V * sizeof(float);
  STGCN_REQUIRE(smem <= 200 * 1024, "layernorm: C*V too large");
  if (smem > 48 * 1024)
    STGCN_CUDA_OK(cudaFuncSetAttribute(k_layernorm_nctv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long frames = (long long)N * T;
  long long blocks = frames < 148 * 8 ? frames : 148 * 8;
  k_layernorm_nctv<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(x, w, b, y, N, C, T, V, eps);
  STGCN_LAUNCH_OK();
  return 0;
}

// ---- BatchNorm (batchnorm.py:13-23, stgcn.py:152) ---------------------------
size_t stgcn_batchnorm_workspace_bytes(int C, int V, int mode) {
  size_t feats = mode == 1 ? (size_t)C * V : (size_t)C;
  return ((sizeof(double) * 2 * feats) + 255) & ~size_t(255);
}

int stgcn_batchnorm_forward(const float *x, const float *w, const float *b, float *y, int N, int C, int T,
                            int V, float eps, int mode, void *workspace, size_t workspace_bytes,
                            void *stream) {
  STGCN_REQUIRE(workspace && workspace_bytes >= stgcn_batchnorm_workspace_bytes(C, V, mode),
                "batchnorm: workspace too small");
  STGCN_REQUIRE(mode == 0 || mode == 1, "batchnorm: mode must be 0 or 1");
  STGCN_REQUIRE((long long)N * T * (mode == 0 ? V : 1) > 1,
                "Expected more than 1 value per channel when computing batch statistics");
  const int feats = mode == 1 ? C * V : C;
  double *sums = static_cast<double *>(workspace);
  cudaStream_t st = as_stream(stream);
  STGCN_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * feats, st));
  int tchunks = cdiv((long long)T * V, 4096);
  k_bn_stats_nctv<<<dim3(tchunks, C, N), 256, sizeof(double) * 2 * V, st>>>(x, C, T, V, mode, 4096, sums,
                                                                        sums + feats);
  STGCN_LAUNCH_OK();
  double inv = 1.0 / ((double)N * T * (mode == 0 ? V : 1));
  k_bn_apply_nctv<<<148 * 8, 256, 0, st>>>(x, w, b, sums, sums + feats, y, N, C, T, V, mode, inv, eps);
  STGCN_LAUNCH_OK();
  return 0;
}

// ---- temporal / pointwise convolution ----------------------------------------
size_t stgcn_conv_workspace_bytes(int N, int c_in, int c_out, int T, int V, int kernel, int stride) {
  const int ci = round4(c_in), co = round4(c_out);
  const int T_out = (T - 1) / stride + 1;
  Bump ws(nullptr, 0);
  ws.take<float>((size_t)N * T * V * ci);
  ws.take<float>((size_t)N * T_out * V * co);
  ws.take<float>((size_t)co * kernel * ci);
  ws.take<float>((size_t)co);
  return ws.peak;
}

int stgcn_conv_forward(const float *x, const float *w, const float *bias, float *y, int N, int c_in,
                       int c_out, int T, int V, int kernel, int stride, void *workspace,
                       size_t workspace_bytes, void *stream) {
  STGCN_REQUIRE(kernel % 2 == 1 && stride >= 1, "conv: kernel must be odd, stride >= 1");
  const int ci = round4(c_in), co = round4(c_out);
  const int T_out = (T - 1) / stride + 1;
  cudaStream_t st = as_stream(stream);
  Bump ws(workspace, workspace_bytes);
  float *xin = ws.take<float>((size_t)N * T * V * ci);
  float *yo = ws.take<float>((size_t)N * T_out * V * co);
  float *wp = ws.take<float>((size_t)co * kernel * ci);
  float *bp = ws.take<float>((size_t)co);
  STGCN_REQUIRE(workspace && !ws.overflow, "conv: workspace too small");
  if (to_ntvc(x, xin, N, c_in, (long long)T * V, ci, st)) return 1;
  long long nw = (long long)co * kernel * ci;
  k_pack_tcn_w<<<cdiv(nw, 256), 256, 0, st>>>(w, wp, c_out, c_in, kernel, co, ci);
  STGCN_LAUNCH_OK();
  if (bias) {
    k_pad_vec<<<cdiv(co, 256), 256, 0, st>>>(bias, bp, c_out, co);
    STGCN_LAUNCH_OK();
  }
  if (launch_gemm(xin, wp, bias ? bp : nullptr, yo, N, T, V, ci, co, kernel, stride, st)) return 1;
  return to_nctv(yo, y, N, c_out, (long long)T_out * V, co, st);
}

// ---- graph convolution (tgcn.py:58-79) -----------------------------------------
size_t stgcn_graphconv_workspace_bytes(int N, int c_in, int c_out, int K, int T, int V) {
  Bump ws(nullptr, 0);
  ws.take<float>((size_t)N * T * V * c_in);
  ws.take<float>((size_t)N * T * V * K * c_out);
  ws.take<float>((size_t)N * T * V * c_out);
  ws.take<int>((size_t)N * (V + 1));
  ws.take<int>((size_t)N * K * V * V);
  ws.take<float>((size_t)N * K * V * V);
  return ws.peak;
}

int stgcn_graphconv_forward(const float *x, const float *w, const float *bias, const float *A,
                            int a_per_sample, float *y, int N, int c_in, int c_out, int K, int T, int V,
                            void *workspace, size_t workspace_bytes, void *stream) {
  cudaStream_t st = as_stream(stream);
  Bump ws(workspace, workspace_bytes);
  float *xin = ws.take<float>((size_t)N * T * V * c_in);
  float *yy = ws.take<float>((size_t)N * T * V * K * c_out);
  float *z = ws.take<float>((size_t)N * T * V * c_out);
  STGCN_REQUIRE(workspace && !ws.overflow, "graphconv: workspace too small");
  AdjCsr csr;
  if (build_csr(A, a_per_sample, N, K, V, c_out, ws, csr, st)) return 1;
  if (to_ntvc(x, xin, N, c_in, (long long)T * V, c_in, st)) return 1;
  if (launch_gemm(xin, w, bias, yy, N, T, V, c_in, K * c_out, 1, 1, st)) return 1;
  FrameArgs a{};
  a.producer = FRAME_ADJ;
  a.frames = (long long)N * T;
  a.frames_per_sample = T;
  a.K = K; a.V = V; a.C = c_out;
  a.y = yy;
  a.adj_ptr = csr.ptr; a.adj_yoff = csr.yoff; a.adj_val = csr.val;
  a.adj_per_sample = a_per_sample;
  a.eps = kEps;
  a.out = z;
  if (launch_frame(a, st)) return 1;
  return to_nctv(z, y, N, c_out, (long long)T * V, c_out, st);
}

// ---- prepared operands -----------------------------------------------------------------
size_t stgcn_model_prepare_bytes(const stgcn_model_desc *m) {
  if (check_model(m)) return 0;
  return model_prepare_layout(*m, nullptr, 0, nullptr);
}

int stgcn_model_prepare(const stgcn_model_desc *m, void *prepared, size_t prepared_bytes, void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(m->num_layers <= 64, "too many layers");
  const size_t need = model_prepare_layout(*m, nullptr, 0, nullptr);
  if (need == 0) return 0;
  STGCN_REQUIRE(prepared && prepared_bytes >= need, "prepare: buffer too small (%zu B given, %zu B needed)",
                prepared_bytes, need);
  LayerPrep P[64];
  model_prepare_layout(*m, prepared, prepared_bytes, P);
  for (int i = 0; i < m->num_layers; ++i)
    if (prep_run(m->layers[i], m->partitions, m->num_joints, P[i], as_stream(stream))) return 1;
  return 0;
}

// ---- ST-GCN layer (stgcn.py:181-193) --------------------------------------------
size_t stgcn_layer_workspace_bytes(const stgcn_layer_desc *d, int K, int V, int N, int T) {
  if (!d) return 0;
  Bump ws(nullptr, 0);
  const int T_out = (T - 1) / d->stride + 1;
  ws.take<float>((size_t)N * T * V * d->c_in);
  ws.take<float>((size_t)N * T_out * V * d->c_out);
  const size_t m0 = ws.mark();
  for (int math = STGCN_MATH_FP32; math <= STGCN_MATH_BF16X3; ++math) {  // largest over arithmetic modes
    layer_forward_ntvc(*d, K, V, math, nullptr, nullptr, N, T, ws, nullptr);
    ws.release(m0);
  }
  return ws.peak;
}

int stgcn_layer_forward(const stgcn_layer_desc *d, int K, int V, int math, const float *x, float *y, int N,
                        int T, void *workspace, size_t workspace_bytes, void *stream) {
  STGCN_REQUIRE(d && !d->rt, "stgcn_layer_forward: descriptor must describe an ST-GCN layer");
  cudaStream_t st = as_stream(stream);
  Bump ws(workspace, workspace_bytes);
  const int T_out = (T - 1) / d->stride + 1;
  float *xin = ws.take<float>((size_t)N * T * V * d->c_in);
  float *out = ws.take<float>((size_t)N * T_out * V * d->c_out);
  STGCN_REQUIRE(workspace && !ws.overflow, "layer: workspace too small");
  if (to_ntvc(x, xin, N, d->c_in, (long long)T * V, d->c_in, st)) return 1;
  if (layer_forward_ntvc(*d, K, V, math, xin, out, N, T, ws, st)) return 1;
  return to_nctv(out, y, N, d->c_out, (long long)T_out * V, d->c_out, st);
}

// ---- ST-GCN model (stgcn.py:80-97) ------------------------------------------------
size_t stgcn_model_workspace_bytes(const stgcn_model_desc *m, int N, int T) {
  if (check_model(m)) return 0;
  return model_chunk_bytes(*m, default_chunk(*m, N, T), T);
}

// Shared body of stgcn_model_forward / stgcn_model_forward_host.  With `x_host` the trial chunks are fed
// from host memory on `copy` (all copies are queued up front, one event per chunk) so that the copy of chunk
// i+1 runs under the compute of chunk i; `x` is then the device staging buffer the chunks land in.
static int model_forward_chunks(const stgcn_model_desc *m, const float *x, float *logits, float *features, int N,
                                int T, void *workspace, size_t workspace_bytes, cudaStream_t st,
                                const float *x_host, cudaStream_t copy) {
  STGCN_REQUIRE(N > 0 && T > 0, "model: empty input");
  // largest trial chunk the caller's workspace can hold (LayerNorm: trials are independent)
  int nc = default_chunk(*m, N, T);
  while (nc > 1 && model_chunk_bytes(*m, nc, T) > workspace_bytes) nc = (nc + 1) / 2;
  STGCN_REQUIRE(workspace && model_chunk_bytes(*m, nc, T) <= workspace_bytes,
                "model: workspace too small (%zu B given, %zu B needed for %d trial(s))", workspace_bytes,
                model_chunk_bytes(*m, nc, T), nc);
  STGCN_REQUIRE(m->norm != STGCN_NORM_BATCHNORM || nc == N,
                "model: BatchNorm needs the whole batch resident; workspace too small");
  const int V = m->num_joints;
  int t_final = T;
  for (int i = 0; i < m->num_layers; ++i) t_final = (t_final - 1) / m->layers[i].stride + 1;
  const int c_last = m->layers[m->num_layers - 1].c_out;
  const size_t per_trial = (size_t)m->in_feat * T * V;
  const int n_chunks = (N + nc - 1) / nc;
  std::vector<cudaEvent_t> fed;
  int rc = 0;
  if (x_host) {
    fed.resize(n_chunks, nullptr);
    cudaEvent_t start;                                     // the staging buffer is free once `st` got here
    STGCN_CUDA_OK(cudaEventCreateWithFlags(&start, cudaEventDisableTiming));
    cudaEventRecord(start, st);
    cudaStreamWaitEvent(copy, start, 0);
    cudaEventDestroy(start);
    for (int c = 0; c < n_chunks && !rc; ++c) {
      const int n0 = c * nc, n = N - n0 < nc ? N - n0 : nc;
      float *dst = const_cast<float *>(x) + (size_t)n0 * per_trial;
      if (cudaEventCreateWithFlags(&fed[c], cudaEventDisableTiming) != cudaSuccess ||
          cudaMemcpyAsync(dst, x_host + (size_t)n0 * per_trial, (size_t)n * per_trial * sizeof(float),
                          cudaMemcpyHostToDevice, copy) != cudaSuccess ||
          cudaEventRecord(fed[c], copy) != cudaSuccess)
        rc = fail("forward_host: staging copy of chunk %d failed: %s", c, cudaGetErrorString(cudaGetLastError()));
    }
  }
  for (int c = 0; c < n_chunks && !rc; ++c) {
    const int n0 = c * nc, n = N - n0 < nc ? N - n0 : nc;
    if (x_host) cudaStreamWaitEvent(st, fed[c], 0);
    Bump ws(workspace, workspace_bytes);
    rc = model_chunk(*m, x + (size_t)n0 * per_trial, logits + (size_t)n0 * m->num_classes,
                     features ? features + (size_t)n0 * c_last * t_final * V : nullptr, n, T, ws, st);
  }
  for (cudaEvent_t e : fed)
    if (e) cudaEventDestroy(e);
  return rc;
}

int stgcn_model_forward(const stgcn_model_desc *m, const float *x, float *logits, float *features, int N,
                        int T, void *workspace, size_t workspace_bytes, void *stream) {
  if (check_model(m)) return 1;
  return model_forward_chunks(m, x, logits, features, N, T, workspace, workspace_bytes, as_stream(stream), nullptr,
                              nullptr);
}

// ---- sliding-window inference (utils/segment_generator.py:109-154, processor.py:374-380) ------------
int stgcn_model_forward_windows(const stgcn_model_desc *m, const float *captures, float *logits, int n_windows,
                                int W, int L_pad, void *workspace, size_t workspace_bytes, void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(captures && logits && workspace && n_windows > 0 && W > 0 && L_pad >= n_windows + W - 1,
                "forward_windows: bad arguments (need L_pad >= n_windows + W - 1)");
  STGCN_REQUIRE(m->norm == STGCN_NORM_LAYERNORM, "forward_windows: LayerNorm only (windows are independent trials)");
  cudaStream_t st = as_stream(stream);
  const int V = m->num_joints;
  // window n = frames [n, n + W) of the padded capture: a (n_windows, C, W, V) VIEW, never materialised
  const long long xs[3] = {(long long)V, (long long)L_pad * V, (long long)V};
  int nc = default_chunk(*m, n_windows, W);
  while (nc > 1 && model_chunk_bytes(*m, nc, W) > workspace_bytes) nc = (nc + 1) / 2;
  STGCN_REQUIRE(model_chunk_bytes(*m, nc, W) <= workspace_bytes, "forward_windows: workspace too small");
  for (int n0 = 0; n0 < n_windows; n0 += nc) {
    const int n = n_windows - n0 < nc ? n_windows - n0 : nc;
    Bump ws(workspace, workspace_bytes);
    if (model_chunk(*m, captures + (size_t)n0 * V, logits + (size_t)n0 * m->num_classes, nullptr, n, W, ws, st,
                    nullptr, nullptr, xs))
      return 1;
  }
  return 0;
}

// ---- T-split forward -----------------------------------------------------------------------
size_t stgcn_model_halo_bytes(const stgcn_model_desc *m, int N) {
  if (check_model(m)) return 0;
  int cmax = 0;
  for (int i = 0; i < m->num_layers; ++i) cmax = m->layers[i].c_out > cmax ? m->layers[i].c_out : cmax;
  return (size_t)2 * N * kHalo * m->num_joints * cmax * sizeof(__nv_bfloat16);
}

size_t stgcn_model_tsplit_workspace_bytes(const stgcn_model_desc *m, int N, int T_local) {
  if (check_model(m)) return 0;
  Bump ws(nullptr, 0);
  stgcn_halo_desc h{};
  model_chunk(*m, nullptr, nullptr, nullptr, N, T_local, ws, nullptr, &h, nullptr);
  return ws.peak;
}

int stgcn_model_forward_tsplit(const stgcn_model_desc *m, const float *x, float *pooled_sums, int N, int T_local,
                               const stgcn_halo_desc *halo, void *workspace, size_t workspace_bytes,
                               void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(halo && pooled_sums && workspace && N > 0 && T_local > 0, "tsplit: null argument or empty chunk");
  STGCN_REQUIRE(m->norm == STGCN_NORM_LAYERNORM && m->math != STGCN_MATH_FP32,
                "tsplit: needs LayerNorm and a tensor-core math mode (batch statistics would span ranks)");
  int total_stride = 1;
  for (int i = 0; i < m->num_layers; ++i) total_stride *= m->layers[i].stride;
  STGCN_REQUIRE(!halo->has_right || T_local % total_stride == 0,
                "tsplit: every chunk but the last must hold a multiple of %d frames (got %d)", total_stride, T_local);
  Bump ws(workspace, workspace_bytes);
  return model_chunk(*m, x, nullptr, nullptr, N, T_local, ws, as_stream(stream), halo, pooled_sums);
}

// ---- RT-ST-GCN continual step (rtstgcn.py:137-157, 528-553, 591-627) ---------------
size_t rtstgcn_state_bytes(const stgcn_model_desc *m, int B) {
  RtLayout L;
  if (check_model(m) || rt_layout(*m, B, L)) return 0;
  return L.total;
}

int rtstgcn_state_reset(const stgcn_model_desc *m, void *state, int B, int first, int count, void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(state && first >= 0 && count >= 0 && first + count <= B, "state_reset: bad stream range");
  RtLayout L;
  if (rt_layout(*m, B, L)) return 1;
  cudaStream_t st = as_stream(stream);
  char *sb = static_cast<char *>(state);
  if (first == 0 && count == B) {
    STGCN_CUDA_OK(cudaMemsetAsync(sb, 0, L.total, st));
    return 0;
  }
  STGCN_CUDA_OK(cudaMemsetAsync(sb + L.counters + sizeof(int32_t) * first, 0, sizeof(int32_t) * count, st));
  for (int i = 0; i < m->num_layers; ++i) {
    const stgcn_layer_desc &d = m->layers[i];
    size_t per = (size_t)m->num_joints * d.c_out * sizeof(float);
    const size_t perf = rt_fifo_bf16(*m, B) ? per / 2 : per;
    int F = d.stride * (d.kernel - 1) + 1;
    for (int f = 0; f < F; ++f)
      STGCN_CUDA_OK(cudaMemsetAsync(sb + L.fifo[i] + ((size_t)f * B + first) * perf, 0, perf * count, st));
    for (int s = 0; s < d.stride; ++s)
      STGCN_CUDA_OK(cudaMemsetAsync(sb + L.acc[i] + ((size_t)s * B + first) * per, 0, per * count, st));
  }
  return 0;
}

size_t rtstgcn_step_workspace_bytes(const stgcn_model_desc *m, int B) {
  if (check_model(m)) return 0;
  Bump ws(nullptr, 0);
  rt_step(*m, nullptr, nullptr, nullptr, B, ws, nullptr);
  return ws.peak;
}

int rtstgcn_step(const stgcn_model_desc *m, const float *x, void *state, float *logits, int B,
                 void *workspace, size_t workspace_bytes, void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(state && workspace && B > 0, "rtstgcn_step: null state/workspace or empty batch");
  Bump ws(workspace, workspace_bytes);
  return rt_step(*m, x, state, logits, B, ws, as_stream(stream));
}

int rtstgcn_step_top5(const stgcn_model_desc *m, const float *x, void *state, float *logits, int32_t *top5, int B,
                      void *workspace, size_t workspace_bytes, void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(state && workspace && top5 && B > 0, "rtstgcn_step_top5: null state/workspace/top5 or empty batch");
  STGCN_REQUIRE(m->num_classes <= 4096, "rtstgcn_step_top5: too many classes");
  Bump ws(workspace, workspace_bytes);
  if (rt_small_supported(*m, B)) {
    // the one-cluster-kernel latency path writes logits only: rank them with the same routine afterwards
    if (rt_step(*m, x, state, logits, B, ws, as_stream(stream))) return 1;
    k_topk5<<<cdiv(B, 8), 256, 0, as_stream(stream)>>>(logits, B, m->num_classes, top5);
    STGCN_LAUNCH_OK();
    return 0;
  }
  return rt_step(*m, x, state, logits, B, ws, as_stream(stream), top5);
}

size_t rtstgcn_layer_state_bytes(const stgcn_layer_desc *d, int V, int B) {
  if (!d) return 0;
  size_t slot = (size_t)B * V * d->c_out * sizeof(float);
  return slot * (size_t)(d->stride * (d->kernel - 1) + 1 + d->stride);
}

size_t rtstgcn_layer_workspace_bytes(const stgcn_layer_desc *d, int K, int V, int B) {
  if (!d) return 0;
  Bump ws(nullptr, 0);
  ws.take<float>((size_t)B * V * d->c_in);
  ws.take<float>((size_t)B * V * d->c_out);
  const size_t m0 = ws.mark();
  for (int math = STGCN_MATH_FP32; math <= STGCN_MATH_BF16X3; ++math) {
    rt_layer_step_ntvc(*d, K, V, math, nullptr, nullptr, nullptr, nullptr, nullptr, B, ws, nullptr);
    ws.release(m0);
  }
  return ws.peak;
}

int rtstgcn_layer_step(const stgcn_layer_desc *d, int K, int V, int math, const float *x, float *y,
                       void *layer_state, int32_t *frame_counter, int B, void *workspace,
                       size_t workspace_bytes, void *stream) {
  STGCN_REQUIRE(d && d->rt, "rtstgcn_layer_step: descriptor must describe an online layer");
  STGCN_REQUIRE(layer_state && frame_counter && workspace, "rtstgcn_layer_step: null state/workspace");
  cudaStream_t st = as_stream(stream);
  Bump ws(workspace, workspace_bytes);
  float *xin = ws.take<float>((size_t)B * V * d->c_in);
  float *out = ws.take<float>((size_t)B * V * d->c_out);
  STGCN_REQUIRE(!ws.overflow, "rt layer: workspace too small");
  const int F = d->stride * (d->kernel - 1) + 1;
  float *fifo = static_cast<float *>(layer_state);
  float *acc = fifo + (size_t)F * B * V * d->c_out;
  if (to_ntvc(x, xin, B, d->c_in, V, d->c_in, st)) return 1;
  if (rt_layer_step_ntvc(*d, K, V, math, xin, out, fifo, acc, frame_counter, B, ws, st)) return 1;
  k_advance_counters<<<cdiv(B, 256), 256, 0, st>>>(frame_counter, 0, B,
                                                   rt_counter_period_of(F, d->stride, 1));
  STGCN_LAUNCH_OK();
  return to_nctv(out, y, B, d->c_out, V, d->c_out, st);
}

// ---- RT-ST-GCN training-time layer (rtstgcn.py:343-389) ---------------------------------------
namespace {
int offline_layer(const stgcn_layer_desc &d, int K, int V, const float *x, float *y, int N, int L, Bump &ws,
                  cudaStream_t st) {
  if (check_layer(d)) return 1;
  STGCN_REQUIRE(d.norm == STGCN_NORM_LAYERNORM, "offline layer: LayerNorm only on the B200 path");
  const long long rows = (long long)N * L * V;
  float *xin = ws.take<float>((size_t)rows * d.c_in);
  float *yy = ws.take<float>((size_t)rows * K * d.c_out);
  float *z = ws.take<float>((size_t)rows * d.c_out);
  float *o = ws.take<float>((size_t)rows * d.c_out);
  float *qr = d.residual == STGCN_RES_CONV ? yy : nullptr;          // reuses the 1x1 output buffer
  float *out = ws.take<float>((size_t)rows * d.c_out);
  AdjCsr csr;
  if (build_csr(d.a_eff, 0, N, K, V, d.c_out, ws, csr, st)) return 1;
  if (ws.measuring()) return 0;
  STGCN_REQUIRE(!ws.overflow, "offline layer: workspace too small");
  if (to_ntvc(x, xin, N, d.c_in, (long long)L * V, d.c_in, st)) return 1;
  if (launch_gemm(xin, d.gcn_w, d.gcn_b, yy, N, L, V, d.c_in, K * d.c_out, 1, 1, st)) return 1;
  FrameArgs a{};
  a.producer = FRAME_ADJ;
  a.frames = (long long)N * L;
  a.frames_per_sample = L;
  a.K = K; a.V = V; a.C = d.c_out;
  a.y = yy;
  a.adj_ptr = csr.ptr; a.adj_yoff = csr.yoff; a.adj_val = csr.val;
  a.eps = kEps;
  a.out = z;
  if (launch_frame(a, st)) return 1;
  const int taps = d.kernel / d.stride;                              // rtstgcn.py:369
  k_causal_tap_sum<<<148 * 8, 256, 0, st>>>(z, o, N, L, V * d.c_out, taps, d.stride);
  STGCN_LAUNCH_OK();
  if (qr && launch_gemm(xin, d.res_w, nullptr, qr, N, L, V, d.c_in, d.c_out, 1, 1, st)) return 1;
  FrameArgs f{};
  f.producer = FRAME_LOAD;
  f.frames = (long long)N * L;
  f.frames_per_sample = L;
  f.K = K; f.V = V; f.C = d.c_out;
  f.a = o;
  f.norm_a = 1; f.na_w = d.n1_w; f.na_b = d.n1_b;
  f.relu_mid = 1;
  if (d.residual == STGCN_RES_IDENTITY) { f.b_mode = B_RAW; f.b = xin; }
  else if (d.residual == STGCN_RES_CONV) { f.b_mode = B_LN; f.b = qr; f.nb_w = d.nr_w; f.nb_b = d.nr_b; }
  f.relu_out = 1;
  f.eps = kEps;
  f.out = out;
  if (launch_frame(f, st)) return 1;
  return to_nctv(out, y, N, d.c_out, (long long)L * V, d.c_out, st);
}
}  // namespace

size_t rtstgcn_offline_layer_workspace_bytes(const stgcn_layer_desc *d, int K, int V, int N, int L) {
  if (!d) return 0;
  Bump ws(nullptr, 0);
  offline_layer(*d, K, V, nullptr, nullptr, N, L, ws, nullptr);
  return ws.peak;
}

int rtstgcn_offline_layer_forward(const stgcn_layer_desc *d, int K, int V, const float *x, float *y, int N, int L,
                                  void *workspace, size_t workspace_bytes, void *stream) {
  STGCN_REQUIRE(d && d->rt, "offline layer: descriptor must describe an RT-ST-GCN layer (rt == 1)");
  STGCN_REQUIRE(workspace && x && y && N > 0 && L > 0, "offline layer: null argument or empty input");
  Bump ws(workspace, workspace_bytes);
  return offline_layer(*d, K, V, x, y, N, L, ws, as_stream(stream));
}

int stgcn_mean_joints_forward(const float *x, float *y, long long rows, int V, void *stream) {
  STGCN_REQUIRE(x && y && rows > 0 && V > 0, "mean_joints: bad arguments");
  k_mean_joints<<<cdiv(rows, 256), 256, 0, as_stream(stream)>>>(x, y, rows, V);
  STGCN_LAUNCH_OK();
  return 0;
}

// ---- CoST-GCN continual step (costgcn.py:81-99, 190-211) ---------------------------------------
size_t costgcn_state_bytes(const stgcn_model_desc *m, int B) {
  CostLayout L;
  if (check_model(m) || cost_layout(*m, B, L)) return 0;
  return L.total;
}

int costgcn_state_reset(const stgcn_model_desc *m, void *state, int B, int first, int count, void *stream) {
  if (check_model(m) || cost_check(*m)) return 1;
  STGCN_REQUIRE(state && first >= 0 && count >= 0 && first + count <= B, "costgcn_state_reset: bad stream range");
  CostLayout L;
  if (cost_layout(*m, B, L)) return 1;
  cudaStream_t st = as_stream(stream);
  const int planes = m->math == STGCN_MATH_BF16X3 ? 2 : 1, V = m->num_joints;
  char *sb = static_cast<char *>(state);
  for (int i = 0; i < m->num_layers && count > 0; ++i) {
    const stgcn_layer_desc &d = m->layers[i];
    const int F = d.stride * (d.kernel - 1) + 1, R2 = cost_res_slots(d);
    const long long tot = (long long)count * V * d.c_out * F;
    k_cost_ring_init<<<cdiv(tot, 256), 256, 0, st>>>(d.n1_b, reinterpret_cast<__nv_bfloat16 *>(sb + L.u[i]), planes, F, B,
                                                    V, d.c_out, first, count);
    STGCN_LAUNCH_OK();
    const size_t per = (size_t)V * d.c_out * sizeof(__nv_bfloat16);
    for (int sl = 0; sl < planes * R2; ++sl)
      STGCN_CUDA_OK(cudaMemsetAsync(sb + L.r[i] + ((size_t)sl * B + first) * per, 0, per * count, st));
  }
  return 0;
}

size_t costgcn_step_workspace_bytes(const stgcn_model_desc *m, int B) {
  if (check_model(m)) return 0;
  Bump ws(nullptr, 0);
  if (cost_step(*m, nullptr, nullptr, 0, nullptr, B, ws, nullptr)) return 0;
  return ws.peak;
}

int costgcn_step(const stgcn_model_desc *m, const float *x, void *state, long long t, float *logits, int B,
                 void *workspace, size_t workspace_bytes, void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(x && state && logits && workspace && B > 0 && t >= 0, "costgcn_step: null argument, empty batch or t < 0");
  Bump ws(workspace, workspace_bytes);
  return cost_step(*m, x, state, t, logits, B, ws, as_stream(stream));
}

// ---- host-buffer entry points --------------------------------------------------------
int stgcn_model_forward_host(const stgcn_model_desc *m, const float *x_host, float *logits_host, int N,
                             int T, void *device_io, void *workspace, size_t workspace_bytes,
                             void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(x_host && logits_host && device_io, "forward_host: null buffer");
  cudaStream_t st = as_stream(stream);
  const size_t nx = (size_t)N * m->in_feat * T * m->num_joints, nl = (size_t)N * m->num_classes;
  float *dx = static_cast<float *>(device_io);
  float *dl = dx + nx;
  // per-device staging stream (created once): chunk copies overlap the previous chunk's kernels
  static std::mutex mu;
  static cudaStream_t copy_streams[kMaxDevices] = {nullptr};
  int dev = 0;
  STGCN_CUDA_OK(cudaGetDevice(&dev));
  STGCN_REQUIRE(dev >= 0 && dev < kMaxDevices, "forward_host: device index %d out of range", dev);
  cudaStream_t copy;
  {
    std::lock_guard<std::mutex> g(mu);
    if (!copy_streams[dev]) STGCN_CUDA_OK(cudaStreamCreateWithFlags(&copy_streams[dev], cudaStreamNonBlocking));
    copy = copy_streams[dev];
  }
  if (model_forward_chunks(m, dx, dl, nullptr, N, T, workspace, workspace_bytes, st, x_host, copy)) {
    cudaStreamSynchronize(copy);
    cudaStreamSynchronize(st);
    return 1;
  }
  STGCN_CUDA_OK(cudaMemcpyAsync(logits_host, dl, nl * sizeof(float), cudaMemcpyDeviceToHost, st));
  STGCN_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int rtstgcn_step_host(const stgcn_model_desc *m, const float *x_host, void *state, float *logits_host,
                      int B, void *device_io, void *workspace, size_t workspace_bytes, void *stream) {
  if (check_model(m)) return 1;
  STGCN_REQUIRE(x_host && logits_host && device_io, "step_host: null buffer");
  cudaStream_t st = as_stream(stream);
  const size_t nx = (size_t)B * m->in_feat * m->num_joints, nl = (size_t)B * m->num_classes;
  float *dx = static_cast<float *>(device_io);
  float *dl = dx + nx;
  STGCN_CUDA_OK(cudaMemcpyAsync(dx, x_host, nx * sizeof(float), cudaMemcpyHostToDevice, st));
  if (rtstgcn_step(m, dx, state, dl, B, workspace, workspace_bytes, stream)) return 1;
  STGCN_CUDA_OK(cudaMemcpyAsync(logits_host, dl, nl * sizeof(float), cudaMemcpyDeviceToHost, st));
  STGCN_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
