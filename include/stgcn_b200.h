/*
 * stgcn_b200.h -- C ABI of the B200-native ST-GCN / RT-ST-GCN forward path.
 *
 * The reference (maximyudayev/Realtime-ST-GCN) has no FFI layer: its hot path is
 * a set of torch.nn.Module.forward methods.  Each entry point below replaces the
 * body of one of them; the Python modules under realtime-st-gcn_b200/models/
 * keep the reference constructor/forward signatures and call these through
 * ctypes (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors use the reference layout (N, C, T, V), fp32, contiguous, unless
 *     stated otherwise; parameter tensors use the reference state_dict layouts;
 *   - the library never allocates or frees memory: the caller (torch) owns all
 *     inputs, outputs, parameters, workspaces and FIFO state;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - return value 0 = success, non-zero = error; stgcn_last_error() returns a
 *     thread-local message.  There is no CPU fallback: a non-sm_100 device or a
 *     missing GPU is an error.
 */
#ifndef STGCN_B200_H
#define STGCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STGCN_ABI_VERSION 2

enum { STGCN_NORM_LAYERNORM = 0, STGCN_NORM_BATCHNORM = 1 };
enum { STGCN_RES_NONE = 0, STGCN_RES_IDENTITY = 1, STGCN_RES_CONV = 2 };
/* arithmetic of the GEMM stages */
enum {
  STGCN_MATH_FP32 = 0,      /* fp32 CUDA-core FMA (exact reference arithmetic); the <= 14-stream
                               continual step (one cluster kernel) uses 3xTF32 MMAs instead   */
  STGCN_MATH_BF16X3 = 1,    /* tcgen05 bf16 hi/lo split, 3 MMAs, fp32 accumulate (1e-4) */
  STGCN_MATH_BF16 = 2       /* tcgen05 single bf16 MMA, fp32 accumulate                 */
};

/* One st_gcn block.  Replaces StgcnLayer.forward (reference
 * models/stgcn/stgcn.py:181-193) when `rt == 0` and OnlineLayer.forward
 * (models/rtstgcn/rtstgcn.py:528-553 + AggregateStgcn.forward :591-627) when
 * `rt == 1`.  Parameter pointers use the reference layouts:
 *   gcn_w  (K*c_out, c_in)      gcn.conv.weight / conv.weight        (1x1)
 *   gcn_b  (K*c_out)            gcn.conv.bias   / conv.bias
 *   a_eff  (K, V, V)            A * edge_importance (or per-sample (N,K,V,V))
 *   n1_w/b (c_out, V) | (c_out) tcn.0 / bn_relu.0   LayerNorm | BatchNorm
 *   tcn_w  (c_out, c_out, G)    tcn.2.weight  (G x 1 temporal conv; ST-GCN only)
 *   tcn_b  (c_out)              tcn.2.bias
 *   n2_w/b                      tcn.3         (ST-GCN only)
 *   res_w  (c_out, c_in)        residual.0.weight
 *   res_b  (c_out) or NULL      residual.0.bias (NULL for RT: conv has no bias)
 *   nr_w/b                      residual.1
 */
typedef struct stgcn_layer_desc {
  int32_t c_in, c_out;
  int32_t kernel;        /* temporal kernel Gamma (odd) */
  int32_t stride;
  int32_t residual;      /* STGCN_RES_* */
  int32_t norm;          /* STGCN_NORM_* */
  int32_t rt;            /* 0: ST-GCN layer, 1: RT-ST-GCN online layer, 2: CoST-GCN layer */
  int32_t a_per_sample;  /* 0: a_eff is (K,V,V); 1: (N,K,V,V) */
  const float *gcn_w, *gcn_b, *a_eff;
  const float *n1_w, *n1_b;
  const float *tcn_w, *tcn_b;
  const float *n2_w, *n2_b;
  const float *res_w, *res_b;
  const float *nr_w, *nr_b;
} stgcn_layer_desc;

/* Whole model.  Replaces models.stgcn.Model.forward (stgcn.py:80-97) and
 * models.rtstgcn.Model.forward with online layers (rtstgcn.py:137-157).
 *   norm_in_w/b : LayerNorm (c_in, V) | BatchNorm1d (V*c_in) (batchnorm.py:13-23)
 *   fcn_in_w (c0, c_in), fcn_in_b (c0); fcn_out_w (classes, c_last), fcn_out_b
 */
typedef struct stgcn_model_desc {
  int32_t in_feat, num_joints, partitions, num_classes, num_layers;
  int32_t norm;          /* STGCN_NORM_* */
  int32_t math;          /* STGCN_MATH_* */
  int32_t reserved;      /* bit 0: do not use the few-streams cluster kernel in rtstgcn_step;
                          * bit 1: the adjacency has at most 6*V non-zeros (tree skeletons): the graph-conv
                          *        stage may use one pre-scaled weight copy per edge (kernels_gcnw.cuh) */
  const float *norm_in_w, *norm_in_b;
  const float *fcn_in_w, *fcn_in_b;
  const float *fcn_out_w, *fcn_out_b;
  const stgcn_layer_desc *layers;   /* HOST pointer to num_layers descriptors */
  /* Optional: operands prepared once per set of weights by stgcn_model_prepare (bf16 hi/lo
   * weight planes, adjacency CSR, bias through the adjacency).  NULL: every forward rebuilds
   * them in its workspace (a few small kernels per layer). */
  const void *prepared;
  size_t prepared_bytes;
} stgcn_model_desc;

/* ---- library ------------------------------------------------------------- */
int stgcn_abi_version(void);
const char *stgcn_last_error(void);
/* 0 if `device` is an sm_100 GPU this build can run on. */
int stgcn_device_check(int device);
/* Number of kernels this library has launched in this process, all devices and threads
 * (bench.py's gpu_launches). */
long long stgcn_launch_count(void);
/* Measurement aid: between begin/end every kernel launch is bracketed by CUDA events on its
 * stream; end() synchronises and returns summed milliseconds and launch counts per kernel class:
 * 0 layout, 1 gemm_1x1, 2 gemm_tcn, 3 frame(adjacency/norm/residual), 4 batchnorm, 5 embed,
 * 6 pool+classifier, 7 misc.  State is per DEVICE (the calling thread's current device): a process that
 * runs one replica thread per GPU profiles each device independently; launches from several threads onto
 * one device are collected under a lock. */
#define STGCN_KERNEL_CLASSES 8
int stgcn_profile_begin(void);
int stgcn_profile_end(float *ms_per_class_host, long long *launches_per_class_host, int n_classes);

/* ---- primitives (models/utils) -------------------------------------------- */
/* LayerNorm over (C,V) per (n,t), unbiased variance, affine (C,V).
 * Replaces LayerNorm.forward, models/utils/layernorm.py:22-28. */
int stgcn_layernorm_forward(const float *x, const float *w, const float *b, float *y,
                            int N, int C, int T, int V, float eps, void *stream);
/* Batch-statistics BatchNorm.  mode 0: per channel over (N,T,V)
 * (nn.BatchNorm2d(track_running_stats=False), stgcn.py:152); mode 1: per (v,c)
 * feature over (N,T), weight index v*C+c (BatchNorm1d.forward,
 * models/utils/batchnorm.py:13-23).  workspace >= stgcn_batchnorm_workspace_bytes. */
size_t stgcn_batchnorm_workspace_bytes(int C, int V, int mode);
int stgcn_batchnorm_forward(const float *x, const float *w, const float *b, float *y,
                            int N, int C, int T, int V, float eps, int mode,
                            void *workspace, size_t workspace_bytes, void *stream);
/* Gamma x 1 convolution over time with stride and zero padding (kernel-1)/2,
 * weight (c_out, c_in, kernel), optional bias.  Replaces the nn.Conv2d calls at
 * stgcn.py:85,95,154-159,166-170 and tgcn.py:71. */
size_t stgcn_conv_workspace_bytes(int N, int c_in, int c_out, int T, int V, int kernel, int stride);
int stgcn_conv_forward(const float *x, const float *w, const float *bias, float *y,
                       int N, int c_in, int c_out, int T, int V, int kernel, int stride,
                       void *workspace, size_t workspace_bytes, void *stream);
/* Graph convolution: 1x1 conv c_in -> K*c_out, contraction with A over (k,v).
 * Replaces ConvTemporalGraphical.forward, models/utils/tgcn.py:58-79. */
size_t stgcn_graphconv_workspace_bytes(int N, int c_in, int c_out, int K, int T, int V);
int stgcn_graphconv_forward(const float *x, const float *w, const float *bias, const float *A,
                            int a_per_sample, float *y, int N, int c_in, int c_out, int K,
                            int T, int V, void *workspace, size_t workspace_bytes, void *stream);

/* ---- prepared operands ------------------------------------------------------ */
/* Bytes of the device buffer stgcn_model_prepare fills (0 if nothing to prepare). */
size_t stgcn_model_prepare_bytes(const stgcn_model_desc *m);
/* Fill `prepared` from the parameters the descriptor points at (enqueued on `stream`); then set
 * m->prepared / m->prepared_bytes.  Must be repeated when a parameter or the adjacency changes. */
int stgcn_model_prepare(const stgcn_model_desc *m, void *prepared, size_t prepared_bytes, void *stream);

/* ---- ST-GCN layer / model --------------------------------------------------- */
size_t stgcn_layer_workspace_bytes(const stgcn_layer_desc *d, int K, int V, int N, int T);
/* x (N,c_in,T,V) -> y (N,c_out,T_out,V), T_out = (T-1)/stride + 1. */
int stgcn_layer_forward(const stgcn_layer_desc *d, int K, int V, int math, const float *x,
                        float *y, int N, int T, void *workspace, size_t workspace_bytes,
                        void *stream);
size_t stgcn_model_workspace_bytes(const stgcn_model_desc *m, int N, int T);
/* x (N,in_feat,T,V) -> logits (N,num_classes); features (optional, may be NULL):
 * pre-pool trunk output in the reference layout (N,c_last,T_final,V).
 * t_halo_left/right: T-split support -- see stgcn_model_forward_tsplit. */
int stgcn_model_forward(const stgcn_model_desc *m, const float *x, float *logits,
                        float *features, int N, int T, void *workspace,
                        size_t workspace_bytes, void *stream);

/* ---- sliding-window inference ------------------------------------------------------ */
/* The reference's continual use of ST-GCN (WindowSegment, utils/segment_generator.py:109-154;
 * processor.py:374-380): every frame is classified from the window of the W frames ending at it,
 * windows are the batch.  captures (in_feat, L_pad, V) is ONE trial already padded with W-1 zero
 * frames in front; window n covers frames [n, n+W).  The windows are read in place through
 * overlapping strides -- the (n_windows, C, W, V) batch the reference materialises with
 * unfold().contiguous() never exists.  logits (n_windows, num_classes).  LayerNorm only. */
int stgcn_model_forward_windows(const stgcn_model_desc *m, const float *captures, float *logits,
                                int n_windows, int W, int L_pad, void *workspace,
                                size_t workspace_bytes, void *stream);

/* ---- T-split: one long trial partitioned along time across ranks ----------------- */
/* Every rank holds a contiguous chunk of T_local frames (a multiple of the trunk's total temporal
 * stride on every rank but the last).  All stages of a layer are frame-local except the Gamma x 1
 * temporal convolution, whose input u = relu(norm1(gcn(x))) needs (kernel-1)/2 frames of the
 * neighbouring chunks (reference: the zero padding of tcn.2, stgcn.py:154-159; the reference's own
 * long-sequence mechanism recomputes a 72-frame halo instead, utils/segment_generator.py:49-54).
 * Between the two stages of every layer the library packs its boundary frames of u into
 * send_left/send_right, calls `exchange(ctx, layer, bytes)` on the host -- which must enqueue, in
 * stream order, send_left -> left rank's recv_right and send_right -> right rank's recv_left (NCCL
 * send/recv) -- and unpacks recv_left/recv_right into the halo frames.  A rank without a left (right)
 * neighbour zero-fills that halo: the convolution's zero padding.  LayerNorm only. */
typedef struct stgcn_halo_desc {
  int32_t has_left, has_right;
  void *send_left, *send_right, *recv_left, *recv_right;   /* device, >= capacity bytes each */
  size_t capacity;
  int (*exchange)(void *ctx, int layer, size_t bytes);     /* host callback, 0 = success */
  void *ctx;
} stgcn_halo_desc;
/* Bytes each of the four staging buffers must hold for N trials. */
size_t stgcn_model_halo_bytes(const stgcn_model_desc *m, int N);
size_t stgcn_model_tsplit_workspace_bytes(const stgcn_model_desc *m, int N, int T_local);
/* x (N,in_feat,T_local,V) -> pooled_sums (N, c_last): sums over this rank's final frames and joints.
 * The caller all-reduces them, divides by (T_final_total * V) and applies fcn_out (stgcn.py:92-95). */
int stgcn_model_forward_tsplit(const stgcn_model_desc *m, const float *x, float *pooled_sums, int N,
                               int T_local, const stgcn_halo_desc *halo, void *workspace,
                               size_t workspace_bytes, void *stream);

/* ---- RT-ST-GCN continual step ----------------------------------------------- */
/* Per-stream FIFO/accumulator state for B concurrent streams (fp32):
 * per layer fifo[F][B][V][c_out] (F = stride*(kernel-1)+1) and acc[stride][B][V][c_out],
 * plus one int32 frame counter per stream.  Replaces the Python attributes
 * fifo/accumulator/fifo_idx/accumulator_idx of AggregateStgcn (rtstgcn.py:576-579). */
size_t rtstgcn_state_bytes(const stgcn_model_desc *m, int B);
/* Zero the state of streams [first, first+count) (all layers). */
int rtstgcn_state_reset(const stgcn_model_desc *m, void *state, int B, int first, int count,
                        void *stream);
size_t rtstgcn_step_workspace_bytes(const stgcn_model_desc *m, int B);
/* One frame for every stream: x (B,in_feat,1,V) -> logits (B,num_classes).
 * From 1024 streams on (STGCN_RT_OVERLAP) the streams are stepped as two halves, the second on an internal
 * per-device stream forked from and joined back into `stream` with events; the call may be captured into a CUDA
 * graph, but the first call on a device creates that stream, so make one un-captured call first (a warm-up
 * launch is needed anyway for the kernels' function attributes). */
int rtstgcn_step(const stgcn_model_desc *m, const float *x, void *state, float *logits, int B,
                 void *workspace, size_t workspace_bytes, void *stream);
/* The same step, and top5 (B, 5) int32 = indices of the five largest logits of every stream, best
 * first: what the reference's Statistics computes with torch.topk(predictions, 5, dim=1)
 * (utils/statistics.py:4-16), produced by the kernel that produces the logits. */
int rtstgcn_step_top5(const stgcn_model_desc *m, const float *x, void *state, float *logits, int32_t *top5,
                      int B, void *workspace, size_t workspace_bytes, void *stream);
/* One OnlineLayer.forward on x (B,c_in,1,V) -> y (B,c_out,1,V); layer_state is the
 * slice for this layer laid out as rtstgcn_layer_state_bytes describes. */
size_t rtstgcn_layer_state_bytes(const stgcn_layer_desc *d, int V, int B);
size_t rtstgcn_layer_workspace_bytes(const stgcn_layer_desc *d, int K, int V, int B);
int rtstgcn_layer_step(const stgcn_layer_desc *d, int K, int V, int math, const float *x,
                       float *y, void *layer_state, int32_t *frame_counter, int B,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---- RT-ST-GCN training-time (whole-sequence) layer ------------------------------ */
/* OfflineLayer.forward (rtstgcn.py:343-389, with the `self.toeplitz` -> `toeplitz` fix): graph
 * convolution, causal sum of kernel/stride taps spaced `stride`, LayerNorm + ReLU, residual, ReLU.
 * x (N,c_in,L,V) -> y (N,c_out,L,V); fp32 CUDA-core arithmetic.  d->rt must be 1 (bias-free,
 * stride-free residual conv); d->a_eff = A * edge_importance. */
size_t rtstgcn_offline_layer_workspace_bytes(const stgcn_layer_desc *d, int K, int V, int N, int L);
int rtstgcn_offline_layer_forward(const stgcn_layer_desc *d, int K, int V, const float *x, float *y, int N,
                                  int L, void *workspace, size_t workspace_bytes, void *stream);
/* Mean over joints, x (N,C,L,V) -> y (N,C,L,1): nn.AvgPool2d((1,V)) of rtstgcn.py:127,149. */
int stgcn_mean_joints_forward(const float *x, float *y, long long rows, int V, void *stream);

/* ---- host-buffer entry points (processor.py:367,380 / :418 equivalents) ----- */
/* x_host/logits_host are HOST buffers (pinned for async copies); device_io must
 * hold N*in_feat*T*V + N*num_classes floats.  The model forward stages the input chunk by chunk
 * (the same trial chunks the forward computes in) on an internal per-device copy stream, so the H2D
 * copy of chunk i+1 runs under the kernels of chunk i; kernels and the D2H of the logits run on
 * `stream`, which is synchronised before returning. */
int stgcn_model_forward_host(const stgcn_model_desc *m, const float *x_host, float *logits_host,
                             int N, int T, void *device_io, void *workspace,
                             size_t workspace_bytes, void *stream);
int rtstgcn_step_host(const stgcn_model_desc *m, const float *x_host, void *state,
                      float *logits_host, int B, void *device_io, void *workspace,
                      size_t workspace_bytes, void *stream);

/* ---- CoST-GCN continual step -----------------------------------------------
 * Replaces models.costgcn.Model.forward on one frame (reference
 * models/costgcn/costgcn.py:81-99) with its StgcnLayer.forward (:190-211): per
 * layer a FIFO of graph-convolved frames, LayerNorm + ReLU, a learnable
 * Gamma x 1 convolution with dilation = stride over the FIFO, second LayerNorm,
 * plus the residual of Gamma/2 frames ago.  Layers are stgcn_layer_desc with
 * rt == 2 (same parameter pointers as an ST-GCN layer; res_b is the bias of
 * the residual conv).  B streams share the caller's frame index t (0, 1, ...);
 * costgcn_state_reset initialises (or re-initialises, for a stream range) the
 * rings and must be called before the first step.  LayerNorm, math in
 * {bf16x3, bf16}, channels in {64,128,256}, sparse adjacency, prepared
 * operands (stgcn_model_prepare) only.
 * x (B, in_feat, 1, V) -> logits (B, num_classes). */
size_t costgcn_state_bytes(const stgcn_model_desc *m, int B);
int costgcn_state_reset(const stgcn_model_desc *m, void *state, int B, int first, int count, void *stream);
size_t costgcn_step_workspace_bytes(const stgcn_model_desc *m, int B);
int costgcn_step(const stgcn_model_desc *m, const float *x, void *state, long long t, float *logits, int B,
                 void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* STGCN_B200_H */
