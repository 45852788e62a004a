"""GPU parity: the CUDA path (through the nn.Module drop-ins -> ctypes -> C ABI) against the
reference's golden outputs and against the CPU oracle on seeded inputs.

Tolerance: north_star's 1e-4 relative (max|diff| / max|ref|) in fp32.
"""
import ctypes

import pytest
import torch

from conftest import load_golden, rel_err
from oracle import stgcn_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
SMALL = dict(in_ch=[16, 16, 32], out_ch=[16, 32, 32], stride=[1, 2, 1])


def _cuda_sd(w, dev):
    return {k: v.to(dev) for k, v in w.items()}


# ------------------------------------------------------------------ primitives
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_layernorm(pkg, cuda, tag):
    from importlib import import_module
    LN = import_module('realtime-st-gcn_b200.models.utils').LayerNorm
    a, w = load_golden('layernorm_' + tag)
    c, v = a['x'].shape[1], a['x'].shape[3]
    ln = LN([c, 1, v]).to(cuda)
    ln.load_state_dict(w)
    assert rel_err(ln(a['x'].to(cuda)), a['y']) < 1e-5


def test_batchnorms(pkg, cuda):
    from importlib import import_module
    U = import_module('realtime-st-gcn_b200.models.utils')
    a, w = load_golden('batchnorm1d')
    bn = U.BatchNorm1d(75).to(cuda)
    bn.load_state_dict(w)
    assert rel_err(bn(a['x'].to(cuda)), a['y']) < 1e-5
    a, w = load_golden('batchnorm2d')
    bn = U.BatchNorm2d(16).to(cuda)
    bn.load_state_dict(w)
    assert rel_err(bn(a['x'].to(cuda)), a['y']) < 1e-5
    with pytest.raises(RuntimeError, match='more than 1 value'):
        U.BatchNorm1d(75).to(cuda)(torch.zeros(1, 3, 1, 25, device=cuda))


def test_conv_matches_torch_cpu(pkg, cuda):
    """Gamma x 1 / 1 x 1 convolution kernel vs torch CPU conv2d (odd channel counts, strides)."""
    from importlib import import_module
    Conv2d = import_module('realtime-st-gcn_b200.models.utils').Conv2d
    g = torch.Generator().manual_seed(3)
    for (ci, co, k, s, n, t, v) in [(3, 64, 1, 1, 2, 7, 25), (16, 16, 9, 1, 2, 20, 25),
                                    (16, 32, 9, 2, 1, 21, 7), (6, 10, 3, 3, 1, 10, 5),
                                    (256, 52, 1, 1, 3, 1, 1)]:
        conv = Conv2d(ci, co, (k, 1), stride=(s, 1), padding=((k - 1) // 2, 0))
        x = torch.randn(n, ci, t, v, generator=g)
        ref = torch.nn.functional.conv2d(x, conv.weight, conv.bias, stride=(s, 1), padding=((k - 1) // 2, 0))
        got = conv.to(cuda)(x.to(cuda))
        assert got.shape == ref.shape
        assert rel_err(got, ref.detach()) < 1e-5


def test_graph_conv(pkg, cuda):
    from importlib import import_module
    T = import_module('realtime-st-gcn_b200.models.utils').ConvTemporalGraphical
    a, w = load_golden('tgcn')
    tg = T(16, 32, 25, 3).to(cuda)
    tg.load_state_dict(w)
    x = a['x'].to(cuda)
    assert rel_err(tg(x, a['A'].to(cuda)), a['y3']) < 1e-5
    assert rel_err(tg(x, a['A4'].to(cuda)), a['y4']) < 1e-5       # dense per-sample adjacency


# ------------------------------------------------------------------ ST-GCN layer
@pytest.mark.parametrize('tag', ['ln_id', 'ln_conv_s2', 'ln_nores', 'ln_64', 'ln_imu_s2', 'bn_id',
                                 'bn_conv_s2'])
def test_stgcn_layer_golden(pkg, cuda, tag):
    from importlib import import_module
    Layer = import_module('realtime-st-gcn_b200.models.stgcn').StgcnLayer
    a, w = load_golden('stgcn_layer_' + tag)
    ci, co, s, res, bn = [int(v) for v in a['meta']]
    v = a['x'].shape[3]
    layer = Layer(ci, co, (9, v), 3, v, stride=s, residual=bool(res),
                  normalization='BatchNorm' if bn else 'LayerNorm').to(cuda)
    layer.load_state_dict(w)
    layer.eval()
    y = layer(a['x'].to(cuda), a['A'].to(cuda))
    assert y.shape == a['y'].shape
    assert rel_err(y, a['y']) < TOL


def test_stgcn_layer_per_sample_adjacency(pkg, cuda):
    """AA-GCN style call: forward(x, A[N,K,V,V]) (reference models/aagcn/aagcn.py:148)."""
    from importlib import import_module
    Layer = import_module('realtime-st-gcn_b200.models.stgcn').StgcnLayer
    a, w = load_golden('stgcn_layer_ln_id')
    g = torch.Generator().manual_seed(5)
    A4 = a['A'].unsqueeze(0) + 0.05 * torch.randn(2, 3, 25, 25, generator=g)
    ref = O.stgcn_layer(a['x'], A4, w)
    layer = Layer(16, 16, (9, 25), 3, 25).to(cuda)
    layer.load_state_dict(w)
    assert rel_err(layer(a['x'].to(cuda), A4.to(cuda)), ref) < TOL


# ------------------------------------------------------------------ ST-GCN model
@pytest.mark.parametrize('tag,norm', [('ln', 'LayerNorm'), ('bn', 'BatchNorm')])
def test_stgcn_model_small(pkg, syn, cuda, tag, norm):
    a, w = load_golden('stgcn_model_small_' + tag)
    m = pkg.Stgcn(**syn.arch_config('st-gcn', normalization=norm, num_classes=12, **SMALL)).to(cuda)
    m.load_state_dict(w)
    m.eval()
    logits, feats = m(a['x'].to(cuda), return_features=True)
    assert logits.shape == (2, 12, 1)
    assert rel_err(feats, a['features']) < TOL
    assert rel_err(logits, a['logits']) < TOL


@pytest.mark.parametrize('tag,norm', [('ln', 'LayerNorm'), ('bn', 'BatchNorm')])
def test_stgcn_model_c1(pkg, syn, cuda, tag, norm):
    """BASELINE config 1: N=1, C=3, T=300, V=25, full 9-layer trunk, vs the reference's output."""
    a, _ = load_golden('stgcn_model_c1_' + tag)
    m = pkg.Stgcn(**syn.arch_config('st-gcn', normalization=norm))
    sd = syn.synth_state_dict(m.state_dict(), int(a['seeds'][0]))
    assert syn.state_digest(sd) == str(a['digest'])
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    x = syn.synth_input((1, 3, 300, 25), int(a['seeds'][1])).to(cuda)
    logits, feats = m(x, return_features=True)
    assert logits.shape == (1, 52, 1) and feats.shape == (1, 256, 75, 25)
    assert rel_err(feats[:, :, [0, 37, 74]], a['features_t0_t37_t74']) < TOL
    assert rel_err(logits, a['logits']) < TOL


def test_stgcn_model_trials_are_independent_and_chunked(pkg, syn, cuda):
    """LayerNorm mode: a batch equals its trials run one by one (the property trial-sharding and
    the internal trial chunking rely on); also exercises a ragged T and the chunk loop."""
    m = pkg.Stgcn(**syn.arch_config('st-gcn', num_classes=12, **SMALL))
    m.load_state_dict(syn.synth_state_dict(m.state_dict(), 5))
    m = m.to(cuda).eval()
    x = syn.synth_input((5, 3, 37, 25), 6).to(cuda)
    full = m(x)
    single = torch.cat([m(x[i:i + 1]) for i in range(5)])
    assert torch.equal(full, single)
    # force the C side to chunk: hand it a workspace that only fits 2 trials
    lib = pkg._lib.load()
    desc, _ = m._descriptor()
    small = lib.stgcn_model_workspace_bytes(ctypes.byref(desc), 2, 37)
    ws = torch.empty(small, dtype=torch.uint8, device=cuda)
    logits = torch.empty(5, 12, device=cuda)
    pkg._lib.check(lib.stgcn_model_forward(ctypes.byref(desc), x.data_ptr(), logits.data_ptr(), None, 5, 37,
                                           ws.data_ptr(), ws.numel(), None))
    torch.cuda.synchronize()
    assert torch.equal(logits, full.squeeze(-1))
    oracle = O.stgcn_model(x.cpu(), {k: v.cpu() for k, v in m.state_dict().items()},
                           dict(layers=3, stride=[1, 2, 1], residual=[1, 1, 1], normalization='LayerNorm'))
    assert rel_err(full, oracle) < TOL


def test_stgcn_model_host_entry(pkg, syn, cuda):
    """stgcn_model_forward_host: pinned host buffers in, host logits out (H2D + forward + D2H)."""
    m = pkg.Stgcn(**syn.arch_config('st-gcn', num_classes=12, **SMALL))
    m.load_state_dict(syn.synth_state_dict(m.state_dict(), 5))
    m = m.to(cuda).eval()
    x = syn.synth_input((3, 3, 30, 25), 8).pin_memory()
    out = torch.empty(3, 12).pin_memory()
    lib = pkg._lib.load()
    desc, _ = m._descriptor()
    ws = torch.empty(lib.stgcn_model_workspace_bytes(ctypes.byref(desc), 3, 30), dtype=torch.uint8, device=cuda)
    io = torch.empty(x.numel() + out.numel(), device=cuda)
    pkg._lib.check(lib.stgcn_model_forward_host(ctypes.byref(desc), x.data_ptr(), out.data_ptr(), 3, 30,
                                                io.data_ptr(), ws.data_ptr(), ws.numel(), None))
    assert torch.equal(out, m(x.to(cuda)).squeeze(-1).cpu())


# ------------------------------------------------------------------ RT-ST-GCN continual
def _online(pkg, syn, cuda, cfg_kw, sd):
    m = pkg.RtStgcn(**syn.arch_config('rt-st-gcn', **cfg_kw))
    m.load_state_dict(sd)
    m = m.to(cuda)
    m.prepare_benchmark({})
    return m.eval()


def test_rt_small(pkg, syn, cuda):
    a, w = load_golden('rtstgcn_small')
    m = _online(pkg, syn, cuda, dict(num_classes=12, **SMALL), w)
    out = m(a['x'].to(cuda))                      # both streams at once, 40 frames
    assert out.shape == (2, 12, 40)
    assert rel_err(out, a['logits']) < TOL
    a, w = load_golden('rtstgcn_small_nores')
    m = _online(pkg, syn, cuda, dict(num_classes=6, in_ch=[16, 16], out_ch=[16, 32], stride=[1, 1],
                                     residual=[0, 1], importance=False), w)
    assert rel_err(m(a['x'].to(cuda)), a['logits']) < TOL


@pytest.mark.parametrize('tag', ['pku', 'imu'])
def test_rt_full(pkg, syn, cuda, tag):
    """BASELINE configs 2/5 trunks vs the reference's own continual loop (48 frames: covers the
    stride-2 FIFO wrap at t = 17 and t = 34)."""
    a, _ = load_golden('rtstgcn_' + tag)
    kw = {} if tag == 'pku' else dict(graph='imu_fogit_ABCD', in_feat=6, num_classes=8)
    cfg = syn.arch_config('rt-st-gcn', **kw)
    sd = syn.synth_state_dict(pkg.RtStgcn(**cfg).state_dict(), int(a['seeds'][0]))
    assert syn.state_digest(sd) == str(a['digest'])
    m = _online(pkg, syn, cuda, kw, sd)
    x = syn.synth_input((2, cfg['in_feat'], 48, cfg['graph']['num_node']), int(a['seeds'][1])).to(cuda)
    out = m(x)
    assert rel_err(out, a['logits']) < TOL


def test_rt_streams_independent_reset_and_many_streams(pkg, syn, cuda):
    """Streams are independent; resetting one stream restarts only that stream; a 300-stream batch
    agrees with the oracle on a sampled subset."""
    _, w = load_golden('rtstgcn_small')
    m = _online(pkg, syn, cuda, dict(num_classes=12, **SMALL), w)
    B, L = 300, 30
    x = syn.synth_input((B, 3, L, 25), 77)
    out = m(x.to(cuda)).cpu()
    cfg = dict(layers=3, stride=[1, 2, 1], residual=[1, 1, 1], importance=True, kernel=9, out_ch=[16, 32, 32])
    pick = [0, 1, 137, 299]
    ref = O.rt_model_run(x[pick], w, cfg)
    assert rel_err(out[pick], ref) < TOL
    # reset stream 137 only, replay: stream 137 restarts from scratch, stream 0 continues
    m.reset_streams(137, 1)
    out2 = m(x[:, :, :5].to(cuda)).cpu()
    assert rel_err(out2[137], ref[2][:, :5]) < TOL
    cont = O.rt_model_run(torch.cat([x[0:1], x[0:1, :, :5]], dim=2), w, cfg)[:, :, L:]
    assert rel_err(out2[0:1], cont) < TOL


def test_rt_online_layer_module(pkg, syn, cuda):
    """OnlineLayer.forward (single layer API) vs the oracle recurrence, stride 2 + residual conv."""
    from importlib import import_module
    R = import_module('realtime-st-gcn_b200.models.rtstgcn')
    _, w = load_golden('rtstgcn_small')
    A = w['A']
    layer = R.OnlineLayer(in_channels=16, out_channels=32, kernel_size=9, num_joints=25, stride=2,
                          num_partitions=3, dropout=0, residual=True, importance=True, graph=A)
    sub = {k[len('st_gcn.1.'):]: v for k, v in w.items() if k.startswith('st_gcn.1.')}
    layer.load_state_dict(sub)
    layer = layer.to(cuda)
    layer.eval_()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 16, 40, 25, generator=g)
    got = torch.cat([layer(x[:, :, t:t + 1].to(cuda), None) for t in range(40)], dim=2)
    st = O.rt_state_init(dict(layers=1, stride=[2], kernel=9, out_ch=[32]), w, 3)
    ref = torch.cat([O.rt_layer_step(x[:, :, t:t + 1], A * w['st_gcn.1.edge_importance'], w, 'st_gcn.1.',
                                     st[0], 32, 2) for t in range(40)], dim=2)
    assert rel_err(got, ref) < TOL


def test_rt_batchnorm_rejected(pkg, syn, cuda):
    m = pkg.RtStgcn(**syn.arch_config('rt-st-gcn', normalization='BatchNorm', num_classes=12, **SMALL)).to(cuda)
    m.prepare_benchmark({})
    with pytest.raises(RuntimeError, match='LayerNorm'):
        m(torch.zeros(1, 3, 1, 25, device=cuda))


# ------------------------------------------------------------------ tensor-core (tcgen05) arithmetic
BF16_TOL = 2e-2      # stated bf16 tolerance (SURVEY 8d anchor; single-pass bf16 operands, fp32 accumulate/statistics)
BF16_TOP1 = 0.98     # and top-1 agreement with the fp32 reference


def _layer_case(cuda, c, stride, residual, v_graph, n, t, seed):
    from importlib import import_module
    pk = import_module('realtime-st-gcn_b200')
    Layer = import_module('realtime-st-gcn_b200.models.stgcn').StgcnLayer
    from oracle import build_adjacency
    g = pk.skeletons.skeleton(v_graph)
    v = g['num_node']
    A = torch.tensor(build_adjacency(**g), dtype=torch.float32)
    gen = torch.Generator().manual_seed(seed)
    A = A * (torch.rand(3, v, v, generator=gen) + 0.5)
    layer = Layer(c, c, (9, v), 3, v, stride=stride, residual=residual)
    sd = pk.synthetic.synth_state_dict(layer.state_dict(), seed)
    layer.load_state_dict(sd)
    x = torch.randn(n, c, t, v, generator=gen)
    ref = O.stgcn_layer(x, A, sd, stride=stride, residual=residual)
    return layer.to(cuda).eval(), x.to(cuda), A.to(cuda), ref


@pytest.mark.parametrize('c,residual,graph,n,t', [
    (64, True, 'pku-mmd', 2, 20), (64, False, 'pku-mmd', 1, 8), (128, True, 'pku-mmd', 2, 37),
    (256, True, 'pku-mmd', 1, 19), (64, True, 'imu_fogit_ABCD', 3, 11), (128, True, 'pku-mmd', 1, 3)])
def test_stgcn_layer_tensor_core(cuda, c, residual, graph, n, t):
    """tcgen05 temporal stage (bf16x3 split = fp32 parity; bf16 = stated tolerance), ragged T
    (tail tiles, T < one tile), 7- and 25-joint graphs, with and without residual."""
    layer, x, A, ref = _layer_case(cuda, c, 1, residual, graph, n, t, 100 + c + t)
    exact = layer(x, A, math='fp32')
    assert rel_err(exact, ref) < 1e-5
    y3 = layer(x, A, math='bf16x3')
    assert rel_err(y3, ref) < TOL, rel_err(y3, ref)
    y1 = layer(x, A, math='bf16')
    assert rel_err(y1, ref) < BF16_TOL, rel_err(y1, ref)
    assert rel_err(y1, ref) > 1e-5       # really ran reduced precision


@pytest.mark.parametrize('ci,co,stride,n,t', [(64, 128, 2, 2, 21), (128, 256, 2, 1, 16), (64, 128, 1, 1, 9)])
def test_stgcn_layer_tensor_core_conv_residual(pkg, cuda, ci, co, stride, n, t):
    """Channel-changing / strided layers: tensor-core graph-conv stage feeding the temporal stage."""
    from importlib import import_module
    from oracle import build_adjacency
    Layer = import_module('realtime-st-gcn_b200.models.stgcn').StgcnLayer
    A = torch.tensor(build_adjacency(**pkg.skeletons.skeleton('pku-mmd')), dtype=torch.float32)
    gen = torch.Generator().manual_seed(ci + co + t)
    A = A * (torch.rand(3, 25, 25, generator=gen) + 0.5)
    layer = Layer(ci, co, (9, 25), 3, 25, stride=stride, residual=True)
    sd = pkg.synthetic.synth_state_dict(layer.state_dict(), 7)
    layer.load_state_dict(sd)
    x = torch.randn(n, ci, t, 25, generator=gen)
    ref = O.stgcn_layer(x, A, sd, stride=stride, residual=True)
    layer = layer.to(cuda).eval()
    y3 = layer(x.to(cuda), A.to(cuda), math='bf16x3')
    assert y3.shape == ref.shape
    assert rel_err(y3, ref) < TOL, rel_err(y3, ref)
    assert rel_err(layer(x.to(cuda), A.to(cuda), math='bf16'), ref) < BF16_TOL


def test_stgcn_layer_golden_tensor_core(cuda):
    from importlib import import_module
    Layer = import_module('realtime-st-gcn_b200.models.stgcn').StgcnLayer
    a, w = load_golden('stgcn_layer_ln_64')
    layer = Layer(64, 64, (9, 25), 3, 25).to(cuda)
    layer.load_state_dict(w)
    y = layer.eval()(a['x'].to(cuda), a['A'].to(cuda), math='bf16x3')
    assert rel_err(y, a['y']) < TOL


@pytest.mark.parametrize('math,tol', [('bf16x3', TOL), ('bf16', BF16_TOL)])
def test_stgcn_model_c1_tensor_core(pkg, syn, cuda, math, tol):
    """BASELINE config 1 on the tensor-core path vs the reference's fp32 output."""
    a, _ = load_golden('stgcn_model_c1_ln')
    cfg = syn.arch_config('st-gcn')
    cfg['math'] = math
    m = pkg.Stgcn(**cfg)
    m.load_state_dict(syn.synth_state_dict(m.state_dict(), int(a['seeds'][0])))
    m = m.to(cuda).eval()
    x = syn.synth_input((1, 3, 300, 25), int(a['seeds'][1])).to(cuda)
    logits, feats = m(x, return_features=True)
    e_f = rel_err(feats[:, :, [0, 37, 74]], a['features_t0_t37_t74'])
    e_l = rel_err(logits, a['logits'])
    print("math=%s  rel_err features %.3e  logits %.3e" % (math, e_f, e_l))
    assert e_f < tol and e_l < tol
    if math == 'bf16':
        assert logits.argmax(1).item() == a['logits'].argmax(1).item()


# ------------------------------------------------------------------ RT continual step on tensor cores
def _rt_case(pkg, syn, cuda, tag, math, small=False):
    a, _ = load_golden('rtstgcn_' + tag)
    kw = {} if tag == 'pku' else dict(graph='imu_fogit_ABCD', in_feat=6, num_classes=8)
    cfg = syn.arch_config('rt-st-gcn', **kw)
    sd = syn.synth_state_dict(pkg.RtStgcn(**cfg).state_dict(), int(a['seeds'][0]))
    cfg['math'] = math
    cfg['small_batch_kernel'] = small
    m = pkg.RtStgcn(**cfg)
    m.load_state_dict(sd)
    m = m.to(cuda)
    m.prepare_benchmark({})
    x = syn.synth_input((2, cfg['in_feat'], 48, cfg['graph']['num_node']), int(a['seeds'][1])).to(cuda)
    return m.eval(), x, a['logits']


@pytest.mark.parametrize('tag', ['pku', 'imu'])
@pytest.mark.parametrize('math,tol', [('bf16x3', TOL), ('bf16', BF16_TOL)])
def test_rt_full_tensor_core(pkg, syn, cuda, tag, math, tol):
    """BASELINE configs 2/5: the continual step with the tcgen05 feature transform and the FIFO /
    accumulator update fused into its epilogue, vs the reference's own continual loop (48 frames)."""
    m, x, ref = _rt_case(pkg, syn, cuda, tag, math)
    out = m(x)
    err = rel_err(out, ref)
    print("rt %s math=%s rel_err %.3e" % (tag, math, err))
    assert err < tol, err
    if math == 'bf16':
        agree = (out.cpu().argmax(1) == ref.argmax(1)).float().mean().item()
        assert agree >= BF16_TOP1, agree     # stated bf16 tolerance: top-1 agreement with fp32


def test_rt_cuda_graph_and_many_streams_tensor_core(pkg, syn, cuda):
    """Graph-replayed steps equal eagerly launched ones bit for bit; 600 streams (several tiles per
    CTA, ragged last tile) agree with the oracle on a sampled subset; per-stream reset works."""
    cfg = syn.arch_config('rt-st-gcn', num_classes=12, in_ch=[64, 64, 128], out_ch=[64, 128, 128],
                          stride=[1, 2, 1])
    sd = syn.synth_state_dict(pkg.RtStgcn(**cfg).state_dict(), 11)
    cfg['math'] = 'bf16x3'
    B, L = 600, 24
    x = syn.synth_input((B, 3, L, 25), 12)
    outs = []
    for graph in (False, True):
        m = pkg.RtStgcn(**cfg)
        m.load_state_dict(sd)
        m = m.to(cuda)
        m.prepare_benchmark({})
        m.enable_cuda_graph(graph)
        outs.append(m(x.to(cuda)).cpu())
    assert torch.equal(outs[0], outs[1])
    ocfg = dict(layers=3, stride=[1, 2, 1], residual=[1, 1, 1], importance=True, kernel=9, out_ch=[64, 128, 128])
    pick = [0, 4, 5, 299, 599]
    ref = O.rt_model_run(x[pick], sd, ocfg)
    assert rel_err(outs[1][pick], ref) < TOL
    # per-stream reset under graph replay: stream 5 restarts, stream 4 continues
    m.reset_streams(5, 1)
    out2 = m(x[:, :, :3].to(cuda)).cpu()
    assert rel_err(out2[5], ref[2][:, :3]) < TOL
    cont = O.rt_model_run(torch.cat([x[4:5], x[4:5, :, :3]], dim=2), sd, ocfg)[:, :, L:]
    assert rel_err(out2[4:5], cont) < TOL


# ------------------------------------------------------------------ T-split (BASELINE config 4)
class _LocalExchange:
    """Two (or more) ranks emulated as threads on ONE GPU: same staging buffers and callback
    contract as tsplit.DistExchange, the swap done by device copies behind a thread barrier."""

    def __init__(self, rank, world, capacity, device, shared):
        import threading  # noqa: F401
        self.rank, self.world, self.capacity, self.shared = rank, world, capacity, shared
        self.has_left, self.has_right = rank > 0, rank < world - 1
        mk = lambda: torch.zeros(capacity, dtype=torch.uint8, device=device)   # noqa: E731
        self.send_left, self.send_right, self.recv_left, self.recv_right = mk(), mk(), mk(), mk()
        self.calls = 0
        shared['ex'][rank] = self

    def __call__(self, layer, nbytes):
        torch.cuda.current_stream().synchronize()
        self.shared['barrier'].wait()
        peers = self.shared['ex']
        if self.has_left:
            self.recv_left[:nbytes].copy_(peers[self.rank - 1].send_right[:nbytes])
        if self.has_right:
            self.recv_right[:nbytes].copy_(peers[self.rank + 1].send_left[:nbytes])
        torch.cuda.current_stream().synchronize()
        self.shared['barrier'].wait()
        self.calls += 1
        return 0

    def all_reduce_sum(self, t):
        torch.cuda.current_stream().synchronize()
        self.shared['sums'][self.rank] = t.clone()
        self.shared['barrier'].wait()
        total = sum(self.shared['sums'][r] for r in range(self.world))
        self.shared['barrier'].wait()
        t.copy_(total)
        return t

    def descriptor(self):
        import importlib
        ts = importlib.import_module('realtime-st-gcn_b200').tsplit

        def cb(_ctx, layer, nbytes):
            try:
                return int(self(layer, nbytes))
            except Exception as e:
                self.error = e
                return 1
        fn = ts.EXCHANGE_FN(cb)
        d = ts.HaloDesc()
        d.has_left, d.has_right = int(self.has_left), int(self.has_right)
        d.send_left, d.send_right = self.send_left.data_ptr(), self.send_right.data_ptr()
        d.recv_left, d.recv_right = self.recv_left.data_ptr(), self.recv_right.data_ptr()
        d.capacity, d.exchange, d.ctx = self.capacity, fn, None
        return d, fn


@pytest.mark.parametrize('world,total_frames,math,tol', [(2, 64, 'bf16x3', TOL), (3, 93, 'bf16x3', TOL),
                                                         (2, 40, 'bf16', BF16_TOL), (2, 4096, 'bf16x3', TOL),
                                                         (4, 4101, 'bf16x3', TOL)])
def test_tsplit_emulated_ranks(pkg, syn, cuda, world, total_frames, math, tol):
    """T-split forward through the C ABI: per-layer halo frames packed, swapped and unpacked in the
    tensor-core operand layout (stride-1 and stride-2 layers, channel-changing layers), pooled sums
    all-reduced; vs the full-sequence oracle."""
    import threading
    kw = dict(num_classes=12, in_ch=[64, 64, 128, 128], out_ch=[64, 128, 128, 256], stride=[1, 2, 1, 2])
    cfg = syn.arch_config('st-gcn', **kw)
    sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 31)
    cfg['math'] = math
    x = syn.synth_input((2, 3, total_frames, 25), 32)
    bounds = pkg.tsplit.chunk_bounds(total_frames, world, 4)
    shared = dict(ex=[None] * world, sums=[None] * world, barrier=threading.Barrier(world))
    out, errs = [None] * world, []

    def run(rank):
        try:
            torch.cuda.set_device(cuda)
            m = pkg.Stgcn(**cfg)
            m.load_state_dict(sd)
            m = m.to(cuda).eval()
            need = pkg._lib.load().stgcn_model_halo_bytes(ctypes.byref(m._descriptor()[0]), 2)
            ex = _LocalExchange(rank, world, need, cuda, shared)
            shared['barrier'].wait()
            a, b = bounds[rank]
            with torch.cuda.stream(torch.cuda.Stream(device=cuda)):
                out[rank] = m.forward_tsplit(x[:, :, a:b].to(cuda), total_frames, ex).cpu()
            assert ex.calls == 4
        except Exception as e:                                  # pragma: no cover
            errs.append(e)
            shared['barrier'].abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errs, errs
    ref = O.stgcn_model(x, sd, dict(layers=4, stride=kw['stride'], residual=[1] * 4, normalization='LayerNorm'))
    for r in range(world):
        assert rel_err(out[r], ref) < tol, (r, rel_err(out[r], ref))
    # and a single rank without neighbours equals the ordinary forward
    m = pkg.Stgcn(**cfg)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    shared1 = dict(ex=[None], sums=[None], barrier=threading.Barrier(1))
    ex = _LocalExchange(0, 1, pkg._lib.load().stgcn_model_halo_bytes(ctypes.byref(m._descriptor()[0]), 2), cuda, shared1)
    single = m.forward_tsplit(x.to(cuda), total_frames, ex)
    assert rel_err(single, m(x.to(cuda))) < 1e-6


# ------------------------------------------------------------------ few-streams cluster kernel
@pytest.mark.parametrize('tag', ['pku', 'imu'])
def test_rt_small_batch_cluster_kernel(pkg, syn, cuda, tag):
    """Latency path (<= 14 streams): the whole continual step in one thread-block-cluster kernel
    (fp32 FMA, distributed-shared-memory exchange per layer) vs the reference's own loop."""
    m, x, ref = _rt_case(pkg, syn, cuda, tag, 'bf16x3', small=True)
    lib = pkg._lib.load()
    m._descriptor()                                          # prepared operands are built once, up front
    n0 = lib.stgcn_launch_count()
    out = m(x)
    assert lib.stgcn_launch_count() - n0 == 48               # ONE kernel per frame
    assert rel_err(out, ref) < 1e-5, rel_err(out, ref)
    # CUDA-graph replay, more streams than one, per-stream reset
    m.enable_cuda_graph(True)
    m.reset_streams()
    assert torch.equal(m(x), out)
    m.reset_streams(1, 1)
    out2 = m(x[:, :, :7])
    assert rel_err(out2[1], ref[1][:, :7]) < 1e-5
    assert rel_err(out2[0], ref[0][:, :7]) > 1e-3            # stream 0 was not reset: it continues


def test_rt_small_batch_matches_batched_path(pkg, syn, cuda):
    """13 streams through the cluster kernel == the same streams through the tensor-core path."""
    cfg = syn.arch_config('rt-st-gcn', num_classes=12, in_ch=[64, 64, 128], out_ch=[64, 128, 128], stride=[1, 2, 1],
                          residual=[1, 1, 0])
    sd = syn.synth_state_dict(pkg.RtStgcn(**cfg).state_dict(), 41)
    x = syn.synth_input((13, 3, 20, 25), 42)
    outs = []
    for small in (True, False):
        cfg['small_batch_kernel'] = small
        m = pkg.RtStgcn(**cfg)
        m.load_state_dict(sd)
        m = m.to(cuda)
        m.prepare_benchmark({})
        outs.append(m(x.to(cuda)).cpu())
    ref = O.rt_model_run(x, sd, dict(layers=3, stride=[1, 2, 1], residual=[1, 1, 0], importance=True, kernel=9,
                                     out_ch=[64, 128, 128]))
    assert rel_err(outs[0], ref) < 1e-5
    assert rel_err(outs[1], ref) < TOL


# ------------------------------------------------------------------ RT-ST-GCN training-time definition
def test_rt_offline_model_and_layers(pkg, syn, cuda):
    """rtstgcn.Model with OfflineLayers on a whole sequence (the definition trained weights refer
    to; reference rtstgcn.py:137-157, 343-389) vs the reference's output."""
    a, w = load_golden('rtstgcn_offline')
    cfg = syn.arch_config('rt-st-gcn', num_classes=12, in_ch=[16, 16, 32], out_ch=[16, 32, 32], stride=[1, 2, 1],
                          residual=[1, 1, 0])
    m = pkg.RtStgcn(**cfg)
    m.load_state_dict(w)
    m = m.to(cuda).eval()
    assert not m.is_online
    x = a['x'].to(cuda)
    logits = m(x)
    assert logits.shape == (2, 12, 30)
    assert rel_err(logits, a['logits']) < TOL
    y1 = m.st_gcn[0](a['h'].to(cuda), m.A)
    assert rel_err(y1, a['y1']) < 1e-5
    y2 = m.st_gcn[1](a['y1'].to(cuda), m.A)
    assert rel_err(y2, a['y2']) < 1e-5


# ------------------------------------------------------------------ host harness (processor.py counterpart)
def test_processor_benchmark_harness(pkg, syn, cuda, tmp_path):
    """benchmark() = prepare_benchmark + frame loop + seconds-per-frame in the reference's
    latency.csv schema (processor.py:395-427, 881-902, 977); predictions equal the plain loop and
    the reference's continual output; top-1/top-5 follow utils/statistics.py."""
    a, _ = load_golden('rtstgcn_pku')
    cfg = syn.arch_config('rt-st-gcn')
    sd = syn.synth_state_dict(pkg.RtStgcn(**cfg).state_dict(), int(a['seeds'][0]))
    m = pkg.RtStgcn(**cfg)
    m.load_state_dict(sd)
    m = m.to(cuda)
    x = syn.synth_input((2, 3, 48, 25), int(a['seeds'][1])).to(cuda)
    labels = a['logits'].argmax(1)                       # the reference's own top-1 as ground truth
    res = pkg.processor.benchmark(m, x, labels, save_dir=str(tmp_path))
    assert rel_err(res['predictions'], a['logits']) < TOL
    assert res['tot'] == 96 and res['top1_cor'] >= 95 and res['top5_cor'] >= res['top1_cor']
    assert 0 < res['latency'] < 0.01
    rows = open(tmp_path / 'latency.csv').read().strip().split('\n')
    assert rows[0] == ',latency_fp32,latency_int8' and rows[1].startswith('0,')
    assert abs(float(rows[1].split(',')[1]) - res['latency']) < 1e-9


# ------------------------------------------------------------------ other graphs / temporal kernels
@pytest.mark.parametrize('graph,c,kernel,stride,t', [
    ('openpose', 64, 9, 1, 23), ('coco', 128, 5, 2, 30), ('lara', 64, 3, 1, 11), ('hugadb', 128, 9, 1, 40),
    ('ntu-edge', 64, 13, 1, 17), ('tp-vicon', 256, 9, 2, 21)])
def test_stgcn_layer_tensor_core_other_graphs(pkg, cuda, graph, c, kernel, stride, t):
    """Every skeleton under data/skeletons (6..24 joints: different frames-per-tile packing) and
    temporal kernels other than 9, on the tensor-core path, vs the oracle."""
    from importlib import import_module
    from oracle import build_adjacency
    Layer = import_module('realtime-st-gcn_b200.models.stgcn').StgcnLayer
    g = pkg.skeletons.skeleton(graph)
    v = g['num_node']
    gen = torch.Generator().manual_seed(c + t + kernel)
    A = torch.tensor(build_adjacency(**g), dtype=torch.float32) * (torch.rand(3, v, v, generator=gen) + 0.5)
    layer = Layer(c, c, (kernel, v), 3, v, stride=stride, residual=True)
    sd = pkg.synthetic.synth_state_dict(layer.state_dict(), 17)
    layer.load_state_dict(sd)
    x = torch.randn(2, c, t, v, generator=gen)
    ref = O.stgcn_layer(x, A, sd, stride=stride, residual=True)
    layer = layer.to(cuda).eval()
    y3 = layer(x.to(cuda), A.to(cuda), math='bf16x3')
    assert y3.shape == ref.shape
    assert rel_err(y3, ref) < TOL, rel_err(y3, ref)


def test_stgcn_cuda_graph_replay(pkg, syn, cuda):
    """Stgcn.enable_cuda_graph: config-1-sized forwards are launch-bound, so the module can replay the captured
    kernel sequence; replays equal the eager forward bit for bit, for new inputs and after a shape change."""
    cfg = syn.arch_config('st-gcn', num_classes=12, in_ch=[64, 64, 128], out_ch=[64, 128, 128], stride=[1, 2, 1])
    m = pkg.Stgcn(**cfg)
    m.load_state_dict(syn.synth_state_dict(m.state_dict(), 71))
    m = m.to(cuda).eval()
    xs = [syn.synth_input((1, 3, 60, 25), 72), syn.synth_input((1, 3, 60, 25), 73), syn.synth_input((2, 3, 37, 25), 74)]
    eager = [m(x.to(cuda)).clone() for x in xs]
    m.enable_cuda_graph(True)
    for x, ref in zip(xs + xs[:1], eager + eager[:1]):
        out = m(x.to(cuda))
        assert torch.equal(out, ref)
    m.enable_cuda_graph(False)
    assert torch.equal(m(xs[1].to(cuda)), eager[1])


# ------------------------------------------------------------------ sliding-window inference
def test_stgcn_sliding_windows(pkg, syn, cuda):
    """WindowSegment semantics (utils/segment_generator.py:109-154): frame i classified from the W
    frames ending at i, zeros before the start.  The in-place strided read equals the materialised
    unfold() batch bit for bit, and matches the oracle on sampled windows."""
    cfg = syn.arch_config('st-gcn', num_classes=12, in_ch=[64, 64, 128], out_ch=[64, 128, 128], stride=[1, 2, 1])
    sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 81)
    m = pkg.Stgcn(**cfg)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    L, W = 45, 16
    cap = syn.synth_input((1, 3, L, 25), 82)
    out = m.forward_windows(cap.to(cuda), W)
    assert out.shape == (1, 12, L)
    padded = torch.nn.functional.pad(cap, (0, 0, W - 1, 0))
    windows = padded.unfold(2, W, 1).permute(0, 2, 1, 4, 3).contiguous().view(L, 3, W, 25)   # the reference's batch
    ref_gpu = m(windows.to(cuda))                                     # (L, classes, 1)
    assert torch.equal(out, ref_gpu.permute(2, 1, 0))
    pick = [0, 1, 15, 16, 44]
    ref = O.stgcn_model(windows[pick], sd, dict(layers=3, stride=[1, 2, 1], residual=[1, 1, 1],
                                                normalization='LayerNorm'))
    assert rel_err(out[0, :, pick].t().unsqueeze(-1), ref) < TOL


def test_stgcn_sliding_windows_shared_first_layer_subprocess(cuda, tmp_path):
    """Sliding windows share their frames: everything before the first temporal convolution is evaluated once per
    frame and the temporal kernel reads its windows out of that one sequence (tensor map with a one-frame trial
    pitch).  The result must equal the per-window evaluation (STGCN_WINDOWS_SHARE=0) bit for bit -- with and
    without a first-layer residual, both arithmetic modes, several window chunks per call -- and both must match
    the oracle.  The switches are read once per process, hence child processes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    got = {}
    for share in ('1', '0'):
        path = str(tmp_path / ('win%s.pt' % share))
        env = dict(os.environ, STGCN_WINDOWS_SHARE=share, STGCN_CHUNK_ROWS='200000')
        out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'check_windows_paths.py'), path], env=env,
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = [l for l in out.stdout.splitlines() if ' err ' in l]
        assert len(lines) == 3, out.stdout[-2000:]
        for l in lines:
            assert float(l.split(' err ')[1]) < (TOL if l.startswith('bf16x3') else BF16_TOL), l
        got[share] = torch.load(path)
    assert set(got['1']) == set(got['0']) and len(got['1']) == 3
    for k in got['1']:
        assert torch.equal(got['1'][k], got['0'][k]), k


# ------------------------------------------------------------------ non-default graph-conv paths
@pytest.mark.parametrize('switches', [{}, {'STGCN_GCNW': '0'}, {'STGCN_GCNW_FUSE': '1'}],
                         ids=['default', 'frame-tile-kernel', 'one-kernel-stage'])
def test_stgcn_model_graphconv_paths_subprocess(cuda, switches):
    """Model forwards default to the per-joint-weight GEMM + streaming LayerNorm (kernels_gcnw.cuh).  The other
    graph-conv forms stay covered at model level: the frame-tile k_gcn_tc2 kernel (STGCN_GCNW=0; also what
    layer-level calls and dense adjacencies use) and the opt-in one-kernel form of the stage
    (STGCN_GCNW_FUSE=1: z through an L2-resident ring to LN warps of the same persistent kernel) -- models of
    growing depth incl. strided / channel-changing layers, both arithmetic modes, against the oracle.  The
    switches are read once per process, so tools/check_graphconv_paths.py runs in a child process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **switches)
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'check_graphconv_paths.py')], env=env, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith(('bf16x3', 'bf16 '))]
    assert len(lines) == 20, out.stdout[-2000:]
    for l in lines:
        feats = float(l.split('feats ')[1].split()[0])
        logits = float(l.split('logits ')[1].split()[0])
        tol = TOL if l.startswith('bf16x3') else BF16_TOL
        assert feats < tol and logits < tol, l


def test_rt_fused_step_subprocess(cuda):
    """The continual step defaults to the split form (tensor-core GEMM + streaming state kernel); the fused
    form (state update inside the GEMM kernel's epilogue, STGCN_RT_SPLIT=0) stays covered: the RT
    tensor-core tests again in a child process with the switch off."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, STGCN_RT_SPLIT='0')
    out = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_parity.py'), '-m', 'gpu',
                          '-q', '-x', '-k', 'rt_full_tensor_core or rt_cuda_graph or rt_online_layer_module or '
                          'rt_small_batch_matches_batched_path'],
                         env=env, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    assert ' passed' in out.stdout and 'failed' not in out.stdout, out.stdout[-1000:]


# ------------------------------------------------------------------ all-taps layer path (large Gamma, BatchNorm)
@pytest.mark.parametrize('math,tol', [('bf16x3', TOL), ('bf16', BF16_TOL)])
def test_stgcn_model_large_temporal_kernel(pkg, syn, cuda, math, tol):
    """The reference's LayerNorm config uses a 69-tap temporal kernel (config/pku-mmd/ln/stgcn_local.json:26):
    beyond the 15 taps of the frame-tile kernel, the temporal convolution runs as the per-joint-tile tcgen05 GEMM
    over the taps (identity, strided / channel-changing and plain layers)."""
    kw = dict(num_classes=12, kernel=69, in_ch=[64, 64, 128], out_ch=[64, 128, 128], stride=[1, 2, 1])
    cfg = syn.arch_config('st-gcn', **kw)
    sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 101)
    cfg['math'] = math
    m = pkg.Stgcn(**cfg)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    x = syn.synth_input((2, 3, 150, 25), 102)
    logits, feats = m(x.to(cuda), return_features=True)
    rl, rf = O.stgcn_model(x, sd, dict(layers=3, stride=kw['stride'], residual=[1] * 3, normalization='LayerNorm'),
                           return_features=True)
    e_l, e_f = rel_err(logits, rl), rel_err(feats, rf)
    print("kernel 69 math=%s rel_err logits %.3e features %.3e" % (math, e_l, e_f))
    assert e_l < tol and e_f < tol


@pytest.mark.parametrize('residual', [[1, 1, 1], [1, 0, 1]])
def test_stgcn_model_batchnorm_tensor_core(pkg, syn, cuda, residual):
    """BatchNorm mode (config/pku-mmd/as_is/stgcn_local.json) with the GEMMs on tcgen05: raw accumulators, batch
    statistics over the whole call, then normalise (two-phase), vs the oracle (batch-statistics BatchNorm)."""
    kw = dict(num_classes=12, normalization='BatchNorm', in_ch=[64, 64, 128], out_ch=[64, 128, 128], stride=[1, 2, 1],
              residual=residual)
    cfg = syn.arch_config('st-gcn', **kw)
    sd = syn.synth_state_dict(pkg.Stgcn(**cfg).state_dict(), 103)
    cfg['math'] = 'bf16x3'
    m = pkg.Stgcn(**cfg)
    m.load_state_dict(sd)
    m = m.to(cuda).eval()
    x = syn.synth_input((3, 3, 44, 25), 104)
    logits, feats = m(x.to(cuda), return_features=True)
    rl, rf = O.stgcn_model(x, sd, dict(layers=3, stride=kw['stride'], residual=residual, normalization='BatchNorm'),
                           return_features=True)
    e_l, e_f = rel_err(logits, rl), rel_err(feats, rf)
    print("BatchNorm tensor-core path rel_err logits %.3e features %.3e" % (e_l, e_f))
    assert e_l < TOL and e_f < TOL


def test_rt_step_top5_matches_topk(pkg, syn, cuda):
    """Top-5 ranked in the logits kernel (batched path) / right after the cluster kernel (latency path) equals
    torch.topk of the logits -- the reference's Statistics (utils/statistics.py:4-16)."""
    for b in (3, 40):
        m, x, _ = _rt_case(pkg, syn, cuda, 'pku', 'bf16x3')
        frames = torch.randn(6, b, 3, 1, 25, device=cuda)
        for t in range(6):
            logits, top5 = m.step_top5(frames[t])
        ref = torch.topk(logits.squeeze(-1), 5, dim=1).indices
        assert torch.equal(top5.long(), ref)
