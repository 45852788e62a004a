"""ctypes binding of ``csrc/libstgcn_b200.so`` (C ABI in ``include/stgcn_b200.h``).

There is no CPU or eager-PyTorch fallback: if the shared library is missing and
cannot be built, or a tensor is not a CUDA fp32 tensor on an sm_100 device, the
call raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p, POINTER

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# STGCN_LIB: alternative build of the same library (e.g. a -DSTGCN_DEBUG_BUILD measurement build)
_LIB_PATH = os.environ.get('STGCN_LIB') or os.path.join(_HERE, 'csrc', 'libstgcn_b200.so')

ABI_VERSION = 2
NORM_LAYERNORM, NORM_BATCHNORM = 0, 1
RES_NONE, RES_IDENTITY, RES_CONV = 0, 1, 2
MATH_FP32, MATH_BF16X3, MATH_BF16 = 0, 1, 2
KERNEL_CLASSES = ['layout', 'gemm_1x1', 'gemm_tcn', 'frame', 'batchnorm', 'embed', 'pool_fc', 'misc']
MATH_NAMES = {'fp32': MATH_FP32, 'bf16x3': MATH_BF16X3, 'bf16': MATH_BF16}


class LayerDesc(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in
                ('c_in', 'c_out', 'kernel', 'stride', 'residual', 'norm', 'rt', 'a_per_sample')] + \
               [(n, c_void_p) for n in
                ('gcn_w', 'gcn_b', 'a_eff', 'n1_w', 'n1_b', 'tcn_w', 'tcn_b', 'n2_w', 'n2_b',
                 'res_w', 'res_b', 'nr_w', 'nr_b')]


class ModelDesc(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in
                ('in_feat', 'num_joints', 'partitions', 'num_classes', 'num_layers', 'norm', 'math',
                 'reserved')] + \
               [(n, c_void_p) for n in
                ('norm_in_w', 'norm_in_b', 'fcn_in_w', 'fcn_in_b', 'fcn_out_w', 'fcn_out_b')] + \
               [('layers', POINTER(LayerDesc)), ('prepared', c_void_p), ('prepared_bytes', c_size_t)]


# name -> (restype, argtypes); every symbol include/stgcn_b200.h declares
_P_LAYER, _P_MODEL = POINTER(LayerDesc), POINTER(ModelDesc)
SYMBOLS = {
    'stgcn_abi_version': (c_int, []),
    'stgcn_last_error': (c_char_p, []),
    'stgcn_device_check': (c_int, [c_int]),
    'stgcn_launch_count': (ctypes.c_longlong, []),
    'stgcn_profile_begin': (c_int, []),
    'stgcn_profile_end': (c_int, [c_void_p, c_void_p, c_int]),
    'stgcn_layernorm_forward': (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_void_p]),
    'stgcn_batchnorm_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'stgcn_batchnorm_forward': (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_int, c_void_p, c_size_t,
                                                                    c_void_p]),
    'stgcn_conv_workspace_bytes': (c_size_t, [c_int] * 7),
    'stgcn_conv_forward': (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_void_p, c_size_t, c_void_p]),
    'stgcn_graphconv_workspace_bytes': (c_size_t, [c_int] * 6),
    'stgcn_graphconv_forward': (c_int, [c_void_p] * 4 + [c_int, c_void_p] + [c_int] * 6 +
                                [c_void_p, c_size_t, c_void_p]),
    'stgcn_model_prepare_bytes': (c_size_t, [_P_MODEL]),
    'stgcn_model_prepare': (c_int, [_P_MODEL, c_void_p, c_size_t, c_void_p]),
    'stgcn_layer_workspace_bytes': (c_size_t, [_P_LAYER, c_int, c_int, c_int, c_int]),
    'stgcn_layer_forward': (c_int, [_P_LAYER, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                    c_void_p, c_size_t, c_void_p]),
    'stgcn_model_workspace_bytes': (c_size_t, [_P_MODEL, c_int, c_int]),
    'stgcn_model_forward': (c_int, [_P_MODEL, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                    c_size_t, c_void_p]),
    'stgcn_model_forward_windows': (c_int, [_P_MODEL, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                            c_size_t, c_void_p]),
    'stgcn_model_halo_bytes': (c_size_t, [_P_MODEL, c_int]),
    'stgcn_model_tsplit_workspace_bytes': (c_size_t, [_P_MODEL, c_int, c_int]),
    'stgcn_model_forward_tsplit': (c_int, [_P_MODEL, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                           c_size_t, c_void_p]),
    'rtstgcn_state_bytes': (c_size_t, [_P_MODEL, c_int]),
    'rtstgcn_state_reset': (c_int, [_P_MODEL, c_void_p, c_int, c_int, c_int, c_void_p]),
    'rtstgcn_step_workspace_bytes': (c_size_t, [_P_MODEL, c_int]),
    'rtstgcn_step': (c_int, [_P_MODEL, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    'rtstgcn_step_top5': (c_int, [_P_MODEL, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                          c_void_p]),
    'rtstgcn_layer_state_bytes': (c_size_t, [_P_LAYER, c_int, c_int]),
    'rtstgcn_layer_workspace_bytes': (c_size_t, [_P_LAYER, c_int, c_int, c_int]),
    'rtstgcn_layer_step': (c_int, [_P_LAYER, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_void_p, c_size_t, c_void_p]),
    'rtstgcn_offline_layer_workspace_bytes': (c_size_t, [_P_LAYER, c_int, c_int, c_int, c_int]),
    'rtstgcn_offline_layer_forward': (c_int, [_P_LAYER, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                              c_void_p, c_size_t, c_void_p]),
    'stgcn_mean_joints_forward': (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p]),
    'stgcn_model_forward_host': (c_int, [_P_MODEL, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                         c_size_t, c_void_p]),
    'costgcn_state_bytes': (c_size_t, [_P_MODEL, c_int]),
    'costgcn_state_reset': (c_int, [_P_MODEL, c_void_p, c_int, c_int, c_int, c_void_p]),
    'costgcn_step_workspace_bytes': (c_size_t, [_P_MODEL, c_int]),
    'costgcn_step': (c_int, [_P_MODEL, c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_int, c_void_p, c_size_t,
                     c_void_p]),
    'rtstgcn_step_host': (c_int, [_P_MODEL, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
}

_lib = None


def lib_path():
    return _LIB_PATH


def load():
    """Load (building in-tree first if needed) the shared library; raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        from .csrc import build as _build
        _build.build()
    try:
        lib = ctypes.CDLL(_LIB_PATH)
    except OSError as e:
        raise RuntimeError("cannot load %s: %s (no CPU fallback exists)" % (_LIB_PATH, e))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)           # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.stgcn_abi_version() != ABI_VERSION:
        raise RuntimeError("libstgcn_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("stgcn_b200: " + load().stgcn_last_error().decode(errors='replace'))


_checked_devices = set()


def require_cuda(*tensors):
    """Every tensor must be a contiguous fp32 CUDA tensor on one sm_100 device."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("stgcn_b200 runs on a B200 GPU only; got a %s tensor (no CPU fallback)"
                               % t.device)
        if t.dtype != torch.float32:
            raise RuntimeError("stgcn_b200 expects float32 tensors, got %s" % t.dtype)
        if not t.is_contiguous():
            raise RuntimeError("stgcn_b200 expects contiguous tensors")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("tensors on different devices: %s vs %s" % (dev, t.device))
    if dev is not None and dev.index not in _checked_devices:
        check(load().stgcn_device_check(dev.index if dev.index is not None else torch.cuda.current_device()))
        _checked_devices.add(dev.index)
    return dev


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


class Workspace:
    """Grow-only device scratch buffer owned by the Python side."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        nbytes = max(int(nbytes), 256)
        if self.buf is None or self.buf.device != device or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.buf


def prepare_model(m, device):
    """Build the prepared operands (bf16 weight planes, adjacency CSR, ...) for ModelDesc ``m``
    once; returns the device buffer that must stay alive as long as ``m`` is used."""
    lib = load()
    nbytes = lib.stgcn_model_prepare_bytes(ctypes.byref(m))
    if nbytes == 0 or m.math == MATH_FP32:
        return None
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    check(lib.stgcn_model_prepare(ctypes.byref(m), ptr(buf), nbytes, stream_ptr(device)))
    m.prepared = buf.data_ptr()
    m.prepared_bytes = nbytes
    return buf
