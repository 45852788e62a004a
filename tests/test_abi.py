"""The C-ABI library loads and exports every symbol include/stgcn_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'stgcn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'\b(?:int|size_t|long long|const char \*)\s*\*?\s*((?:stgcn|rtstgcn|costgcn)_\w+)\s*\(', text)
    return sorted(set(names))


def test_header_symbols_exported(pkg):
    names = _declared_symbols()
    assert len(names) >= 20
    lib = pkg._lib.load()
    raw = ctypes.CDLL(pkg._lib.lib_path())
    for n in names:
        assert hasattr(raw, n), "missing export: " + n
    assert set(names) == set(pkg._lib.SYMBOLS), "ctypes table out of sync with the header"
    assert lib.stgcn_abi_version() == 2


def test_struct_layout_matches_header(pkg):
    # 8 int32 + 13 pointers / 8 int32 + 6 pointers + layers + prepared + prepared_bytes (LP64)
    assert ctypes.sizeof(pkg._lib.LayerDesc) == 8 * 4 + 13 * 8
    assert ctypes.sizeof(pkg._lib.ModelDesc) == 8 * 4 + 9 * 8


def test_sizing_calls_run_without_gpu(pkg, syn):
    """workspace/state sizing is pure host arithmetic: usable on a CPU-only box."""
    lib = pkg._lib.load()
    m = pkg.RtStgcn(**syn.arch_config('rt-st-gcn'))
    m._swap_layers_for_inference()
    desc, _ = m._descriptor()
    per_stream = lib.rtstgcn_state_bytes(ctypes.byref(desc), 4096) / 4096
    # 1.69 MB of fp32 FIFO + accumulator state per stream (PKU graph)
    assert 1.68e6 < per_stream < 1.70e6
    assert lib.rtstgcn_step_workspace_bytes(ctypes.byref(desc), 1) > 0
    s = pkg.Stgcn(**syn.arch_config('st-gcn'))
    d2, _ = s._descriptor()
    assert lib.stgcn_model_workspace_bytes(ctypes.byref(d2), 1, 300) > 0


def test_no_cpu_fallback(pkg, syn):
    import pytest
    import torch
    m = pkg.Stgcn(**syn.arch_config('st-gcn', in_ch=[16], out_ch=[16], stride=[1]))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.zeros(1, 3, 8, 25))
