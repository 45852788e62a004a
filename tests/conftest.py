import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu on the GPU box)")


@pytest.fixture(scope='session')
def pkg():
    return importlib.import_module('realtime-st-gcn_b200')


@pytest.fixture(scope='session')
def syn(pkg):
    return pkg.synthetic


def load_golden(name):
    """-> (arrays dict of torch tensors, weights dict keyed by state_dict key)."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    arrays, weights = {}, {}
    for k in z.files:
        v = z[k]
        t = torch.from_numpy(v) if v.dtype.kind in 'fiu' else v
        if k.startswith('w:'):
            weights[k[2:]] = t
        else:
            arrays[k] = t
    return arrays, weights


def rel_err(a, b):
    """max |a-b| / max |b|  (the 'relative' of north_star's 1e-4 bound)."""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope='session')
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device('cuda:0')
