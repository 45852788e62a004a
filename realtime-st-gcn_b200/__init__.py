"""realtime-st-gcn_b200: B200-native (sm_100a) forward path of ST-GCN / RT-ST-GCN.

Drop-in for the reference's ``models/stgcn`` and ``models/rtstgcn`` st_gcn blocks
(same nn.Module constructors, forward signatures, state_dict keys and the
``models/utils`` partitioned adjacency), backed by hand-written CUDA behind the C
ABI in ``include/stgcn_b200.h``.  The directory name is not a Python identifier;
import it with ``importlib.import_module('realtime-st-gcn_b200')`` or through the
``rtstgcn_b200`` alias module at the repository root.
"""
from . import _lib            # noqa: F401
from . import synthetic       # noqa: F401
from . import skeletons       # noqa: F401
from . import tsplit          # noqa: F401
from . import processor       # noqa: F401
from . import segment         # noqa: F401
from . import checkpoint      # noqa: F401
from .models import MODELS, Stgcn, RtStgcn, CostGcn   # noqa: F401

__all__ = ['MODELS', 'Stgcn', 'RtStgcn', 'CostGcn', 'synthetic', 'skeletons', 'tsplit', 'processor', 'segment', 'checkpoint']
