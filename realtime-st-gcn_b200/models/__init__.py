"""Model registry mirroring the reference ``models/__init__.py`` for the hot models (and CoST-GCN)."""
from .stgcn import Model as Stgcn
from .rtstgcn import Model as RtStgcn
from .costgcn import Model as CostGcn

MODELS = {
    'st-gcn': Stgcn,
    'rt-st-gcn': RtStgcn,
    'co-st-gcn': CostGcn,
}
