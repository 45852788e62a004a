"""Model registry mirroring the reference ``models/__init__.py`` for the two hot models."""
from .stgcn import Model as Stgcn
from .rtstgcn import Model as RtStgcn

MODELS = {
    'st-gcn': Stgcn,
    'rt-st-gcn': RtStgcn,
}
