from .stgcn import Model, StgcnLayer
