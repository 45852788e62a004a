from .costgcn import Model, StgcnLayer
