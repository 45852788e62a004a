from .tgcn import ConvTemporalGraphical
from .graph import Graph
from .layernorm import LayerNorm
from .batchnorm import BatchNorm1d, BatchNorm2d
from .conv import Conv2d
