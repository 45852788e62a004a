from .rtstgcn import Model, OfflineLayer, OnlineLayer, AggregateStgcn
