"""Skeleton graph -> partitioned, normalised adjacency ``A[K, V, V]``.

Drop-in for the reference ``models/utils/graph.py`` ``Graph`` class (same
constructor arguments, same ``.A`` / ``.num_node`` / ``.hop_dis`` attributes and
``get_adjacency_raw()``), host-side float64 numpy, run once per model.

Semantics kept from the reference:
  * hop distance equals the reference's Floyd-Warshall result
    (graph.py:182-205), here computed by breadth-first search per joint; the
    diagonal is 0 only for joints whose self-loop is listed in ``edge``
    (otherwise the shortest closed walk, 2, or inf for an isolated joint);
  * 'spatial' partitions {same, closer, farther} by hop distance to the centre
    joint (graph.py:143-164); 'distance' one partition per hop; 'uniform'
    returns the reference's never-filled all-zero ``(1, V, V)`` (graph.py:134);
  * per-partition normalisation with ``D = rowsum + alpha``: symmetric
    ``D^-1/2 A D^-1/2`` (graph.py:238-243) or ``A D^-1`` (graph.py:219-224);
  * the result is transposed to ``A[k, v, w]`` with the contraction index in
    the middle (graph.py:179).
"""
from collections import deque

import numpy as np


class Graph:
    def __init__(self, num_node, edge, center, strategy='spatial',
                 normalization='symmetric', max_hop=1, dilation=1, alpha=0.001):
        self.max_hop = max_hop
        self.dilation = dilation
        self.num_node = num_node
        self.edge = edge
        self.center = center
        self.alpha = alpha

        self.hop_dis = self.get_hop_distance()
        self._A = self.get_adjacency('spatial')
        scale = self.normalize_sym if normalization == 'symmetric' else self.normalize_nonsym
        self.A = self.normalize_adjacency(self.get_adjacency(strategy), scale)

    def __str__(self):
        return str(self.A)

    # ------------------------------------------------------------------ #
    def get_hop_distance(self):
        V = self.num_node
        nbrs = [set() for _ in range(V)]
        loops = set()
        for i, j in self.edge:
            if i == j:
                loops.add(i)
            else:
                nbrs[i].add(j)
                nbrs[j].add(i)
        dist = np.full((V, V), np.inf)
        for src in range(V):
            seen = {src: 0}
            queue = deque([src])
            while queue:
                u = queue.popleft()
                for w in nbrs[u]:
                    if w not in seen:
                        seen[w] = seen[u] + 1
                        queue.append(w)
            for w, d in seen.items():
                if w != src:
                    dist[src, w] = d
            if src in loops:
                dist[src, src] = 0
            elif nbrs[src]:
                dist[src, src] = 2          # out and back along one bone
        return dist

    def get_adjacency_raw(self):
        """Un-normalised spatial partitions ``(3, V, V)`` (self, close, far)."""
        return self._A

    def get_adjacency(self, strategy):
        V = self.num_node
        hop = self.hop_dis
        hops = list(range(0, self.max_hop + 1, self.dilation))
        linked = np.isin(hop, hops).astype(np.float64)

        if strategy == 'uniform':
            return np.zeros((1, V, V))
        if strategy == 'distance':
            return np.stack([linked * (hop == h) for h in hops])
        if strategy == 'spatial':
            to_center = hop[:, self.center]
            # rel[i, j] < 0: j closer to the centre than i; > 0: farther
            with np.errstate(invalid='ignore'):
                rel = np.sign(to_center[None, :] - to_center[:, None])
            rel = np.where(np.isnan(rel), 0.0, rel)      # inf == inf counts as "same"
            parts = []
            for h in hops:
                at_h = linked * (hop == h)
                same, close, far = at_h * (rel == 0), at_h * (rel < 0), at_h * (rel > 0)
                parts += [same] if h == 0 else [close, far]
            return np.stack(parts)
        raise ValueError("Strategy Does Not Exist.")

    def normalize_adjacency(self, A, foo):
        A = np.stack([foo(a) for a in A])
        return np.ascontiguousarray(A.transpose(0, 2, 1))

    def normalize_nonsym(self, A):
        d = np.power(np.sum(A, 1) + self.alpha, -1.0)
        d[np.isinf(d)] = 0
        return A * d[None, :]

    def normalize_sym(self, A):
        d = np.power(np.sum(A, 1) + self.alpha, -0.5)
        d[np.isinf(d)] = 0
        return (d[:, None] * A) * d[None, :]
