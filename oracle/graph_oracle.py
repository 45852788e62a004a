"""Oracle restatement of the skeleton-graph adjacency construction.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows reference ``models/utils/graph.py``:
  * hop distance by Floyd-Warshall relaxation over the edge list   (:182-205)
  * partitioning 'uniform' | 'distance' | 'spatial'                 (:129-170)
  * per-partition degree normalisation with alpha                   (:209-243)
  * final transpose(0, 2, 1)                                        (:173-179)

Everything is float64 numpy, like the reference.
"""
import numpy as np


def hop_distance(num_node, edge):
    """All-pairs hop count, reference graph.py:182-205 (Floyd-Warshall).

    Note the diagonal is only zero when the self-loop is listed in ``edge``;
    otherwise the relaxation leaves the length of the shortest closed walk.
    """
    d = np.full((num_node, num_node), np.inf)
    for i, j in edge:
        if i == j:
            d[i, i] = 0
        else:
            d[i, j] = 1
            d[j, i] = 1
    for k in range(num_node):
        for i in range(num_node):
            for j in range(num_node):
                via = d[i, k] + d[k, j]
                if via < d[i, j]:
                    d[i, j] = via
    return d


def partition(hop, center, strategy, max_hop=1, dilation=1):
    """Un-normalised partitioned adjacency, reference graph.py:129-170."""
    V = hop.shape[0]
    hops = list(range(0, max_hop + 1, dilation))
    adj = np.zeros((V, V))
    for h in hops:
        adj[hop == h] = 1
    if strategy == 'uniform':
        # reference allocates zeros and never fills them (graph.py:134-135)
        return np.zeros((1, V, V))
    if strategy == 'distance':
        out = np.zeros((len(hops), V, V))
        for n, h in enumerate(hops):
            out[n][hop == h] = adj[hop == h]
        return out
    if strategy == 'spatial':
        parts = []
        for h in hops:
            root = np.zeros((V, V))
            close = np.zeros((V, V))
            far = np.zeros((V, V))
            for i in range(V):
                for j in range(V):
                    if hop[i, j] != h:
                        continue
                    dj, di = hop[j, center], hop[i, center]
                    if dj == di:
                        root[i, j] = adj[i, j]
                    elif dj < di:
                        close[i, j] = adj[i, j]
                    else:
                        far[i, j] = adj[i, j]
            if h == 0:
                parts.append(root)
            else:
                parts.extend([close, far])
        return np.stack(parts)
    raise ValueError("Strategy Does Not Exist.")


def normalize_sym(a, alpha):
    """D^-1/2 A D^-1/2 with D = rowsum + alpha, reference graph.py:227-243."""
    dl = np.power(np.sum(a, 1) + alpha, -0.5)
    dl[np.isinf(dl)] = 0
    dn = np.eye(a.shape[0]) * dl
    return np.dot(np.dot(dn, a), dn)


def normalize_nonsym(a, alpha):
    """A D^-1 with D = rowsum + alpha, reference graph.py:209-224."""
    dl = np.power(np.sum(a, 1) + alpha, -1)
    dl[np.isinf(dl)] = 0
    dn = np.eye(a.shape[0]) * dl
    return np.dot(a, dn)


def build_adjacency(num_node, edge, center, strategy='spatial',
                    normalization='symmetric', max_hop=1, dilation=1,
                    alpha=0.001):
    """Reference ``Graph(...).A`` (graph.py:78-89): float64 ``(K, V, V)``."""
    hop = hop_distance(num_node, edge)
    a = partition(hop, center, strategy, max_hop, dilation)
    fn = normalize_sym if normalization == 'symmetric' else normalize_nonsym
    for k in range(a.shape[0]):
        a[k] = fn(a[k], alpha)
    return np.ascontiguousarray(a.transpose(0, 2, 1))
