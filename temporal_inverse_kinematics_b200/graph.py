"""Skeleton graph -> normalised adjacency stack A (K, V, V).

Drop-in for ``mmskeleton.ops.st_gcn.Graph`` (reference mmskeleton/ops/st_gcn/graph.py:4-133): same
constructor arguments, same attributes (``A``, ``num_node``, ``edge``, ``center``, ``hop_dis``,
``max_hop``, ``dilation``) and bit-identical float64 ``A``.  Host-side numpy, built once per model.
"""
import numpy as np

_LAYOUTS = {
    # name: (num_node, 1-based?, neighbour links, centre)
    "openpose": (18, False, [(4, 3), (3, 2), (7, 6), (6, 5), (13, 12), (12, 11), (10, 9), (9, 8), (11, 5), (8, 2),
                             (5, 1), (2, 1), (0, 1), (15, 0), (14, 0), (17, 15), (16, 14)], 1),
    "ntu-rgb+d": (25, True, [(1, 2), (2, 21), (3, 21), (4, 3), (5, 21), (6, 5), (7, 6), (8, 7), (9, 21), (10, 9),
                             (11, 10), (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18),
                             (20, 19), (22, 23), (23, 8), (24, 25), (25, 12)], 20),
    "ntu_edge": (24, True, [(1, 2), (3, 2), (4, 3), (5, 2), (6, 5), (7, 6), (8, 7), (9, 2), (10, 9), (11, 10),
                            (12, 11), (13, 1), (14, 13), (15, 14), (16, 15), (17, 1), (18, 17), (19, 18), (20, 19),
                            (21, 22), (22, 8), (23, 24), (24, 12)], 2),
    "coco": (17, True, [(16, 14), (14, 12), (17, 15), (15, 13), (12, 13), (6, 12), (7, 13), (6, 7), (8, 6), (9, 7),
                        (10, 8), (11, 9), (2, 3), (2, 1), (3, 1), (4, 2), (5, 3), (4, 6), (5, 7)], 0),
}


def _hops(num_node, edge, max_hop):
    """All-pairs hop count up to max_hop (inf beyond), by frontier expansion of the boolean adjacency."""
    adj = np.zeros((num_node, num_node), dtype=bool)
    for i, j in edge:
        adj[i, j] = adj[j, i] = True
    hop = np.full((num_node, num_node), np.inf)
    reach = np.eye(num_node, dtype=bool)
    hop[reach] = 0
    frontier = reach
    for d in range(1, max_hop + 1):
        frontier = (frontier.astype(np.int64) @ adj.astype(np.int64)) > 0
        new = frontier & np.isinf(hop)
        hop[new] = d
    return hop


class Graph:
    def __init__(self, layout="openpose", strategy="uniform", max_hop=1, dilation=1):
        if layout not in _LAYOUTS:
            raise ValueError("Do Not Exist This Layout.")
        self.max_hop, self.dilation = max_hop, dilation
        n, one_based, links, centre = _LAYOUTS[layout]
        self.num_node, self.center = n, centre
        self.edge = [(i, i) for i in range(n)] + [((a - 1, b - 1) if one_based else (a, b)) for a, b in links]
        self.hop_dis = _hops(n, self.edge, max_hop)
        self.A = self._adjacency(strategy)

    def __str__(self):
        return str(self.A)

    def _adjacency(self, strategy):
        n, hop = self.num_node, self.hop_dis
        valid = list(range(0, self.max_hop + 1, self.dilation))
        binary = np.isin(hop, valid).astype(np.float64)
        deg = binary.sum(axis=0)
        inv = np.zeros(n)
        nz = deg > 0
        inv[nz] = deg[nz] ** (-1)
        norm = binary * inv[None, :]                     # A . D^-1
        if strategy == "uniform":
            return norm[None].copy()
        if strategy == "distance":
            return np.stack([np.where(hop == h, norm, 0.0) for h in valid])
        if strategy == "spatial":
            dc = hop[:, self.center]                      # hop distance of every node to the centre
            planes = []
            for h in valid:
                on = hop == h                             # on[j, i]
                same = on & (dc[:, None] == dc[None, :])
                outer = on & (dc[:, None] > dc[None, :])
                inner = on & ~same & ~outer
                if h == 0:
                    planes.append(np.where(same, norm, 0.0))
                else:
                    planes.append(np.where(same | outer, norm, 0.0))
                    planes.append(np.where(inner, norm, 0.0))
            return np.stack(planes)
        raise ValueError("Do Not Exist This Strategy")
