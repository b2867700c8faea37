"""
Host-side lowering: networkx trees / rate matrices / observation maps ->
dense arrays and the small integer "programs" the CUDA kernels walk.

Node numbering follows the reference's own lowering for its native calls:
DFS preorder from the root (raoteh/sampler/_mcy_dense.py:153,
_density.py:104-140), so parent index < child index, and per-edge matrices
are indexed by the CHILD node with the root slot unused (_density.py:143-180).
"""
from __future__ import annotations

import numpy as np

# ---- op codes of the upward program (mirrored in csrc/rt_common.cuh) -------
OP_MSG_SLOT = 0    # node=c, a=stack slot of L_c, b=store index of c   acc *= P_c . L_c
OP_MSG_OBS = 1     # node=c, a=obs slot of leaf c           acc *= P_c . obs_c
OP_MSG_ONES = 2    # node=c (unobserved leaf)               acc *= P_c . 1
OP_APPLY_OBS = 3   # node=v, a=obs slot                     acc *= obs_v
OP_STORE = 4       # node=v, a=stack slot, b=store index    rescale, park L_v
OP_ROOT = 5        # node=root, b=store index               rescale, combine with pi
OP_FLAG_FRESH = 1 << 8   # on OP_MSG_SLOT: L_c is the partial stored by the previous OP_STORE
OP_FLAG_KEEP_ON_CHIP = 1 << 9   # on OP_STORE: the next op consumes this partial FRESH
OP_FLAG_PARK = 1 << 10          # on OP_STORE: a later, non-fresh op reads it back from its slot

MISSING = 255


class TreeSchedule(object):
    """Rooted tree in DFS-preorder arrays plus traversal programs."""

    def __init__(self, parent, length, nodes=None):
        self.parent = np.asarray(parent, dtype=np.int32)
        self.length = np.asarray(length, dtype=np.float64)
        self.n = n = len(self.parent)
        if n < 1 or self.parent[0] != -1:
            raise ValueError('node 0 must be the root (parent -1)')
        if n > 1 and not (self.parent[1:] < np.arange(1, n)).all():
            raise ValueError('nodes must be in preorder (parent < child)')
        self.nodes = list(range(n)) if nodes is None else list(nodes)
        self.node_index = dict((v, i) for i, v in enumerate(self.nodes))
        self.children = [[] for _ in range(n)]
        for b in range(1, n):
            self.children[self.parent[b]].append(b)
        self.is_leaf = np.array([len(c) == 0 for c in self.children])
        self.leaves = np.nonzero(self.is_leaf)[0].astype(np.int32)
        self.internal = np.nonzero(~self.is_leaf)[0].astype(np.int32)
        # store index of internal nodes: rank in preorder (root -> 0)
        self.store_index = np.full(n, -1, dtype=np.int32)
        self.store_index[self.internal] = np.arange(len(self.internal), dtype=np.int32)
        self.n_store = len(self.internal)
        self.depth = np.zeros(n, dtype=np.int32)
        for b in range(1, n):
            self.depth[b] = self.depth[self.parent[b]] + 1

    # -- construction from the reference's tree type -----------------------
    @classmethod
    def from_nx(cls, T, root):
        """T: undirected weighted nx.Graph; preorder as
        raoteh/sampler/_mcy_dense.py:153 (nx.dfs_preorder_nodes)."""
        import networkx as nx
        if root not in T:
            raise ValueError('the specified root is not in the tree')
        nodes = list(nx.dfs_preorder_nodes(T, root))
        if len(nodes) != T.number_of_nodes():
            raise ValueError('the tree is not connected')
        index = dict((v, i) for i, v in enumerate(nodes))
        parent = np.full(len(nodes), -1, dtype=np.int32)
        length = np.zeros(len(nodes), dtype=np.float64)
        for a, b in nx.dfs_edges(T, root):
            parent[index[b]] = index[a]
            w = T[a][b].get('weight', None)
            length[index[b]] = 0.0 if w is None else float(w)
        return cls(parent, length, nodes)

    @property
    def n_edges(self):
        return self.n - 1

    # -- upward program -----------------------------------------------------
    def up_program(self, obs_slot=None):
        """Postorder program over internal nodes with a minimal slot stack.

        obs_slot[n]: row of the observation array per node, -1 = unobserved.
        Returns (ops int32 [n_ops,4], n_slots).
        """
        n = self.n
        if obs_slot is None:
            obs_slot = np.full(n, -1, dtype=np.int32)
        need = np.zeros(n, dtype=np.int64)
        order_children = [None] * n
        for v in range(n - 1, -1, -1):
            ch = self.children[v]
            if not ch:
                continue
            internal = sorted([c for c in ch if not self.is_leaf[c]],
                              key=lambda c: -need[c])
            leaf = [c for c in ch if self.is_leaf[c]]
            peak = 1
            for i, c in enumerate(internal):
                peak = max(peak, need[c] + i)
            peak = max(peak, len(internal))
            need[v] = peak
            order_children[v] = internal + leaf
        ops = []
        free = []
        n_slots = [0]
        slot_of = {}

        def alloc():
            if free:
                return free.pop()
            n_slots[0] += 1
            return n_slots[0] - 1

        # iterative postorder honouring order_children
        stack = [(0, 0)]
        while stack:
            v, i = stack.pop()
            ch = order_children[v]
            internal = [c for c in ch if not self.is_leaf[c]]
            if i < len(internal):
                stack.append((v, i + 1))
                stack.append((internal[i], 0))
                continue
            # the child stored last is still on chip: consume it first (FRESH)
            if internal:
                ch = [internal[-1]] + internal[:-1] + [c for c in ch if self.is_leaf[c]]
            for k, c in enumerate(ch):
                if not self.is_leaf[c]:
                    fresh = OP_FLAG_FRESH if (k == 0 and c == internal[-1]) else 0
                    ops.append((OP_MSG_SLOT | fresh, c, slot_of[c],
                                int(self.store_index[c])))
                elif obs_slot[c] >= 0:
                    ops.append((OP_MSG_OBS, c, int(obs_slot[c]), 0))
                else:
                    ops.append((OP_MSG_ONES, c, 0, 0))
            if obs_slot[v] >= 0:
                ops.append((OP_APPLY_OBS, v, int(obs_slot[v]), 0))
            for c in ch:
                if not self.is_leaf[c]:
                    free.append(slot_of.pop(c))
            if v == 0:
                ops.append((OP_ROOT, v, 0, int(self.store_index[v])))
            else:
                s = alloc()
                slot_of[v] = s
                ops.append((OP_STORE, v, s, int(self.store_index[v])))
        ops = [list(o) for o in ops]
        for i, o in enumerate(ops):
            if (o[0] & 0xff) != OP_STORE:
                continue
            nxt = ops[i + 1] if i + 1 < len(ops) else None
            fresh_next = (nxt is not None and (nxt[0] & 0xff) == OP_MSG_SLOT and
                          (nxt[0] & OP_FLAG_FRESH) and nxt[1] == o[1])
            o[0] |= OP_FLAG_KEEP_ON_CHIP if fresh_next else OP_FLAG_PARK
        return np.asarray(ops, dtype=np.int32).reshape(-1, 4), max(1, n_slots[0])

    # -- downward program -----------------------------------------------------
    def down_program(self, obs_slot=None):
        """Edges grouped by depth of the child (parents before children).

        Returns (edges int32 [n_edges,4], level_ptr): each row is
        (child node, parent store idx, child store idx or -1, child obs slot),
        rows sorted by level; level_ptr[l]..level_ptr[l+1] are the edges whose
        child is at depth l+1.
        """
        n = self.n
        if obs_slot is None:
            obs_slot = np.full(n, -1, dtype=np.int32)
        order = sorted(range(1, n), key=lambda b: (self.depth[b], b))
        rows = []
        level_ptr = [0]
        cur = 1
        for b in order:
            while self.depth[b] > cur:
                level_ptr.append(len(rows))
                cur += 1
            rows.append((b, int(self.store_index[self.parent[b]]),
                         int(self.store_index[b]), int(obs_slot[b])))
        level_ptr.append(len(rows))
        return (np.asarray(rows, dtype=np.int32).reshape(-1, 4),
                np.asarray(level_ptr, dtype=np.int32))


def subdivide_long_branches(sched, max_length):
    """Split every branch longer than max_length into equal pieces joined by new degree-two
    nodes (appended labels: None).  Returns (new TreeSchedule, pieces) where pieces[c] lists, for
    the original child node c, the new-schedule nodes of its pieces from the parent end to the
    child end (the last one is the image of c), and image[v] is the new index of old node v.
    The per-branch event counts of the sampling kernels are uint8, so a branch must not expect
    more than ~255 candidate events (omega * length); this keeps long branches usable."""
    n = sched.n
    if not (max_length > 0):
        raise ValueError('max_length must be positive')
    kpieces = np.ones(n, dtype=np.int64)
    for c in range(1, n):
        kpieces[c] = max(1, int(np.ceil(sched.length[c] / max_length)))
    if int(kpieces.max()) == 1:
        return sched, dict((c, [c]) for c in range(1, n)), np.arange(n)
    # preorder of the subdivided tree: walk the old preorder, inserting the chain before each node
    parent, length, nodes = [], [], []
    image = np.zeros(n, dtype=np.int64)
    pieces = {}
    for v in range(n):
        if v == 0:
            parent.append(-1)
            length.append(0.0)
            nodes.append(sched.nodes[0])
            image[0] = 0
            continue
        k = int(kpieces[v])
        prev = int(image[sched.parent[v]])
        chain = []
        for i in range(k):
            parent.append(prev)
            length.append(sched.length[v] / k)
            nodes.append(sched.nodes[v] if i == k - 1 else ('__piece__', sched.nodes[v], i))
            prev = len(parent) - 1
            chain.append(prev)
        image[v] = prev
        pieces[v] = chain
    # the old preorder with chains inserted keeps parent < child only if children follow their
    # parent's image, which holds because image[parent] was assigned before v was visited
    return TreeSchedule(np.asarray(parent, dtype=np.int32), np.asarray(length), nodes), pieces, image


# ---------------------------------------------------------------------------
# rate matrices
# ---------------------------------------------------------------------------
def dense_rate_matrix(Q_sparse, states=None):
    """nx.DiGraph without diagonal -> (states, dense Q with diagonal).

    raoteh/sampler/_util.py:27-53 (get_dense_rate_matrix) and
    _density.py:32-54 (rate_matrix_to_numpy_array).
    """
    if states is None:
        states = sorted(Q_sparse)
    index = dict((s, i) for i, s in enumerate(states))
    S = len(states)
    Q = np.zeros((S, S), dtype=np.float64)
    for sa, sb, d in Q_sparse.edges(data=True):
        if sa == sb:
            continue
        Q[index[sa], index[sb]] = d['weight']
    Q -= np.diag(Q.sum(axis=1))
    return list(states), Q


def check_square_dense(M):
    """raoteh/sampler/_density.py:78-101"""
    if M is None:
        raise ValueError('the matrix is None')
    shape = getattr(M, 'shape', None)
    if shape is None:
        if hasattr(M, 'number_of_nodes'):
            raise ValueError('expected an ndarray but found a graph object')
        raise ValueError('expected an ndarray')
    if len(shape) != 2:
        raise ValueError('expected len(M.shape) == 2')
    if shape[0] != shape[1]:
        raise ValueError('expected the array to be square')


# ---------------------------------------------------------------------------
# observations
# ---------------------------------------------------------------------------
def allowed_sets_to_mask(sched, node_to_allowed_states, nstates, state_index=None):
    """dict node -> set of allowed states  =>  uint64 bitmask [n] for ONE site.

    Missing node or map None = unrestricted (raoteh/sampler/_mcy_dense.py:43-54,
    _mcy.py:108-130).  Nodes not in the tree are ignored.
    """
    full = (1 << nstates) - 1
    mask = np.full(sched.n, full, dtype=np.uint64)
    if node_to_allowed_states is not None:
        for v, allowed in node_to_allowed_states.items():
            i = sched.node_index.get(v)
            if i is None:
                continue
            m = 0
            for s in allowed:
                k = s if state_index is None else state_index.get(s)
                if k is not None and 0 <= k < nstates:
                    m |= (1 << k)
            mask[i] = m
    return mask
