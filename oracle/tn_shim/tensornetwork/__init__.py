"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the slice of the third-party package ``tensornetwork==0.4.3``
(pinned at /root/reference/requirements.txt:3; not vendored, not installable
here) that OQuPy's hot path calls.  With this package on ``sys.path`` the
reference's own ``oqupy/backends/node_array.py``, ``tempo_backend.py``,
``pt_tempo_backend.py``, ``process_tensor.py`` and ``system_dynamics.py`` run
UNMODIFIED from /root/reference, which is how the golden vectors under
``tests/golden/`` were produced (``tests/golden/make_golden.py``).

Restated from the published tensornetwork 0.4.3 semantics (source absent):
  * ``Node`` / ``Edge`` graph objects, ``edge ^ edge`` (connect), ``node @ node``
    (contract_between: all shared edges; result edge order = remaining edges of
    the left operand followed by those of the right operand),
  * ``contract(edge)``, ``copy``, ``replicate_nodes``, ``remove_node``,
    ``contractors.greedy`` / ``contractors.optimal``, ``flatten_edges``,
    ``split_edge``,
  * ``split_node_full_svd`` -> ``numpy.linalg.svd(full_matrices=False)`` and the
    tail-norm truncation rule of ``backends/numpy/decompositions.py::svd``:
        keep = count_nonzero( sqrt(cumsum(s[::-1]**2)) > eps * s[0] )
    (mirrored in-repo by /root/reference/oqupy/mps_mpo.py:452-457).
Call sites in the reference: node_array.py:163,192,197,218,223,262,285,395,519,541;
system_dynamics.py:638,645,649,697; process_tensor.py:398,403.
"""
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

__version__ = "0.4.3-oracle-shim"

# hook: tests may install a callable(matrix, s, keep) to log every SVD operand
SVD_LOG_HOOK = None


class Edge:
    """An edge joins (node1, axis1) with (node2, axis2); dangling if node2 None."""

    def __init__(self, node1, axis1, name=None, node2=None, axis2=None):
        self.node1 = node1
        self.axis1 = axis1
        self.node2 = node2
        self.axis2 = axis2
        self.name = name if name is not None else "__unnamed_edge__"

    def is_dangling(self) -> bool:
        return self.node2 is None

    def is_trace(self) -> bool:
        return self.node1 is self.node2

    @property
    def dimension(self) -> int:
        return self.node1.tensor.shape[self.axis1]

    def get_nodes(self):
        return [self.node1, self.node2]

    def update_axis(self, old_axis, old_node, new_axis, new_node) -> None:
        if self.node1 is old_node and self.axis1 == old_axis:
            self.node1, self.axis1 = new_node, new_axis
        elif self.node2 is old_node and self.axis2 == old_axis:
            self.node2, self.axis2 = new_node, new_axis
        else:
            raise ValueError("Edge is not attached to (old_node, old_axis).")

    def disconnect(self, edge1_name=None, edge2_name=None):
        return disconnect(self, edge1_name, edge2_name)

    def __xor__(self, other: "Edge") -> "Edge":
        return connect(self, other)

    def __lt__(self, other):  # tensornetwork orders edges by name/signature
        return id(self) < id(other)


class Node:
    def __init__(self, tensor, name=None, axis_names=None, backend=None):
        if isinstance(tensor, Node):
            tensor = tensor.tensor
        self.tensor = np.asarray(tensor)
        self.name = name if name is not None else "__unnamed_node__"
        self.backend = backend
        self.edges = [Edge(self, i) for i in range(self.tensor.ndim)]
        self.axis_names = axis_names

    # -- introspection -----------------------------------------------------
    def get_tensor(self):
        return self.tensor

    def set_tensor(self, tensor):
        self.tensor = tensor

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return self.tensor.dtype

    def get_rank(self) -> int:
        return self.tensor.ndim

    def get_dimension(self, axis) -> int:
        return self.tensor.shape[axis]

    def get_all_edges(self):
        return list(self.edges)

    def get_all_nondangling(self):
        return {e for e in self.edges if not e.is_dangling()}

    def get_all_dangling(self):
        return [e for e in self.edges if e.is_dangling()]

    def get_edge(self, key):
        return self.edges[key]

    def __getitem__(self, key):
        return self.edges[key]

    def add_edge(self, edge, axis, override=False):
        self.edges[axis] = edge

    # -- manipulation ------------------------------------------------------
    def reorder_edges(self, edge_order: Sequence[Edge]) -> "Node":
        if set(map(id, edge_order)) != set(map(id, self.edges)) or \
                len(edge_order) != len(self.edges):
            raise ValueError("edge_order is not a permutation of node's edges.")
        ids = [id(e) for e in self.edges]
        perm = [ids.index(id(e)) for e in edge_order]
        return self.reorder_axes(perm)

    def reorder_axes(self, perm: Sequence[int]) -> "Node":
        old_edges = list(self.edges)
        updates = []  # (edge, end, new_axis): collect first, trace edges have 2 ends
        for new_axis, old_axis in enumerate(perm):
            edge = old_edges[old_axis]
            if edge.node1 is self and edge.axis1 == old_axis:
                updates.append((edge, 1, new_axis))
            else:
                updates.append((edge, 2, new_axis))
        self.tensor = np.transpose(self.tensor, perm)
        self.edges = [old_edges[p] for p in perm]
        for edge, end, new_axis in updates:
            if end == 1:
                edge.axis1 = new_axis
            else:
                edge.axis2 = new_axis
        return self

    def copy(self, conjugate: bool = False) -> "Node":
        tensor = np.conj(self.tensor) if conjugate else self.tensor
        new = Node(tensor, name=self.name, backend=self.backend)
        for i, edge in enumerate(self.edges):
            new.edges[i].name = edge.name
        return new

    def __matmul__(self, other: "Node") -> "Node":
        return contract_between(self, other)


# -- edge operations -----------------------------------------------------------

def connect(edge1: Edge, edge2: Edge, name=None) -> Edge:
    if edge1 is edge2:
        raise ValueError("Cannot connect an edge to itself.")
    if not (edge1.is_dangling() and edge2.is_dangling()):
        raise ValueError("Only dangling edges can be connected.")
    if edge1.dimension != edge2.dimension:
        raise ValueError(
            f"Cannot connect edges of unequal dimension "
            f"({edge1.dimension} != {edge2.dimension}).")
    n1, a1 = edge1.node1, edge1.axis1
    n2, a2 = edge2.node1, edge2.axis1
    new = Edge(n1, a1, name, n2, a2)
    n1.add_edge(new, a1, override=True)
    n2.add_edge(new, a2, override=True)
    return new


def disconnect(edge: Edge, edge1_name=None, edge2_name=None) -> Tuple[Edge, Edge]:
    if edge.is_dangling():
        raise ValueError("Cannot disconnect a dangling edge.")
    e1 = Edge(edge.node1, edge.axis1, edge1_name)
    e2 = Edge(edge.node2, edge.axis2, edge2_name)
    edge.node1.add_edge(e1, edge.axis1, override=True)
    edge.node2.add_edge(e2, edge.axis2, override=True)
    return e1, e2


def get_shared_edges(node1: Node, node2: Node):
    return [e for e in dict.fromkeys(node1.edges)
            if not e.is_dangling() and
            ((e.node1 is node1 and e.node2 is node2) or
             (e.node1 is node2 and e.node2 is node1))]


def _adopt(new_node: Node, donors: List[Tuple[Node, List[int]]]) -> None:
    """Re-point the surviving edges of the donor nodes to new_node, in order."""
    pos = 0
    for donor, axes in donors:
        for ax in axes:
            edge = donor.edges[ax]
            edge.update_axis(ax, donor, pos, new_node)
            new_node.edges[pos] = edge
            pos += 1


def _contract_trace(node: Node, edges: List[Edge]) -> Node:
    tensor = node.tensor
    keep = list(range(tensor.ndim))
    letters = list(range(tensor.ndim))
    for e in edges:
        letters[e.axis2] = letters[e.axis1]
        keep.remove(e.axis1)
        keep.remove(e.axis2)
    new_tensor = np.einsum(tensor, letters, [letters[k] for k in keep])
    new = Node(new_tensor, backend=node.backend)
    _adopt(new, [(node, keep)])
    return new


def contract_between(node1: Node, node2: Node, name=None,
                     allow_outer_product: bool = False,
                     output_edge_order=None, axis_names=None) -> Node:
    if node1 is node2:
        traces = [e for e in dict.fromkeys(node1.edges) if e.is_trace()]
        if not traces:
            raise ValueError("No trace edges to contract.")
        new = _contract_trace(node1, traces)
    else:
        shared = get_shared_edges(node1, node2)
        if not shared and not allow_outer_product:
            raise ValueError(
                f"No edges found between nodes {node1.name} and {node2.name} "
                "and allow_outer_product=False.")
        axes1, axes2 = [], []
        for e in shared:
            if e.node1 is node1:
                axes1.append(e.axis1)
                axes2.append(e.axis2)
            else:
                axes1.append(e.axis2)
                axes2.append(e.axis1)
        new_tensor = np.tensordot(node1.tensor, node2.tensor, [axes1, axes2])
        new = Node(new_tensor, name=name, backend=node1.backend)
        rem1 = [i for i in range(node1.tensor.ndim) if i not in axes1]
        rem2 = [i for i in range(node2.tensor.ndim) if i not in axes2]
        _adopt(new, [(node1, rem1), (node2, rem2)])
    if output_edge_order is not None:
        new.reorder_edges(list(output_edge_order))
    return new


def contract(edge: Edge, name=None, axis_names=None) -> Node:
    """Contract ONE edge (tensordot over that axis pair only)."""
    if edge.is_dangling():
        raise ValueError("Cannot contract a dangling edge.")
    if edge.is_trace():
        return _contract_trace(edge.node1, [edge])
    n1, n2 = edge.node1, edge.node2
    new_tensor = np.tensordot(n1.tensor, n2.tensor, [[edge.axis1], [edge.axis2]])
    new = Node(new_tensor, name=name, backend=n1.backend)
    rem1 = [i for i in range(n1.tensor.ndim) if i != edge.axis1]
    rem2 = [i for i in range(n2.tensor.ndim) if i != edge.axis2]
    _adopt(new, [(n1, rem1), (n2, rem2)])
    return new


def outer_product(node1: Node, node2: Node, name=None) -> Node:
    new_tensor = np.tensordot(node1.tensor, node2.tensor, 0)
    new = Node(new_tensor, name=name, backend=node1.backend)
    _adopt(new, [(node1, list(range(node1.tensor.ndim))),
                 (node2, list(range(node2.tensor.ndim)))])
    return new


# -- network copies --------------------------------------------------------------

def copy(nodes: Iterable[Node], conjugate: bool = False
         ) -> Tuple[Dict[Node, Node], Dict[Edge, Edge]]:
    nodes = list(nodes)
    node_dict = {n: n.copy(conjugate) for n in nodes}
    edge_dict: Dict[Edge, Edge] = {}
    seen = []
    for n in nodes:
        for e in n.edges:
            if any(e is s for s in seen):
                continue
            seen.append(e)
            if e.is_dangling():
                edge_dict[e] = node_dict[n].edges[e.axis1]
                continue
            in1, in2 = e.node1 in node_dict, e.node2 in node_dict
            if in1 and in2:
                new = Edge(node_dict[e.node1], e.axis1, e.name,
                           node_dict[e.node2], e.axis2)
                node_dict[e.node1].edges[e.axis1] = new
                node_dict[e.node2].edges[e.axis2] = new
                edge_dict[e] = new
            elif in1:
                edge_dict[e] = node_dict[e.node1].edges[e.axis1]
            else:
                edge_dict[e] = node_dict[e.node2].edges[e.axis2]
    return node_dict, edge_dict


def replicate_nodes(nodes: Iterable[Node], conjugate: bool = False) -> List[Node]:
    nodes = list(nodes)
    node_dict, _ = copy(nodes, conjugate)
    return [node_dict[n] for n in nodes]


def remove_node(node: Node):
    """Detach node from the network; returns (by_name, by_axis) broken edges."""
    broken_by_axis = {}
    for i, e in enumerate(list(node.edges)):
        if e.is_dangling():
            continue
        e1, e2 = disconnect(e)
        # keep the edge that lives on the *other* node
        broken_by_axis[i] = e2 if e1.node1 is node else e1
    return {}, broken_by_axis


# -- edge reshaping (PT-TEBD only: pt_tebd_backend.py:424-433) ------------------

def flatten_edges(edges: List[Edge], new_edge_name=None) -> Edge:
    """Merge several edges shared by the same node set into one edge."""
    edges = list(edges)
    if all(e.is_dangling() for e in edges):
        node = edges[0].node1
        axes = [e.axis1 for e in edges]
        rest = [i for i in range(node.tensor.ndim) if i not in axes]
        perm = rest + axes
        old_edges = [node.edges[i] for i in rest]
        t = np.transpose(node.tensor, perm)
        t = t.reshape(t.shape[:len(rest)] + (-1,))
        node.tensor = t
        for pos, e in enumerate(old_edges):
            old_axis = rest[pos]
            e.update_axis(old_axis, node, pos, node)
        new = Edge(node, len(rest), new_edge_name)
        node.edges = old_edges + [new]
        return new
    n1, n2 = edges[0].node1, edges[0].node2
    new_halves = []
    dis = [disconnect(e) for e in edges]
    for node in (n1, n2):
        mine = [h for pair in dis for h in pair if h.node1 is node]
        new_halves.append(flatten_edges(mine))
    return connect(new_halves[0], new_halves[1], new_edge_name)


def split_edge(edge: Edge, shape: Tuple[int, ...], new_edge_names=None) -> List[Edge]:
    """Inverse of flatten_edges on one (dangling or connected) edge."""
    shape = tuple(int(s) for s in shape)
    if int(np.prod(shape)) != edge.dimension:
        raise ValueError("Edge dimension does not match the requested shape.")

    def _split_dangling(e: Edge) -> List[Edge]:
        node, ax = e.node1, e.axis1
        nd = node.tensor.ndim
        rest = [i for i in range(nd) if i != ax]
        perm = rest + [ax]
        old_edges = [node.edges[i] for i in rest]
        t = np.transpose(node.tensor, perm)
        t = t.reshape(t.shape[:-1] + shape)
        node.tensor = t
        for pos, oe in enumerate(old_edges):
            if oe.node1 is node and oe.axis1 == rest[pos]:
                oe.axis1 = pos
            elif oe.node2 is node and oe.axis2 == rest[pos]:
                oe.axis2 = pos
        new_edges = [Edge(node, len(rest) + k) for k in range(len(shape))]
        node.edges = old_edges + new_edges
        return new_edges

    if edge.is_dangling():
        return _split_dangling(edge)
    e1, e2 = disconnect(edge)
    a = _split_dangling(e1)
    b = _split_dangling(e2)
    return [connect(x, y) for x, y in zip(a, b)]


# -- SVD split --------------------------------------------------------------------

def _svd_truncated(matrix, max_singular_values, max_truncation_err, relative):
    u, s, vh = np.linalg.svd(matrix, full_matrices=False)
    if max_singular_values is None:
        max_singular_values = s.size
    if max_truncation_err is not None:
        trunc_errs = np.sqrt(np.cumsum(np.square(s[::-1])))
        if relative:
            abs_err = max_truncation_err * s[0]
        else:
            abs_err = max_truncation_err
        num_err = int(np.count_nonzero((trunc_errs > abs_err).astype(np.int32)))
    else:
        num_err = max_singular_values
    keep = min(max_singular_values, num_err)
    if SVD_LOG_HOOK is not None:
        SVD_LOG_HOOK(matrix, s, keep)
    s = s.astype(matrix.dtype)
    return u[:, :keep], s[:keep], vh[:keep, :], s[keep:]


def split_node_full_svd(node: Node, left_edges: List[Edge], right_edges: List[Edge],
                        max_singular_values=None, max_truncation_err=None,
                        relative=False, left_name=None, singular_values_name=None,
                        right_name=None, left_edge_name=None, right_edge_name=None):
    left_edges, right_edges = list(left_edges), list(right_edges)
    node.reorder_edges(left_edges + right_edges)
    nl = len(left_edges)
    ldims, rdims = node.tensor.shape[:nl], node.tensor.shape[nl:]
    mat = node.tensor.reshape(int(np.prod(ldims, dtype=np.int64)),
                              int(np.prod(rdims, dtype=np.int64)))
    u, s, vh, rest = _svd_truncated(mat, max_singular_values,
                                    max_truncation_err, relative)
    k = s.shape[0]
    left = Node(u.reshape(tuple(ldims) + (k,)), name=left_name, backend=node.backend)
    sing = Node(np.diag(s), name=singular_values_name, backend=node.backend)
    right = Node(vh.reshape((k,) + tuple(rdims)), name=right_name, backend=node.backend)
    for i, e in enumerate(left_edges):
        e.update_axis(i, node, i, left)
        left.edges[i] = e
    for i, e in enumerate(right_edges):
        e.update_axis(nl + i, node, i + 1, right)
        right.edges[i + 1] = e
    connect(left.edges[-1], sing.edges[0], left_edge_name)
    connect(sing.edges[1], right.edges[0], right_edge_name)
    return left, sing, right, rest


def split_node(node, left_edges, right_edges, max_singular_values=None,
               max_truncation_err=None, relative=False, **_):
    left, sing, right, rest = split_node_full_svd(
        node, left_edges, right_edges, max_singular_values,
        max_truncation_err, relative)
    sq = Node(np.sqrt(sing.tensor), backend=node.backend)
    # U*sqrt(S), sqrt(S)*Vh
    l_edges = left.edges[:-1]
    r_edges = right.edges[1:]
    sq2 = Node(np.sqrt(sing.tensor), backend=node.backend)
    e_l, _ = disconnect(left.edges[-1])
    _, e_r = disconnect(right.edges[0])
    connect(e_l, sq.edges[0])
    connect(sq2.edges[1], e_r)
    new_left = contract_between(left, sq)
    new_right = contract_between(sq2, right)
    connect(new_left.edges[-1], new_right.edges[0])
    del l_edges, r_edges
    return new_left, new_right, rest


from . import contractors  # noqa: E402,F401  (tn.contractors.greedy / optimal)
