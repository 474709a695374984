"""TEST INFRASTRUCTURE ONLY. ``tn.contractors.greedy`` / ``optimal`` restated.

tensornetwork 0.4.3 asks opt_einsum for a pairwise path and contracts node
pairs with ``contract_between``; for the 2-5 node networks OQuPy builds
(node_array.py:395,519; pt_tebd_backend.py:508,539,550) any pairwise order gives
the same tensor up to floating-point summation order.  Here: repeatedly contract
the connected pair whose result is smallest.
"""
import numpy as np


def _result_size(a, b, shared):
    dims = 1
    for n in (a, b):
        for e in n.edges:
            if not any(e is s for s in shared):
                dims *= e.dimension
    return dims


def _contract_all(nodes, output_edge_order=None, ignore_edge_order=False):
    import tensornetwork as tn
    nodes = list(nodes)
    if len(nodes) == 0:
        raise ValueError("No nodes given.")
    # self-traces first
    for i, n in enumerate(nodes):
        if any(e.is_trace() for e in n.edges):
            nodes[i] = tn.contract_between(n, n)
    while len(nodes) > 1:
        best = None
        for i in range(len(nodes)):
            for j in range(i + 1, len(nodes)):
                shared = tn.get_shared_edges(nodes[i], nodes[j])
                if not shared:
                    continue
                size = _result_size(nodes[i], nodes[j], shared)
                if best is None or size < best[0]:
                    best = (size, i, j)
        if best is None:  # disconnected network: outer product
            i, j = 0, 1
            new = tn.outer_product(nodes[i], nodes[j])
        else:
            _, i, j = best
            new = tn.contract_between(nodes[i], nodes[j])
        nodes = [n for k, n in enumerate(nodes) if k not in (i, j)] + [new]
    final = nodes[0]
    if output_edge_order is not None:
        final.reorder_edges(list(output_edge_order))
    elif not ignore_edge_order and final.get_rank() > 1:
        raise ValueError("output_edge_order must be given when the result "
                         "has more than one dangling edge.")
    return final


def greedy(nodes, output_edge_order=None, ignore_edge_order=False):
    return _contract_all(nodes, output_edge_order, ignore_edge_order)


def optimal(nodes, output_edge_order=None, ignore_edge_order=False,
            memory_limit=None):
    return _contract_all(nodes, output_edge_order, ignore_edge_order)


auto = greedy
branch = greedy
