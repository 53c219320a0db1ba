"""TEST INFRASTRUCTURE.  The compiled reference (oracle/_ref) on all host cores: its transform() is a naive loop
over (grid point, multi-index) pairs (transform.cpp:115-146), linear in the number of grid points, so a d-dimensional
transform is cut into slabs of the outer grid axis -- the forward sum is additive over slabs, the inverse produces
each slab's values independently -- and the slabs run in a process pool (the reference itself is single-threaded)."""
import multiprocessing as mp
import os

import numpy as np


def _slab_job(args):
    degree, f_slab, nodes, weights, forward, index_set, a, b = args
    from oracle import ref
    nd, wt = [nodes[0][a:b]] + list(nodes[1:]), [weights[0][a:b]] + list(weights[1:])
    return ref.transform(degree, f_slab, nd, wt, forward, index_set=index_set)


def reference_transform_slabs(degree, data, nodes, weights, forward, index_set, rows_per_slab=None, processes=None):
    """data: grid values (n_0, ..., n_{d-1}) for forward, coefficient vector for the inverse."""
    n0 = len(nodes[0])
    procs = processes or os.cpu_count() or 1
    rows = rows_per_slab or max(1, -(-n0 // (4 * procs)))
    bounds = [(a, min(a + rows, n0)) for a in range(0, n0, rows)]
    nodes, weights = [np.asarray(v, dtype=float) for v in nodes], [np.asarray(v, dtype=float) for v in weights]
    if forward:
        data = np.asarray(data, dtype=float).reshape([len(v) for v in nodes])
        jobs = [(degree, data[a:b].ravel(), nodes, weights, True, index_set, a, b) for a, b in bounds]
    else:
        coef = np.asarray(data, dtype=float)
        jobs = [(degree, coef, nodes, weights, False, index_set, a, b) for a, b in bounds]
    with mp.get_context("spawn").Pool(min(procs, len(jobs))) as pool:
        parts = pool.map(_slab_job, jobs)
    if forward:
        return np.sum(parts, axis=0)
    return np.concatenate(parts).reshape([len(v) for v in nodes])
