"""Host-pointer batch transform (`hb_transform_batch`): the copy/compute pipeline, and its CUDA-graph replay for
repeated calls on the same page-locked buffers, against the eager path on pageable buffers (bitwise) and the
compiled reference."""
import numpy as np
import pytest

from conftest import gauss_hermite, rel_err

pytestmark = pytest.mark.gpu


def _pinned(shape):
    import torch
    return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()


@pytest.mark.parametrize("index_set,n,degree,batch", [("cube", 129, 64, 12), ("triangle", 128, 70, 9)])
def test_graph_replay_matches_eager_bitwise(index_set, n, degree, batch):
    from hermipy_b200 import core
    x, w = gauss_hermite(n)
    nodes, weights = [x, x], [w, w]
    npol = core.iterator_size(2, degree, index_set)
    rng = np.random.default_rng(n)
    f_pin, c_pin, g_pin = _pinned((batch, n * n)), _pinned((batch, npol)), _pinned((batch, n * n))
    assert f_pin.nbytes >= 1 << 20
    for call in range(5):                       # call 0 eager, call 1 captured, calls 2+ replayed
        f = rng.standard_normal((batch, n * n))
        f_pin[:] = f
        core.transform_batch(degree, f_pin, nodes, weights, True, index_set=index_set, out=c_pin)
        want_c = core.transform_batch(degree, f, nodes, weights, True, index_set=index_set)       # pageable: eager
        assert np.array_equal(c_pin, want_c), call
        core.transform_batch(degree, c_pin, nodes, weights, False, index_set=index_set, out=g_pin)
        want_g = core.transform_batch(degree, want_c, nodes, weights, False, index_set=index_set)
        assert np.array_equal(g_pin, want_g), call
        if call == 2:
            # a larger batch through the same entry point grows the staging buffers: captured graphs that hold
            # the old addresses must be re-captured, not replayed
            big = rng.standard_normal((3 * batch, n * n))
            got_big = core.transform_batch(degree, big, nodes, weights, True, index_set=index_set)
            assert np.array_equal(got_big[:batch], core.transform_batch(degree, big[:batch], nodes, weights, True,
                                                                        index_set=index_set))


def test_graph_replay_against_reference(ref):
    from hermipy_b200 import core
    n, degree, batch = 96, 40, 16
    x, w = gauss_hermite(n)
    nodes, weights = [x, x], [w, w]
    npol = core.iterator_size(2, degree, "triangle")
    rng = np.random.default_rng(3)
    f_pin, c_pin = _pinned((batch, n * n)), _pinned((batch, npol))
    for call in range(3):
        f_pin[:] = rng.standard_normal((batch, n * n))
        core.transform_batch(degree, f_pin, nodes, weights, True, index_set="triangle", out=c_pin)
    want = ref.transform(degree, f_pin[5].reshape(n, n), nodes, weights, True, index_set="triangle")
    assert rel_err(c_pin[5], want) < 1e-12


def test_alternating_buffers_and_plans_keep_separate_graphs():
    from hermipy_b200 import core
    n, batch = 128, 10
    x, w = gauss_hermite(n)
    nodes, weights = [x, x], [w, w]
    rng = np.random.default_rng(8)
    bufs = []
    for degree in (30, 50):
        npol = core.iterator_size(2, degree, "cube")
        for _ in range(2):
            bufs.append((degree, _pinned((batch, n * n)), _pinned((batch, npol))))
    for call in range(4):
        for degree, f_pin, c_pin in bufs:
            f_pin[:] = rng.standard_normal((batch, n * n))
            core.transform_batch(degree, f_pin, nodes, weights, True, index_set="cube", out=c_pin)
            want = core.transform_batch(degree, f_pin.copy(), nodes, weights, True, index_set="cube")
            assert np.array_equal(c_pin, want), (call, degree)


@pytest.mark.parametrize("dim,n,degree,batch,index_set", [(1, 257, 128, 4096, "triangle"), (1, 64, 63, 20000, "cube"),
                                                          (3, 48, 20, 5, "triangle"), (3, 32, 31, 6, "cube")])
def test_graph_replay_other_dimensions(dim, n, degree, batch, index_set):
    """1-D batches (even and odd row lengths: with and without the re-pitch kernels) and 3-D grids through the
    pinned-buffer path, forward and inverse, against the eager path."""
    from hermipy_b200 import core
    x, w = gauss_hermite(n)
    nodes, weights = [x] * dim, [w] * dim
    npol = core.iterator_size(dim, degree, index_set)
    rng = np.random.default_rng(dim * 1000 + n)
    f_pin, c_pin, g_pin = _pinned((batch, n ** dim)), _pinned((batch, npol)), _pinned((batch, n ** dim))
    for call in range(4):
        f = rng.standard_normal((batch, n ** dim))
        f_pin[:] = f
        core.transform_batch(degree, f_pin, nodes, weights, True, index_set=index_set, out=c_pin)
        want_c = core.transform_batch(degree, f, nodes, weights, True, index_set=index_set)
        assert np.array_equal(c_pin, want_c), call
        core.transform_batch(degree, c_pin, nodes, weights, False, index_set=index_set, out=g_pin)
        want_g = core.transform_batch(degree, want_c, nodes, weights, False, index_set=index_set)
        assert np.array_equal(g_pin, want_g), call
