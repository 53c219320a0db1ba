"""CPU tests (-m "not gpu"): pin the oracle.  The NumPy restatement (oracle/port.py) is checked against
the golden vectors generated from the reference's own object code (tests/golden/, oracle/make_golden.py),
against the known answers captured during the survey, and against oracle/_ref itself when present."""
import os

import numpy as np
import pytest

from conftest import rel_err
from oracle import port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SETS = ["triangle", "cube", "cross", "cross_nc", "rectangle"]


def gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_known_answer_enumerations():
    # captured from the compiled reference during the survey
    kat = {
        ("triangle", 2, 3): "(0,0)(0,1)(1,0)(0,2)(1,1)(2,0)(0,3)(1,2)(2,1)(3,0)",
        ("triangle", 3, 2): "(0,0,0)(0,0,1)(0,1,0)(1,0,0)(0,0,2)(0,1,1)(1,0,1)(0,2,0)(1,1,0)(2,0,0)",
        ("cube", 2, 2): "(0,0)(1,0)(0,1)(1,1)(2,0)(2,1)(0,2)(1,2)(2,2)",
        ("cross", 2, 4): "(0,0)(1,0)(0,1)(1,1)(0,2)(1,2)(2,0)(2,1)(0,3)(1,3)(3,0)(3,1)(0,4)(1,4)(2,2)(4,0)(4,1)",
        ("cross_nc", 2, 4): "(0,0)(0,1)(1,0)(0,2)(1,1)(2,0)(0,3)(1,2)(2,1)(3,0)(0,4)(1,3)(2,2)(3,1)(4,0)(1,4)(4,1)",
        ("rectangle", 2, 4): "(0,0)(1,0)(2,0)(0,1)(1,1)(2,1)(3,0)(3,1)(4,0)(4,1)(0,2)(1,2)(2,2)(3,2)(4,2)",
    }
    for (s, d, N), text in kat.items():
        got = "".join("(" + ",".join(str(v) for v in m) + ")" for m in port.list_indices(d, N, s))
        assert got == text, (s, d, N)


def test_port_enumerations_match_golden():
    g = gold("index_sets")
    for key in g.files:
        if key.startswith("sizes"):
            continue
        s, d, N = key.rsplit("_", 2)
        d, N = int(d[1:]), int(N[1:])
        if d * N > 60:
            continue  # pure-Python enumeration: small cases only
        assert np.array_equal(port.list_indices(d, N, s), g[key]), key


def test_port_hermite_eval_matches_golden():
    g = gold("hermite_eval")
    got = port.hermite_eval(g["x"], 40)
    assert np.array_equal(got, g["values_o2"])         # strict IEEE build of the reference: bit-exact
    assert rel_err(got, g["values"]) < 1e-15            # the -Ofast build the reference ships
    kat = [1, 1.5, 0.88388347648318444, -0.4592793267718458, -1.1099250396986275, -0.33376843347971064]
    assert np.allclose(port.hermite_eval([1.5], 5)[0], kat, rtol=1e-15, atol=0)


def test_port_transform_matches_golden():
    g = gold("transform")
    c = port.transform(20, g["t1_f"], [g["t1_nodes"]], [g["t1_weights"]], True)
    assert rel_err(c, g["t1_fwd"]) < 1e-13
    assert rel_err(port.transform(20, g["t1_fwd"], [g["t1_nodes"]], [g["t1_weights"]], False), g["t1_bwd"]) < 1e-13
    nodes, weights = [g["t2_nodes0"], g["t2_nodes1"]], [g["t2_weights0"], g["t2_weights1"]]
    for s in SETS:
        assert rel_err(port.transform(10, g["t2_f"], nodes, weights, True, index_set=s), g["t2_fwd_" + s]) < 1e-13
        assert rel_err(port.transform(10, g["t2_fwd_" + s], nodes, weights, False, index_set=s),
                       g["t2_bwd_" + s]) < 1e-13
    for s in ("triangle", "cube"):
        assert rel_err(port.transform(6, g["t3_f"], [g["t3_nodes"]] * 3, [g["t3_weights"]] * 3, True, index_set=s),
                       g["t3_fwd_" + s]) < 1e-13


def test_port_triple_products_and_varf_match_golden():
    g = gold("varf")
    assert np.array_equal(port.triple_products(6), g["T6_o2"])
    assert rel_err(port.triple_products(6), g["T6"]) < 1e-15
    assert np.array_equal(port.triple_products(30)[::7], g["T30_o2_k0"])
    assert rel_err(port.triple_products(40), port.triple_products_closed_form(40)) < 1e-12
    A = port.varf(12, g["v1_f"], [g["v1_nodes"]], [g["v1_weights"]])
    assert rel_err(A, g["v1_varf_N12"]) < 1e-13
    assert np.array_equal(A != 0, g["v1_varf_N12"] != 0)
    for s in SETS:
        A = port.varf(5, g["v2_f"], [g["v2_nodes"]] * 2, [g["v2_weights"]] * 2, index_set=s)
        assert rel_err(A, g["v2_varf_" + s]) < 1e-13, s
        assert rel_err(port.varfd(2, 5, 0, g["v2_varf_" + s], index_set=s), g["v2_varfd0_" + s]) < 1e-15


def test_port_tensorize_project_match_golden():
    g = gold("tensorize_project")
    for s in SETS:
        v, m = list(g["in_v_" + s]), list(g["in_m_" + s])
        assert np.array_equal(port.tensorize_vecs(v, [[0], [1], [2]], s), g["tens_v_" + s])
        assert np.array_equal(port.tensorize_mats(m, [[0], [1]], s), g["tens_m_" + s])
        assert np.array_equal(port.tensorize_vecs([g["in_vxy_" + s], v[1]], [[0, 2], [1]], s), g["tens_dirs_" + s])
        assert np.array_equal(port.project(g["tens_v_" + s], 3, [0, 2], s), g["proj_v02_" + s])
        assert np.array_equal(port.project(g["tens_m_" + s], 2, [1], s), g["proj_m1_" + s])


def test_port_matches_compiled_reference(ref):
    """Where the reference's own object code is available (build container and, via the prebuilt .so,
    the GPU box), the restatement is compared with it directly on fresh seeded inputs."""
    rng = np.random.default_rng(11)
    import scipy.special as sp
    x, w = sp.roots_hermitenorm(17)
    w = w / np.sqrt(2 * np.pi)
    f = rng.standard_normal((17, 17))
    for s in SETS:
        assert np.array_equal(port.list_indices(3, 5, s), ref.iterator_list_indices(3, 5, s))
        c = ref.transform(7, f.ravel(), [x, x], [w, w], True, index_set=s)
        assert rel_err(port.transform(7, f, [x, x], [w, w], True, index_set=s), c) < 1e-13
    X, Y = np.meshgrid(x, x, indexing="ij")
    assert rel_err(port.varf(4, X * Y + Y ** 2, [x, x], [w, w]), ref.varf(4, (X * Y + Y ** 2).ravel(), [x, x], [w, w])) < 1e-13
    assert np.array_equal(port.triple_products(25), ref.triple_products(25, variant="_o2"))
