"""Generates tests/golden/*.npz from the REFERENCE'S OWN OBJECT CODE (oracle/_ref/libhermite_ref.so,
built by oracle/Makefile from /root/reference/cpp/src/hermite/*.cpp).  Run in the build container,
where /root/reference exists:

    make -C oracle ref && python oracle/make_golden.py

The fixtures travel with the repository; nothing reads /root/reference at test time.
Inputs are seeded; nodes/weights are stored in the fixture so that consumers need no quadrature code.
"""
import os
import sys

import numpy as np
import scipy.special as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SETS = ["triangle", "cube", "cross", "cross_nc", "rectangle"]


def gh(n):
    x, w = sp.roots_hermitenorm(n)
    return x, w / np.sqrt(2 * np.pi)


def main():
    os.makedirs(OUT, exist_ok=True)
    # --- index sets (bit-exact contract) ---
    idx = {}
    for s in SETS:
        for dim, degree in [(1, 5), (2, 9), (3, 6), (4, 3), (2, 30)]:
            if s == "cross" and degree == 0:
                continue
            idx["%s_d%d_N%d" % (s, dim, degree)] = ref.iterator_list_indices(dim, degree, s).astype(np.int16)
    sizes = {"%s_d%d_N%d" % (s, d, N): ref.iterator_size(d, N, s)
             for s in SETS for d in (1, 2, 3) for N in (1, 7, 40, 200)}
    idx["sizes_keys"] = np.array(sorted(sizes))
    idx["sizes_vals"] = np.array([sizes[k] for k in sorted(sizes)], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "index_sets.npz"), **idx)

    # --- recurrence ---
    xs = np.array([-7.5, -2.0, -0.3, 0.0, 0.7, 1.5, 3.1, 9.0])
    np.savez_compressed(os.path.join(OUT, "hermite_eval.npz"), x=xs,
                        values=np.array([ref.hermite_eval(x, 40) for x in xs]),
                        values_o2=np.array([ref.hermite_eval(x, 40, variant="_o2") for x in xs]))

    # --- transforms ---
    tr = {}
    x, w = gh(41)
    f = np.exp(-x ** 2 / 8) * np.cos(3 * x)
    tr["t1_nodes"], tr["t1_weights"], tr["t1_f"] = x, w, f
    tr["t1_fwd"] = ref.transform(20, f, [x], [w], True)
    tr["t1_bwd"] = ref.transform(20, tr["t1_fwd"], [x], [w], False)
    x2, w2 = gh(24)
    y2, v2 = gh(19)
    X, Y = np.meshgrid(x2, y2, indexing="ij")
    f2 = np.exp(-(X ** 2 + X * Y + Y ** 2) / 10)
    tr["t2_nodes0"], tr["t2_weights0"], tr["t2_nodes1"], tr["t2_weights1"], tr["t2_f"] = x2, w2, y2, v2, f2
    for s in SETS:
        c = ref.transform(10, f2.ravel(), [x2, y2], [w2, v2], True, index_set=s)
        tr["t2_fwd_" + s] = c
        tr["t2_bwd_" + s] = ref.transform(10, c, [x2, y2], [w2, v2], False, index_set=s)
    x3, w3 = gh(9)
    X, Y, Z = np.meshgrid(x3, x3, x3, indexing="ij")
    f3 = np.exp(-(X ** 2 + Y ** 2 + Z ** 2 + X * Z) / 9) * (1 + 0.3 * Y)
    tr["t3_nodes"], tr["t3_weights"], tr["t3_f"] = x3, w3, f3
    for s in ("triangle", "cube"):
        tr["t3_fwd_" + s] = ref.transform(6, f3.ravel(), [x3] * 3, [w3] * 3, True, index_set=s)
    np.savez_compressed(os.path.join(OUT, "transform.npz"), **tr)

    # --- triple products and varf ---
    vf = {"T6": ref.triple_products(6), "T6_o2": ref.triple_products(6, variant="_o2"),
          "T30_o2_k0": ref.triple_products(30, variant="_o2")[::7]}
    x, w = gh(41)
    vf["v1_nodes"], vf["v1_weights"] = x, w
    f1 = x ** 4 / 4 - x ** 2 / 2
    vf["v1_f"] = f1
    vf["v1_varf_N12"] = ref.varf(12, f1, [x], [w])
    x, w = gh(21)
    X, Y = np.meshgrid(x, x, indexing="ij")
    V = X ** 4 / 4 - X ** 2 / 2 + Y ** 4 / 4 - Y ** 2 / 2 + X * Y / 2
    vf["v2_nodes"], vf["v2_weights"], vf["v2_f"] = x, w, V
    for s in SETS:
        A = ref.varf(5, V.ravel(), [x, x], [w, w], index_set=s)
        vf["v2_varf_" + s] = A
        vf["v2_varfd0_" + s] = ref.varfd(2, 5, 0, A, index_set=s)
    np.savez_compressed(os.path.join(OUT, "varf.npz"), **vf)

    # --- tensorize / project ---
    rng = np.random.default_rng(7)
    tp = {}
    for s in SETS:
        n1 = ref.iterator_size(1, 4, s)
        n2 = ref.iterator_size(2, 4, s)
        v = [rng.standard_normal(n1) for _ in range(3)]
        m = [rng.standard_normal((n1, n1)) for _ in range(2)]
        vxy = rng.standard_normal(n2)
        tp["in_v_" + s] = np.array(v)
        tp["in_m_" + s] = np.array(m)
        tp["in_vxy_" + s] = vxy
        tp["tens_v_" + s] = ref.tensorize(v, index_set=s)
        tp["tens_m_" + s] = ref.tensorize(m, index_set=s)
        tp["tens_dirs_" + s] = ref.tensorize({frozenset({0, 2}): vxy, frozenset({1}): v[1]}, index_set=s)
        tp["proj_v02_" + s] = ref.project(tp["tens_v_" + s], 3, [0, 2], index_set=s)
        tp["proj_m1_" + s] = ref.project(tp["tens_m_" + s], 2, [1], index_set=s)
    np.savez_compressed(os.path.join(OUT, "tensorize_project.npz"), **tp)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
