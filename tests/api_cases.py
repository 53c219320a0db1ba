"""Object-layer parity cases, written once and run against two implementations of the hermipy API:

  * the reference's UNMODIFIED Python package on top of its own C++ (oracle/make_golden_api.py, in the build
    container) -> tests/golden/api.npz;
  * this repository's hermipy_b200 (tests/test_api_golden.py): on the CUDA backend (-m gpu) and, for the host
    logic, on the oracle backend (tests/oracle_backend.py).

`run(hm)` takes the package (`import hermipy as hm` / `import hermipy_b200 as hm`) and returns
{case name: ndarray}.  Quadratures are built from explicit nodes and weights so that both sides see
bit-identical inputs.  The shapes follow the reference's own tests (tests/test_quad.py,
tests/test_tensorize_project.py, tests/test_series.py, tests/test_equations.py) and BASELINE.json configs[0].
"""
import numpy as np
import numpy.polynomial.hermite_e as herm
import sympy as sym


def gh(n):
    x, w = herm.hermegauss(n)        # fine for n <= 370 (SURVEY.md finding 3)
    return x, w / np.sqrt(2 * np.pi)


def quad_gh(hm, n_points, **kwargs):
    rules = [gh(n) for n in n_points]
    return hm.Quad([r[0] for r in rules], [r[1] for r in rules], **kwargs)


def wl2(quad, values):
    """Grid values times sqrt(quadrature weight): the inverse transform evaluates He_k/sqrt(k!) un-weighted, which
    reaches 1e45 and beyond at the outer nodes where the coefficients' rounding noise is amplified alike (for the
    reference as much as for any other implementation), so evaluations are compared in the discrete L2(w) sense."""
    w = np.ones(1)
    for wi in quad.weights:
        w = np.multiply.outer(w, np.asarray(wi)).ravel()
    return np.asarray(values) * np.sqrt(w)


def dense(m):
    return np.asarray(m.todense()) if hasattr(m, "todense") else np.asarray(m)


def case_fokker_planck_1d(hm, out):
    """BASELINE configs[0]: 1-D Fokker-Planck, degree 100, 201 nodes (tests/test_equations.py:31-141)."""
    x, f = sym.symbols('x', real=True), sym.Function('f')
    beta, m, s2, degree = sym.Rational(3), sym.Rational(1, 10), sym.Rational(1, 5), 100
    quad = quad_gh(hm, [2 * degree + 1], mean=[float(m)], cov=[[float(s2)]])
    for name, Vp in (("gaussian", x * x / 4), ("bistable", x ** 4 / 4 - x ** 2 / 2)):
        Vq = sym.Rational(1, 2) * (x - m) * (x - m) / (beta * s2)
        forward = sym.diff(sym.diff(Vp, x) * f(x) + 1 / beta * sym.diff(f(x), x), x)
        factor = sym.exp(-beta / 2 * (Vq + Vp))
        backward = (forward.subs(f(x), f(x) * factor).doit() / factor).expand()
        out["fp1d/%s/operator" % name] = dense(quad.discretize_op(backward, degree).matrix)
        out["fp1d/%s/operator_sparse" % name] = dense(quad.discretize_op(backward, degree, sparse=True).matrix)
        out["fp1d/%s/factor_grid" % name] = quad.discretize(factor)
    series = quad.transform('exp(-x*x/8)*cos(3*x)', degree)
    out["fp1d/transform"] = series.coeffs
    out["fp1d/eval"] = wl2(quad, quad.eval(series))
    out["fp1d/varf_x4"] = dense(quad.varf('x**4/4 - x**2/2', degree).matrix)
    out["fp1d/integrate_flat"] = np.array([quad.integrate(sym.exp(-beta * (x ** 4 / 4 - x ** 2 / 2)), flat=True)])
    # a quadrature with a weight factor: transform divides by it, eval multiplies by it
    quad_w = quad_gh(hm, [120], mean=[.3], cov=[[.7]], factor=sym.exp(-x * x / 5))
    s = quad_w.transform('exp(-x*x/3)*(1 + x)', 40)
    out["fp1d/weighted_transform"] = s.coeffs
    out["fp1d/weighted_eval"] = wl2(quad_w, quad_w.eval(s))
    other = quad_gh(hm, [90], mean=[-1.], cov=[[.1]])
    out["fp1d/eval_other_grid"] = wl2(other, other.eval(quad_gh(hm, [100], mean=[2.], cov=[[2.]]).transform('x', 50)))


def case_2d(hm, out):
    """2-D transforms / varf / varfd / discretize_op on every index set, tensorize on and off
    (tests/test_quad.py:111-192, :234-330; tests/test_tensorize_project.py:184-199)."""
    x, y, f = sym.symbols('x y', real=True) + (sym.Function('f'),)
    quad = quad_gh(hm, [44, 44], mean=[.2, -.1], cov=[[.8, 0], [0, 1.3]])
    bump = 'exp(-(x*x + x*y + y*y)/10)'                         # not separable
    split = x * x * sym.cos(x) + sym.exp(y / 3) * x + sym.sqrt(2) + 2    # a sum of separable terms
    poly = 'x**4/4 - x**2/2 + y**4/4 - y**2/2 + x*y/2'
    for index_set in ("triangle", "cube", "cross", "cross_nc", "rectangle"):
        degree = 12
        s = quad.transform(bump, degree, index_set=index_set)
        out["2d/%s/transform" % index_set] = s.coeffs
        out["2d/%s/eval" % index_set] = wl2(quad, quad.eval(s))
        out["2d/%s/varf_poly" % index_set] = dense(quad.varf(poly, degree, index_set=index_set).matrix)
    for tensorize in (False, True):
        tag = "tensorized" if tensorize else "direct"
        out["2d/transform_split_" + tag] = quad.transform(split, 10, tensorize=tensorize).coeffs
        out["2d/varf_split_" + tag] = dense(quad.varf(split, 10, tensorize=tensorize).matrix)
        out["2d/varfd_split_" + tag] = dense(quad.varfd(split, 10, [0, 1, 1], tensorize=tensorize).matrix)
        out["2d/integrate_split_" + tag] = np.array([quad.integrate(split, tensorize=tensorize)])
    out["2d/varf_sparse"] = dense(quad.varf(poly, 10, sparse=True, index_set="cube").matrix)
    # a McKean-Vlasov-type generator (tests/test_equations.py:202-424), coefficients polynomial
    op = sym.diff(f(x, y), x, x) / 2 + sym.diff((x ** 3 - x + y / 2) * f(x, y), x) \
        + sym.diff(f(x, y), y, y) + sym.diff(y * f(x, y), y)
    out["2d/operator"] = dense(quad.discretize_op(op, 10, index_set="triangle").matrix)
    out["2d/operator_cube_sparse"] = dense(quad.discretize_op(op, 8, index_set="cube", sparse=True).matrix)


def case_integrate(hm, out):
    """Moments under a general Gaussian, dimension 3 (tests/test_quad.py:37-108)."""
    rng = np.random.default_rng(11)
    a = rng.random((3, 3))
    mean, cov = rng.random(3), a.T @ a
    quad = quad_gh(hm, [8, 8, 8], mean=mean, cov=cov)
    vals = [quad.integrate('1')]
    for i in range(3):
        vals.append(quad.integrate('x{}'.format(i)))
        for j in range(3):
            vals.append(quad.integrate('x{}*x{}'.format(i, j)))
    out["integrate/moments"] = np.array(vals)
    out["integrate/mean_cov"] = np.concatenate([mean, cov.ravel()])
    q1 = quad_gh(hm, [50, 50], mean=[.5, 1.], cov=[[3, 1], [1, 6]])
    out["integrate/flat"] = np.array([q1.integrate(quad_gh(hm, [50, 50]).position.weight(), flat=True)])


def case_series_algebra(hm, out):
    """Series products / projections and Varf tensorisation through the operators of the object layer
    (tests/test_series.py:49-89, tests/test_tensorize_project.py:34-158)."""
    degree = 9
    qx, qy = quad_gh(hm, [40], dirs=[0]), quad_gh(hm, [40], dirs=[1])
    qxy = qx * qy
    for index_set in ("triangle", "cube", "cross"):
        sx = qx.transform('exp(-x*x/9)*(1+x)', degree, index_set=index_set)
        sy = qy.transform('cos(y) + y*y', degree, index_set=index_set)
        sxy = qxy.transform('exp(-(x*x+y*y)/9)*(1 + x*y)', degree, index_set=index_set, tensorize=False)
        out["series/%s/outer" % index_set] = (sx * sy).coeffs
        out["series/%s/contract" % index_set] = (sxy * sy).coeffs
        out["series/%s/project0" % index_set] = sxy.project(0).coeffs
        out["series/%s/project1" % index_set] = sxy.project(1).coeffs
        vx = qx.varf('1 + x*x', degree, index_set=index_set)
        vy = qy.varf('y', degree, index_set=index_set)
        out["series/%s/varf_outer" % index_set] = dense((vx * vy).matrix)
        out["series/%s/varf_apply" % index_set] = (vx * vy)(sxy).coeffs
    tri = qxy.transform('exp(-(x*x+y*y)/9)*(1 + x*y)', degree, tensorize=False)
    out["series/to_cross"] = tri.to_cross(4).coeffs
    out["series/subdegree"] = tri.subdegree(5).coeffs
    q3 = quad_gh(hm, [12, 12, 12])
    s3 = q3.transform('exp(-(x*x + y*y + z*z)/8)*(x + y*z)', 6, tensorize=False)
    out["series/3d_project_02"] = s3.project([0, 2]).coeffs
    v3 = q3.varf('x*y + z', 5, tensorize=False)
    # (the reference's Varf.project reads a non-existent attribute, varf.py:139: go through core.project)
    out["series/3d_varf_project_12"] = dense(hm.core.project(v3.matrix, 3, [1, 2], index_set="triangle"))


def case_fourier(hm, out):
    """Fourier x Hermite quadrature (examples/langevin_2d.py; transform.cpp:90-109, varf.cpp:99-162)."""
    q = hm.Quad.fourier(32, dirs=[0]) * quad_gh(hm, [32], dirs=[1])     # the reference needs equal point counts
    s = q.transform('cos(x)*y + sin(2*x) + 1', 8, index_set="cube", tensorize=False)
    out["fourier/transform"] = s.coeffs
    out["fourier/eval"] = wl2(q, q.eval(s))
    out["fourier/varf"] = dense(q.varf('cos(x) + y*y', 6, index_set="cube", tensorize=False).matrix)
    out["fourier/varfd"] = dense(q.varfd('1 + y', 6, [0, 1], index_set="cube", tensorize=False).matrix)
    out["fourier/integrate"] = np.array([q.integrate('cos(x)**2 * y*y', tensorize=False)])


CASES = [case_fokker_planck_1d, case_2d, case_integrate, case_series_algebra, case_fourier]


def run(hm, only=None):
    saved = dict(hm.settings)
    hm.settings.update({'cache': False, 'sparse': False, 'tensorize': True, 'debug': False, 'trails': False})
    out = {}
    try:
        for case in CASES:
            if only is None or case.__name__ in only:
                case(hm, out)
    finally:
        hm.settings.update(saved)
    return out
