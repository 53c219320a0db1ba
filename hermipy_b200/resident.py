"""Assembled Galerkin matrices that stay in HBM, and their consumers.

Reference: hermipy/varf.py:127-134 (`Varf.__call__`), :170-226 (`solve`) and :228-237 (`eigs`).  There the
matrix is a NumPy / SciPy object; at the sizes this library assembles (C2 cube: 13 GB, C3 cube: 207 GB
row-sharded over 8 GPUs) a host copy costs more than the assembly, so `Quad.varf(..., resident=True)`,
`Quad.varfd(..., resident=True)` and `Quad.discretize_op(..., resident=True)` return a `ResidentVarf`: the
same object surface (`__call__`, `+ - *`, `solve`, `eigs`, `multi_indices`), with the matrix a row block per
rank on the device.  `solve` runs a device-side restarted GMRES (`device_gmres`: matrix, Krylov basis and vectors in
HBM, the product A x is `hb_dev_matvec`, csrc/hb_matvec.cu, HBM-bound; per step one Hessenberg column crosses PCIe);
`eigs` without shift runs a device-side Krylov-Schur iteration (`device_eigs`: thick-restarted Arnoldi, basis and
products in HBM, the ncv x ncv projected problem on the host with LAPACK); with `sigma` (shift-invert) it is SciPy's
ARPACK, as in the reference, on a device LU.  Dense direct solves run on the device through torch.linalg (cuSOLVER), which is library code
beside the path, not part of it.
"""
import ctypes as C

import numpy as np
import numpy.linalg as la
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist

from . import core
from . import function as func
from . import series as hs
from . import stats
from ._lib import check, lib

_NUMBER = (int, float, np.integer, np.floating)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_matvec(A, x, out=None):
    """out = A @ x for a row-major float64 CUDA matrix (any leading dimension) through hb_dev_matvec."""
    if A.dtype != torch.float64 or x.dtype != torch.float64 or A.stride(1) != 1 or not x.is_contiguous():
        raise ValueError("device_matvec: float64 operands, matrix rows contiguous")
    if x.numel() != A.shape[1]:
        raise ValueError("device_matvec: vector has %d entries, matrix has %d columns" % (x.numel(), A.shape[1]))
    if out is None:
        out = torch.empty(A.shape[0], dtype=torch.float64, device=A.device)
    check(lib().hb_dev_matvec(A.shape[0], A.shape[1], A.data_ptr(), A.stride(0), x.data_ptr(), out.data_ptr(),
                              _stream()))
    return out


def device_gmres(matvec, b, x0=None, rtol=1e-5, atol=0.0, restart=20, maxiter=None):
    """Restarted GMRES with every vector resident in HBM (same stopping rule and defaults as
    scipy.sparse.linalg.gmres: ||b - A x|| <= max(rtol ||b||, atol), restart 20, at most `maxiter` restart cycles).
    `matvec`: device vector -> device vector (hb_dev_matvec for a resident matrix).  The Krylov basis lives in one
    device array; orthogonalisation is classical Gram-Schmidt applied twice (two matrix-vector products with the basis
    per step, no loop over the basis vectors); per step only the j + 2 Hessenberg entries cross PCIe -- the small
    least-squares problem is solved on the host.  Returns (x on the device, info): info = 0 converged, > 0 the number
    of inner iterations spent without reaching the tolerance."""
    n = b.numel()
    dev, dt = b.device, b.dtype
    x = torch.zeros(n, dtype=dt, device=dev) if x0 is None else x0.clone()
    maxiter = 10 * n if maxiter is None else int(maxiter)
    m = max(1, min(int(restart), n))
    target = max(float(rtol) * float(torch.linalg.vector_norm(b)), float(atol))
    V = torch.empty((m + 1, n), dtype=dt, device=dev)
    spent = 0
    for _ in range(maxiter):
        r = b - matvec(x) if (x0 is not None or spent) else b.clone()
        beta = float(torch.linalg.vector_norm(r))
        if beta <= target:
            return x, 0
        V[0] = r / beta
        H = np.zeros((m + 1, m))
        g = np.zeros(m + 1)
        g[0] = beta
        k_used, y = 0, None
        for j in range(m):
            w = matvec(V[j])
            Vj = V[:j + 1]
            h = Vj @ w                         # classical Gram-Schmidt, twice
            w = w - Vj.t() @ h
            h2 = Vj @ w
            w = w - Vj.t() @ h2
            hn = torch.linalg.vector_norm(w)
            col = torch.cat((h + h2, hn.reshape(1))).cpu().numpy()
            H[:j + 2, j] = col
            spent += 1
            k_used = j + 1
            # least squares of the small Hessenberg system (host, (j+2) x (j+1))
            y, _, _, _ = np.linalg.lstsq(H[:j + 2, :j + 1], g[:j + 2], rcond=None)
            res = float(np.linalg.norm(g[:j + 2] - H[:j + 2, :j + 1] @ y))
            if col[-1] > 0:
                V[j + 1] = w / hn
            if res <= target or col[-1] == 0:
                break
        x = x + V[:k_used].t() @ torch.from_numpy(y).to(dev)
        if res <= target:
            # the recurrence residual can drift from the true one: confirm it
            if float(torch.linalg.vector_norm(b - matvec(x))) <= target * (1 + 1e-8) + 1e-300:
                return x, 0
    return x, spent


def _wanted_order(theta, which):
    """indices of Ritz values, most wanted first, for ARPACK's `which` codes (general, non-symmetric problem)"""
    key = {"LM": -np.abs(theta), "SM": np.abs(theta), "LR": -theta.real, "SR": theta.real,
           "LI": -np.abs(theta.imag), "SI": np.abs(theta.imag)}
    if which not in key:
        raise ValueError("eigs: which must be one of %s" % sorted(key))
    return np.argsort(key[which], kind="stable")


def device_eigs(matvec, n, k=6, which="LM", ncv=None, tol=0.0, maxiter=None, v0=None, device=None, dtype=torch.float64):
    """k eigenpairs of a real non-symmetric operator by the Krylov-Schur method (Stewart 2001: Arnoldi with thick
    restarts through a reordered real Schur form) -- the algorithm family ARPACK's implicitly restarted Arnoldi belongs
    to, with the same meaning of k / which / ncv / tol / maxiter / v0 as scipy.sparse.linalg.eigs (standard problem, no
    shift).  The Krylov basis (ncv + 1 vectors) and every operator product stay in HBM; the small projected matrix
    (ncv x ncv) is handled on the host with LAPACK, and per Arnoldi step one column of it crosses PCIe.
    Returns (eigenvalues (k,), eigenvectors as a device tensor (n, k) -- complex when a conjugate pair is among them)."""
    import scipy.linalg as sla
    if not 0 < k < n - 1:
        raise ValueError("eigs: need 0 < k < n - 1")
    m = min(n, max(2 * k + 1, 20)) if ncv is None else int(ncv)
    if not k + 1 < m <= n:
        raise ValueError("eigs: need k + 1 < ncv <= n")
    maxiter = 10 * n if maxiter is None else int(maxiter)
    eps = float(np.finfo(np.float64).eps)
    tol = eps if not tol else float(tol)
    dev = device if device is not None else (v0.device if torch.is_tensor(v0) else "cpu")
    V = torch.zeros((m + 1, n), dtype=dtype, device=dev)
    if v0 is None:
        g = torch.Generator(device="cpu").manual_seed(0)
        v = torch.rand(n, generator=g, dtype=dtype).to(dev) - 0.5
    else:
        v = (v0 if torch.is_tensor(v0) else torch.from_numpy(np.ascontiguousarray(v0, dtype=np.float64))).to(dev)
    V[0] = v / torch.linalg.vector_norm(v)
    H = np.zeros((m + 1, m))
    p = 0                      # columns of the current decomposition that are in (reordered) Schur form

    def expand(j0):
        """Arnoldi steps j0 .. m-1: A V_m = V_m H_m + H[m, m-1] v_m e^T (classical Gram-Schmidt, twice)"""
        for j in range(j0, m):
            w = matvec(V[j])
            Vj = V[:j + 1]
            h = Vj @ w
            w = w - Vj.t() @ h
            h2 = Vj @ w
            w = w - Vj.t() @ h2
            hn = torch.linalg.vector_norm(w)
            col = torch.cat((h + h2, hn.reshape(1))).cpu().numpy()
            H[:j + 2, j] = col
            if col[-1] <= eps * max(1.0, np.abs(col[:-1]).max()):      # invariant subspace: restart direction
                g2 = torch.Generator(device="cpu").manual_seed(j + 1)
                w = torch.rand(n, generator=g2, dtype=dtype).to(dev) - 0.5
                w = w - Vj.t() @ (Vj @ w)
                hn = torch.linalg.vector_norm(w)
                H[j + 1, j] = 0.0
            V[j + 1] = w / hn

    expand(0)
    theta, order, Q, T = None, None, None, None
    for _ in range(maxiter):
        Hm = H[:m, :m]
        beta_row = H[m, :m].copy()             # the residual row: zero except the last entry right after expand()
        T, Q = sla.schur(Hm, output="real")
        # eigenvalue of every diagonal position of the quasi-triangular T (2 x 2 blocks: a conjugate pair)
        theta = np.zeros(m, dtype=complex)
        i = 0
        while i < m:
            if i + 1 < m and T[i + 1, i] != 0.0:
                pair = np.linalg.eigvals(T[i:i + 2, i:i + 2])
                theta[i], theta[i + 1] = pair[0], pair[1]
                i += 2
            else:
                theta[i] = T[i, i]
                i += 1
        order = _wanted_order(theta, which)
        # keep the k wanted values plus a share of the rest (thick restart), never cutting a conjugate pair
        keep = min(m - 2, k + max(1, (m - k) // 2))
        sel = np.zeros(m, dtype=bool)
        sel[order[:keep]] = True
        for i in range(m - 1):
            if T[i + 1, i] != 0.0 and (sel[i] or sel[i + 1]):
                sel[i] = sel[i + 1] = True
        # move the selected eigenvalues to the leading block (LAPACK dtrsen: orthogonal reordering of the Schur form)
        T, Q, _wr, _wi, pk, _s, _sep, info = sla.lapack.dtrsen(sel.astype(np.int32), T, Q, job="N")
        if info != 0:
            raise RuntimeError("device_eigs: reordering of the Schur form failed (dtrsen info = %d)" % info)
        pk = int(pk)
        b = beta_row @ Q                        # residual row in the Schur basis
        # Ritz residuals of the wanted pairs: |b . y| for the eigenvectors y of the leading block
        wv, Y = np.linalg.eig(T[:pk, :pk])
        resid = np.abs(b[:pk] @ Y)
        idx = _wanted_order(wv, which)[:k]
        # ARPACK's test, |residual| <= tol max(eps^(2/3), |theta|), with a floor at the rounding level of the projected
        # problem (residuals of converged pairs stagnate at ~eps ||A||: a pair cannot be resolved below that)
        floor = 20.0 * eps * max(1.0, float(np.abs(theta).max()))
        done = np.all(resid[idx] <= np.maximum(tol * np.maximum(eps ** (2.0 / 3.0), np.abs(wv[idx])), floor))
        if done:
            Qk = torch.from_numpy(np.ascontiguousarray(Q[:, :pk])).to(dev)
            basis = V[:m].t() @ Qk              # (n, pk) Schur vectors
            vals, Yk = wv[idx], Y[:, idx]
            if np.all(np.abs(vals.imag) <= 0) and np.all(np.abs(Yk.imag) <= 0):
                vecs = basis @ torch.from_numpy(np.ascontiguousarray(Yk.real)).to(dev)
                return vals.real.astype(np.float64), vecs
            Yc = torch.from_numpy(np.ascontiguousarray(Yk)).to(dev)
            return vals, basis.to(Yc.dtype) @ Yc
        # thick restart: V <- [V Q[:, :pk], v_{m+1}], H <- [[T_pk], [b_pk]]
        Qk = torch.from_numpy(np.ascontiguousarray(Q[:, :pk])).to(dev)
        Vnew = (V[:m].t() @ Qk).t().contiguous()
        V[pk] = V[m]
        V[:pk] = Vnew
        H[:] = 0.0
        H[:pk, :pk] = T[:pk, :pk]
        H[pk, :pk] = b[:pk]
        p = pk
        expand(p)
    raise RuntimeError("device_eigs: no convergence in %d restarts (k = %d, ncv = %d)" % (maxiter, k, m))


class ResidentMatrix:
    """Rows [row_begin, row_begin + block.shape[0]) of an n x n matrix; `block` is a float64 tensor on the
    compute device.  With a process group the row blocks of all ranks (contiguous, in rank order) form the
    matrix and a product ends in one all-gather of the result vector -- the only exchange this path has."""

    def __init__(self, block, n=None, row_begin=0, group=None, matvec=device_matvec):
        self.block = block
        self.n = int(block.shape[1] if n is None else n)
        self.row_begin = int(row_begin)
        self.group = group
        self._matvec = matvec
        self.sharded = group is not None and dist.get_world_size(group) > 1
        if not self.sharded and block.shape[0] != self.n:
            raise ValueError("ResidentMatrix: a row block needs a process group")
        if self.sharded:
            rows = torch.tensor([block.shape[0]], dtype=torch.int64, device=block.device)
            counts = [torch.zeros_like(rows) for _ in range(dist.get_world_size(group))]
            dist.all_gather(counts, rows, group=group)
            self.row_counts = [int(c.item()) for c in counts]
            if sum(self.row_counts) != self.n:
                raise ValueError("ResidentMatrix: row blocks hold %d rows, matrix has %d" % (sum(self.row_counts), self.n))

    shape = property(lambda self: (self.n, self.n))

    def _like(self, block):
        out = ResidentMatrix.__new__(ResidentMatrix)
        out.__dict__.update(self.__dict__)
        out.block = block
        return out

    # ---- products ------------------------------------------------------------------------------
    def matvec_device(self, x):
        """x: full vector on the device -> full result vector on the device (every rank)."""
        y = self._matvec(self.block, x)
        if not self.sharded:
            return y
        return self._gather_rows(y)

    def _gather_rows(self, local):
        """Concatenate the ranks' row blocks (vector entries or matrix rows); ragged blocks are padded to the
        tallest one for the collective, which wants equal contributions."""
        world, tall = len(self.row_counts), max(self.row_counts)
        tail = tuple(local.shape[1:])
        if local.shape[0] != tall:
            padded = torch.zeros((tall,) + tail, dtype=local.dtype, device=local.device)
            padded[:local.shape[0]] = local
            local = padded
        full = torch.empty((world * tall,) + tail, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local.contiguous(), group=self.group)
        if len(set(self.row_counts)) == 1:
            return full
        return torch.cat([full[r * tall:r * tall + c] for r, c in enumerate(self.row_counts)])

    def matvec(self, x):
        """Host vector in, host vector out (complex input: real and imaginary parts separately)."""
        x = np.asarray(x).reshape(-1)
        if np.iscomplexobj(x):
            return self.matvec(x.real) + 1j * self.matvec(x.imag)
        xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.block.device)
        return self.matvec_device(xd).cpu().numpy()

    def linear_operator(self):
        return spla.LinearOperator(self.shape, matvec=self.matvec, dtype=np.float64)

    # ---- element-wise algebra (device, library kernels) --------------------------------------
    def __add__(self, other):
        if not isinstance(other, ResidentMatrix) or other.shape != self.shape or other.row_begin != self.row_begin:
            raise ValueError("Invalid argument!")
        return self._like(self.block + other.block)

    def __mul__(self, number):
        return self._like(self.block * float(number))

    def diagonal_shift(self, sigma):
        """A - sigma I on this rank's rows (new buffer)."""
        out = self.block.clone()
        rows = out.shape[0]
        idx = torch.arange(rows, device=out.device)
        out[idx, idx + self.row_begin] -= float(sigma)
        return self._like(out)

    def to_host(self):
        """The full matrix as an ndarray (gathers the row blocks; meant for sizes the host can hold)."""
        if not self.sharded:
            return self.block.cpu().numpy()
        return self._gather_rows(self.block).cpu().numpy()


class ResidentVarf:
    """`Varf` whose matrix is a `ResidentMatrix` (see the module docstring)."""

    def __init__(self, matrix, position, factor=1, index_set="triangle"):
        if not isinstance(matrix, ResidentMatrix):
            matrix = ResidentMatrix(matrix)
        self.is_sparse = False
        self.matrix = matrix
        self.position = position
        self.index_set = index_set
        self.degree = core.iterator_get_degree(position.dim, matrix.n, index_set=index_set)
        self.factor = func.Function(factor, dirs=position.dirs)

    def _like(self, matrix):
        return ResidentVarf(matrix, self.position, factor=self.factor, index_set=self.index_set)

    def _compatible(self, other):
        return isinstance(other, ResidentVarf) and self.position == other.position and \
            self.index_set == other.index_set and self.factor == other.factor

    def __add__(self, other):
        if isinstance(other, _NUMBER):                   # element-wise, as ndarray + number (0: sum() start)
            return self if other == 0 else self._like(self.matrix._like(self.matrix.block + float(other)))
        if not self._compatible(other):
            raise ValueError("Invalid argument!")
        return self._like(self.matrix + other.matrix)

    __radd__ = __add__

    def __mul__(self, other):
        if not isinstance(other, _NUMBER):
            raise ValueError("Invalid argument: {}".format(other))
        return self._like(self.matrix * other)

    __rmul__ = __mul__

    def __sub__(self, other):
        return self + other * (-1)

    def __neg__(self):
        return self * (-1)

    def __call__(self, series):
        if not isinstance(series, hs.Series) or not self.position == series.position or \
                self.index_set != series.index_set:
            raise ValueError("Invalid argument!")
        return hs.Series(self.matrix.matvec(series.coeffs), self.position, factor=self.factor,
                         index_set=self.index_set)

    def multi_indices(self):
        return core.iterator_list_indices(self.position.dim, self.degree, index_set=self.index_set)

    def to_varf(self):
        """Host copy as a plain `Varf`."""
        from .varf import Varf
        return Varf(self.matrix.to_host(), self.position, factor=self.factor, index_set=self.index_set)

    # ---- solvers ---------------------------------------------------------------------------------
    def _bordered_operator(self, border):
        """[[A, b], [b^T, 0]] as a LinearOperator (varf.py:200-206 builds this matrix explicitly)."""
        n = self.matrix.n

        def apply(v):
            v = np.asarray(v).reshape(-1)
            top = self.matrix.matvec(v[:n]) + border * v[n]
            return np.append(top, border @ v[:n])
        return spla.LinearOperator((n + 1, n + 1), matvec=apply, dtype=np.float64)

    def solve(self, series, use_gmres=True, remove0=False, remove_vec=None, preconditioner=None, quiet=False,
              **kwargs):
        """Solve A u = series (varf.py:170-226).  GMRES (the default here: the matrix never leaves the device
        and is not copied) drives `hb_dev_matvec`; `use_gmres=False` factorises a device copy with
        torch.linalg.solve.  `remove0` / `remove_vec` border the system exactly as the reference does."""
        if not self.position == series.position or self.index_set != series.index_set:
            raise ValueError("Invalid argument: position and index set must coincide!")
        if preconditioner is not None:
            raise ValueError("ResidentVarf.solve: no preconditioner is available for a resident dense matrix")
        bordered = remove0 or remove_vec is not None
        border = None
        if bordered:
            if remove0:
                [e], [remove_vec] = self.eigs(k=1, which='SM')
                if not quiet:
                    print("Removing eigenspace with eigenvalue {}.".format(abs(e)))
            else:
                remove_vec = remove_vec / np.sqrt(float(remove_vec * remove_vec))
                e = float(remove_vec * self(remove_vec))
                if not quiet:
                    print("Removing vector with value {}.".format(abs(e)))
            if abs(e) > 0.01:
                print("[Hermipy:varf:solve warning] Value of e: {}".format(e))
            border = np.real(np.asarray(remove_vec.coeffs)).astype(np.float64)
        vector = np.append(series.coeffs, 0) if bordered else np.asarray(series.coeffs, dtype=np.float64)

        if use_gmres:
            # device-side GMRES: the matrix, the Krylov basis and every vector stay in HBM; per step only the
            # Hessenberg column crosses PCIe (the reference hands the host matrix to SciPy's GMRES, varf.py:212-219)
            dev = self.matrix.block.device
            n = self.matrix.n
            rhs = torch.from_numpy(np.ascontiguousarray(vector, dtype=np.float64)).to(dev)
            if bordered:
                bd = torch.from_numpy(border).to(dev)

                def product(v):
                    top = self.matrix.matvec_device(v[:n].contiguous()) + bd * v[n]
                    return torch.cat((top, (bd @ v[:n]).reshape(1)))
            else:
                def product(v):
                    return self.matrix.matvec_device(v.contiguous())
            known = {k: kwargs.pop(k) for k in ("rtol", "atol", "restart", "maxiter") if k in kwargs}
            if "tol" in kwargs:                       # older SciPy spelling
                known.setdefault("rtol", kwargs.pop("tol"))
            x0 = kwargs.pop("x0", None)
            if x0 is not None:
                x0 = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float64)).to(dev)
            if kwargs:
                raise ValueError("ResidentVarf.solve: unsupported GMRES options %s" % sorted(kwargs))
            sol_dev, info = device_gmres(product, rhs, x0=x0, **known)
            solution = sol_dev.cpu().numpy()
            if info != 0 and not quiet:
                print("[Hermipy:varf:solve warning] GMRES returned info = {}".format(info))
        else:
            if self.matrix.sharded:
                raise ValueError("ResidentVarf.solve: the direct solve needs the whole matrix on one device")
            A = self.matrix.block
            dev = A.device
            if bordered:
                b = torch.from_numpy(border).to(dev)
                A = torch.cat((torch.cat((A, b.reshape(1, -1))),
                               torch.cat((b, torch.zeros(1, dtype=torch.float64, device=dev))).reshape(-1, 1)), dim=1)
            rhs = torch.from_numpy(np.ascontiguousarray(vector, dtype=np.float64)).to(dev)
            solution = torch.linalg.solve(A, rhs).cpu().numpy()
        if bordered:
            solution = np.array(solution[0:-1])
        return hs.Series(solution, position=self.position, factor=self.factor, index_set=self.index_set)

    @stats.log_stats()
    def eigs(self, **kwargs):
        """Eigenpairs of the resident matrix with the arguments of scipy.sparse.linalg.eigs (varf.py:228-237).
        Standard problem without shift: the device-side Krylov-Schur driver (`driver='arpack'` forces SciPy's ARPACK on
        the device product).  With `sigma` (shift-invert): ARPACK on a device LU of A - sigma I
        (torch.linalg.lu_factor)."""
        sigma = kwargs.get('sigma')
        driver = kwargs.pop('driver', None)
        plain = sigma is None and not (set(kwargs) - {'k', 'which', 'ncv', 'tol', 'maxiter', 'v0'})
        if driver not in (None, 'device', 'arpack'):
            raise ValueError("ResidentVarf.eigs: driver must be 'device' or 'arpack'")
        if driver == 'device' and not plain:
            raise ValueError("ResidentVarf.eigs: the device driver handles the standard problem without shift")
        if plain and driver != 'arpack':
            # device-side Krylov-Schur: basis and products in HBM, the projected ncv x ncv problem on the host
            dev = self.matrix.block.device
            v0 = kwargs.get('v0')
            values, vectors = device_eigs(lambda v: self.matrix.matvec_device(v.contiguous()), self.matrix.n,
                                          k=kwargs.get('k', 6), which=kwargs.get('which', 'LM'), ncv=kwargs.get('ncv'),
                                          tol=kwargs.get('tol', 0.0), maxiter=kwargs.get('maxiter'),
                                          v0=None if v0 is None else torch.from_numpy(np.ascontiguousarray(v0, dtype=np.float64)).to(dev),
                                          device=dev)
            vectors = vectors.cpu().numpy()
            series = []
            for v in vectors.T:
                v = v if la.norm(np.imag(v)) > 1e-15 else np.real(v)
                series.append(hs.Series(v, self.position, factor=self.factor, index_set=self.index_set))
            return values, series
        if sigma is not None and 'OPinv' not in kwargs:
            if self.matrix.sharded or np.iscomplex(sigma):
                raise ValueError("ResidentVarf.eigs: shift-invert needs a real shift and the whole matrix on one device")
            shifted = self.matrix.diagonal_shift(sigma).block
            LU, piv = torch.linalg.lu_factor(shifted)
            del shifted

            def opinv(v):
                v = np.asarray(v).reshape(-1)
                if np.iscomplexobj(v):
                    return opinv(v.real) + 1j * opinv(v.imag)
                rhs = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(LU.device).reshape(-1, 1)
                return torch.linalg.lu_solve(LU, piv, rhs).reshape(-1).cpu().numpy()
            kwargs = dict(kwargs, OPinv=spla.LinearOperator(self.matrix.shape, matvec=opinv, dtype=np.float64))
        values, vectors = spla.eigs(self.matrix.linear_operator(), **kwargs)
        series = []
        for v in vectors.T:
            v = v if la.norm(np.imag(v)) > 1e-15 else np.real(v)
            series.append(hs.Series(v, self.position, factor=self.factor, index_set=self.index_set))
        return values, series
