"""Resident-data API: a Plan owns the device tables for one (grid, degree, index set) and works on
torch CUDA tensors in place (torch is used only for device memory and streams).  Thin wrapper over
the hb_plan_* entry points of the C ABI."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import c_dp, c_ip, check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dgemm(M, N, K, A, lda, B, ldb, Cout, ldc, batch=1, strideA=0, strideB=0, strideC=0, a_mmajor=False):
    """C[b] = A[b] @ B[b] on device pointers (ints or tensors) through the DMMA/TMA kernel."""
    ptr = lambda t: t.data_ptr() if hasattr(t, "data_ptr") else int(t)
    check(lib().hb_dev_dgemm(M, N, K, batch, ptr(A), lda, strideA, int(a_mmajor), ptr(B), ldb, strideB, ptr(Cout), ldc,
                             strideC, _stream()))


def dgemm_scatter(M, N, K, A, lda, B, ldb, blocks, block_rows, ldc, batch=1, strideA=0, strideB=0, strideC=0,
                  a_mmajor=False):
    """Like dgemm, but rows [i*block_rows, (i+1)*block_rows) of the result go to blocks[i] (device pointers or
    tensors, possibly peer-GPU memory)."""
    ptr = lambda t: t.data_ptr() if hasattr(t, "data_ptr") else int(t)
    arr = (C.c_void_p * len(blocks))(*[ptr(b) for b in blocks])
    check(lib().hb_dev_dgemm_scatter(M, N, K, batch, ptr(A), lda, strideA, int(a_mmajor), ptr(B), ldb, strideB, arr,
                                     len(blocks), block_rows, ldc, strideC, _stream()))


def varfd_device(dim, degree, direction, var, do_fourier=0, index_set="triangle", out=None):
    """core.varfd on a resident row block (`var`: rows x n_polys float64 CUDA tensor); returns a new block."""
    if out is None:
        out = torch.empty_like(var)
    check(lib().hb_dev_varfd(int(dim), int(degree), int(direction), var.data_ptr(), var.stride(0), var.shape[0],
                             var.shape[1], int(do_fourier), index_set.encode(), out.data_ptr(), out.stride(0),
                             _stream()))
    return out


class Plan:
    def __init__(self, degree, nodes, weights, do_fourier=None, index_set="triangle", sqrt_weights=None):
        """sqrt_weights: optional list (one entry per axis, None to skip an axis) of sqrt(w_i) in extended range for
        the Gram-form varf at high degree (hermipy_b200.lib.hermegauss(n, sqrt_weights=True) supplies them): weights
        below 1e-308 underflow to 0 in double precision while sqrt(w) phi_k is still O(1) there."""
        nodes = [np.ascontiguousarray(n, dtype=np.float64).ravel() for n in nodes]
        weights = [np.ascontiguousarray(w, dtype=np.float64).ravel() for w in weights]
        self.dim = len(nodes)
        self.n_pts = [len(n) for n in nodes]
        self.degree = int(degree)
        self.index_set = index_set
        n_pts = np.ascontiguousarray(self.n_pts, dtype=np.int32)
        nd, wt = np.concatenate(nodes), np.concatenate(weights)
        four = np.ascontiguousarray([0] * self.dim if do_fourier is None else do_fourier, dtype=np.int32)
        self._h = C.c_void_p()
        torch.cuda.current_device()  # make sure the CUDA context of this process exists
        check(lib().hb_plan_create(C.byref(self._h), self.dim, n_pts.ctypes.data_as(c_ip), self.degree,
                                   nd.ctypes.data_as(c_dp), wt.ctypes.data_as(c_dp), four.ctypes.data_as(c_ip),
                                   index_set.encode()))
        self.n_polys = int(lib().hb_plan_n_polys(self._h))
        self.grid_pitch = int(lib().hb_plan_grid_pitch(self._h))
        self.grid_size = int(lib().hb_plan_grid_size(self._h))
        self.coef_stride = (self.n_polys + 1) & ~1
        for a, sw in enumerate(sqrt_weights or []):
            if sw is not None:
                sw = np.ascontiguousarray(sw, dtype=np.float64).ravel()
                if len(sw) != self.n_pts[a]:
                    raise ValueError("sqrt_weights[%d] has %d entries, the axis has %d nodes" % (a, len(sw), self.n_pts[a]))
                check(lib().hb_plan_set_sqrt_weights(self._h, a, sw.ctypes.data_as(c_dp)))

    def __del__(self):
        # at interpreter shutdown the module globals may already be gone: then the process-exit cleanup of the
        # library frees the plan
        try:
            if getattr(self, "_h", None) and _lib is not None and _lib._lib is not None:
                _lib._lib.hb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # --- layout helpers -------------------------------------------------------------------------
    def pack_grid(self, f):
        """numpy (batch, n_0, ..., n_{d-1}) -> CUDA tensor (batch, grid_size) in the padded device layout."""
        f = np.asarray(f, dtype=np.float64).reshape((-1,) + tuple(self.n_pts))
        pad = self.grid_pitch - self.n_pts[-1]
        if pad:
            f = np.concatenate([f, np.zeros(f.shape[:-1] + (pad,))], axis=-1)
        return torch.from_numpy(np.ascontiguousarray(f.reshape(f.shape[0], -1))).cuda()

    def unpack_grid(self, t):
        a = t.cpu().numpy().reshape((-1,) + tuple(self.n_pts[:-1]) + (self.grid_pitch,))
        return a[..., :self.n_pts[-1]]

    def empty_grid(self, batch):
        return torch.zeros((batch, self.grid_size), dtype=torch.float64, device="cuda")

    def empty_coef(self, batch):
        return torch.zeros((batch, self.coef_stride), dtype=torch.float64, device="cuda")

    # --- compute --------------------------------------------------------------------------------
    def forward(self, f, out=None):
        batch = f.shape[0]
        out = self.empty_coef(batch) if out is None else out
        check(lib().hb_plan_forward(self._h, f.data_ptr(), f.stride(0), batch, out.data_ptr(), out.stride(0),
                                    _stream()))
        return out

    def backward(self, coef, out=None):
        batch = coef.shape[0]
        out = self.empty_grid(batch) if out is None else out
        check(lib().hb_plan_backward(self._h, coef.data_ptr(), coef.stride(0), batch, out.data_ptr(),
                                     out.stride(0), _stream()))
        return out

    def varf(self, f, mode="gram", out=None, row_begin=0, row_end=None):
        """f: one grid function in device layout (grid_size,).  Returns rows [row_begin,row_end) of the matrix."""
        row_end = self.n_polys if row_end is None else row_end
        if out is None:
            out = torch.empty((row_end - row_begin, self.n_polys), dtype=torch.float64, device="cuda")
        check(lib().hb_plan_varf(self._h, f.data_ptr(), {"reference": 0, "gram": 1}[mode], out.data_ptr(),
                                 out.stride(0), row_begin, row_end, _stream()))
        return out

    def varf_sparse(self, f, row_begin=0, row_end=None, fetch=True):
        """Rows [row_begin,row_end) of the reference-algebra varf of a polynomial-like f as a scipy CSR matrix (rows
        relative to row_begin), assembled from the band of retained triple products: no n_polys^2 buffer.  Row sharding
        of a sparse matrix = one call per rank with its own range.  fetch=False returns only nnz (timing)."""
        import scipy.sparse as ss
        row_end = self.n_polys if row_end is None else row_end
        nnz = C.c_longlong(0)
        check(lib().hb_plan_varf_sp(self._h, f.data_ptr(), 0, row_begin, row_end, C.byref(nnz), _stream()))
        if not fetch:
            return nnz.value
        rows = row_end - row_begin
        indptr = np.zeros(rows + 1, dtype=np.int64)
        indices = np.zeros(max(nnz.value, 1), dtype=np.int32)
        data = np.zeros(max(nnz.value, 1))
        check(lib().hb_sparse_fetch(indptr.ctypes.data_as(_lib.c_llp), indices.ctypes.data_as(c_ip), data.ctypes.data_as(c_dp)))
        return ss.csr_matrix((data[:nnz.value], indices[:nnz.value], indptr), shape=(rows, self.n_polys))

    def discretize(self, body, translation=None, dilation=None, out=None):
        """Evaluate the C expression `body` (in v[i]) on this plan's grid, directly in device layout."""
        out = self.empty_grid(1)[0] if out is None else out
        tr = None if translation is None else np.ascontiguousarray(translation, dtype=np.float64)
        di = None if dilation is None else np.ascontiguousarray(dilation, dtype=np.float64)
        check(lib().hb_plan_discretize(self._h, str(body).encode(),
                                       None if tr is None else tr.ctypes.data_as(c_dp),
                                       None if di is None else di.ctypes.data_as(c_dp), out.data_ptr(), _stream()))
        return out

    def set_timing(self, on=True):
        check(lib().hb_plan_set_timing(self._h, int(on)))

    def timings(self):
        """[(tag, ms, padded_flops)] for every GEMM launch of the last call (needs set_timing(True))."""
        out = []
        for i in range(lib().hb_plan_timing_count(self._h)):
            tag, ms, fl = C.c_int(0), C.c_double(0), C.c_double(0)
            check(lib().hb_plan_timing_get(self._h, i, C.byref(tag), C.byref(ms), C.byref(fl)))
            out.append((tag.value, ms.value, fl.value))
        return out

    def table(self, axis, kind):
        """(device pointer, rows, cols, ld) of a basis table; kind: 0 phi[i][k], 1 w*phi[i][k], 2 phi[k][i], 3 w*phi[k][i]."""
        ptr, r, c, ld = C.c_void_p(), C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        check(lib().hb_plan_table(self._h, axis, kind, C.byref(ptr), C.byref(r), C.byref(c), C.byref(ld)))
        return ptr.value, r.value, c.value, ld.value

    def varf_info(self):
        k, p = C.c_int(0), C.c_int(0)
        check(lib().hb_plan_varf_info(self._h, C.byref(k), C.byref(p)))
        return k.value, p.value
