"""ctypes binding of libhermite_b200.so (C ABI declared in include/hermite_b200.h).

The library is built in-tree by `hermipy_b200/csrc/Makefile` (see __graft_entry__.build()).  There is
no CPU fallback: if the shared library is missing, or no CUDA device is usable, the compute entry
points raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhermite_b200.so")

HB_OK, HB_ERR_INVALID, HB_ERR_INDEX_SET, HB_ERR_FOURIER_DEGREE, HB_ERR_DEGREE, HB_ERR_CUDA, HB_ERR_ALIGN, \
    HB_ERR_NOMEM = range(8)
VARF_REFERENCE, VARF_GRAM = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_up = C.POINTER(C.c_uint32)
c_llp = C.POINTER(C.c_longlong)


_CODE_NAMES = ["HB_OK", "HB_ERR_INVALID", "HB_ERR_INDEX_SET", "HB_ERR_FOURIER_DEGREE", "HB_ERR_DEGREE", "HB_ERR_CUDA",
               "HB_ERR_ALIGN", "HB_ERR_NOMEM"]


class HermiteB200Error(RuntimeError):
    def __init__(self, code, text):
        self.code = code
        name = _CODE_NAMES[code] if 0 <= code < len(_CODE_NAMES) else "?"
        super().__init__("libhermite_b200 status %d (%s): %s" % (code, name, text))


class DegreeError(HermiteB200Error, ValueError):
    """n_polys matches no degree (the reference prints "Can't find x" and exits, lib.cpp:93-99)."""


class GemmDesc(C.Structure):
    """hb_gemm_desc of include/hermite_b200.h"""
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int), ("batch_lo", C.c_int),
                ("a_mmajor", C.c_int),
                ("A", C.c_void_p), ("lda", C.c_longlong), ("strideA", C.c_longlong), ("strideA_lo", C.c_longlong),
                ("B", C.c_void_p), ("ldb", C.c_longlong), ("strideB", C.c_longlong), ("strideB_lo", C.c_longlong),
                ("C", C.c_void_p), ("ldc", C.c_longlong), ("strideC", C.c_longlong), ("strideC_lo", C.c_longlong),
                ("n_blocks", C.c_int), ("block_rows", C.c_int), ("row_offset", C.c_int), ("row_offset_lo", C.c_int),
                ("blocks", C.POINTER(C.c_void_p)),
                ("signal_flags", C.POINTER(C.c_void_p)), ("n_signal", C.c_int), ("signal_state", C.c_void_p),
                ("wait_flags", C.c_void_p), ("n_wait", C.c_int), ("wait_epoch", C.c_void_p)]


_lib = None

# name -> (restype, argtypes); every declaration of include/hermite_b200.h appears here.
SIGNATURES = {
    "hb_last_error": (C.c_char_p, []),
    "hb_version": (C.c_char_p, []),
    "hb_device_count": (C.c_int, [c_ip]),
    "hb_index_size": (C.c_int, [C.c_char_p, C.c_int, C.c_int, c_llp]),
    "hb_index_get_degree": (C.c_int, [C.c_char_p, C.c_int, C.c_longlong, c_ip]),
    "hb_index_list": (C.c_int, [C.c_char_p, C.c_int, C.c_int, c_up, C.c_size_t]),
    "hb_triangle_index": (C.c_int, [c_up, C.c_int, c_up]),
    "hb_cube_index": (C.c_int, [c_up, C.c_int, c_up]),
    "hb_index_positions": (C.c_int, [C.c_char_p, C.c_int, C.c_int, c_up, C.c_size_t, c_ip]),
    "hb_hermite_eval": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp]),
    "hb_discretize": (C.c_int, [C.c_char_p, C.c_int, c_ip, c_dp, c_dp, c_dp, c_dp, C.c_size_t]),
    "hb_plan_discretize": (C.c_int, [C.c_void_p, C.c_char_p, c_dp, c_dp, C.c_void_p, C.c_void_p]),
    "hb_transform": (C.c_int, [C.c_int, c_ip, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_int, C.c_char_p, c_dp,
                               C.c_size_t]),
    "hb_transform_batch": (C.c_int, [C.c_int, c_ip, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_int, C.c_char_p,
                                     c_dp, C.c_size_t]),
    "hb_transform_weighted": (C.c_int, [C.c_int, c_ip, C.c_int, C.c_int, c_dp, C.c_char_p, C.c_char_p, c_dp, c_dp, c_dp,
                                        c_dp, c_ip, C.c_char_p, c_dp, C.c_size_t]),
    "hb_eval_weighted": (C.c_int, [C.c_int, c_ip, C.c_int, c_dp, C.c_size_t, c_dp, c_dp, c_dp, C.c_char_p, c_dp, c_dp, c_ip,
                                   C.c_char_p, c_dp, C.c_size_t]),
    "hb_varf_function": (C.c_int, [C.c_int, c_ip, C.c_int, C.c_char_p, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_char_p, C.c_int,
                                   c_dp, C.c_size_t]),
    "hb_varf_function_sp": (C.c_int, [C.c_int, c_ip, C.c_int, C.c_char_p, c_dp, c_dp, c_dp, c_dp, c_ip, C.c_char_p,
                                      C.c_int, c_llp]),
    "hb_integrate": (C.c_int, [C.c_int, c_ip, c_dp, c_dp, c_dp, c_ip, c_dp]),
    "hb_triple_products": (C.c_int, [C.c_int, c_dp]),
    "hb_triple_products_fourier": (C.c_int, [C.c_int, c_dp]),
    "hb_varf": (C.c_int, [C.c_int, c_ip, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_char_p, C.c_int, c_dp, C.c_size_t]),
    "hb_varf_sp": (C.c_int, [C.c_int, c_ip, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_char_p, C.c_int, c_llp]),
    "hb_sparse_fetch": (C.c_int, [c_llp, c_ip, c_dp]),
    "hb_tensorize_mats_sp": (C.c_int, [C.c_int, C.POINTER(c_llp), C.POINTER(c_ip), C.POINTER(c_dp), c_ip, c_ip, c_ip,
                                       C.c_char_p, c_llp, c_llp]),
    "hb_project_mat_sp": (C.c_int, [c_llp, c_ip, c_dp, C.c_size_t, C.c_int, c_ip, C.c_int, C.c_char_p, c_llp, c_llp]),
    "hb_varfd_sp": (C.c_int, [C.c_int, C.c_int, C.c_int, c_llp, c_ip, c_dp, C.c_size_t, C.c_int, C.c_char_p, c_llp]),
    "hb_plan_set_sqrt_weights": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "hb_plan_varf_sp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, c_llp, C.c_void_p]),
    "hb_varfd": (C.c_int, [C.c_int, C.c_int, C.c_int, c_dp, C.c_size_t, C.c_int, C.c_char_p, c_dp]),
    "hb_tensorize_vecs": (C.c_int, [C.c_int, C.POINTER(c_dp), c_ip, c_ip, c_ip, C.c_char_p, c_dp, C.c_size_t]),
    "hb_tensorize_mats": (C.c_int, [C.c_int, C.POINTER(c_dp), c_ip, c_ip, c_ip, C.c_char_p, c_dp, C.c_size_t]),
    "hb_project_vec": (C.c_int, [c_dp, C.c_size_t, C.c_int, c_ip, C.c_int, C.c_char_p, c_dp, C.c_size_t]),
    "hb_project_mat": (C.c_int, [c_dp, C.c_size_t, C.c_int, c_ip, C.c_int, C.c_char_p, c_dp, C.c_size_t]),
    "hb_inner": (C.c_int, [c_dp, C.c_size_t, c_dp, C.c_size_t, c_ip, C.c_int, c_ip, C.c_int, C.c_char_p, c_dp,
                           C.c_size_t]),
    "hb_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_ip, C.c_int, c_dp, c_dp, c_ip, C.c_char_p]),
    "hb_plan_destroy": (None, [C.c_void_p]),
    "hb_plan_n_polys": (C.c_longlong, [C.c_void_p]),
    "hb_plan_grid_pitch": (C.c_longlong, [C.c_void_p]),
    "hb_plan_grid_size": (C.c_longlong, [C.c_void_p]),
    "hb_plan_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong,
                                  C.c_void_p]),
    "hb_plan_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong,
                                   C.c_void_p]),
    "hb_plan_varf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong,
                               C.c_longlong, C.c_void_p]),
    "hb_launch_count": (C.c_longlong, [C.c_int]),
    "hb_plan_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "hb_plan_timing_count": (C.c_int, [C.c_void_p]),
    "hb_plan_timing_get": (C.c_int, [C.c_void_p, C.c_int, c_ip, c_dp, c_dp]),
    "hb_dev_dgemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_int,
                               C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong,
                               C.c_void_p]),
    "hb_dev_dgemm_scatter": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong,
                                       C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.POINTER(C.c_void_p), C.c_int,
                                       C.c_int, C.c_longlong, C.c_longlong, C.c_void_p]),
    "hb_dev_dgemm_ex": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hb_dev_parity_pass": (C.c_int, [C.c_int, C.c_int, c_ip, c_ip, c_llp, c_llp, c_llp, C.c_longlong, C.c_longlong,
                                     C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_plan_parity_table": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), c_llp, c_llp, c_llp, c_llp,
                                       C.c_void_p]),
    "hb_dev_matvec": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "hb_dev_varfd": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_size_t, C.c_int,
                               C.c_char_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "hb_plan_table": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), c_llp, c_llp, c_llp]),
    "hb_plan_varf_info": (C.c_int, [C.c_void_p, c_ip, c_ip]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `make -C hermipy_b200/csrc` (or __graft_entry__.build()); "
                "hermipy_b200 has no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != HB_OK:
        text = lib().hb_last_error().decode(errors="replace")
        if rc == HB_ERR_DEGREE:
            raise DegreeError(rc, text)
        raise HermiteB200Error(rc, text)
