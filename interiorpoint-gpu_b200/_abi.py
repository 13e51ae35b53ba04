"""ctypes binding of ``libipm_b200.so`` (the C-ABI declared in ``include/ipm_b200.h``).

The product path has NO CPU fallback: if the library is missing or no sm_100 device is visible,
``lib()`` / ``require_device()`` raise.  Arguments are raw device pointers (``tensor.data_ptr()``) and
sizes; every call is asynchronous on the given CUDA stream and returns a status code that is mapped to a
Python exception here.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPM_B200_LIB") or os.path.join(_HERE, "lib", "libipm_b200.so")  # override: A/B builds

_dp = C.c_void_p  # device pointer
_i = C.c_int
_d = C.c_double
_ll = C.c_longlong

# name -> (restype, argtypes).  Keep in sync with include/ipm_b200.h (tests/test_abi.py checks the header).
SIGNATURES = {
    "ipm_abi_version": (_i, []),
    "ipm_device_ok": (_i, []),
    "ipm_last_cuda_error": (C.c_char_p, []),
    "ipm_launch_count": (C.c_ulonglong, []),
    "ipm_device_fault": (C.c_uint, []),
    "ipm_clear_device_fault": (None, []),
    "ipm_set_spin_limit": (C.c_uint, [C.c_uint]),
    "ipm_gemm_tn_f64": (_i, [_dp, _i, _dp, _i, _dp, _d, _d, _dp, _i, _i, _i, _i, _i, _dp]),
    "ipm_hess_i8_ws_bytes": (_ll, [_i, _i, _i]),
    "ipm_hess_i8_prepare": (_i, [_dp, _i, _i, _i, _dp]),
    "ipm_hess_i8_f64": (_i, [_dp, _i, _i, _i, _dp, _d, _dp, _i, _i, _dp, _dp]),
    "ipm_hess_i8_scatter_f64": (_i, [_dp, _i, _i, _i, _dp, _i, _dp, C.POINTER(_dp), C.POINTER(_dp), _i, _i, _i, C.c_uint,
                                     _dp]),
    "ipm_gemv_n_f64": (_i, [_dp, _i, _i, _i, _dp, _dp, _d, _d, _dp]),
    "ipm_gemv_t_ws_doubles": (_ll, [_i, _i, _i]),
    "ipm_gemv_t_f64": (_i, [_dp, _i, _i, _i, _dp, _i, _i, _dp, _i, _d, _d, _dp, _ll, _dp]),
    "ipm_syrk_scatter_f64": (_i, [_dp, _i, _dp, _i, _i, _d, _dp, _i, C.POINTER(_dp), C.POINTER(_dp), _i, _i, _i, C.c_uint,
                                  _dp]),
    "ipm_hess_reduce_bcast_f64": (_i, [_dp, _dp, C.POINTER(_dp), C.POINTER(_dp), _i, _i, _i, _i, _i, C.c_uint, C.c_uint,
                                       _dp, _i, _d, _dp]),
    "ipm_hess_reduce_bcast_pull_f64": (_i, [C.POINTER(_dp), _dp, C.POINTER(_dp), C.POINTER(_dp), _i, _i, _i, _i, _i, C.c_uint,
                                            C.c_uint, _dp, _i, _d, _dp]),
    "ipm_l2_persist": (_i, [_dp, C.c_ulonglong, C.POINTER(_d), _dp]),
    "ipm_csr_gemv_f64": (_i, [_dp, _dp, _dp, _i, _dp, _dp, _d, _d, _dp]),
    "ipm_sparse_syrk_f64": (_i, [_i, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _i, _dp]),
    "ipm_dots_f64": (_i, [_i, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_i), _dp, _dp]),
    "ipm_axpy_dev_f64": (_i, [_i, _dp, _dp, _dp, _dp]),
    "ipm_symmetrize_upper_f64": (_i, [_dp, _i, _i, _dp]),
    "ipm_cg_ws_doubles": (_ll, [_i]),
    "ipm_cg_descent_x0_f64": (_i, [_i, _dp, _dp, _dp, _dp, _dp, _dp]),
    "ipm_cg_solve_f64": (_i, [_dp, _i, _i, _dp, _dp, _dp, _d, _i, _d, _dp, _dp]),
    "ipm_potrf_upper_f64": (_i, [_dp, _i, _i, _dp, _dp]),
    "ipm_potrf_upper_dag_f64": (_i, [_dp, _i, _i, _dp, _dp]),
    "ipm_potrf_peer_prog_words": (_i, []),
    "ipm_potrf_upper_peer_f64": (_i, [C.POINTER(_dp), _i, _i, C.POINTER(_dp), C.POINTER(_dp), _i, _i, C.c_uint, _i, _dp]),
    "ipm_potrf_trsm_upper_f64": (_i, [_dp, _i, _i, _dp, _i, _i, _dp, _dp]),
    "ipm_trsm_upper_t_f64": (_i, [_dp, _i, _i, _dp, _i, _i, _dp]),
    "ipm_trsv_upper_f64": (_i, [_dp, _i, _i, _dp, _i, _dp, _dp]),
    "ipm_lin_barrier_ws_doubles": (_ll, []),
    "ipm_lin_barrier_eval_f64": (_i, [_i, _i, _dp, _dp, _dp, _dp, _dp, _dp, _i, _d, _dp, _dp, _dp, _dp, _dp, _dp,
                                      _dp]),
    "ipm_lasso_partials_doubles": (_ll, [_i, _i]),
    "ipm_lasso_admm_step_f64": (_i, [_dp, _i, _i, _i, _dp, _dp, _d, _dp, _dp, _dp, _dp, _i, _i, _i, _i, _dp, _dp,
                                     _dp]),
    "ipm_lasso_steps_ws_bytes": (_ll, [_i, _i]),
    "ipm_lasso_steps_norms_offset": (_ll, [_i, _i]),
    "ipm_lasso_admm_steps_f64": (_i, [_dp, _i, _i, _i, _dp, _dp, _d, _dp, _dp, _dp, _dp, _i, _i, _i, _i, _i, _d, _d, _dp,
                                      _dp]),
    "ipm_lasso_objective_f64": (_i, [_dp, _i, _i, _dp, _i, _i, _i, _dp, _i, _i, _dp, _dp]),
    "ipm_scale_shift_f64": (_i, [_dp, _i, _dp, _i, _i, _i, _d, _d, _dp]),
    "ipm_cone_eval_f64": (_i, [_i, _dp, _i, _dp, _dp, _dp, _d, _i, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "ipm_cone_grad_rows_f64": (_i, [_i, _i, _dp, _dp, _i, _dp, _i, _dp, _dp, _dp, _i, _dp]),
    "ipm_cone_ls_coeffs_f64": (_i, [_i, _dp, _dp, _dp, _dp, _dp, _dp, _i, _dp, _dp, _dp]),
    "ipm_lin_grad_f64": (_i, [_i, _d, _dp, _dp, _dp, _dp, _i, _dp, _dp, _dp, _dp, _dp]),
    "ipm_hess_finish_f64": (_i, [_dp, _i, _i, _dp, _dp, _dp, _d, _dp]),
    "ipm_scale_copy_upper_f64": (_i, [_dp, _i, _dp, _i, _i, _d, _dp]),
    "ipm_ls_feas_lin_f64": (_i, [_i, _i, _dp, _dp, _dp, _i, _i, _i, _dp, _i, _dp, _dp, _dp]),
    "ipm_ls_feas_poly_f64": (_i, [_i, _dp, _dp, _dp, _dp, _i, _dp, _i, _dp]),
    "ipm_ls_armijo_f64": (_i, [_i, _dp, _dp, _dp, _dp, _i, _dp, _dp, _dp, _d, _d, _i, _dp, _dp, _i, _i, _dp, _dp]),
    "ipm_ls_logsum_f64": (_i, [_i, _dp, _dp, _dp, _dp, _i, _dp, _dp, _dp]),
    "ipm_ls_residual_f64": (_i, [_i, _i, _dp, _dp, _dp, _dp, _dp, _dp, _i, _dp, _d, _dp, _i, _i, _dp, _dp]),
    "ipm_trial_point_f64": (_i, [_i, _dp, _dp, _dp, _dp, _dp]),
    "ipm_lincomb3_f64": (_i, [_i, _d, _dp, _d, _dp, _d, _dp, _dp, _dp]),
    "ipm_table_lookup_f64": (_i, [_dp, _i, _dp, _dp, _dp]),
    "ipm_vec_op_f64": (_i, [_i, _i, _dp, _dp, _dp, _d, _dp]),
}

IPM_OK, IPM_ERR_ARG, IPM_ERR_CUDA, IPM_ERR_NO_DEVICE = 0, -1, -2, -3

_lib = None


class IpmError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IpmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the B200 engine has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def require_device():
    rc = lib().ipm_device_ok()
    if rc != IPM_OK:
        raise IpmError("no sm_100 (B200) CUDA device visible: the ipm_b200 engine has no CPU fallback")


def check(rc, what=""):
    if rc == IPM_OK:
        return
    if rc == IPM_ERR_ARG:
        raise IpmError(f"{what}: invalid argument")
    if rc == IPM_ERR_CUDA:
        raise IpmError(f"{what}: CUDA error: {lib().ipm_last_cuda_error().decode()}")
    if rc == IPM_ERR_NO_DEVICE:
        raise IpmError(f"{what}: no sm_100 device / driver entry point")
    raise IpmError(f"{what}: status {rc}")


FAULTS = {1: "tile-DAG Cholesky: a column-progress counter never arrived",
          2: "persistent GEMM: a stream-K partial never arrived",
          3: "row-sharded Hessian: a peer's partial tile never arrived",
          4: "row-sharded Hessian: the owners' final tiles never arrived",
          5: "triangular solve: a solution block was never published",
          6: "persistent ADMM kernel: a neighbour panel never arrived",
          7: "distributed Cholesky: a peer's tiles never arrived"}


def check_device_fault():
    """Raise if a device-side wait gave up (include/ipm_b200.h: ipm_device_fault).  Call after a stream sync."""
    code = lib().ipm_device_fault()
    if code:
        lib().ipm_clear_device_fault()
        raise IpmError(f"device watchdog fired (code {code}): {FAULTS.get(code, 'unknown wait')}")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def call(name, *args):
    check(getattr(lib(), name)(*args), name)
