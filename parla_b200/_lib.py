"""ctypes binding of ``libparla_b200.so`` (the C ABI declared in ``include/parla_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a symbol cannot be
resolved, importing a kernel wrapper raises.  Build with ``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C parla_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PARLA_B200_LIB", os.path.join(_HERE, "libparla_b200.so"))

c_i64, c_u64, c_int, c_dbl, c_sz, c_vp = C.c_int64, C.c_uint64, C.c_int, C.c_double, C.c_size_t, C.c_void_p

# name -> (restype, argtypes).  Pointers are passed as integers (torch .data_ptr()).
SIGNATURES = {
    "pla_version": (c_int, []),
    "pla_last_error": (C.c_char_p, []),
    "pla_num_sms": (c_int, []),
    "pla_launch_count": (C.c_longlong, []),
    "pla_note_launches": (None, [C.c_longlong]),
    "pla_dmma_probe": (c_int, [c_int, c_int, c_vp, c_vp]),
    "pla_stream_pass_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "pla_stream_pass_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_int,
                                    c_vp, c_vp, c_sz, c_vp]),
    "pla_stream_pass_parts_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_int,
                                          c_vp, c_vp, c_sz, c_vp, c_vp]),
    "pla_stream_pass_peer_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_int,
                                         c_vp, c_vp, c_sz, c_vp, c_int, c_int, c_i64, C.c_uint32, c_vp]),
    "pla_peer_exchange_bytes": (c_sz, [c_int, c_i64]),
    "pla_peer_alloc": (c_int, [c_sz, c_vp]),
    "pla_peer_free": (c_int, [c_vp]),
    "pla_peer_export": (c_int, [c_vp, c_vp]),
    "pla_peer_import": (c_int, [c_vp, c_vp]),
    "pla_peer_close": (c_int, [c_vp]),
    "pla_trsv_upper_f64": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp]),
    "pla_trtri_diag_f64": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "pla_trtri_merge_f64": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp]),
    "pla_lsqr_init_f64": (c_int, [c_i64, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_dbl, c_int, c_vp, c_vp, c_vp, c_vp,
                                  c_vp, c_vp, c_vp]),
    "pla_lsqr_step_f64": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pla_lsqr_fused_step_f64": (c_int, [c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pla_lsqr_ridge_f64": (c_int, [c_i64, c_dbl, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_vp, c_vp, c_vp]),
    "pla_lsqr_under_init_f64": (c_int, [c_i64, c_vp, c_vp, c_dbl, c_dbl, c_dbl, c_int, c_vp, c_vp, c_vp]),
    "pla_lsqr_under_init2_f64": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp]),
    "pla_lsqr_under_head_f64": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pla_lsqr_under_tail_f64": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pla_lsqr_under_long_f64": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_sz, c_vp]),
    "pla_pcg_residual_f64": (c_int, [c_i64, c_vp, c_vp, c_dbl, c_vp, c_vp, c_vp, c_vp, c_int, c_dbl, c_vp]),
    "pla_pcg_direction_f64": (c_int, [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "pla_pcg_update_f64": (c_int, [c_i64, c_vp, c_dbl, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_vp]),
    "pla_sjlt_rmatvec_f64": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_dbl, c_vp, c_vp]),
    "pla_gauss_rmatvec_f64": (c_int, [c_i64, c_i64, c_u64, c_i64, c_dbl, c_vp, c_vp, c_vp]),
    "pla_srct_weights_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_vp, c_i64, c_vp]),
    "pla_gather_rows_scale_f64": (c_int, [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp]),
    "pla_sjlt_plan_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pla_sjlt_plan_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pla_sjlt_plan_f64": (c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "pla_sjlt_apply_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_dbl, c_vp, c_i64, c_vp,
                                   c_i64, c_int, c_vp, c_sz, c_vp]),
    "pla_sjlt_apply_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "pla_sjlt_plan_status": (c_int, [c_vp, C.POINTER(c_i64)]),
    "pla_sjlt_generate": (c_int, [c_i64, c_i64, c_i64, c_u64, c_i64, c_vp, c_vp, c_vp]),
    "pla_gemm_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pla_gemm_f64": (c_int, [c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_vp, c_i64, c_vp, c_i64, c_dbl, c_vp, c_i64,
                             c_vp, c_sz, c_vp]),
    "pla_sketch_gauss_workspace_bytes": (c_sz, [c_i64, c_i64, c_i64]),
    "pla_sketch_gauss_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_u64, c_i64, c_dbl, c_dbl, c_vp, c_i64,
                                     c_vp, c_sz, c_vp]),
    "pla_philox_normal_fill_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_u64, c_i64, c_i64, c_dbl, c_vp]),
    "pla_qr_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "pla_geqrf_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "pla_qr_factor_block_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_int, c_i64, c_vp, c_sz, c_vp]),
    "pla_qr_apply_block_f64": (c_int, [c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "pla_orgqr_f64": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp, c_sz, c_vp]),
    "pla_sumsq_workspace_bytes": (c_sz, [c_i64]),
    "pla_sumsq_f64": (c_int, [c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
}

_lib = None


class ParlaB200Error(RuntimeError):
    pass


def load():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ParlaB200Error(
            f"{LIB_PATH} not found: the CUDA extension is not built. There is no CPU fallback; "
            "run `make -C parla_b200/csrc` (or __graft_entry__.build()).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pla_last_error().decode(errors="replace")
        raise ParlaB200Error(f"{what} failed (rc={rc}): {msg}")
