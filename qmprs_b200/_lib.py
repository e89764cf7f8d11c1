"""ctypes loader for libqmprs_b200.so (the C ABI declared in include/qmprs_b200.h).

There is no fallback: if the shared library is missing or CUDA is unavailable the
package raises.  ``build.py`` compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QM_B200_LIB: another build of the same library (A/B runs of kernel variants on one box); the default is the in-tree one
LIB_PATH = os.environ.get("QM_B200_LIB") or os.path.join(_HERE, "libqmprs_b200.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_d = ctypes.c_double
_ip = ctypes.POINTER(ctypes.c_int)
_llp = ctypes.POINTER(ctypes.c_longlong)

# name -> (restype, argtypes); mirrors include/qmprs_b200.h one to one
SIGNATURES = {
    "qm_zgemm": (_i, [_i, _i, _i, _d, _d, _vp, _ll, _vp, _ll, _d, _d, _vp, _ll, _i, _ll, _ll, _ll, _i, _vp]),
    "qm_svd_work_bytes": (_ll, [_i, _i]),
    "qm_svd": (_i, [_i, _i, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _d, _i, _ip, _i, _vp]),
    "qm_svd_static": (_i, [_i, _i, _vp, _ll, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _d, _i, _vp, _i, _vp]),
    "qm_svd_small_fits": (_i, [_i, _i, _i]),
    "qm_svd_small": (_i, [_i, _i, _vp, _ll, _ll, _vp, _ll, _ll, _vp, _ll, _vp, _ll, _ll, _d, _i, _i, _i, _vp, _vp]),
    "qm_expect_ints": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "qm_expect_not_close": (_i, [_vp, _d, _vp, _vp]),
    "qm_transpose": (_i, [_vp, _ll, _vp, _ll, _ll, _ll, _i, _vp]),
    "qm_qr": (_i, [_i, _i, _vp, _ll, _vp, _vp]),
    "qm_qr_work_bytes": (_ll, [_i, _i]),
    "qm_qr_blocked": (_i, [_i, _i, _vp, _ll, _vp, _vp, _ll, _vp]),
    "qm_qr_formq_blocked": (_i, [_i, _i, _vp, _ll, _vp, _vp, _ll, _vp, _ll, _vp]),
    "qm_qr_formq": (_i, [_i, _i, _vp, _ll, _vp, _vp, _ll, _vp]),
    "qm_qr_finish": (_i, [_i, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp]),
    "qm_trim": (_i, [_vp, _i, _d, _i, _i, _vp, _vp, _vp]),
    "qm_scale_copy": (_i, [_vp, _ll, _vp, _ll, _i, _i, _vp, _vp, _i, _i, _vp]),
    "qm_theta_gate": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "qm_site_gate": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "qm_chi2_select": (_i, [_vp, _vp, _ll, _d, _d, _vp, _vp, _vp, _i, _d, _vp, _vp]),
    "qm_chi2_first": (_i, [_vp, _vp, _vp]),
    "qm_complete_unitaries": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _d, _vp]),
    "qm_split_absorb": (_i, [_vp, _ll, _vp, _vp, _ll, _i, _i, _i, _d, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "qm_theta_small": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp]),
    "qm_chi2_env": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "qm_chi2_bond": (_i, [_vp, _i, _vp, _vp, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _vp]),
    "qm_zero_overlap": (_i, [ctypes.POINTER(_vp), _ip, _i, _d, _vp, _vp, _vp]),
    "qm_split_absorb_batch": (_i, [_vp, _ll, _vp, _vp, _ll, _i, _i, _i, _d, _i, _i, _i, _vp, _vp, _vp, _i, _llp, _vp]),
    "qm_theta_small_batch": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _i, _llp, _vp]),
    "qm_chi2_env_batch": (_i, [_vp, _vp, _i, _i, _vp, _i, _llp, _vp]),
    "qm_chi2_bond_batch": (_i, [_vp, _i, _vp, _vp, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _i, _llp, _vp]),
    "qm_zero_overlap_batch": (_i, [ctypes.POINTER(_vp), _ip, _i, _d, _vp, _vp, _i, _llp, _vp]),
    "qm_site_gate_batch": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _ll, _vp]),
    "qm_chi2_first_batch": (_i, [_vp, _vp, _i, _ll, _ll, _vp]),
    "qm_complete_unitaries_batch": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _d, _i, _ll, _vp]),
    "qm_expect_ints_batch": (_i, [_vp, _vp, _i, _i, _vp, _i, _ll, _vp]),
    "qm_reverse3": (_i, [_vp, _vp, _i, _i, _vp]),
    "qm_conj_scale_copy": (_i, [_vp, _vp, _ll, _i, _d, _vp]),
    "qm_vdot_out_doubles": (_i, []),
    "qm_vdot": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "qm_div_sqrt": (_i, [_vp, _ll, _vp, _vp]),
    "qm_apply_gate": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "qm_circuit_state": (_i, [_vp, _i, _vp, _ip, _ip, _i, _vp]),
    "qm_sweep_work_bytes": (_ll, []),
    "qm_sweep": (_i, [_vp, _vp, _i, _vp, _ip, _ip, _i, _vp, _vp, _vp]),
    "qm_circuit_states": (_i, [_vp, _i, _vp, _ip, _ip, _i, _vp]),
    "qm_sweep_stored": (_i, [_vp, _vp, _i, _vp, _ip, _ip, _i, _vp, _vp, _vp, _vp]),
    "qm_sweeps_small": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "qm_sweeps_persist_work_bytes": (_ll, [_i]),
    "qm_sweeps_persist": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "qm_version": (_i, []),
    "qm_set_pdl": (_i, [_i]),
    "qm_launch_count": (_ll, []),
    "qm_prof_num_classes": (_i, []),
    "qm_prof_class_name": (ctypes.c_char_p, [_i]),
    "qm_prof_begin": (_i, []),
    "qm_prof_end": (_i, [ctypes.POINTER(_d), ctypes.POINTER(_ll)]),
    "qm_prof_work_get": (_i, [ctypes.POINTER(_d)]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library and bind every declared symbol (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m qmprs_b200.build` "
            "(nvcc, sm_100a).  qmprs_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
