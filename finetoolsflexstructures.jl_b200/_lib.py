"""ctypes binding of libfsgpu.so (the C ABI declared in include/fsgpu.h).

This is the Python stand-in for the Julia `ccall` layer (julia/FlexStructuresGPU.jl):
one thin stub per exported symbol, plain pointers and sizes.  The library is REQUIRED:
importing this module raises if `libfsgpu.so` has not been built -- there is no CPU
fallback of any operator.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FSGPU_LIB", os.path.join(_HERE, "libfsgpu.so"))


class FsgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fsgpu error {code}: {msg}")
        self.code = code


OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_DOF_RANGE, ERR_SINGULAR = range(6)

# enum fsgpu_target
SPARSE, SPARSE_SYMM, SPARSE_DIAG, FFBLOCK, FFBLOCK_DIAG, CSR_SYMM = range(6)
CSYS_CYLINDRICAL, CSYS_SPHERICAL, CSYS_NORMAL_AXIS = 1, 2, 3


class ShellParams(C.Structure):
    _fields_ = [
        ("Dps", C.c_double * 9),
        ("Dt", C.c_double * 4),
        ("rho", C.c_double),
        ("stab_alpha", C.c_double),
        ("drilling_stiffness_scale", C.c_double),
        ("transv_shear_formulation", C.c_int32),
        ("reserved", C.c_int32),
    ]


class BeamParams(C.Structure):
    _fields_ = [
        ("E", C.c_double),
        ("nu", C.c_double),
        ("rho", C.c_double),
        ("mass_type", C.c_int32),
        ("reserved", C.c_int32),
    ]


_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_dbl = C.c_double
_P = C.POINTER

# name -> argtypes (restype is int unless listed in _RESTYPE)
_SIGNATURES = {
    "fsgpu_create": [_P(_vp), C.c_int],
    "fsgpu_destroy": [_vp],
    "fsgpu_last_error": [],
    "fsgpu_version": [],
    "fsgpu_host_alloc": [_P(_vp), _i64],
    "fsgpu_host_free": [_vp],
    "fsgpu_set_stream": [_vp, _vp],
    "fsgpu_sync": [_vp],
    "fsgpu_launch_count": [_vp],
    "fsgpu_last_kernel_ms": [_vp, _P(_dbl)],
    "fsgpu_set_deterministic": [_vp, C.c_int],
    "fsgpu_scatter_path": [_vp, _P(C.c_int)],
    "fsgpu_measure_peaks": [_vp, _P(_dbl), _P(_dbl)],
    "fsgpu_set_mesh": [_vp, _i32, _i64, _vp, _i64, _vp],
    "fsgpu_set_dofnums": [_vp, _vp, _i64, _i64],
    "fsgpu_set_normals": [_vp, _vp, _vp],
    "fsgpu_associategeometry": [_vp, _dbl, _vp, _i32],
    "fsgpu_associategeometry_dirs": [_vp, _dbl, _vp, _i32],
    "fsgpu_associategeometry_csys": [_vp, _dbl, _i32, _vp, _vp, _i32],
    "fsgpu_set_layup_csys": [_vp, _i32, _vp, _vp],
    "fsgpu_normals_accumulate": [_vp, _vp, _i32, _P(_vp)],
    "fsgpu_normals_finish": [_vp, _dbl, _vp, _P(_vp)],
    "fsgpu_get_normals": [_vp, _vp, _vp],
    "fsgpu_set_thickness": [_vp, _vp, _i64],
    "fsgpu_set_stab_factor": [_vp, _vp, _i64],
    "fsgpu_element_sizes": [_vp, _vp],
    "fsgpu_set_rule": [_vp, _i32, _vp, _vp, _vp],
    "fsgpu_set_layup": [_vp, _i32, _vp, _vp, _vp, _i64],
    "fsgpu_set_beam_sections": [_vp] + [_vp] * 8,
    "fsgpu_set_state": [_vp, _vp, _vp],
    "fsgpu_symbolic": [_vp, _i32, _P(_i64), _P(_i64), _P(_i64)],
    "fsgpu_t3ff_stiffness": [_vp, _P(ShellParams)],
    "fsgpu_t3ff_mass": [_vp, _P(ShellParams)],
    "fsgpu_q4rs_stiffness": [_vp, _P(ShellParams)],
    "fsgpu_q4rs_mass": [_vp, _P(ShellParams)],
    "fsgpu_t3ffcomp_stiffness": [_vp, _P(ShellParams)],
    "fsgpu_t3ffcomp_mass": [_vp, _P(ShellParams)],
    "fsgpu_q4rscomp_stiffness": [_vp, _P(ShellParams)],
    "fsgpu_q4rscomp_mass": [_vp, _P(ShellParams)],
    "fsgpu_corotbeam_stiffness": [_vp, _P(BeamParams)],
    "fsgpu_corotbeam_geostiffness": [_vp, _P(BeamParams)],
    "fsgpu_corotbeam_mass": [_vp, _P(BeamParams)],
    "fsgpu_corotbeam_restoringforce": [_vp, _P(BeamParams), _i32],
    "fsgpu_set_velocity": [_vp, _vp],
    "fsgpu_corotbeam_gyroscopic": [_vp, _P(BeamParams)],
    "fsgpu_corotbeam_distribloads": [_vp, _P(BeamParams), _vp, _i64, _i32],
    "fsgpu_shell_mass_diag": [_vp, _P(ShellParams), _i32, _i32],
    "fsgpu_shell_resultants": [_vp, _P(ShellParams), _i32, _i32, _vp, _vp, _i64, _vp],
    "fsgpu_shell_nodal_field": [_vp, _P(ShellParams), _i32, _i32, _vp, _vp, _i64, _vp],
    "fsgpu_update_rotation_field": [_vp, _vp, _vp],
    "fsgpu_element_matrices": [_vp, _i32, _i32, _vp, _vp],
    "fsgpu_element_vectors": [_vp, _P(BeamParams), _vp],
    "fsgpu_result_size": [_vp, _P(_i64), _P(_i64), _P(_i64)],
    "fsgpu_fetch_matrix": [_vp, _vp, _vp, _vp],
    "fsgpu_result_size_uplo": [_vp, _i32, _P(_i64)],
    "fsgpu_fetch_matrix_uplo": [_vp, _i32, _vp, _vp, _vp],
    "fsgpu_fetch_vector": [_vp, _vp, _i64],
    "fsgpu_result_device": [_vp, _P(_vp), _P(_vp), _P(_vp)],
    "fsgpu_vector_device": [_vp, _P(_vp), _P(_i64)],
    "fsgpu_result_block": [_vp, _i64, _i64, _vp, _P(_i64), _vp, _vp, _vp],
    "fsgpu_coo_to_csc": [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _P(_i64), _vp, _vp, _vp],
    "fsgpu_explicit_create": [_P(_vp), _vp, _i64, _vp, _vp, _vp, _vp, _dbl, _dbl],
    "fsgpu_explicit_create_from_ctx": [_P(_vp), _vp, _dbl, _dbl],
    "fsgpu_explicit_destroy": [_vp],
    "fsgpu_explicit_set_state": [_vp, _vp, _vp],
    "fsgpu_explicit_set_load": [_vp, _vp],
    "fsgpu_explicit_set_timestep": [_vp, _dbl, _dbl],
    "fsgpu_explicit_start": [_vp, _dbl],
    "fsgpu_explicit_step": [_vp, _i64, _vp],
    "fsgpu_explicit_get_state": [_vp, _vp, _vp, _vp],
    "fsgpu_explicit_spmv": [_vp, _vp, _vp],
    "fsgpu_explicit_omega_max": [_vp, _i32, _P(_dbl)],
    "fsgpu_explicit_kinetic_energy": [_vp, _P(_dbl)],
    "fsgpu_explicit_device_state": [_vp, _P(_vp), _P(_vp), _P(_vp), _P(_vp)],
    "fsgpu_explicit_step_begin": [_vp],
    "fsgpu_explicit_step_end": [_vp, _dbl],
    "fsgpu_d2h_bytes": [_vp, _P(_i64)],
    "fsgpu_explicit_layout": [_vp, _P(_i64), _P(_i64), _P(_i64), _P(_i64)],
    "fsgpu_explicit_create_dist": [_P(_vp), _vp, _i32, _i32, _i64, _i64, _vp, _vp, _dbl, _dbl],
    "fsgpu_explicit_export": [_vp, _vp],
    "fsgpu_explicit_connect": [_vp, _vp],
    "fsgpu_explicit_dist_info": [_vp, _P(_i64), _P(_i64), _P(_i64), _P(_i64), _P(_i32)],
}
EXPLICIT_BLOB_BYTES = 1024
_RESTYPE = {"fsgpu_last_error": C.c_char_p, "fsgpu_launch_count": _i64}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with ./build.sh (or __graft_entry__.build()). "
            "The GPU path has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, C.c_int)
    return lib


lib = load()


def check(rc):
    if rc != OK:
        raise FsgpuError(rc, lib.fsgpu_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array (or None); torch tensors / ints pass through as raw addresses."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["ALIGNED"] + (["F"] if order == "F" else ["C"]))


def i64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.int64), requirements=["ALIGNED"] + (["F"] if order == "F" else ["C"]))
