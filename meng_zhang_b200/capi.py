"""ctypes binding of libannp_b200.so (include/annp_b200.h).

The shared library is the product; this module only loads it and declares the prototypes.  There is
no Python or CPU implementation behind it: if the library is missing, import of this module's
`lib()` raises, and on a box without an sm_100 GPU every compute entry point returns
ANNP_B200_ENODEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ANNP_B200_LIB selects another build of the same library (kernel experiments); there is still no non-CUDA path behind it
LIB_PATH = os.environ.get("ANNP_B200_LIB") or os.path.join(HERE, "lib", "libannp_b200.so")

MAX_SF, MAX_NOD, MAX_LAYERS, MAX_ELEMENTS, MAX_NEIGH = 64, 32, 6, 4, 384
ABI_VERSION = 3
VARIANT_FE, VARIANT_NI, VARIANT_ANNA_ADP = 0, 1, 2
MAX_GPARAMS = 32
VARIANT_FLAG_GENERIC = 0x100
VARIANT_FLAG_NOPAIR = 0x200
SCATTER_GATHER, SCATTER_FIXED = 0, 1

OK, ENOMEM, ENODEVICE, EINVAL, ECUDA, ESTATE, EOVERFLOW, EIO, ECOMM = 0, -3, -4, -20, -21, -22, -23, -24, -25

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int64_p = C.POINTER(C.c_int64)


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("ntypes", C.c_int), ("nelements", C.c_int),
        ("ntl", C.c_int), ("nhl", C.c_int), ("nnod", C.c_int), ("nsf", C.c_int), ("npsf", C.c_int), ("ntsf", C.c_int),
        ("flagsym", C.c_int), ("flagact", C.c_int * MAX_LAYERS),
        ("e_scale", C.c_double), ("e_shift", C.c_double), ("e_atom", C.c_double), ("cut", C.c_double),
        ("sfnor_scal", c_double_p), ("sfnor_avg", c_double_p), ("cutsq", c_double_p), ("map", c_int_p),
        ("weights", c_double_p), ("bias", c_double_p),
        ("variant", C.c_int), ("sym_coerad", c_double_p), ("sym_coeang", c_double_p),
    ]


class Potential(C.Structure):
    _fields_ = [
        ("nelements", C.c_int), ("ntl", C.c_int), ("nhl", C.c_int), ("nnod", C.c_int), ("nsf", C.c_int),
        ("npsf", C.c_int), ("ntsf", C.c_int), ("flagsym", C.c_int), ("flagact", C.c_int * MAX_LAYERS),
        ("cut", C.c_double), ("e_scale", C.c_double), ("e_shift", C.c_double), ("e_atom", C.c_double),
        ("id_elem", C.c_int * MAX_ELEMENTS), ("mass", C.c_double * MAX_ELEMENTS),
        ("elements", (C.c_char * 16) * MAX_ELEMENTS),
        ("sfnor_cov", C.c_double * MAX_SF), ("sfnor_avg", C.c_double * MAX_SF),
        ("weight_all", c_double_p), ("bias_all", c_double_p),
        ("has_sym_coeff", C.c_int), ("sym_coerad", (C.c_double * 3) * MAX_SF), ("sym_coeang", (C.c_double * 4) * MAX_SF),
    ]


class AnnaPotential(C.Structure):
    _fields_ = [
        ("nelements", C.c_int), ("ntl", C.c_int), ("nhl", C.c_int), ("nnod", C.c_int), ("nout", C.c_int), ("nsf", C.c_int),
        ("npsf", C.c_int), ("ntsf", C.c_int), ("flagsym", C.c_int), ("flagact", C.c_int * MAX_LAYERS),
        ("cut", C.c_double), ("e_base", C.c_double), ("e_scal", C.c_double), ("ngp", C.c_int),
        ("gparams", C.c_double * MAX_GPARAMS),
        ("id_elem", C.c_int * MAX_ELEMENTS), ("mass", C.c_double * MAX_ELEMENTS), ("elements", (C.c_char * 16) * MAX_ELEMENTS),
        ("weight_all", c_double_p), ("bias_all", c_double_p),
    ]


class AnnaParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("ntypes", C.c_int), ("nelements", C.c_int),
        ("ntl", C.c_int), ("nhl", C.c_int), ("nnod", C.c_int), ("nout", C.c_int), ("nsf", C.c_int), ("npsf", C.c_int),
        ("ntsf", C.c_int), ("ngp", C.c_int), ("flagsym", C.c_int), ("flagact", C.c_int * MAX_LAYERS),
        ("e_base", C.c_double), ("cut", C.c_double),
        ("cutsq", c_double_p), ("map", c_int_p), ("weights", c_double_p), ("bias", c_double_p), ("gparams", c_double_p),
    ]


class NhConfig(C.Structure):
    _fields_ = [
        ("tstat", C.c_int), ("pstat", C.c_int), ("t_start", C.c_double), ("t_stop", C.c_double), ("t_damp", C.c_double),
        ("p_flag", C.c_int * 3), ("p_start", C.c_double * 3), ("p_stop", C.c_double * 3), ("p_damp", C.c_double * 3),
        ("tchain", C.c_int), ("pchain", C.c_int), ("mtk", C.c_int), ("dt", C.c_double), ("mass", C.c_double),
        ("natoms_total", C.c_double), ("tdof", C.c_double), ("nsteps_ramp", C.c_longlong),
    ]


class NhState(C.Structure):
    _fields_ = [
        ("step", C.c_longlong), ("t_current", C.c_double), ("t_target", C.c_double), ("p_current", C.c_double * 3),
        ("boxlo", C.c_double * 3), ("boxhi", C.c_double * 3), ("omega_dot", C.c_double * 3),
        ("ke_tensor", C.c_double * 6), ("virial", C.c_double * 6),
        ("eta", C.c_double * 8), ("eta_dot", C.c_double * 8), ("etap", C.c_double * 8), ("etap_dot", C.c_double * 8),
        ("extended_energy", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("inum", C.c_int), ("nall", C.c_int), ("max_neigh_list", C.c_int), ("max_neigh_cut", C.c_int),
        ("avg_neigh_cut", C.c_double), ("sum_triplets", C.c_double), ("kernel_launches", C.c_longlong),
        ("last_force_kernel_ms", C.c_float), ("force_kernel_ms_total", C.c_double), ("force_kernel_samples", C.c_int),
        ("overflow_pass_atoms", C.c_longlong), ("tile_capacity", C.c_int), ("stage_cycles", C.c_double * 8),
    ]


# every symbol include/annp_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "annp_b200_abi_version": (C.c_int, []),
    "annp_b200_device_count": (C.c_int, []),
    "annp_b200_weights_per_element": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "annp_b200_bias_per_element": (C.c_size_t, [C.c_int, C.c_int]),
    "annp_b200_read_potential": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(Potential), C.c_char_p, C.c_int]),
    "annp_b200_free_potential": (None, [C.POINTER(Potential)]),
    "anna_b200_read_potential": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(AnnaPotential), C.c_char_p, C.c_int]),
    "anna_b200_free_potential": (None, [C.POINTER(AnnaPotential)]),
    "anna_b200_init": (C.c_int, [C.POINTER(AnnaParams), C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]),
    "annp_b200_init": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]),
    "annp_b200_clear": (None, [C.c_void_p]),
    "annp_b200_neigh_build_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double]),
    "annp_b200_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "annp_b200_host_unregister": (C.c_int, [C.c_void_p]),
    "annp_b200_host_alloc": (C.c_void_p, [C.c_size_t]),
    "annp_b200_host_free": (None, [C.c_void_p]),
    "annp_b200_bytes": (C.c_double, [C.c_void_p]),
    "annp_b200_last_error": (C.c_char_p, [C.c_void_p]),
    "annp_b200_neigh": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p, c_int_p, C.POINTER(c_int_p)]),
    "annp_b200_neigh_csr": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p, c_int64_p, c_int_p]),
    "annp_b200_compute": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, c_int_p, C.c_int, C.c_int,
                                    c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "annp_b200_neigh_build": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, c_double_p, c_double_p, C.c_double, C.c_void_p]),
    "annp_b200_compute_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_send_lists_count": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_double_p, c_double_p, C.c_double, C.c_int, c_int_p, c_int_p, C.c_void_p]),
    "annp_b200_send_lists_fill": (C.c_int, [C.c_void_p, c_double_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_comm_unique_id": (C.c_int, [C.c_char_p]),
    "annp_b200_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "annp_b200_comm_destroy": (None, [C.c_void_p]),
    "annp_b200_set_halo_peers": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_int_p]),
    "annp_b200_halo_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_halo_reverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_allreduce_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "annp_b200_peer_export": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    "annp_b200_peer_open": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_peer_close": (C.c_int, [C.c_void_p]),
    "annp_b200_set_halo": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_halo_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_halo_unpack_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_max_displacement_sq": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_nve_initial": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_nve_final": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_nh_create": (C.c_int, [C.POINTER(NhConfig), c_double_p, c_double_p, C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_int]),
    "annp_b200_nh_destroy": (None, [C.c_void_p]),
    "annp_b200_nh_reduce": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_nh_setup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_nh_initial": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "annp_b200_nh_final_kick": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_nh_final_scale": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "annp_b200_nh_set_box": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int_p, C.c_void_p]),
    "annp_b200_nh_get_state": (C.c_int, [C.c_void_p, C.POINTER(NhState), C.c_void_p]),
    "annp_b200_fp64_peak_tflops": (C.c_double, [C.c_void_p, C.c_int]),
    "annp_b200_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "annp_b200_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "annp_b200_set_scatter": (C.c_int, [C.c_void_p, C.c_int]),
    "annp_b200_basis_matrices": (C.c_int, [C.c_int, c_double_p, c_double_p]),
    "annp_b200_debug_descriptors": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "annp_b200_debug_neighbors": (C.c_longlong, [C.c_void_p, c_int64_p, c_int_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libannp_b200.so (built by `make -C meng_zhang_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() - there is no fallback path")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.annp_b200_abi_version() != ABI_VERSION:
            raise ImportError("libannp_b200.so ABI version mismatch")
        _lib = L
    return _lib


class AnnpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"annp_b200 error {code}: {msg}")
        self.code = code
