"""ctypes binding of libcg_b200.so (include/cg_b200.h).  No torch types cross this boundary: pointers and sizes only.

The library is built in-tree by `make -C coulombgalore_b200/csrc` (or __graft_entry__.build()).  There is no
fallback of any kind: if the shared object is missing, importing this module raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CG_B200_LIB: tuning builds (make LIB=... EXTRA_NVFLAGS=-D...); the default is the in-tree library
LIB_PATH = os.environ.get("CG_B200_LIB") or os.path.join(HERE, "libcg_b200.so")

CG_OK = 0
STATUS = {0: "CG_OK", -1: "CG_ERR_INVALID", -2: "CG_ERR_NO_DEVICE", -3: "CG_ERR_CUDA", -4: "CG_ERR_ALLOC",
          -5: "CG_ERR_UNSUPPORTED", -6: "CG_ERR_STATE"}

# cg_scheme_id
(PLAIN, EWALD, EWALDT, REACTIONFIELD, WOLF, POISSON, QPOTENTIAL, FANOURGAKIS, ZERODIPOLE, ZAHN, FENNELL, QPOTENTIAL5,
 SPLINE) = range(13)
MODE_ION, MODE_DIPOLE, MODE_MULTIPOLE, MODE_MULTIPOLE_QUAD = 0, 1, 2, 3
ENERGY, FORCE, FIELD, POTENTIAL = 1, 2, 4, 8
MAX_POISSON_C = 12
SPLINE_MAX_KNOTS = 64

_dp = C.POINTER(C.c_double)


class SchemeDesc(C.Structure):
    """cg_scheme_desc"""
    _fields_ = [
        ("functor", C.c_int32), ("scheme", C.c_int32), ("order", C.c_int32), ("C", C.c_int32), ("D", C.c_int32),
        ("yukawa_map", C.c_int32), ("shifted", C.c_int32), ("reserved0", C.c_int32),
        ("cutoff", C.c_double), ("inverse_cutoff", C.c_double), ("cutoff_squared", C.c_double),
        ("debye_length", C.c_double), ("inverse_debye_length", C.c_double),
        ("eta", C.c_double), ("zeta", C.c_double), ("erfc_eta", C.c_double), ("exp_eta2", C.c_double),
        ("F0", C.c_double), ("rf_a", C.c_double), ("rf_b", C.c_double), ("eps_rf", C.c_double), ("eps_r", C.c_double),
        ("reduced_kappa", C.c_double), ("yukawa_denom", C.c_double), ("binomCDC", C.c_double),
        ("poisson_w", C.c_double * MAX_POISSON_C),
        ("self_energy_prefactor", C.c_double * 2), ("self_field_prefactor", C.c_double * 2),
        ("T0", C.c_double), ("chi", C.c_double),
        ("spline_knots", C.c_int32 * 4), ("spline_r2", _dp * 4), ("spline_c", _dp * 4),
    ]


class CgError(RuntimeError):
    def __init__(self, rc, where, detail=""):
        self.rc = rc
        super().__init__("%s failed: %s%s" % (where, STATUS.get(rc, str(rc)), (" -- " + detail) if detail else ""))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "coulombgalore_b200: %s is missing -- build it with `make -C coulombgalore_b200/csrc` "
            "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    D = C.POINTER(SchemeDesc)
    H = C.c_void_p
    sig = {
        "cg_scheme_plain": [C.c_double, D],
        "cg_scheme_reactionfield": [C.c_double, C.c_double, C.c_double, C.c_int, D],
        "cg_scheme_poisson": [C.c_double, C.c_int, C.c_int, C.c_double, D],
        "cg_scheme_qpotential": [C.c_double, C.c_int, D],
        "cg_scheme_qpotential5": [C.c_double, D],
        "cg_scheme_fanourgakis": [C.c_double, D],
        "cg_scheme_fennell": [C.c_double, C.c_double, D],
        "cg_scheme_zerodipole": [C.c_double, C.c_double, D],
        "cg_scheme_zahn": [C.c_double, C.c_double, D],
        "cg_scheme_wolf": [C.c_double, C.c_double, D],
        "cg_scheme_ewald": [C.c_double, C.c_double, C.c_double, C.c_double, D],
        "cg_scheme_ewaldt": [C.c_double, C.c_double, C.c_double, C.c_double, D],
        "cg_scheme_create": [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), _dp, D],
        "cg_scheme_splined": [D, C.c_double, _dp, D],
        "cg_scheme_splined_tables": [D, C.POINTER(C.c_int32), C.POINTER(_dp), C.POINTER(_dp), D],
        "cg_srf_host": [D, C.c_int64, _dp, _dp],
        "cg_device_count": [],
        "cg_create": [D, C.c_int, C.POINTER(H)],
        "cg_destroy": [H],
        "cg_set_stream": [H, C.c_void_p],
        "cg_set_precision": [H, C.c_int],
        "cg_share_system": [H, H],
        "cg_set_shard": [H, C.c_int, C.c_int],
        "cg_set_system": [H, C.c_int64] + [C.c_void_p] * 7 + [_dp, C.c_int],
        "cg_set_system_device": [H, C.c_int64] + [C.c_void_p] * 7 + [_dp, C.c_int],
        "cg_eval": [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _dp, C.POINTER(C.c_int64)],
        "cg_eval_begin": [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p],
        "cg_eval_end": [H, _dp, C.POINTER(C.c_int64)],
        "cg_eval_device": [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
        "cg_set_quadrupoles": [H, C.c_void_p],
        "cg_set_quadrupoles_device": [H, C.c_void_p],
        "cg_self_terms": [H, _dp, C.c_void_p],
        "cg_self_terms_device": [H, C.c_void_p, C.c_void_p],
        "cg_eval_multi_device": [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)],
        "cg_eval_multi_begin": [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_void_p)],
        "cg_eval_multi_end": [C.POINTER(C.c_void_p), C.c_int, _dp, C.POINTER(C.c_int64)],
        "cg_halo_collect_owned": [H, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_void_p)],
        "cg_tail_plan_probe": [C.c_int, C.c_int, C.POINTER(C.c_int)],
        "cg_collect_evaluated": [H, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_int64,
                                 C.POINTER(C.c_int64)],
        "cg_dipole_moments": [H, _dp],
        "cg_dipole_moments_device": [H, C.c_void_p],
        "cg_read_scalars": [H, C.c_void_p, _dp],
        "cg_dipole_torque": [H, C.c_void_p, C.c_void_p],
        "cg_dipole_torque_device": [H, C.c_void_p, C.c_void_p],
        "cg_srf_device": [H, C.c_int64, C.c_void_p, C.c_void_p],
        "cg_last_timings": [H, _dp],
        "cg_timings_history": [H, C.c_int, _dp],
        "cg_last_counters": [H, C.POINTER(C.c_int64)],
        "cg_peer_alloc": [C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p],
        "cg_peer_free": [C.c_void_p],
        "cg_peer_open": [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)],
        "cg_peer_close": [C.c_void_p],
        "cg_peer_gather": [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_void_p)],
        "cg_set_grid_count": [H, C.c_int64],
        "cg_set_count_device": [H, C.c_void_p],
        "cg_halo_set_count_mirror": [H, C.c_void_p],
        "cg_shard_window": [H, _dp, C.c_int, C.c_int, C.c_int64, _dp],
        "cg_layout_probe": [C.c_double, _dp, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int), _dp],
        "cg_halo_publish": [H, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), _dp, _dp, C.c_double, C.c_void_p],
        "cg_halo_pull": [H, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_void_p), C.c_int64,
                         C.c_void_p],
        "cg_halo_publish_step": [H, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), _dp, _dp, C.c_double, C.c_void_p, C.c_int64,
                                 C.c_int, C.POINTER(C.c_void_p), C.c_void_p],
        "cg_halo_pull_step": [H, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_void_p), C.c_int64,
                              C.c_void_p, C.c_int64],
        "cg_halo_exchange_step": [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p), _dp, _dp, _dp, C.c_void_p,
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_int, C.c_int],
        "cg_generate_lattice_device": [H, C.c_int32, C.c_double, C.c_uint64] + [C.c_void_p] * 7,
        "cg_fp64_peak_probe": [H, _dp, _dp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)  # AttributeError here == the library does not export what the header declares
        fn.argtypes = args
        fn.restype = C.c_int
    R = C.c_void_p
    p7 = [C.c_void_p] * 7
    rsig = {
        "cg_recip_create": [C.c_double] * 4 + [C.c_int, _dp, C.c_int, C.POINTER(R)],
        "cg_recip_destroy": [R],
        "cg_recip_set_stream": [R, C.c_void_p],
        "cg_recip_num_kvectors": [R],
        "cg_recip_kvectors": [R, _dp, _dp],
        "cg_recip_get_q": [R, _dp],
        "cg_recip_set_q": [R, _dp],
        "cg_recip_update": [R, C.c_int64] + p7 + [C.c_int],
        "cg_recip_update_device": [R, C.c_int64] + p7 + [C.c_int],
        "cg_recip_energy": [R, _dp],
        "cg_recip_energy_device": [R, C.c_void_p],
        "cg_recip_force_field": [R, C.c_int64] + p7 + [C.c_void_p, C.c_void_p],
        "cg_recip_force_field_device": [R, C.c_int64] + p7 + [C.c_void_p, C.c_void_p],
        "cg_recip_surface": [R, C.c_int64] + p7 + [_dp, _dp],
        "cg_recip_surface_device": [R, C.c_int64] + p7 + [C.c_void_p],
    }
    for name, args in rsig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.cg_recip_last_error.argtypes = [R]
    lib.cg_recip_last_error.restype = C.c_char_p
    lib.cg_recip_q_device.argtypes = [R]
    lib.cg_recip_q_device.restype = C.c_void_p
    lib.cg_halo_block_bytes.argtypes = [C.c_int, C.c_int, C.c_int64]
    lib.cg_halo_block_bytes.restype = C.c_size_t
    lib.cg_spline_storage_doubles.argtypes = []
    lib.cg_spline_storage_doubles.restype = C.c_int64
    lib.cg_last_error.argtypes = [H]
    lib.cg_last_error.restype = C.c_char_p
    lib.cg_version.argtypes = []
    lib.cg_version.restype = C.c_char_p
    return lib


lib = _load()

# every symbol include/cg_b200.h declares (checked by tests/test_capi_symbols.py against the header text)
EXPORTED = ["cg_scheme_plain", "cg_scheme_reactionfield", "cg_scheme_poisson", "cg_scheme_qpotential",
            "cg_scheme_qpotential5", "cg_scheme_fanourgakis", "cg_scheme_fennell", "cg_scheme_zerodipole",
            "cg_scheme_zahn", "cg_scheme_wolf", "cg_scheme_ewald", "cg_scheme_ewaldt", "cg_scheme_create", "cg_spline_storage_doubles",
            "cg_scheme_splined", "cg_scheme_splined_tables", "cg_srf_host", "cg_device_count", "cg_create",
            "cg_destroy", "cg_last_error", "cg_version", "cg_set_stream", "cg_set_precision", "cg_share_system", "cg_set_shard",
            "cg_set_system", "cg_set_system_device", "cg_eval", "cg_eval_begin", "cg_eval_end", "cg_eval_device", "cg_eval_multi_device", "cg_eval_multi_begin", "cg_eval_multi_end", "cg_set_quadrupoles", "cg_set_quadrupoles_device", "cg_self_terms", "cg_self_terms_device", "cg_dipole_moments", "cg_dipole_moments_device", "cg_read_scalars", "cg_dipole_torque", "cg_dipole_torque_device", "cg_srf_device", "cg_last_timings", "cg_timings_history",
            "cg_last_counters", "cg_peer_alloc", "cg_peer_free", "cg_peer_open", "cg_peer_close", "cg_peer_gather", "cg_set_grid_count", "cg_set_count_device", "cg_shard_window", "cg_halo_block_bytes", "cg_layout_probe", "cg_tail_plan_probe", "cg_halo_publish",
            "cg_halo_pull", "cg_halo_set_count_mirror", "cg_halo_publish_step", "cg_halo_pull_step", "cg_halo_exchange_step", "cg_halo_collect_owned", "cg_collect_evaluated", "cg_generate_lattice_device", "cg_recip_create", "cg_recip_destroy", "cg_recip_last_error",
            "cg_recip_set_stream", "cg_recip_num_kvectors", "cg_recip_kvectors", "cg_recip_q_device", "cg_recip_get_q", "cg_recip_set_q",
            "cg_recip_update", "cg_recip_update_device", "cg_recip_energy", "cg_recip_energy_device", "cg_recip_force_field",
            "cg_recip_force_field_device", "cg_recip_surface", "cg_recip_surface_device", "cg_fp64_peak_probe"]


def check(rc, where, handle=None):
    if rc != CG_OK:
        detail = ""
        try:
            msg = lib.cg_last_error(handle)
            detail = msg.decode() if msg else ""
        except Exception:  # pragma: no cover
            pass
        raise CgError(rc, where, detail)
