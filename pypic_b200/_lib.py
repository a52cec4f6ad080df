"""ctypes binding of libpic_b200.so (the C ABI declared in include/pic_b200.h).

There is NO CPU fallback: if the library cannot be loaded, or a call returns a
non-zero status, a PicError is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpic_b200.so")

PIC_OK, PIC_ERR_CUDA, PIC_ERR_ARG, PIC_ERR_NODEVICE, PIC_ERR_RANGE = 0, -1, -2, -3, -4
PIC_PCR_SMEM_MAX = 6144


class PicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libpic_b200 error %d: %s" % (code, msg))
        self.code = code


class DDParams(C.Structure):
    _fields_ = [("N", C.c_int64), ("n_split", C.c_int64), ("Ng", C.c_int32), ("flags", C.c_int32),
                ("dx", C.c_double), ("dt", C.c_double), ("L", C.c_double), ("p2c", C.c_double),
                ("q", C.c_double * 2), ("m", C.c_double * 2)]


class DDPrologue(C.Structure):
    """pic_dd_prologue (include/pic_b200.h)."""
    _fields_ = [("log", C.c_void_p), ("next_log", C.c_void_p), ("log_cap", C.c_int32), ("philox", C.c_int32),
                ("sigma", C.c_double * 2), ("seed", C.c_uint64), ("step", C.c_uint64), ("global_offset", C.c_int64),
                ("slot", C.c_void_p), ("orig_of_draw", C.c_void_p), ("xd", C.c_void_p), ("ud", C.c_void_p),
                ("vd", C.c_void_p), ("wd", C.c_void_p), ("n_draws", C.c_int64), ("corr", C.c_void_p),
                ("x0", C.c_void_p), ("u0", C.c_void_p), ("v0", C.c_void_p), ("w0", C.c_void_p), ("active", C.c_void_p),
                ("orig", C.c_void_p),
                ("Es", C.c_void_p), ("E0", C.c_void_p), ("wall_cum", C.c_void_p), ("stats", C.c_void_p),
                ("nstats", C.c_int64), ("ctl", C.c_void_p)]


class PypicParams(C.Structure):
    _fields_ = [("N", C.c_int64), ("Ng", C.c_int32), ("flags", C.c_int32),
                ("dx", C.c_double), ("dt", C.c_double), ("L", C.c_double), ("p2c", C.c_double),
                ("q", C.c_double), ("m", C.c_double)]


class LParams(C.Structure):
    _fields_ = [("N", C.c_int64), ("n_split", C.c_int64), ("Ng", C.c_int32), ("flags", C.c_int32),
                ("dx", C.c_double), ("dt", C.c_double), ("L", C.c_double), ("p2c", C.c_double),
                ("q", C.c_double * 2), ("m", C.c_double * 2)]


class GCParams(C.Structure):
    _fields_ = [("N", C.c_int64), ("ng", C.c_int32), ("flags", C.c_int32),
                ("dx", C.c_double), ("dt", C.c_double), ("length", C.c_double),
                ("B", C.c_double * 3), ("Eyz", C.c_double * 2)]


P = C.c_void_p          # device or host pointer passed as an integer address
I64, I32, F64 = C.c_int64, C.c_int, C.c_double
R7 = C.c_void_p * 7

# name -> argtypes ; every function returns int
_SIGS = {
    "pic_device_info": [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, I32],
    "pic_dev_read": [P, P, I64, P],
    "pic_dev_write": [P, P, I64, P],
    "pic_dev_zero": [P, I64, P],
    "pic_dev_copy": [P, P, I64, P],
    "pic_stream_sync": [P],
    "pic_dev_dd_step_begin": [P, P, I32, P, P, I64, P, P],
    "pic_dev_dd_step_prologue": [C.POINTER(DDParams), C.POINTER(DDPrologue), P],
    "pic_host_release": [],
    "pic_dev_smooth": [P, P, I32, I32, P],
    "pic_dev_differentiate": [P, P, I32, F64, I32, P],
    "pic_dev_integrate_field": [P, P, I32, F64, I32, P],
    "pic_dev_shift_extreme": [P, P, I32, I32, P],
    "pic_dev_tridiag_pcr": [P, P, P, P, P, I32, P, P],
    "pic_dev_poisson_periodic": [P, P, I32, F64, I32, P, P],
    "pic_dev_poisson_dirichlet": [P, P, I32, F64, P, P],
    "pic_dev_newton_boltzmann": [P, P, I32, F64, F64, F64, I32, F64, I32, P, P],
    "pic_dev_newton_boltzmann_l": [P, P, I32, F64, F64, F64, I32, I32, P, P, P],
    "pic_dev_dd_interpolate": [P, P, P, I64, I32, F64, P, P],
    "pic_dev_dd_weight": [P, P, P, P, P, I64, I32, F64, F64, F64, P, P],
    "pic_dev_dd_picard_iter": [C.POINTER(DDParams), P, P, P, P, P, P, P, I32, P, P],
    "pic_dev_dd_picard_iter2": [C.POINTER(DDParams), P, P, P, P, P, P, P, P, I32, P, P],
    "pic_dev_dd_picard_iter3": [C.POINTER(DDParams), P, P, P, P, P, P, P, P, I32, P, P, P],
    "pic_dev_dd_picard_iter4": [C.POINTER(DDParams), P, P, P, P, P, P, P, P, I32, P, P, P, I32, P, I32, P],
    "pic_dev_dd_picard_iter5": [C.POINTER(DDParams), P, P, P, P, P, P, P, P, I32, P, P, P, I32, P, I32, P, P],
    "pic_dev_dd_apply_draws3": [P, P, P, P, P, P, I64, P, P, P, P, P, P, P],
    "pic_dev_dd_sort_by_cell2": [C.POINTER(DDParams), P, P, P, P, P, P, P, P],
    "pic_dev_dd_apply_draws2": [P, P, P, P, P, P, I64, P, P, P, P, P, P],
    "pic_dev_dd_reinject_philox_log": [C.POINTER(DDParams), P, I32, P, P, P, P, P, P, C.POINTER(C.c_double * 2),
                                       C.c_uint64, C.c_uint64, I64, P],
    "pic_dev_dd_reinject_philox2": [C.POINTER(DDParams), P, P, P, P, P, P, C.POINTER(C.c_double * 2), C.c_uint64,
                                    C.c_uint64, I64, P, I32, P],
    "pic_dev_dd_thermostat_philox": [C.POINTER(DDParams), P, P, P, P, P, F64, C.POINTER(C.c_double * 2), C.c_uint64,
                                     C.c_uint64, I64, P],
    "pic_dev_gather_i32": [P, P, P, I64, P],
    "pic_dev_scatter_f64": [P, P, P, I64, P],
    "pic_dev_scatter_i8": [P, P, P, I64, P],
    "pic_dev_invert_perm": [P, P, I64, P],
    "pic_dev_moments": [P, I64, P, P],
    "pic_dev_dd_commit_u": [C.POINTER(DDParams), P, P, P, P, P, P, P, I32, P, P],
    "pic_dev_dd_commit_u2": [C.POINTER(DDParams), P, P, P, P, P, P, P, I32, P, P, P],
    "pic_dev_dd_j1_finish": [C.POINTER(DDParams), P, P, P, P, P],
    "pic_dev_debug_cta_timer": [P],
    "pic_dev_selftest_div": [F64, C.c_uint64, C.c_uint64, P, P],
    "pic_dev_dd_field_update": [C.POINTER(DDParams), P, P, P, P, P, P, P, P],
    "pic_dev_dd_field_update2": [C.POINTER(DDParams), P, P, P, P, P, P, P, P, P, P, F64, I32, P],
    "pic_slab_message_len": [I32],
    "pic_slab_work_len": [],
    "pic_dev_slab_pack": [C.POINTER(DDParams), I32, I32, I32, I32, I32, P, P, P, P, P, P],
    "pic_dev_slab_field_update": [C.POINTER(DDParams), I32, I32, I32, I32, I32, P, P, P, P, P, P, P, P, P, P, P],
    "pic_dev_slab_finish": [C.POINTER(DDParams), I32, I32, I32, I32, I32, P, P, P, P, P, P, F64, I32, P],
    "pic_p2p_alloc": [I64, I32, C.POINTER(C.c_void_p), P],
    "pic_p2p_open": [P, C.POINTER(C.c_void_p)],
    "pic_p2p_set_timeout": [I32],
    "pic_p2p_close": [P],
    "pic_p2p_free": [P],
    "pic_dev_p2p_reduce": [P, I32, I32, C.c_uint32, I64, P, P, P],
    "pic_dev_dd_field_update_p2p": [C.POINTER(DDParams), P, I32, I32, C.c_uint32, P, P, P, P, P, P, P, P, P, P, F64, I32, P, P],
    "pic_dev_dd_apply_draws": [P, P, P, P, P, I64, P, P, P, P, P, P],
    "pic_dev_dd_reinject_philox": [C.POINTER(DDParams), P, P, P, P, P, C.POINTER(C.c_double * 2), C.c_uint64,
                                   C.c_uint64, I64, P],
    "pic_dev_sum_sq": [P, I64, F64, P, P],
    "pic_dev_dd_sort_by_cell": [C.POINTER(DDParams), P, P, P, P, P, P, P, P, P, P],
    "pic_dev_dd_sort_by_cell_stable": [C.POINTER(DDParams), P, P, P, P, P, I64, C.POINTER(C.c_int), P],
    "pic_dev_dd_sort_by_cell_stable2": [C.POINTER(DDParams), P, P, P, P, P, P, I32, P, I64, C.POINTER(C.c_int), P],
    "pic_dev_sort_perm_by_cell": [C.POINTER(DDParams), P, P, P, P, P],
    "pic_dev_sort_by_cell_payload": [C.POINTER(DDParams), P, P, P, P, P, P, P, P, P, P, P],
    "pic_dev_soa_permute": [P, I64, P, P, I32, P, P, I32, P, P, I32, P],
    "pic_dev_pypic_interpolate": [P, P, P, I64, I32, F64, P, P],
    "pic_dev_pypic_weight": [P, P, P, P, I64, I32, F64, F64, P, P],
    "pic_dev_pypic_weight_fixed": [P, P, P, P, I64, I32, F64, F64, F64, P, P],
    "pic_dev_pypic_picard_iter": [C.POINTER(PypicParams), P, P, P, P, P, P, I32, P, P],
    "pic_dev_pypic_field_update": [C.POINTER(PypicParams), P, P, P, P, P, P, P, P],
    "pic_dev_pypic_picard_iter2": [C.POINTER(PypicParams), P, P, P, P, P, P, P, I32, P, P],
    "pic_dev_pypic_picard_iter3": [C.POINTER(PypicParams), P, P, P, P, P, P, P, I32, P, P, P],
    "pic_dev_pypic_picard_iter_qm": [C.POINTER(PypicParams), P, P, P, P, P, P, P, P, P, I32, P, P, P],
    "pic_dev_pypic_field_update2": [C.POINTER(PypicParams), P, P, P, P, P, P, P, P, P, P, F64, I32, P],
    "pic_dev_pypic_j1_repair": [C.POINTER(PypicParams), P, P, P, P, P, P, I32, P, P, P],
    "pic_dev_pypic_j1_finish": [C.POINTER(PypicParams), P, P, P, P],
    "pic_dev_wrap_periodic": [P, I64, F64, P],
    "pic_dev_l_interpolate": [P, P, P, I64, I32, F64, P, P],
    "pic_dev_l_weight": [P, P, P, P, I64, I32, F64, F64, P, P],
    "pic_dev_l_weight_bounded": [P, P, P, P, I64, I32, F64, F64, P, P],
    "pic_dev_l_push_implicit": [P, P, P, P, P, P, P, P, I64, I32, F64, F64, P, P],
    "pic_dev_l_outside_flags": [P, P, I64, F64, P],
    "pic_dev_l_push_deposit": [C.POINTER(LParams), P, P, P, P, P, P],
    "pic_dev_l_deposit_fixed": [C.POINTER(LParams), P, P, P, P],
    "pic_dev_l_field_solve": [C.POINTER(LParams), P, P, P, P, P, P, P],
    "pic_dev_gc_interpolate": [P, P, P, I64, I32, F64, P, P],
    "pic_dev_gc_weight": [P, P, P, P, P, P, I64, I32, F64, P, P],
    "pic_dev_gc_push_boris": [C.POINTER(GCParams), C.POINTER(R7), P, P, P, P, P, P, P, P, P],
    "pic_dev_gc_push_boris_uniform": [C.POINTER(GCParams), C.POINTER(R7), F64, F64, F64, P, P, P, P, P, P, P, P],
    "pic_dev_gc_push_boris_uniform2": [C.POINTER(GCParams), C.POINTER(R7), F64, F64, F64, I32, F64, P, P, P, P, P, P, P, P],
    "pic_dev_gc_push_boris_mixed": [C.POINTER(GCParams), C.POINTER(R7), P, P, P, I32, F64, P, P, P, P, P, P, P, P, P],
    "pic_dev_gc_post_push": [P, P, P, P, P, P, P, I32, F64, F64, F64, C.POINTER(C.c_double * 4), I32, P, P, P, P, I64, P, P],
    "pic_dev_gc_uniform_finish": [P, P, P, I32, F64, P],
    "pic_dev_gc_deposit_idx": [P, P, I64, F64, F64, I32, P, P, P],
    "pic_dev_gc_apply_bcs": [P, P, P, I64, F64, P],
    "pic_dev_gc_to_gc": [C.POINTER(GCParams), C.POINTER(R7), P, P, P, P],
    "pic_dev_gc_to_6d": [C.POINTER(GCParams), C.POINTER(R7), P, P, P, P, P, P, P],
    "pic_dev_gc_push_rk4": [C.POINTER(GCParams), C.POINTER(R7), P, P, P, P, P, P],
    "pic_dev_gc_push_rk4_uniform": [C.POINTER(GCParams), C.POINTER(R7), F64, F64, P, P, P, P],
    "pic_dev_gc_n0_update": [P, P, P, I32, F64, F64, F64, F64, P, P],
    "pic_dev_gc_decide": [P, P, P, P, I64, I64, P, P, P, P],
    "pic_dev_compact_flags": [P, I64, I32, P, P, P, P],
    "pic_dev_gather_f64": [P, P, P, I64, P],
    "pic_dev_gather_i8": [P, P, P, I64, P],
    "pic_dev_init_uniform_maxwellian": [P, P, P, P, I64, I64, F64, F64, C.POINTER(C.c_double * 2),
                                        C.POINTER(C.c_double * 2), C.c_uint64, C.c_uint64, I64, P],
    "pic_dev_pypic_perturb_positions": [P, I64, P, P, I32, C.c_uint64, I64, P],
    "pic_dev_gc_iead_hist": [P, P, P, P, P, P, I32, I64, P, I32, P, I32, P, P],
    "pic_mt_jump_poly": [C.c_uint64, P],
    "pic_mt_jump": [P, C.POINTER(C.c_int32), P],
    "pic_mt_skip": [P, C.POINTER(C.c_int32), C.c_uint64],
    "pic_mt_sheath_draws": [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), I64, P, F64, P, P, P, P],
    "pic_mt_sheath_thermostat": [P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double), I64, I64, F64, F64,
                                 F64, I64, P, P, P, P, C.POINTER(C.c_int64)],
    "pic_host_pypic_interpolate_p": [P, P, I32, I64, F64, P],
    "pic_host_pypic_weight_current_p": [P, P, P, I32, I32, I64, F64, P],
    "pic_host_pypic_weight_density_p": [P, P, I32, I32, I64, F64, P],
    "pic_host_dd_interpolateField": [P, P, I32, I64, F64, P],
    "pic_host_dd_weightCurrents": [P, P, P, F64, I32, I64, F64, F64, P, P],
    "pic_host_dd_weightDensities": [P, P, F64, I32, I64, F64, P, P],
    "pic_host_dd_step": [C.POINTER(DDParams), P, P, P, F64, I32, P, P, P, P, P, C.POINTER(C.c_int),
                         C.POINTER(C.c_double)],
    "pic_host_dd_step_batches": [C.POINTER(DDParams), I32, P, P, P, F64, I32, P, P, P, P, P, P, P],
}
EXPORTS = sorted(list(_SIGS) + ["pic_last_error", "pic_version"])

_lib = None


def load():
    """Loads the shared library (once).  Raises PicError if it is missing."""
    global _lib, LIB_PATH
    if _lib is not None:
        return _lib
    if os.environ.get("PIC_LIB_PATH"):             # A/B runs of differently built libraries (tools/gpu_runs)
        LIB_PATH = os.environ["PIC_LIB_PATH"]
    if not os.path.isfile(LIB_PATH):
        raise PicError(PIC_ERR_NODEVICE,
                       "%s not found -- run `python -m pypic_b200.build` (or __graft_entry__.build()); "
                       "there is no CPU fallback" % LIB_PATH)
    try:
        import torch  # noqa: F401  -- loads torch's libcudart first so both share one CUDA runtime
    except Exception:
        pass
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.pic_last_error.restype = C.c_char_p
    lib.pic_last_error.argtypes = []
    lib.pic_version.restype = C.c_int
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Calls an ABI function and raises PicError on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise PicError(rc, lib.pic_last_error().decode())
    return rc
