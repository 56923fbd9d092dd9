"""ctypes binding of libclimaseaice_b200.so (include/climaseaice_b200.h).

The shared library is the product; this module only loads it and mirrors its structs.  There is
no fallback: if the library is missing, import of the compute entry points raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
import os

LIB_PATH = Path(os.environ.get("CSI_B200_LIB", str(_HERE / "libclimaseaice_b200.so")))  # override only for kernel-variant experiments

ABI_VERSION = 1
PERIODIC, BOUNDED, FOLDED = 0, 1, 2
STRESS_NONE, STRESS_CONST, STRESS_FIELD, STRESS_SEMI_IMPLICIT = 0, 1, 2, 3
REPLACEMENT_PRESSURE, ICE_STRENGTH = 0, 1
CORIOLIS_NONE, CORIOLIS_FPLANE, CORIOLIS_SPHERICAL = 0, 1, 2
BC_DEFAULT, BC_VALUE = 0, 1
RK3, FE = 0, 1
SOLVER_AUTO, SOLVER_UNFUSED, SOLVER_FUSED = 0, 1, 2
METRIC_REGULAR, METRIC_J, METRIC_IJ = 0, 1, 2
FD_NONE, FD_FIELDS, FD_STRESS_BALANCE = 0, 1, 2

ERRORS = {-1: "CSI_ERR_ARG", -2: "CSI_ERR_SHAPE", -3: "CSI_ERR_UNSUPPORTED", -4: "CSI_ERR_NO_DEVICE", -5: "CSI_ERR_NCCL_MISSING"}


class CsiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libclimaseaice_b200: {ERRORS.get(code, code)}: {msg}")
        self.code = code


class csi_array(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("nx_tot", C.c_int32), ("ny_tot", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32)]


class csi_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("Nx", C.c_int32), ("Ny", C.c_int32), ("Hx", C.c_int32), ("Hy", C.c_int32),
        ("topo_x", C.c_int32), ("topo_y", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double),
        ("immersed_mask", C.c_void_p),
        ("ice_compressive_strength", C.c_double), ("ice_compaction_hardening", C.c_double),
        ("yield_curve_eccentricity", C.c_double), ("minimum_plastic_stress", C.c_double),
        ("min_relaxation_parameter", C.c_double), ("max_relaxation_parameter", C.c_double),
        ("relaxation_strength", C.c_double),
        ("pressure_formulation", C.c_int32), ("substeps", C.c_int32),
        ("minimum_mass", C.c_double), ("minimum_concentration", C.c_double), ("ice_density", C.c_double),
        ("coriolis_kind", C.c_int32), ("top_stress_kind", C.c_int32),
        ("coriolis_f", C.c_double), ("top_tau_x", C.c_double), ("top_tau_y", C.c_double),
        ("bottom_stress_kind", C.c_int32), ("u_south_north_bc", C.c_int32),
        ("rho_e", C.c_double), ("Cd", C.c_double), ("ue_const", C.c_double), ("ve_const", C.c_double),
        ("u_south_north_value", C.c_double),
        ("v_west_east_bc", C.c_int32), ("advection_order", C.c_int32),
        ("v_west_east_value", C.c_double),
        ("timestepper", C.c_int32), ("solver_impl", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("exchange_every", C.c_int32), ("partition_x", C.c_int32),
        ("immersed_drag_u", C.c_double), ("immersed_drag_v", C.c_double),
        ("metric_kind", C.c_int32), ("serial_exchange", C.c_int32), ("metrics", C.POINTER(C.c_double) * 12),
        ("free_drift_kind", C.c_int32), ("reserved3_", C.c_int32), ("top_rho_e", C.c_double), ("top_Cd", C.c_double),
        ("coriolis_f_ff", C.POINTER(C.c_double)),
        ("fold_target", C.POINTER(C.c_int32) * 4), ("fold_source", C.POINTER(C.c_int32) * 4), ("fold_count", C.c_int32 * 4),
        ("fold_sign_velocity", C.c_double), ("fold_sign_external", C.c_double),
    ]


FIELD_NAMES = ("u", "v", "h", "a", "s11", "s22", "s12", "zeta_f", "zeta_c", "delta", "alpha", "un", "vn", "P",
               "top_x", "top_y", "ue", "ve", "Gh", "Ga", "hm", "am", "um", "vm", "hs", "Ghs", "hsm", "fd_u", "fd_v")
# (face_x, face_y) of every field, same order
FIELD_LOC = dict(u=(1, 0), v=(0, 1), h=(0, 0), a=(0, 0), s11=(0, 0), s22=(0, 0), s12=(1, 1), zeta_f=(1, 1),
                 zeta_c=(0, 0), delta=(0, 0), alpha=(0, 0), un=(1, 0), vn=(0, 1), P=(0, 0), top_x=(1, 0),
                 top_y=(0, 1), ue=(1, 0), ve=(0, 1), Gh=(0, 0), Ga=(0, 0), hm=(0, 0), am=(0, 0), um=(1, 0), vm=(0, 1),
                 hs=(0, 0), Ghs=(0, 0), hsm=(0, 0), fd_u=(1, 0), fd_v=(0, 1))


class csi_fields(C.Structure):
    _fields_ = [(n, csi_array) for n in FIELD_NAMES]


TOP_FLUX_BALANCE, TOP_PRESCRIBED = 0, 1
BOTTOM_EQUILIBRIUM, BOTTOM_PRESCRIBED = 0, 1
FLUX_CONST, FLUX_ARRAY, FLUX_RADIATIVE_EMISSION, FLUX_CONDUCTIVE, FLUX_LINEAR = 0, 1, 2, 3, 4


class csi_thermo_config(C.Structure):
    _fields_ = [
        ("density", C.c_double), ("heat_capacity", C.c_double), ("liquid_density", C.c_double),
        ("liquid_heat_capacity", C.c_double), ("reference_latent_heat", C.c_double), ("reference_temperature", C.c_double),
        ("liquidus_freshwater_melting_temperature", C.c_double), ("liquidus_slope", C.c_double),
        ("top_heat_bc", C.c_int32), ("snow_top_heat_bc", C.c_int32), ("bottom_heat_bc", C.c_int32), ("layered", C.c_int32),
        ("ice_conductivity", C.c_double), ("snow_conductivity", C.c_double),
        ("bottom_salinity", C.c_double), ("bottom_temperature", C.c_double),
        ("n_top_terms", C.c_int32), ("top_term_kind", C.c_int32 * 2), ("reserved_", C.c_int32),
        ("top_flux_const", C.c_double), ("emissivity", C.c_double), ("stefan_boltzmann_constant", C.c_double),
        ("emission_reference_temperature", C.c_double), ("bottom_flux_const", C.c_double),
        ("snowfall", C.c_double), ("snow_density", C.c_double), ("ice_consolidation_thickness", C.c_double), ("ice_salinity", C.c_double),
        ("secant_tolerance", C.c_double), ("secant_maxiters", C.c_int32), ("reserved2_", C.c_int32),
        ("linear_coefficient", C.c_double), ("linear_temperature", C.c_double),
        ("linear_times_concentration", C.c_int32), ("reserved3_", C.c_int32),
    ]


THERMO_FIELD_NAMES = ("h", "a", "hs", "Tu", "Tus", "S", "hc", "Qtop", "Qbot", "Sb", "Tb", "snowfall", "rho_s",
                      "mf_ice", "mf_snow", "mf_snowfall")


class csi_thermo_fields(C.Structure):
    _fields_ = [(n, csi_array) for n in THERMO_FIELD_NAMES]


# every symbol include/climaseaice_b200.h declares
EXPORTS = (
    "csi_version", "csi_last_error", "csi_create", "csi_destroy", "csi_evp_substeps",
    "csi_compute_tracer_tendencies", "csi_dynamic_time_step", "csi_cache_current_fields", "csi_update_state",
    "csi_fill_halos", "csi_time_step", "csi_cell_advection_timescale", "csi_diagnostics", "csi_time_step_host",
    "csi_evp_substeps_host", "csi_last_transfer_bytes", "csi_nccl_unique_id", "csi_comm_init", "csi_exchange_halos", "csi_exchange_halos_async", "csi_wait_halos", "csi_launch_count", "csi_fused_stats",
    "csi_last_elapsed_ms", "csi_time_dominant_kernel", "csi_thermodynamic_time_step", "csi_attach_thermodynamics", "csi_selftest_math", "csi_measure_fp64_rate", "csi_host_exp", "csi_host_div_by_const", "csi_host_halo_width",
)

_lib = None


def lib():
    """Load the shared library (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(libclimaseaice_b200 has no Python/CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    H = C.c_void_p
    L.csi_version.restype = C.c_int
    L.csi_last_error.restype = C.c_char_p
    L.csi_last_error.argtypes = [H]
    L.csi_create.argtypes = [C.POINTER(csi_config), C.POINTER(H)]
    L.csi_destroy.argtypes = [H]
    L.csi_evp_substeps.argtypes = [H, C.POINTER(csi_fields), C.c_double, C.c_int32, C.c_void_p]
    L.csi_compute_tracer_tendencies.argtypes = [H, C.POINTER(csi_fields), C.c_void_p]
    L.csi_dynamic_time_step.argtypes = [H, C.POINTER(csi_fields), C.c_double, C.c_void_p]
    L.csi_cache_current_fields.argtypes = [H, C.POINTER(csi_fields), C.c_void_p]
    L.csi_update_state.argtypes = [H, C.POINTER(csi_fields), C.c_void_p]
    L.csi_thermodynamic_time_step.argtypes = [H, C.POINTER(csi_thermo_config), C.POINTER(csi_thermo_fields), C.c_double, C.c_void_p]
    L.csi_attach_thermodynamics.argtypes = [H, C.POINTER(csi_thermo_config), C.POINTER(csi_thermo_fields)]
    L.csi_fill_halos.argtypes = [H, C.POINTER(csi_array), C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.csi_time_step.argtypes = [H, C.POINTER(csi_fields), C.c_double, C.c_int32, C.c_void_p]
    L.csi_cell_advection_timescale.argtypes = [H, C.POINTER(csi_fields), C.POINTER(C.c_double), C.c_void_p]
    L.csi_diagnostics.argtypes = [H, C.POINTER(csi_fields), C.POINTER(C.c_double), C.c_void_p]
    L.csi_time_step_host.argtypes = [H, C.POINTER(csi_fields), C.c_double, C.c_int32, C.c_int32]
    L.csi_evp_substeps_host.argtypes = [H, C.POINTER(csi_fields), C.c_double, C.c_int32]
    L.csi_last_transfer_bytes.argtypes = [H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.csi_nccl_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.csi_comm_init.argtypes = [H, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]
    L.csi_exchange_halos.argtypes = [H, C.POINTER(csi_array), C.c_int32, C.c_int32, C.c_void_p]
    L.csi_exchange_halos_async.argtypes = [H, C.POINTER(csi_array), C.c_int32, C.c_int32, C.c_void_p]
    L.csi_wait_halos.argtypes = [H, C.c_void_p]
    L.csi_launch_count.restype = C.c_int64
    L.csi_launch_count.argtypes = [H]
    L.csi_fused_stats.argtypes = [H, C.POINTER(C.c_int64)]
    L.csi_last_elapsed_ms.restype = C.c_double
    L.csi_last_elapsed_ms.argtypes = [H]
    L.csi_time_dominant_kernel.argtypes = [H, C.POINTER(csi_fields), C.c_double, C.c_int32, C.POINTER(C.c_double), C.c_char_p,
                                           C.POINTER(C.c_int32), C.c_void_p]
    L.csi_selftest_math.argtypes = [C.c_int64, C.c_uint64, C.c_int32, C.POINTER(C.c_uint64)]
    L.csi_measure_fp64_rate.argtypes = [C.c_int32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.csi_host_exp.restype = C.c_double
    L.csi_host_exp.argtypes = [C.c_double]
    L.csi_host_div_by_const.restype = C.c_double
    L.csi_host_div_by_const.argtypes = [C.c_double, C.c_double]
    L.csi_host_halo_width.argtypes = [C.c_int32]
    _lib = L
    return L


def check(rc, handle=None):
    if rc != 0:
        msg = lib().csi_last_error(handle)
        raise CsiError(rc, msg.decode() if msg else "")
