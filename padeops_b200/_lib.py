"""ctypes loader for libpadeops_b200.so — prototypes mirror include/padeops_b200.h one to one."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "lib", "libpadeops_b200.so")
_lib = None

c_dp = C.c_void_p  # field pointers travel as raw addresses (host or device)


class PadeOpsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"padeops_b200 error {code}: {msg}")
        self.code = code


class DecompInfo(C.Structure):
    _fields_ = [(nm, C.c_int * 3) for nm in ("xst", "xen", "xsz", "yst", "yen", "ysz", "zst", "zen", "zsz")]


def library_path():
    return _SO


def build_library(force=False):
    """Compile every CUDA source for sm_100a into lib/libpadeops_b200.so (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-s", "clean"])
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-s", "-j8", "all"])
    return _SO


_OPS7 = [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
_PROTOS = {
    "pdo_last_error": (C.c_char_p, []),
    "pdo_version": (C.c_int, []),
    "pdo_launch_count": (C.c_int64, []),
    "pdo_malloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "pdo_free": (C.c_int, [C.c_void_p]),
    "pdo_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdo_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "pdo_stream_sync": (C.c_int, [C.c_void_p]),
    "pdo_cd10_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]),
    "pdo_cd10_destroy": (C.c_int, [C.c_void_p]),
    "pdo_cd10_getsize": (C.c_int, [C.c_void_p]),
    "pdo_cd06_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]),
    "pdo_cd06_destroy": (C.c_int, [C.c_void_p]),
    "pdo_cd06_getsize": (C.c_int, [C.c_void_p]),
    "pdo_cf90_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "pdo_cf90_destroy": (C.c_int, [C.c_void_p]),
    "pdo_gaussian_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "pdo_gaussian_destroy": (C.c_int, [C.c_void_p]),
    "pdo_cd10_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pdo_cd06_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "pdo_cf90_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "pdo_gaussian_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "pdo_plan_on_first_call": (C.c_int, [C.c_int]),
    "pdo_cd06stagg_init_periodic": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_double]),
    "pdo_cd06stagg_destroy": (C.c_int, [C.c_void_p]),
    "pdo_derivatives_init": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                       C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p,
                                       C.c_char_p]),
    "pdo_derivatives_destroy": (C.c_int, [C.c_void_p]),
    "pdo_filters_init": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                   C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]),
    "pdo_filters_destroy": (C.c_int, [C.c_void_p]),
    "pdo_comm_unique_id": (C.c_int, [C.c_char_p]),
    "pdo_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_char_p]),
    "pdo_comm_finalize": (C.c_int, []),
    "pdo_comm_register_buffer": (C.c_int, [C.c_void_p, C.c_size_t]),
    "pdo_comm_deregister_buffer": (C.c_int, [C.c_void_p]),
    "pdo_comm_p2p_enabled": (C.c_int, []),
    "pdo_comm_rank": (C.c_int, []),
    "pdo_comm_size": (C.c_int, []),
    "pdo_decomp_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pdo_decomp_destroy": (C.c_int, [C.c_void_p]),
    "pdo_decomp_get_info": (C.c_int, [C.c_void_p, C.POINTER(DecompInfo)]),
    "pdo_decomp_info_for": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(DecompInfo)]),
    "pdo_p_maxval": (C.c_int, [C.c_double, C.POINTER(C.c_double)]),
    "pdo_p_sum": (C.c_int, [C.c_double, C.POINTER(C.c_double)]),
    "pdo_fft3d_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                 C.c_int]),
    "pdo_fft3d_destroy": (C.c_int, [C.c_void_p]),
    "pdo_fft3d_get_complex_output_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "pdo_fft3d_get_spectral_info": (C.c_int, [C.c_void_p, C.POINTER(DecompInfo)]),
    "pdo_fft3d_get_physical_info": (C.c_int, [C.c_void_p, C.POINTER(DecompInfo)]),
    "pdo_fft3d_fft3_x2z": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_fft3d_ifft3_z2x": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_fft3d_fft2_x2y": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_fft3d_ifft2_y2x": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_void_p]),
    "pdo_poisson_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdo_poisson_destroy": (C.c_int, [C.c_void_p]),
    "pdo_poisson_solve": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
}
for _t, _fns in (("cd10", ("dd1", "dd2", "dd3", "d2d1", "d2d2", "d2d3")), ("cd06", ("dd1", "dd2", "dd3")),
                 ("cf90", ("filter1", "filter2", "filter3")), ("gaussian", ("filter1", "filter2", "filter3"))):
    for _f in _fns:
        _PROTOS[f"pdo_{_t}_{_f}"] = (C.c_int, _OPS7)
for _f in ("ddz_E2C", "ddz_C2E", "interpz_E2C", "interpz_C2E", "d2dz2_C2C", "d2dz2_E2E"):
    _PROTOS[f"pdo_cd06stagg_{_f}"] = (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_void_p])
for _f in ("ddx", "ddy", "ddz", "d2dx2", "d2dy2", "d2dz2"):
    _PROTOS[f"pdo_derivatives_{_f}"] = (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_void_p])
for _f in ("filterx", "filtery", "filterz"):
    _PROTOS[f"pdo_filters_{_f}"] = (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_void_p])
for _f in ("x_to_y", "y_to_x", "y_to_z", "z_to_y"):
    _PROTOS[f"pdo_transpose_{_f}"] = (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_void_p])

class IgridParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
                ("Re", C.c_double), ("is_inviscid", C.c_int), ("dealias_fact", C.c_double), ("t_divergence_check", C.c_int),
                ("time_stepping_scheme", C.c_int), ("p_row", C.c_int), ("p_col", C.c_int), ("use_d2dz2_c2c", C.c_int),
                ("compute_all_gradients", C.c_int), ("rotational_advection", C.c_int), ("fourier_collocation_z", C.c_int),
                ("wall_bounded", C.c_int), ("top_wall", C.c_int), ("bot_wall", C.c_int), ("no_stokes_pressure", C.c_int)]


_PROTOS.update({
    "pdo_spectral_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_double]),
    "pdo_spectral_destroy": (C.c_int, [C.c_void_p]),
    "pdo_spectral_get_physical_info": (C.c_int, [C.c_void_p, C.POINTER(DecompInfo)]),
    "pdo_spectral_get_spectral_info": (C.c_int, [C.c_void_p, C.POINTER(DecompInfo)]),
    "pdo_spectral_fft": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_spectral_ifft": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_void_p]),
    "pdo_spectral_mtimes_ik1_oop": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_spectral_mtimes_ik2_oop": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_spectral_mtimes_ik1_ip": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_mtimes_ik2_ip": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_dealias": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_dealias_edgefield": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_take_fft1d_z2z_ip": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_take_ifft1d_z2z_ip": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_debug_ztables": (C.c_int, [C.c_int, C.c_double, c_dp]),
    "pdo_debug_fft_plan": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "pdo_debug_fft_tables": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "pdo_lstsq_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "pdo_lstsq_destroy": (C.c_int, [C.c_void_p]),
    "pdo_lstsq_filter1": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_void_p]),
    "pdo_lstsq_filter2": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_void_p]),
    "pdo_lstsq_filter3": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_void_p]),
    "pdo_padepoisson_init2": (C.c_int, [C.POINTER(C.c_void_p), C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "pdo_padepoisson_init3": (C.c_int, [C.POINTER(C.c_void_p), C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double]),
    "pdo_igrid_enable_sgs": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "pdo_debug_sgs_point": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp]),
    "pdo_hit_forcing_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int]),
    "pdo_hit_forcing_destroy": (C.c_int, [C.c_void_p]),
    "pdo_hit_forcing_set_wavenumbers": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pdo_hit_forcing_get_wavenumbers": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pdo_hit_forcing_get_rhs": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, C.c_int, C.c_void_p]),
    "pdo_igrid_enable_hit_forcing": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int]),
    "pdo_igrid_hit_forcing": (C.c_void_p, [C.c_void_p]),
    "pdo_debug_hit_draw": (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pdo_cd06stagg_init_nonperiodic": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pdo_cd06stagg_ddz_C2C": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pdo_cd06stagg_ddz_E2E": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "pdo_debug_stagg_np_host": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_longlong]),
    "pdo_debug_stagg_np_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp]),
    "pdo_decomp_write_one": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int, C.c_char_p]),
    "pdo_decomp_read_one": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int, C.c_char_p]),
    "pdo_io_write_block": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, c_dp, C.c_int]),
    "pdo_io_read_block": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, c_dp]),
    "pdo_io_format_g15_5": (C.c_int, [C.c_double, C.c_char_p]),
    "pdo_igrid_dump_restart": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "pdo_igrid_read_restart": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "pdo_igrid_dump_full_field": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int]),
    "pdo_spectral_ddz_c2c_real_ip": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_ddz_c2c_complex_ip": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_shiftz_e2c": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_spectral_shiftz_c2e": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_ops_periodic_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]),
    "pdo_ops_periodic_destroy": (C.c_int, [C.c_void_p]),
    "pdo_ops_periodic_spect": (C.c_void_p, [C.c_void_p]),
    "pdo_ops_periodic_ddx": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_ops_periodic_ddy": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_ops_periodic_ddz": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_ops_periodic_ddz_cmplx2cmplx": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_ops_periodic_solve_poisson": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_ops_periodic_dealias_field": (C.c_int, [C.c_void_p, c_dp, C.c_void_p]),
    "pdo_ops_periodic_write_field3d": (C.c_int, [C.c_void_p, c_dp, C.c_char_p, C.c_int, C.c_int, C.c_char_p]),
    "pdo_ops_periodic_read_field3d": (C.c_int, [C.c_void_p, c_dp, C.c_char_p, C.c_int, C.c_int, C.c_char_p]),
    "pdo_spectral_get_tables": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "pdo_pade6stagg_init": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_double, C.c_int, C.c_int]),
    "pdo_pade6stagg_init2": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_double, C.c_int, C.c_int, C.c_void_p]),
    "pdo_pade6stagg_destroy": (C.c_int, [C.c_void_p]),
    "pdo_pade6stagg_get_modified_wavenumbers": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int]),
    "pdo_padepoisson_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pdo_padepoisson_destroy": (C.c_int, [C.c_void_p]),
    "pdo_padepoisson_pressure_projection": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, C.c_void_p]),
    "pdo_padepoisson_get_pressure": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "pdo_padepoisson_get_pressure_and_update_rhs": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "pdo_padepoisson_divergence_check": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "pdo_igrid_init": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(IgridParams), c_dp, c_dp, c_dp]),
    "pdo_igrid_destroy": (C.c_int, [C.c_void_p]),
    "pdo_igrid_time_advance": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "pdo_igrid_get_field": (C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_void_p]),
    "pdo_igrid_get_decomp_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(DecompInfo)]),
    "pdo_igrid_get_state": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "pdo_igrid_compute_delta_t": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.c_void_p]),
    "pdo_igrid_max_divergence": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
})
for _f in ("ddz_C2E", "ddz_E2C", "interpz_C2E", "interpz_E2C", "d2dz2_C2C", "d2dz2_E2E"):
    _PROTOS[f"pdo_pade6stagg_{_f}"] = (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_int, C.c_int, C.c_int, C.c_void_p])
_PROTOS.update({
    "pdo_operators_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_char_p, C.c_int]),
    "pdo_operators_destroy": (C.c_int, [C.c_void_p]),
    "pdo_operators_zmode": (C.c_int, [C.c_void_p]),
    "pdo_operators_ddx": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_operators_ddy": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_operators_ddz": (C.c_int, [C.c_void_p, c_dp, c_dp, C.c_void_p]),
    "pdo_operators_gradient": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "pdo_operators_curl": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "pdo_operators_divergence": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp, C.c_void_p]),
    "pdo_operators_filter3d": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]),
})
_PROTOS["pdo_debug_zslab_emulate"] = (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, C.c_longlong, C.c_int, C.c_int, C.c_void_p])
_PROTOS["pdo_debug_cd10_generic"] = (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.c_void_p])

class ChunkTables(C.Structure):
    """Mirror of pdo::ChunkTables (csrc/tables.h): kMaxChunk = 32, kMaxW = 16."""
    _fields_ = [("l1", C.c_double * 32), ("l2", C.c_double * 32), ("ginv", C.c_double * 32), ("ug", C.c_double * 32),
                ("bg", C.c_double * 32), ("V", (C.c_double * 2) * 32), ("U", (C.c_double * 2) * 32), ("G", (C.c_double * 4) * 33),
                ("b1", C.c_double), ("b2", C.c_double), ("W", C.c_int), ("dense", C.c_int), ("n", C.c_int), ("M", C.c_int),
                ("P", C.c_int), ("BW", C.c_int)]


_PROTOS["pdo_debug_np_chunk_tables"] = (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.POINTER(C.c_int)])
_PROTOS["pdo_debug_ctma_config"] = (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)])
_PROTOS["pdo_debug_igrid_bcs"] = (C.c_int, [C.c_int, C.c_int, C.c_void_p])
_PROTOS["pdo_debug_np_fast"] = (C.c_int, [C.c_int])
_PROTOS["pdo_debug_np_rows"] = (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_dp])
_PROTOS["pdo_debug_np_line_host"] = (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_longlong, C.c_longlong])
_PROTOS["pdo_debug_chunk_tables"] = (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int])
_PROTOS["pdo_debug_transpose_emulate"] = (C.c_int, [C.c_int] * 8 + [C.c_void_p, C.c_void_p, C.c_void_p])
_PROTOS["pdo_debug_set_variant"] = (C.c_int, [C.c_int, C.c_int])
_PROTOS["pdo_debug_last_variant"] = (C.c_int, [])

EXPORTED = sorted(k for k in _PROTOS if not k.startswith("pdo_debug"))


_HOOKS_SO = os.path.join(_HERE, "lib", "libpadeops_b200_testhooks.so")


class _Library:
    """The product library plus, for tests only, the separate hooks library: attribute access to a pdo_debug_* name loads
    libpadeops_b200_testhooks.so (extern "C" wrappers around pdo::hooks::*, csrc/testhooks.cu) on first use.  The product
    library itself exports no pdo_debug_* symbol, and nothing in this package's public classes touches one."""

    def __init__(self, main):
        self._main = main
        self._hooks = None

    def __getattr__(self, name):
        if name.startswith("pdo_debug"):
            if self._hooks is None:
                if not os.path.exists(_HOOKS_SO):
                    raise PadeOpsError(-1, f"{_HOOKS_SO} not found (test hooks; built by the same make as the library)")
                H = C.CDLL(_HOOKS_SO, mode=C.RTLD_GLOBAL)
                for nm, (res, args) in _PROTOS.items():
                    if nm.startswith("pdo_debug"):
                        fn = getattr(H, nm)
                        fn.restype = res
                        fn.argtypes = args
                self._hooks = H
            return getattr(self._hooks, name)
        return getattr(self._main, name)


def lib():
    """Load the CUDA library.  Fails loudly if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise PadeOpsError(-1, f"{_SO} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(padeops_b200 has no CPU fallback)")
        # Map torch's CUDA libraries first when torch is around: libnccl.so.2 / libcufft.so.11 are resolved by
        # SONAME, and torch refuses to import after an older system NCCL has been mapped under that name.
        try:
            import torch  # noqa: F401
        except ImportError:
            pass
        L = C.CDLL(_SO, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _PROTOS.items():
            if name.startswith("pdo_debug"):
                continue
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = _Library(L)
    return _lib


def check(rc):
    if rc != 0:
        raise PadeOpsError(rc, lib().pdo_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Raw address of a torch tensor (device or host) or a numpy array."""
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def stream_ptr(stream=None):
    if stream is None:
        try:
            import torch
            if torch.cuda.is_available():
                return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        except Exception:
            pass
        return C.c_void_p(0)
    if hasattr(stream, "cuda_stream"):
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))
