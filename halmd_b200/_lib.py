"""ctypes binding of the C ABI declared in include/halmd_b200.h (kernel level) and
include/halmd_b200_modules.h (module level).

The shared library is built in-tree by ``halmd_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no CPU fallback: if the library is missing the import fails loudly, and every compute
entry point returns an error code that is raised as :class:`HxError`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HALMD_B200_LIB: development aid (kernel build variants for A/B timing); the default is the in-tree build
LIB_PATH = os.environ.get("HALMD_B200_LIB") or os.path.join(_HERE, "libhalmd_b200.so")

PLACEHOLDER = 0xFFFFFFFF
LAYOUT_COLUMN, LAYOUT_ROW = 0, 1
FORCE_POSITIONS_REDUCED = 0x100     # OR-ed into `layout` of the pair-force calls (include/halmd_b200.h)
FORCE_SCALAR = 0x200
NSHELL = 8
THERMO_DOUBLES = 8 + 8 * 296     # words of the output / scratch buffer of hx_thermodynamics


def FORCE_LANES(lanes):
    return int(lanes) << 12

TRUNC_SHARP, TRUNC_SHIFTED, TRUNC_SMOOTH_R4, TRUNC_FORCE_SHIFTED = 0, 1, 2, 3
CORR_MSD, CORR_MQD, CORR_VACF = 0, 1, 2
MAX_PAIR_TYPES = 16


class HxError(RuntimeError):
    pass


class LjParams(C.Structure):
    _fields_ = [
        ("ntype1", C.c_uint), ("ntype2", C.c_uint), ("truncation", C.c_int), ("rri_smooth", C.c_float),
        ("epsilon", C.c_float * MAX_PAIR_TYPES), ("sigma2", C.c_float * MAX_PAIR_TYPES),
        ("r_cut", C.c_float * MAX_PAIR_TYPES), ("rr_cut", C.c_float * MAX_PAIR_TYPES),
        ("en_cut", C.c_float * MAX_PAIR_TYPES), ("force_cut", C.c_float * MAX_PAIR_TYPES),
        ("hard_core", C.c_int), ("r_core", C.c_float * MAX_PAIR_TYPES),
    ]


class ShellRows(C.Structure):
    _fields_ = [("rows_dev", C.c_void_p), ("ends_dev", C.c_void_p), ("pitch", C.c_uint), ("delta", C.c_float * NSHELL)]


class VerletTail(C.Structure):
    _fields_ = [("position_out_dev", C.c_void_p), ("velocity_dev", C.c_void_p), ("image_dev", C.c_void_p),
                ("array_size", C.c_size_t), ("timestep", C.c_float), ("max_rr_word_in", C.c_uint)]


CTL_WORDS = 8
DD_MAX_RANKS, DD_IPC_HANDLE_BYTES = 32, 64


class Rand48(C.Structure):
    _fields_ = [("blocks", C.c_uint), ("threads", C.c_uint), ("a", C.c_uint16 * 3), ("c", C.c_uint16 * 3)]


class DdSteps(C.Structure):
    """hx_dd_steps (include/halmd_b200.h): everything one chunk of decomposed MD steps needs, device pointers as integers."""
    _fields_ = [
        ("position_dev", C.c_void_p * 2), ("current", C.c_int),
        ("image_dev", C.c_void_p), ("velocity_dev", C.c_void_p), ("force_dev", C.c_void_p), ("r0_dev", C.c_void_p),
        ("n_own", C.c_size_t), ("n_ghost", C.c_size_t), ("array_size", C.c_size_t),
        ("timestep", C.c_float), ("box_length", C.c_float * 3),
        ("nranks", C.c_int), ("rank", C.c_int),
        ("send_index_dev", C.c_void_p * 2), ("send_count", C.c_size_t * 2),
        ("peer_ghost_dev", (C.c_void_p * 2) * 2),
        ("peer_inbox_dev", C.c_void_p), ("inbox_dev", C.c_void_p), ("timeout_cycles", C.c_longlong),
        ("ctl_dev", C.c_void_p), ("skin_half", C.c_float), ("step_base", C.c_uint64), ("pending_finalize", C.c_int),
        ("potential", C.POINTER(LjParams)), ("shells", C.POINTER(ShellRows)), ("g_neighbour_dev", C.c_void_p),
        ("size", C.c_uint), ("stride", C.c_size_t),
        ("rng", C.POINTER(Rand48)), ("rng_state_dev", C.c_void_p), ("andersen_scratch_dev", C.c_void_p),
        ("andersen_scratch_bytes", C.c_size_t), ("sqrt_temperature", C.c_float), ("coll_prob", C.c_float),
    ]


MAX_WALLS, MAX_WALL_SPECIES = 8, 4


class HooverChain(C.Structure):
    _fields_ = [("xi", C.c_double * 2), ("v_xi", C.c_double * 2), ("mass_xi", C.c_double * 2),
                ("temperature", C.c_double), ("en_kin_target_2", C.c_double), ("dof", C.c_double)]


class PlanarWall(C.Structure):
    _fields_ = [
        ("nwall", C.c_uint), ("nspecies", C.c_uint), ("smoothing", C.c_float),
        ("surface_normal", (C.c_float * 3) * MAX_WALLS), ("offset", C.c_float * MAX_WALLS),
        ("epsilon", C.c_float * (MAX_WALL_SPECIES * MAX_WALLS)), ("sigma", C.c_float * (MAX_WALL_SPECIES * MAX_WALLS)),
        ("wetting", C.c_float * (MAX_WALL_SPECIES * MAX_WALLS)), ("cutoff", C.c_float * (MAX_WALL_SPECIES * MAX_WALLS)),
    ]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA extension first "
        "(python -c 'import __graft_entry__ as g; g.build()' or make -C halmd_b200/csrc). "
        "halmd_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

P = C.c_void_p
F3 = C.POINTER(C.c_float)
D3 = C.POINTER(C.c_double)
U3 = C.POINTER(C.c_uint)
SZ = C.c_size_t

_protos = {
    "hx_version": (C.c_int, []),
    "hx_last_error_string": (C.c_char_p, []),
    "hx_device_count": (C.c_int, []),
    "hx_launch_count": (C.c_uint64, []),
    "hx_array_size": (SZ, [SZ]),
    "hx_state_init": (C.c_int, [P, P, P, P, P, P, SZ, SZ, P]),
    "hx_binning_plan": (C.c_int, [D3, SZ, C.c_float, C.c_float, C.c_double, U3, F3, U3]),
    "hx_binning_update": (C.c_int, [P, SZ, F3, U3, C.c_uint, P, P, P, P, P]),
    "hx_device_rcp": (C.c_int, [F3, F3, C.c_uint]),
    "hx_hilbert_depth": (C.c_uint, [D3]),
    "hx_hilbert_keys": (C.c_int, [P, SZ, F3, C.c_uint, P, P]),
    "hx_iota_u32": (C.c_int, [P, SZ, P]),
    "hx_sort_scratch_bytes": (SZ, [SZ]),
    "hx_sort_pairs_stable_u32": (C.c_int, [P, P, SZ, C.c_uint, P, SZ, P]),
    "hx_state_permute": (C.c_int, [P, SZ, SZ, P, P, P, P, P, P, P, P, P, P]),
    "hx_displacement_zero": (C.c_int, [P, P, SZ, P]),
    "hx_displacement_max": (C.c_int, [P, P, SZ, F3, P, C.c_int, P]),
    "hx_displacement_exceeds": (C.c_int, [C.c_uint32, C.c_double]),
    "hx_nlist_plan": (C.c_int, [C.c_float, C.c_float, SZ, C.c_double, C.c_double, C.c_uint, U3]),
    "hx_nlist_build": (C.c_int, [P, U3, C.c_uint, P, SZ, F3, C.c_uint, C.c_uint, F3, P, C.c_uint, SZ,
                                 C.c_int, C.c_int, P, P, P]),
    "hx_potential_lj_init": (C.c_int, [C.POINTER(LjParams), C.c_uint, C.c_uint, F3, F3, C.c_int, F3, C.c_float]),
    "hx_potential_lj_init_hard_core": (C.c_int, [C.POINTER(LjParams), C.c_uint, C.c_uint, F3, F3, F3, C.c_int, F3, C.c_float]),
    "hx_pair_force": (C.c_int, [C.POINTER(LjParams), P, P, P, C.c_uint, SZ, C.c_int, SZ, F3, C.c_int,
                                C.c_float, P, P, P, SZ, P]),
    "hx_planar_wall_init": (C.c_int, [C.POINTER(PlanarWall), C.c_uint, C.c_uint, F3, F3, F3, F3, F3, F3, C.c_float]),
    "hx_external_planar_wall": (C.c_int, [C.POINTER(PlanarWall), P, SZ, F3, C.c_int, P, P, P, SZ, P, P]),
    "hx_pair_force_guarded": (C.c_int, [C.POINTER(LjParams), P, P, P, C.c_uint, SZ, C.c_int, SZ, F3, C.c_int,
                                        C.c_float, P, P, P, SZ, P, C.c_float, P]),
    "hx_shell_rows_plan": (C.c_int, [C.c_float, C.POINTER(ShellRows)]),
    "hx_shell_rows_words": (SZ, [SZ, C.c_uint]),
    "hx_count_max": (C.c_int, [P, SZ, P, P]),
    "hx_shell_rows_build": (C.c_int, [C.POINTER(LjParams), P, SZ, P, P, P, SZ, F3, C.POINTER(ShellRows), P]),
    "hx_pair_force_shells": (C.c_int, [C.POINTER(LjParams), P, P, C.POINTER(ShellRows), P, SZ, F3, C.c_int, C.c_int, P, P,
                                       C.c_float, P]),
    "hx_pair_force_shells_verlet": (C.c_int, [C.POINTER(LjParams), P, C.POINTER(ShellRows), P, SZ, F3, C.c_int, P, P, C.c_float,
                                              C.POINTER(VerletTail), P]),
    "hx_verlet_integrate": (C.c_int, [P, P, P, P, SZ, C.c_float, F3, P]),
    "hx_verlet_finalize": (C.c_int, [P, P, SZ, C.c_float, P]),
    "hx_verlet_finalize_integrate": (C.c_int, [P, P, P, P, SZ, SZ, C.c_float, F3, C.c_int, P, P, P]),
    "hx_verlet_finalize_guarded": (C.c_int, [P, P, SZ, C.c_float, P, P]),
    "hx_rand48_seed": (C.c_int, [C.POINTER(Rand48), C.c_uint, C.c_uint, C.c_uint, P, P]),
    "hx_rand48_uniform": (C.c_int, [C.POINTER(Rand48), P, P, SZ, P]),
    "hx_andersen_finalize": (C.c_int, [P, P, SZ, SZ, C.c_float, C.c_float, C.c_float, C.POINTER(Rand48), P, P]),
    "hx_hoover_integrate": (C.c_int, [P, P, P, P, SZ, C.c_float, C.c_double, F3, P]),
    "hx_hoover_rescale": (C.c_int, [P, SZ, C.c_double, P]),
    "hx_hoover_chain_init": (C.c_int, [C.POINTER(HooverChain), SZ, C.c_uint, C.c_double, C.c_double]),
    "hx_hoover_chain_set_temperature": (C.c_int, [C.POINTER(HooverChain), C.c_double, C.c_double]),
    "hx_hoover_chain_propagate": (C.c_double, [C.POINTER(HooverChain), C.c_double, C.c_double]),
    "hx_hoover_chain_energy": (C.c_double, [C.POINTER(HooverChain), SZ]),
    "hx_andersen_scratch_bytes": (SZ, [SZ]),
    "hx_andersen_finalize2": (C.c_int, [P, P, SZ, SZ, C.c_float, C.c_float, C.c_float, C.POINTER(Rand48), P, P, SZ, P, P]),
    "hx_andersen_finalize_guarded": (C.c_int, [P, P, SZ, SZ, C.c_float, C.c_float, C.c_float,
                                               C.POINTER(Rand48), P, P, P]),
    "hx_sample_positions": (C.c_int, [P, P, P, SZ, F3, P, P]),
    "hx_sample_velocities": (C.c_int, [P, P, SZ, P, P]),
    "hx_msd_scratch_bytes": (SZ, [C.c_uint]),
    "hx_msd_block": (C.c_int, [P, P, C.c_uint, SZ, P, P, SZ, P]),
    "hx_correlate_block": (C.c_int, [C.c_int, P, P, C.c_uint, SZ, P, P, SZ, P]),
    "hx_density_modes": (C.c_int, [P, SZ, P, C.c_uint, P, P]),
    "hx_thermodynamics": (C.c_int, [P, P, P, SZ, SZ, P, P]),
    "hx_wavevector_pick": (C.c_int, [D3, C.c_uint, D3, C.c_double, C.c_uint, U3, D3, C.c_uint, U3]),
    "hx_isf_accumulate": (C.c_int, [D3, D3, U3, C.c_uint, C.c_double, D3]),
    "hx_ctl_reset": (C.c_int, [P, P]),
    "hx_nlist_build2": (C.c_int, [P, U3, C.c_uint, P, SZ, SZ, F3, C.c_uint, C.c_uint, F3, P, C.c_uint, SZ,
                                  C.c_int, C.c_int, P, P, P]),
    "hx_binning_update_slab": (C.c_int, [P, SZ, F3, U3, C.c_uint, C.c_uint, C.c_uint, P, P, P, P, P]),
    "hx_nlist_build_slab": (C.c_int, [P, U3, C.c_uint, C.c_uint, C.c_uint, P, SZ, SZ, F3, C.c_uint, C.c_uint, F3, P, C.c_uint, SZ,
                                      C.c_int, C.c_int, P, P, P]),
    "hx_pair_force2": (C.c_int, [C.POINTER(LjParams), P, P, SZ, P, C.c_uint, SZ, C.c_int, SZ, F3, C.c_int,
                                 C.c_float, P, P, P, SZ, P, C.c_float, P]),
    "hx_verlet_finalize_integrate_range": (C.c_int, [P, P, P, P, SZ, SZ, SZ, C.c_float, F3, C.c_int, C.c_int, P, P, P]),
    "hx_dd_owner_keys": (C.c_int, [P, SZ, C.c_float, C.c_float, C.c_int, C.c_int, P, P, P]),
    "hx_dd_ghost_keys": (C.c_int, [P, SZ, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, P, P, P]),
    "hx_dd_permute": (C.c_int, [P, SZ, SZ, P, P, P, P, P, P, P, P, P]),
    "hx_gather_float4": (C.c_int, [P, P, SZ, P, P, P]),
    "hx_gather_u32": (C.c_int, [P, P, SZ, P, P]),
    "hx_dd_alloc": (C.c_int, [SZ, C.POINTER(C.c_void_p)]),
    "hx_dd_free": (C.c_int, [P]),
    "hx_dd_ipc_export": (C.c_int, [P, C.c_char_p]),
    "hx_dd_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "hx_dd_ipc_close": (C.c_int, [P]),
    "hx_dd_enqueue_steps": (C.c_int, [C.POINTER(DdSteps), C.c_uint, P]),
}

KERNEL_SYMBOLS = tuple(_protos)

PP = C.POINTER(C.c_void_p)
U64 = C.c_uint64
_module_protos = {
    "hxm_particle_create": (C.c_int, [SZ, C.c_uint, P, PP]),
    "hxm_particle_destroy": (None, [P]),
    "hxm_particle_nparticle": (SZ, [P]),
    "hxm_particle_array_size": (SZ, [P]),
    "hxm_particle_nspecies": (C.c_uint, [P]),
    "hxm_particle_set": (C.c_int, [P, C.c_char_p, P]),
    "hxm_particle_get": (C.c_int, [P, C.c_char_p, P]),
    "hxm_particle_device_ptr": (P, [P, C.c_char_p]),
    "hxm_particle_aux_enable": (C.c_int, [P]),
    "hxm_potential_lj_create": (C.c_int, [C.c_uint, C.c_uint, F3, F3, C.c_int, F3, C.c_float, PP]),
    "hxm_potential_lj_hard_core_create": (C.c_int, [C.c_uint, C.c_uint, F3, F3, F3, C.c_int, F3, C.c_float, PP]),
    "hxm_potential_destroy": (None, [P]),
    "hxm_potential_r_cut": (C.c_int, [P, F3]),
    "hxm_potential_params": (C.POINTER(LjParams), [P]),
    "hxm_binning_create": (C.c_int, [P, D3, F3, C.c_uint, C.c_uint, C.c_double, C.c_double, PP]),
    "hxm_binning_destroy": (None, [P]),
    "hxm_binning_ncell": (C.c_int, [P, U3]),
    "hxm_binning_cell_size": (C.c_uint, [P]),
    "hxm_binning_update": (C.c_int, [P, P]),
    "hxm_neighbour_create": (C.c_int, [P, D3, F3, C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_int, C.c_int, P, PP]),
    "hxm_neighbour_destroy": (None, [P]),
    "hxm_neighbour_size": (C.c_uint, [P]),
    "hxm_neighbour_stride": (SZ, [P]),
    "hxm_neighbour_layout": (C.c_int, [P]),
    "hxm_neighbour_check": (C.c_int, [P, C.POINTER(C.c_int)]),
    "hxm_neighbour_set_shell_rows": (C.c_int, [P, C.c_int]),
    "hxm_neighbour_updates": (SZ, [P]),
    "hxm_neighbour_mean_length": (C.c_double, [P]),
    "hxm_neighbour_get": (C.c_int, [P, P, P]),
    "hxm_neighbour_device_ptr": (P, [P]),
    "hxm_neighbour_named_ptr": (P, [P, C.c_char_p]),
    "hxm_neighbour_ctl_read": (C.c_int, [P, C.POINTER(C.c_uint32)]),
    "hxm_force_pair_trunc_create": (C.c_int, [P, P, D3, P, C.c_float, PP]),
    "hxm_force_destroy": (None, [P]),
    "hxm_force_apply": (C.c_int, [P]),
    "hxm_density_modes": (C.c_int, [P, F3, C.c_uint, P]),
    "hxm_force_add_planar_wall": (C.c_int, [P, C.POINTER(PlanarWall)]),
    "hxm_integrator_verlet_create": (C.c_int, [P, P, D3, C.c_double, PP]),
    "hxm_integrator_andersen_create": (C.c_int, [P, P, D3, C.c_double, C.c_double, C.c_double, C.c_uint, PP]),
    "hxm_integrator_destroy": (None, [P]),
    "hxm_integrator_integrate": (C.c_int, [P]),
    "hxm_integrator_finalize": (C.c_int, [P]),
    "hxm_integrator_set_timestep": (C.c_int, [P, C.c_double]),
    "hxm_integrator_set_temperature": (C.c_int, [P, C.c_double]),
    "hxm_integrator_hoover_create": (C.c_int, [P, P, D3, C.c_double, C.c_double, C.c_double, PP]),
    "hxm_integrator_hoover_state": (C.c_int, [P, C.POINTER(C.c_double)]),
    "hxm_integrator_hoover_set_mass": (C.c_int, [P, C.POINTER(C.c_double)]),
    "hxm_core_create": (C.c_int, [P, PP]),
    "hxm_core_destroy": (None, [P]),
    "hxm_core_mdstep": (C.c_int, [P]),
    "hxm_core_run": (C.c_int, [P, U64]),
    "hxm_core_step": (U64, [P]),
    "hxm_core_set_chunk": (C.c_int, [P, C.c_uint, C.c_uint]),
    "hxm_core_set_fused_verlet": (C.c_int, [P, C.c_int]),
    "hxm_core_stats": (C.c_int, [P, C.POINTER(U64), C.POINTER(U64), C.POINTER(U64)]),
    "hxm_core_thermodynamics": (C.c_int, [P, D3]),
    "hxm_blocking_create": (C.c_int, [P, C.c_double, C.c_double, C.c_uint, C.c_uint, C.c_uint, F3, C.c_uint, U3, C.c_uint, PP]),
    "hxm_blocking_destroy": (None, [P]),
    "hxm_blocking_levels": (C.c_uint, [P]),
    "hxm_blocking_block_size": (C.c_uint, [P]),
    "hxm_blocking_intervals": (C.c_int, [P, C.POINTER(U64)]),
    "hxm_blocking_sample": (C.c_int, [P]),
    "hxm_blocking_finalise": (C.c_int, [P]),
    "hxm_blocking_msd": (C.c_int, [P, D3]),
    "hxm_blocking_isf": (C.c_int, [P, D3]),
    "hxm_blocking_mqd": (C.c_int, [P, D3]),
    "hxm_blocking_vacf": (C.c_int, [P, D3]),
    "hxm_blocking_trace": (C.c_int, [P, C.c_int, C.POINTER(U64), C.c_size_t, C.POINTER(C.c_size_t)]),
    "hxm_core_run_sampling": (C.c_int, [P, U64, P]),
}
MODULE_SYMBOLS = tuple(_module_protos)
_protos.update(_module_protos)

for _name, (_res, _args) in _protos.items():
    _fn = getattr(lib, _name)   # AttributeError here = header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc, what=""):
    """Raise HxError for a non-zero return code of an hx_* call."""
    if rc != 0:
        msg = lib.hx_last_error_string().decode(errors="replace")
        raise HxError(f"{what or 'hx call'} failed (code {rc}): {msg}")


def f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def u3(v):
    return (C.c_uint * 3)(*[int(x) for x in v])


def farr(v):
    v = [float(x) for x in v]
    return (C.c_float * len(v))(*v)
