"""ctypes binding of libmarlgrid_b200.so (include/marlgrid_b200.h).

There is NO CPU fallback: if the CUDA library is missing the import fails loudly.
"""
import ctypes
import os

from .config import MgConfig, MgState

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmarlgrid_b200.so")

EXPORTS = (
    "mg_version", "mg_build_info", "mg_sizeof_config", "mg_config_validate", "mg_obs_bytes_per_env", "mg_init", "mg_sync_derived", "mg_reset", "mg_step",
    "mg_obs_encode", "mg_obs_rgb", "mg_step_fused", "mg_step_fused_rgb", "mg_rollout_fused", "mg_rollout_persistent", "mg_rollout_policy", "mg_policy_act", "mg_rollout_fused_rr", "mg_random_actions",
    "mg_los_batch", "mg_engine_create", "mg_engine_destroy", "mg_engine_reset", "mg_engine_step", "mg_engine_copy_only", "mg_host_alloc",
    "mg_host_free", "mg_launch_count", "mg_pregen_words_per_env", "mg_pregen_run", "mg_pregen_drain", "mg_pregen_set_auto", "mg_pregen_stats", "mg_debug_set_mid_event", "mg_debug_force_two_kernels", "mg_debug_force_general_fused",
)

_lib = None


class MarlgridLibraryError(ImportError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and os.environ.get("MG_NO_AUTOBUILD") != "1":
        # a fresh checkout, or csrc edited since the library was linked: (re)compile in tree (nvcc cross-compiles sm_100a
        # without a GPU).  Under torchrun the ranks would race, so there the library must have been built beforehand.
        try:
            from . import build as _build

            if not os.path.exists(LIB_PATH) or _build.needs_build():
                _build.build()
        except Exception as e:  # noqa: BLE001 -- a missing library is reported below; a stale one must not pass silently
            if os.path.exists(LIB_PATH):
                raise MarlgridLibraryError(f"{LIB_PATH} is older than marlgrid_b200/csrc and rebuilding it failed: {e}") from e
    if not os.path.exists(LIB_PATH):
        raise MarlgridLibraryError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `python -m marlgrid_b200.build`). marlgrid_b200 has no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise MarlgridLibraryError(f"{LIB_PATH} does not export {name}")
    P, I64, U64, I = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int
    CFG, ST = ctypes.POINTER(MgConfig), ctypes.POINTER(MgState)
    L.mg_version.restype = I
    L.mg_build_info.restype = ctypes.c_char_p
    L.mg_config_validate.argtypes = [CFG]
    L.mg_obs_bytes_per_env.argtypes = [CFG, I]
    L.mg_obs_bytes_per_env.restype = I64
    L.mg_init.argtypes = [CFG, ST, P]
    L.mg_sync_derived.argtypes = [CFG, ST, P]
    L.mg_reset.argtypes = [CFG, ST, P, P]
    L.mg_step.argtypes = [CFG, ST, P, P, P, I, P]
    L.mg_obs_encode.argtypes = [CFG, ST, P, P]
    L.mg_obs_rgb.argtypes = [CFG, ST, P, P, P]
    L.mg_step_fused.argtypes = [CFG, ST, P, P, P, P, I, P]
    L.mg_step_fused_rgb.argtypes = [CFG, ST, P, P, P, P, P, I, P]
    L.mg_rollout_fused.argtypes = [CFG, ST, P, I64, P, P, P, I, P]
    L.mg_rollout_persistent.argtypes = [CFG, ST, P, I64, P, P, P, I, P]
    L.mg_rollout_policy.argtypes = [CFG, ST, P, I64, P, P, P, P, I, P]
    L.mg_policy_act.argtypes = [CFG, ST, P, P, P, P]
    L.mg_rollout_fused_rr.argtypes = [CFG, ST, I, P, I64, P, P, P, I, P]
    L.mg_random_actions.argtypes = [P, I64, I, U64, U64, P]
    L.mg_los_batch.argtypes = [P, P, I64, I, I, I, P]
    L.mg_engine_create.argtypes = [ctypes.POINTER(P), CFG, I64, I64, U64, I, I, P, I64]
    L.mg_engine_destroy.argtypes = [P]
    L.mg_engine_destroy.restype = None
    L.mg_engine_reset.argtypes = [P, P]
    L.mg_engine_step.argtypes = [P, P, P, P, P, I]
    L.mg_engine_copy_only.argtypes = [P, P, P, P, P]
    L.mg_host_alloc.argtypes = [I64]
    L.mg_host_alloc.restype = P
    L.mg_host_free.argtypes = [P]
    L.mg_host_free.restype = None
    L.mg_launch_count.restype = I64
    L.mg_pregen_run.argtypes = [CFG, ST, P]
    L.mg_pregen_set_auto.argtypes = [I]
    L.mg_pregen_set_auto.restype = None
    L.mg_pregen_stats.argtypes = [P, I]
    L.mg_debug_set_mid_event.argtypes = [P]
    L.mg_debug_set_mid_event.restype = None
    L.mg_debug_force_two_kernels.argtypes = [I]
    L.mg_debug_force_two_kernels.restype = None
    L.mg_debug_force_general_fused.argtypes = [I]
    L.mg_debug_force_general_fused.restype = None
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is ctypes.c_int and name not in ("mg_version",):
            fn.restype = I
    if L.mg_sizeof_config() != ctypes.sizeof(MgConfig):
        raise MarlgridLibraryError(f"MgConfig layout mismatch: C {L.mg_sizeof_config()} vs ctypes {ctypes.sizeof(MgConfig)}")
    _lib = L
    return L


def check(code, what):
    """0 = ok; >0 = cudaError_t; <0 = MG_E_*."""
    if code == 0:
        return
    if code > 0:
        raise RuntimeError(f"{what}: CUDA error {code}")
    if code == -1:
        raise ValueError(f"{what}: invalid configuration (MG_E_CONFIG): field out of range, (2*max_steps+1)*n_agents >= 65536 (16-bit arrival "
                         "stamps), or a grid / tile atlas too large for the 227 KB of shared memory one CTA stages 32 envs in")
    raise ValueError(f"{what}: invalid argument (code {code})")
