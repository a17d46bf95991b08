"""ctypes binding of the C-ABI in include/brb.h (libbrb_cuda.so, built in-tree by __graft_entry__.build()).

There is deliberately no fallback: if the shared library is missing or no CUDA device is usable, the
calls raise.  The oracle under oracle/ is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import pathlib
import shutil
import subprocess

from .model import BrbModelConsts

CSRC = pathlib.Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libbrb_cuda.so"
_EXPERIMENT_LIB = "BRB_EXPERIMENT_LIB"   # kernel-tuning experiments only (scripts/): alternative build of the SAME sources
SOURCES = ("brb_kernels.cu", "brb_cabi.cu", "brb_policy.cu", "brb_policy_tc.cu")
HEADERS = (CSRC / "brb_internal.h", CSRC / "brb_policy_layout.h", CSRC / "brb_chol8.inc", CSRC / "brb_chol8z.inc", CSRC / "brb_chol6.inc", CSRC / "brb_schur6.inc", CSRC / "brb_schur6w.inc", CSRC / "brb_env03.cuh",
           CSRC.parent.parent / "include" / "brb.h")
# -ftz=true: no denormal handling around MUFU.RSQ / RCP (the state is O(1); 0.895 -> 0.859 ms per step at 65,536 robots)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-ftz=true",
              "-Xcompiler", "-fPIC", "-shared"]

DONE_ROW_WORDS = 10   # include/brb.h BRB_DONE_ROW_WORDS
POLICY_NPARAM = 9413   # include/brb.h BRB_POLICY_NPARAM
NSTATS = 12
STAT_NAMES = ("substeps", "contact_substeps", "solves", "nonconverged", "unsupported", "episodes", "env_steps", "contact_slots", "coupled_substeps", "block_contact_substeps", "coupled_fallbacks", "coupled_solves")

EXPORTS = (
    "brb_version", "brb_strerror", "brb_model_create", "brb_model_destroy", "brb_env_create", "brb_env_destroy",
    "brb_env_reset_all", "brb_env_step", "brb_env_step_host", "brb_env_step_host_compact", "brb_env_host_layout", "brb_env_get_state", "brb_env_set_state",
    "brb_env_get_elapsed", "brb_env_get_stats", "brb_env_num_envs", "brb_env_num_launches", "brb_fp32_peak_flops", "brb_policy_act", "brb_ppo_grad", "brb_adam_clip_step", "brb_random_permutation", "brb_policy_value_masked", "brb_ppo_tc_fault",
    "brb_comm_create", "brb_comm_export", "brb_comm_open", "brb_comm_destroy", "brb_comm_grad", "brb_comm_fault", "brb_comm_allreduce_adam",
)


class BrbError(RuntimeError):
    pass


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libbrb_cuda.so (cross-compiles without a GPU)."""
    srcs = [CSRC / s for s in SOURCES]
    deps = srcs + list(HEADERS)
    if not force and LIB_PATH.exists() and all(LIB_PATH.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB_PATH)] + [str(s) for s in srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise BrbError(f"nvcc failed:\n{res.stderr[-4000:]}")
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load libbrb_cuda.so; raises BrbError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = pathlib.Path(os.environ.get(_EXPERIMENT_LIB, LIB_PATH))
    if not path.exists():
        raise BrbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the env step)")
    L = C.CDLL(str(path))
    vp, i64, u64 = C.c_void_p, C.c_int64, C.c_uint64
    L.brb_version.restype = C.c_int
    L.brb_strerror.restype = C.c_char_p
    L.brb_strerror.argtypes = [C.c_int]
    L.brb_model_create.argtypes = [C.POINTER(BrbModelConsts), vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.brb_model_destroy.argtypes = [vp]
    L.brb_model_destroy.restype = None
    L.brb_env_create.argtypes = [vp, i64, u64, i64, C.POINTER(vp)]
    L.brb_env_destroy.argtypes = [vp]
    L.brb_env_destroy.restype = None
    L.brb_env_reset_all.argtypes = [vp, vp, vp, vp]
    L.brb_env_step.argtypes = [vp] * 11
    L.brb_env_step_host.argtypes = [vp] * 9
    L.brb_env_step_host_compact.argtypes = [vp] * 7 + [i64]
    L.brb_env_host_layout.argtypes = [vp, C.POINTER(i64 * 3), C.POINTER(i64)]
    L.brb_env_get_state.argtypes = [vp] * 5
    L.brb_env_set_state.argtypes = [vp] * 4
    L.brb_env_get_elapsed.argtypes = [vp] * 3
    L.brb_env_get_stats.argtypes = [vp, C.POINTER(u64 * NSTATS)]
    L.brb_env_num_envs.argtypes = [vp]
    L.brb_env_num_envs.restype = i64
    L.brb_env_num_launches.argtypes = [vp]
    L.brb_env_num_launches.restype = i64
    L.brb_fp32_peak_flops.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.brb_policy_act.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp]
    L.brb_ppo_grad.argtypes = [vp] * 7 + [i64, vp, C.c_float, C.c_float, C.c_float, vp, vp, vp]
    L.brb_ppo_tc_fault.argtypes = [C.c_int]
    L.brb_comm_create.argtypes = [C.c_int, C.c_int, C.c_int, i64, C.POINTER(vp)]
    L.brb_comm_export.argtypes = [vp, vp]
    L.brb_comm_open.argtypes = [vp, vp]
    L.brb_comm_destroy.argtypes = [vp]
    L.brb_comm_destroy.restype = None
    L.brb_comm_grad.argtypes = [vp, i64]
    L.brb_comm_grad.restype = vp
    L.brb_comm_fault.argtypes = [vp]
    L.brb_comm_allreduce_adam.argtypes = [vp, vp, vp, vp] + [C.c_float] * 4 + [i64, C.c_float, vp, vp]
    L.brb_policy_value_masked.argtypes = [vp, vp, vp, i64, vp, vp]
    L.brb_adam_clip_step.argtypes = [vp] * 4 + [i64] + [C.c_float] * 4 + [i64, C.c_float, C.c_float, vp, vp]
    L.brb_random_permutation.argtypes = [vp, i64, u64, vp]
    _lib = L
    return L


def check(code: int, what: str) -> None:
    if code != 0:
        raise BrbError(f"{what} failed: {lib().brb_strerror(code).decode()} ({code})")
