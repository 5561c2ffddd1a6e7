"""ctypes binding of libg2048.so (include/g2048.h).  No CPU fallback: if the library is missing and
cannot be built, or a compute call is made without a CUDA device, the call raises."""
from __future__ import annotations

import ctypes as C
import os
import re
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SO_PATH = os.path.join(_HERE, "libg2048.so")
SOURCES = [os.path.join(_HERE, "csrc", "g2048.cu")]
DEPENDS = SOURCES + [os.path.join(_HERE, "csrc", "g2048_device.cuh"), os.path.join(ROOT, "include", "g2048.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off", "-diag-suppress", "550,177"]


class G2048Error(RuntimeError):
    pass


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise G2048Error("nvcc not found: libg2048.so cannot be built (there is no CPU fallback)")


def _stale() -> bool:
    return (not os.path.exists(SO_PATH)) or any(os.path.getmtime(d) > os.path.getmtime(SO_PATH) for d in DEPENDS)


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None, out: str | None = None) -> str:
    """Compile the CUDA library in-tree for sm_100a (cross-compiles without a GPU).  Safe under torchrun: one process
    compiles (file lock) into a temporary file that is renamed into place, the others wait and load the finished file."""
    target = out or SO_PATH
    if not (force or out or _stale()):
        return target
    import fcntl
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or out or _stale():       # nobody built it while this process waited for the lock
                tmp = f"{target}.{os.getpid()}.tmp"
                cmd = [_nvcc()] + NVCC_FLAGS + (extra_flags or []) + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
                res = subprocess.run(cmd, capture_output=True, text=True)
                if res.returncode != 0:
                    if os.path.exists(tmp):
                        os.unlink(tmp)
                    raise G2048Error("nvcc failed:\n" + res.stdout + res.stderr)
                os.replace(tmp, target)
                if verbose:
                    print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return target


def declared_symbols() -> list[str]:
    """Every entry point include/g2048.h declares."""
    text = open(os.path.join(ROOT, "include", "g2048.h")).read()
    return sorted(set(re.findall(r"G2048_API\s+[\w\s\*]+?\b(g2048_\w+)\s*\(", text)))


_lib = None
vp, i64, u64, i32, f32, f64, sz = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_float, C.c_double, C.c_size_t
_SIG = {
    "g2048_version": (i32, []),
    "g2048_last_error": (C.c_char_p, []),
    "g2048_init": (i32, [i32]),
    "g2048_device_count": (i32, []),
    "g2048_host_tables": (None, [vp] * 6),
    "g2048_host_alloc": (vp, [sz]),
    "g2048_host_free": (None, [vp]),
    "g2048_env_reset": (i32, [vp, vp, vp, vp, i64, u64, u64, u64, vp]),
    "g2048_env_step": (i32, [vp] * 10 + [i64, i32, u64, u64, u64, vp]),
    "g2048_move_trial": (i32, [vp, vp, vp, vp, vp, i64, vp]),
    "g2048_legal_mask": (i32, [vp, vp, i64, vp]),
    "g2048_pack_i64": (i32, [vp, vp, i64, vp, vp]),
    "g2048_unpack_i64": (i32, [vp, vp, i64, vp]),
    "g2048_encode_onehot": (i32, [vp, vp, i64, i32, vp]),
    "g2048_dqn_env_step": (i32, [vp] * 11 + [i32, i64, f64, C.c_uint32, u64, u64, u64, u64, vp]),
    "g2048_select_action": (i32, [vp, vp, vp, i64, f64, u64, u64, u64, vp]),
    "g2048_rollout_random": (i32, [vp, vp, vp, i64, i64, i32, u64, u64, u64, vp, vp]),
    "g2048_rollout_qlearn": (i32, [vp, vp, vp, vp, u64, i64, i64, i32, f32, f32, f64, u64, u64, u64, vp, vp]),
    "g2048_qlearn_scratch_bytes": (sz, [i64]),
    "g2048_qlearn_step": (i32, [vp, vp, vp, vp, u64, i64, i32, f32, f32, f64, i32, i32, u64, u64, u64, vp, vp, vp, vp,
                                vp, sz, vp]),
    "g2048_rollout_qlearn_sharded": (i32, [vp, vp, vp, vp, i32, u64, i64, i64, i32, f32, f32, f64, u64, u64, u64, vp, vp]),
    "g2048_qtable_lookup_sharded": (i32, [vp, i32, u64, vp, i64, vp, vp, i32, vp]),
    "g2048_qlearn_emit": (i32, [vp, vp, vp, vp, u64, i64, i32, f32, f64, u64, u64, u64, vp, vp, vp]),
    "g2048_qtable_apply_records": (i32, [vp, u64, vp, vp, i32, f32, i32, vp, sz, vp]),
    "g2048_qlearn_emit_owned": (i32, [vp, vp, vp, vp, i32, u64, i64, i32, f32, f64, u64, u64, u64, u64, i32, vp, vp, vp, vp,
                                      vp, i32, vp]),
    "g2048_qtable_apply_owned": (i32, [vp, u64, vp, vp, i32, i32, f32, vp, sz, vp]),
    "g2048_peer_read_u64": (i32, [vp, i32, vp, vp]),
    "g2048_peer_memset": (i32, [vp, i32, sz, vp]),
    "g2048_peer_alloc": (i32, [sz, vp, vp]),
    "g2048_peer_open": (i32, [vp, vp]),
    "g2048_peer_close": (i32, [vp]),
    "g2048_peer_free": (i32, [vp]),
    "g2048_peer_barrier": (i32, [vp, i32, i32, u64, u64, vp, vp]),
    "g2048_routed_buffer_bytes": (sz, [i32, i64]),
    "g2048_routed_create": (vp, [i32, i32, i64, i64, vp, vp, u64]),
    "g2048_routed_destroy": (None, [vp]),
    "g2048_routed_prime": (i32, [vp, vp, i64, vp]),
    "g2048_routed_step": (i32, [vp, vp, vp, vp, i64, i32, f32, f32, f64, u64, u64, u64, vp, vp, vp]),
    "g2048_qtable_bytes": (sz, [u64]),
    "g2048_qtable_clear": (i32, [vp, u64, vp]),
    "g2048_qtable_lookup": (i32, [vp, u64, vp, i64, vp, vp, i32, vp]),
    "g2048_choose_action": (i32, [vp, u64, vp, vp, i64, f64, u64, u64, u64, vp]),
    "g2048_qtable_update": (i32, [vp, u64, vp, vp, vp, vp, vp, i64, f32, f32, i32, vp, sz, vp]),
    "g2048_qtable_apply_targets": (i32, [vp, u64, vp, vp, vp, i64, f32, i32, vp, sz, vp]),
    "g2048_qtable_size": (i32, [vp, u64, vp, vp]),
    "g2048_qtable_probe_stats": (i32, [vp, u64, vp, vp]),
    "g2048_qtable_export": (i32, [vp, u64, vp, vp, i64, vp, vp]),
    "g2048_ctx_create": (vp, [i32, i64, u64]),
    "g2048_ctx_destroy": (None, [vp]),
    "g2048_ctx_env_reset": (i32, [vp, vp, vp, vp, vp, i64, u64, u64, u64]),
    "g2048_ctx_env_step": (i32, [vp] * 10 + [i64, i32, u64, u64, u64]),
    "g2048_ctx_legal_mask": (i32, [vp, vp, vp, i64]),
    "g2048_ctx_move_trial": (i32, [vp, vp, vp, vp, vp, vp, i64]),
    "g2048_ctx_rollout_random": (i32, [vp, vp, vp, vp, i64, i64, i32, u64, u64, u64, vp]),
    "g2048_ctx_rollout_qlearn": (i32, [vp, vp, vp, vp, i64, i64, i32, f32, f32, f64, u64, u64, u64, vp]),
    "g2048_ctx_qtable_lookup": (i32, [vp, vp, i64, vp, vp, i32]),
    "g2048_ctx_choose_action": (i32, [vp, vp, vp, i64, f64, u64, u64, u64]),
    "g2048_ctx_qtable_update": (i32, [vp, vp, vp, vp, vp, vp, i64, f32, f32, i32]),
    "g2048_ctx_qtable_size": (i64, [vp]),
    "g2048_ctx_qtable_export": (i64, [vp, vp, vp, i64]),
    "g2048_ctx_qtable_clear": (i32, [vp]),
    "g2048_ctx_table": (vp, [vp]),
    "g2048_ctx_table_capacity": (u64, [vp]),
    "g2048_ctx_stream": (vp, [vp]),
}


def lib():
    """Load (building first if needed) libg2048.so and attach the prototypes."""
    global _lib
    if _lib is None:
        # G2048_LIB: load another build of the same library (kernel experiments; tools/ only)
        path = os.environ.get("G2048_LIB") or build()
        try:
            handle = C.CDLL(path)
        except OSError as e:  # e.g. libcudart missing
            raise G2048Error(f"cannot load {path}: {e} (there is no CPU fallback)") from e
        for name, (res, args) in _SIG.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().g2048_last_error().decode(errors="replace")
        raise G2048Error(f"{what or 'g2048'} failed with code {rc}: {msg}")


_inited: set[int] = set()


def init(device: int = 0) -> None:
    """g2048_init(device): raises if there is no usable CUDA device."""
    if device not in _inited:
        L = lib()
        if L.g2048_device_count() <= device:
            raise G2048Error(f"CUDA device {device} not available: the g2048 hot path has no CPU fallback")
        check(L.g2048_init(device), "g2048_init")
        _inited.add(device)
