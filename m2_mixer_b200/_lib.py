"""ctypes binding of libm2b200.so (the C ABI declared in include/m2b200.h).

The library is built in-tree by ``m2_mixer_b200.build`` (nvcc, sm_100a).  There is NO CPU fallback: if the library
is missing and cannot be built, importing the ops fails loudly; if a call returns a non-zero status it raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libm2b200.so")

FP32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2

vp, i32, i64, f32, sz, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint64
PP = C.POINTER(C.c_void_p)
PI64 = C.POINTER(C.c_int64)
PI32 = C.POINTER(C.c_int)
PF32 = C.POINTER(C.c_float)

# name -> (restype, argtypes).  Order and meaning follow include/m2b200.h exactly.
PROTOTYPES = {
    "m2b200_abi_version": (i32, []),
    "m2b200_status_string": (C.c_char_p, [i32]),
    "m2b200_cast_bf16": (i32, [vp, i64, vp, i64, i32, i32, vp]),
    "m2b200_cast_bf16_multi": (i32, [vp, i32, vp]),
    "m2b200_gemm": (i32, [i32, vp, i32, i64, vp, i32, i64, i32, i32, i32, i32, i64, i64, vp, i32, i32, vp, i64, i64, vp,
                          i32, i64, i64, i32, i32, vp]),
    "m2b200_token_mix_fwd_workspace_bytes": (sz, [i32] * 5),
    "m2b200_token_mix_fwd": (i32, [vp] * 8 + [i32] * 5 + [f32, u64, vp, sz, vp]),
    "m2b200_token_mix_bwd_workspace_bytes": (sz, [i32] * 5),
    "m2b200_token_mix_bwd": (i32, [vp] * 14 + [i32] * 5 + [f32, u64, vp, sz, vp]),
    "m2b200_channel_mix_workspace_bytes": (sz, [i32] * 5),
    "m2b200_channel_mix_fwd": (i32, [vp] * 9 + [i32, vp] + [i32] * 4 + [f32, u64, vp, sz, vp]),
    "m2b200_channel_mix_bwd": (i32, [vp] * 9 + [i32] + [vp] * 7 + [i32] * 4 + [f32, u64, vp, sz, vp]),
    "m2b200_layernorm_fwd": (i32, [vp] * 4 + [i32] * 3 + [i64, vp]),
    "m2b200_layernorm_bwd": (i32, [vp, i64] + [vp] * 6 + [i32] * 3 + [vp]),
    "m2b200_linear_workspace_bytes": (sz, [i32] * 5),
    "m2b200_linear_fwd": (i32, [vp, vp, vp, i32, vp, i32, vp] + [i32] * 4 + [f32, u64, vp, sz, vp]),
    "m2b200_linear_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp] + [i32] * 4 + [f32, u64, vp, sz, vp]),
    "m2b200_dropout_mask": (i32, [vp, i32, i32, i64, f32, u64, i32, vp]),
    "m2b200_set_dropout_epoch_ptr": (None, [vp]),
    "m2b200_dropout_epoch_advance": (i32, [vp, vp]),
    "m2b200_patch_embed_workspace_bytes": (sz, [vp] + [i32] * 9),
    "m2b200_patch_embed_fwd": (i32, [vp, i32, vp, vp, i32, vp, vp] + [i32] * 7 + [vp, sz, vp]),
    "m2b200_patch_embed_bwd": (i32, [vp, vp, i32, vp, vp] + [i32] * 7 + [vp, sz, vp]),
    "m2b200_patch_gather": (i32, [vp, vp] + [i32] * 5 + [vp]),
    "m2b200_copy_tokens": (i32, [vp, i64, vp, i64, i32, i64, i32, vp]),
    "m2b200_add": (i32, [vp, vp, vp, i64, vp]),
    "m2b200_fuse2_fwd": (i32, [vp, vp, vp, i64, i32, vp]),
    "m2b200_fuse2_max_bwd": (i32, [vp, vp, vp, vp, vp, i64, vp]),
    "m2b200_gate_fwd": (i32, [vp, vp, vp, vp, i64, vp]),
    "m2b200_gate_bwd": (i32, [vp] * 7 + [i64, vp]),
    "m2b200_mean_pool_fwd": (i32, [vp, vp, i32, i32, i32, vp]),
    "m2b200_mean_pool_bwd": (i32, [vp, vp, i32, i32, i32, vp]),
    "m2b200_heads_loss_fwd": (i32, [PP, PI64, PI32, PI32, PP, PP, i32, i32, i32, i32, vp, vp, PF32, vp, vp, vp, vp]),
    "m2b200_heads_loss_bwd": (i32, [PP, PI64, PI32, PI32, PP, PP, i32, i32, i32, i32, vp, vp, PF32, vp, f32, vp, PP, PI64,
                                    PI32, PP, PP, vp]),
    "m2b200_launch_count": (C.c_ulonglong, []),
    "m2b200_profile_enable": (None, [i32]),
    "m2b200_profile_collect": (sz, [C.c_char_p, sz]),
    "m2b200_adam_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp, vp]),
}

_lib = None


class M2B200Error(RuntimeError):
    pass


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the CUDA library.  Raises if it cannot be produced: no fallback path exists."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    stale = False
    if os.path.exists(LIB_PATH) and not os.environ.get("M2B200_SKIP_DIGEST"):
        # an edited .cu / .cuh / header must never run against yesterday's binary: compare the source digest with the one
        # the library was built from (build.py writes it next to the objects)
        stamp = os.path.join(_build.OBJ, "digest.txt")
        stale = not os.path.exists(stamp) or open(stamp).read() != _build._digest()
    if not os.path.exists(LIB_PATH) or stale or os.environ.get("M2B200_REBUILD"):
        if not build_if_missing:
            raise M2B200Error(f"{LIB_PATH} is missing or older than its sources (run `python -m m2_mixer_b200.build`)")
        _build.build(force=bool(os.environ.get("M2B200_REBUILD")))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError here == header/library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().m2b200_status_string(status).decode()
        raise M2B200Error(f"{what} failed: status {status} ({msg})")


def launch_count() -> int:
    return int(load().m2b200_launch_count())


class profile:
    """``with profile() as p: ...`` then ``p.table`` = {kernel name: (launches, total_ms)} measured with CUDA events
    on the launching stream."""

    def __enter__(self):
        load().m2b200_profile_enable(1)
        self.table = {}
        return self

    def __exit__(self, *exc):
        lib = load()
        lib.m2b200_profile_enable(0)
        buf = C.create_string_buffer(1 << 16)
        lib.m2b200_profile_collect(buf, len(buf))
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split(",")
            self.table[name] = (int(n), float(ms))
        return False
