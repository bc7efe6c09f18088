"""ctypes binding of librepmode_b200.so (the C ABI in include/repmode_b200.h) and its in-tree nvcc build.

There is deliberately no CPU or PyTorch fallback: if the shared library is missing or a call fails, the
caller gets a RuntimeError carrying mode_last_error().
"""
import ctypes
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.path.join(HERE, "librepmode_b200.so")
SOURCES = ["mode_abi.cu", "reparam.cu", "conv_simt.cu", "bn.cu", "conv_umma.cu", "conv_pair.cu", "wgrad_split.cu", "wgrad_deep.cu", "peer.cu", "predict.cu", "optim.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]

MODE_F32, MODE_F16 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_UMMA = 0, 1, 2
IMPL_UMMA_SINGLE, IMPL_UMMA_PAIR = 3, 4     # force the single-CTA / the CTA-pair (cta_group::2) tcgen05 kernel
IMPL_WGRAD_SPLIT, IMPL_WGRAD_DEEP = 5, 6     # mode_conv3d_wgrad: force wgrad_split.cu (r1 kernel, A/B arm) / wgrad_deep.cu (default)

_lock = threading.Lock()
_lib = None


class ModeLayer(ctypes.Structure):
    _fields_ = [("k5", ctypes.c_void_p), ("k3", ctypes.c_void_p), ("k1", ctypes.c_void_p), ("a3", ctypes.c_void_p),
                ("a5", ctypes.c_void_p), ("gate_w", ctypes.c_void_p), ("gate_b", ctypes.c_void_p),
                ("ci", ctypes.c_int32), ("co", ctypes.c_int32), ("num_tasks", ctypes.c_int32)]


class ModeReparamItem(ctypes.Structure):
    _fields_ = [("layer", ModeLayer), ("g_out", ctypes.c_void_p), ("w_fwd", ctypes.c_void_p), ("w_dgrad", ctypes.c_void_p)]


REPARAM_GROUP_MAX = 24


class ModePlanes(ctypes.Structure):
    _fields_ = [("rows_per_plane", ctypes.c_int64), ("D", ctypes.c_int32), ("own_lo", ctypes.c_int32),
                ("own_hi", ctypes.c_int32), ("valid_lo", ctypes.c_int32), ("valid_hi", ctypes.c_int32),
                ("m_global", ctypes.c_int64)]


class ModeRowMap(ctypes.Structure):
    _fields_ = [("d2s", ctypes.c_int32), ("D", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("pitch", ctypes.c_int64)]


class ModeHaloPush(ctypes.Structure):
    _fields_ = [("lo_dst", ctypes.c_void_p), ("lo_signal", ctypes.c_void_p), ("hi_dst", ctypes.c_void_p),
                ("hi_signal", ctypes.c_void_p), ("bytes", ctypes.c_int64), ("ticket", ctypes.c_void_p)]


class ModePeerPush(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("dst", ctypes.c_void_p * 8), ("signal", ctypes.c_void_p * 8),
                ("ticket", ctypes.c_void_p)]


class ModePeerGather(ctypes.Structure):
    _fields_ = [("slots", ctypes.c_void_p), ("world", ctypes.c_int32), ("signal", ctypes.c_void_p),
                ("expect", ctypes.c_void_p)]


class ModeConvOpts(ctypes.Structure):
    _fields_ = [("Dx", ctypes.c_int32), ("x_off", ctypes.c_int32), ("ep_scale", ctypes.c_void_p),
                ("ep_shift", ctypes.c_void_p), ("relu", ctypes.c_int32), ("y16", ctypes.c_void_p),
                ("Dy16", ctypes.c_int32), ("y16_off", ctypes.c_int32), ("y16_scale", ctypes.c_float),
                ("stats_push", ctypes.POINTER(ModePeerPush)), ("splitk_ws", ctypes.c_void_p),
                ("splitk_ws_bytes", ctypes.c_int64)]


class ModeCaps(ctypes.Structure):
    _fields_ = [("sm_major", ctypes.c_int32), ("sm_minor", ctypes.c_int32), ("sm_count", ctypes.c_int32),
                ("smem_per_block_optin", ctypes.c_int32), ("tmem_columns", ctypes.c_int32),
                ("abi_version", ctypes.c_int32)]


def _needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "repmode_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a into repmode_b200/librepmode_b200.so (in-tree)."""
    if not force and not _needs_build():
        return SO_PATH
    # One builder at a time ACROSS processes (torchrun starts N ranks that may all find stale sources), and the library only
    # ever appears complete: nvcc writes a private temporary that is renamed over SO_PATH.
    import fcntl
    with open(SO_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _needs_build():          # another rank built it while we waited
                return SO_PATH
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = f"{SO_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
            res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, SO_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO_PATH


_vp, _i32, _i64, _f32, _f64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double
SIGNATURES = {
    "mode_last_error": (ctypes.c_char_p, []),
    "mode_version": (ctypes.c_int, []),
    "mode_launch_count": (_i64, []),
    "mode_query": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ModeCaps)]),
    "mode_poll_error": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int32)]),
    "mode_debug_profile": (ctypes.c_int, [_vp]),
    "mode_reparam_fwd_grouped": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, ctypes.c_int, _f32, _vp]),
    "mode_reparam_fwd": (ctypes.c_int, [ctypes.POINTER(ModeLayer), _vp, _vp, _i32, _vp, _vp, _vp, ctypes.c_int, _f32,
                                        _vp, _vp]),
    "mode_packed_weight_elems": (_i64, [_i32, _i32]),
    "mode_packed_weight_elems_f16": (_i64, [_i32, _i32]),
    "mode_reparam_bwd_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "mode_reparam_bwd": (ctypes.c_int, [ctypes.POINTER(ModeLayer), _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _vp]),
    "mode_reparam_bwd_partial": (ctypes.c_int, [ctypes.POINTER(ModeLayer), _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _f32, _vp,
                                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mode_conv3d_wgrad_partial_layout": (ctypes.c_int, [ctypes.c_int, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "mode_conv3d": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp,
                                   _i32, _i32, _i32, _vp]),
    "mode_conv3d_ex": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp,
                                      _i32, _i32, _i32, ctypes.POINTER(ModeConvOpts), _vp]),
    "mode_conv3d_wgrad_ex": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp,
                                            _vp, _i32, _i32, _i32, _vp]),
    "mode_peer_enable_access": (ctypes.c_int, [_i32]),
    "mode_peer_arena_alloc": (ctypes.c_int, [_i64, ctypes.POINTER(ctypes.c_void_p), _vp]),
    "mode_peer_arena_open": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_void_p)]),
    "mode_peer_arena_close": (ctypes.c_int, [_vp]),
    "mode_peer_arena_free": (ctypes.c_int, [_vp]),
    "mode_peer_put": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "mode_peer_wait": (ctypes.c_int, [_vp, _vp, _i32, _vp]),
    "mode_peer_sum_slots": (ctypes.c_int, [_vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mode_conv3d_wgrad_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "mode_adam_chunk_elems": (_i64, []),
    "mode_adam_step": (ctypes.c_int, [_vp, _i32, _vp, _i32, _f64, _f64, _f64, _f64, _f64, _vp, _vp, _vp]),
    "mode_blend_accumulate": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                             _i32, _vp]),
    "mode_blend_finalize": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i64, _vp]),
    "mode_conv3d_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, ctypes.c_int]),
    "mode_conv3d_wgrad": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp,
                                         _vp, _i32, _vp]),
    "mode_bn_stats": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "mode_bn_finalize": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mode_bn_finalize_ex": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp,
                                           ctypes.POINTER(ModePeerGather), _vp]),
    "mode_bn_relu_bwd_reduce_ex": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                                  ctypes.POINTER(ModePeerPush), _vp]),
    "mode_bn_relu_bwd_apply_ex": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, ctypes.POINTER(ModePeerGather), ctypes.POINTER(ModeHaloPush), _vp]),
    "mode_cast_f16_ex": (ctypes.c_int, [_vp, _vp, _i64, _f32, _vp, ctypes.POINTER(ModeHaloPush), _vp]),
    "mode_bn_apply_relu": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _f32, _vp, _vp]),
    "mode_bn_finalize_apply_relu": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp,
                                                   _vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp, _vp]),
    "mode_bn_relu_bwd_reduce_v2": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "mode_bn_relu_bwd_apply_v2": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, _vp, _vp]),
    "mode_bn_relu_bwd_reduce": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mode_bn_relu_bwd_apply": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                              _vp, _vp]),
    "mode_bn_bwd_workspace_bytes": (_i64, [_i32]),
    "mode_bn_relu_bwd": (ctypes.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mode_cast_f16": (ctypes.c_int, [_vp, _vp, _i64, _f32, _vp, _vp]),
    "mode_cast_f16_pad": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "mode_cast_f16_cat": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp, _i64, _vp]),
    "mode_amax": (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    "mode_amax_multi": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "mode_f16_scale": (ctypes.c_int, [_vp, _f32, _vp, _vp]),
}


def load():
    """Load (building first if the sources are newer) and return the ctypes library handle."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _needs_build() and os.environ.get("REPMODE_NO_BUILD", "0") != "1":
            if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
                build()
            elif not os.path.exists(SO_PATH):
                raise RuntimeError(f"{SO_PATH} is missing and nvcc is not available: run __graft_entry__.build()")
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == the .so does not export the ABI
            fn.restype = res
            fn.argtypes = args
        if lib.mode_version() != 1:
            raise RuntimeError("librepmode_b200.so ABI version mismatch")
        _lib = lib
        return lib


def poll_error(what="repmode_b200"):
    """Synchronise the current device and raise if a kernel raised the device error flag since the last poll: a tcgen05
    pipeline or peer-exchange wait that timed out (codes 1-6, 41-44: the kernel bailed out and its output is incomplete) or
    a task id outside [0, num_tasks) (code 50: clamped).  Called once per Model.do_train_iter / predict, which synchronise
    anyway."""
    code = ctypes.c_int32(0)
    check(load().mode_poll_error(ctypes.byref(code)), "mode_poll_error")
    if code.value == 50:
        raise IndexError(f"{what}: a task id was outside [0, num_tasks) (the reference raises IndexError at RepMode.py:47)")
    if code.value != 0:
        raise RuntimeError(f"{what}: device error flag {code.value} (a kernel's pipeline / exchange wait timed out; "
                           "results of this step are incomplete)")


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {load().mode_last_error().decode()}")
