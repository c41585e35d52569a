"""ctypes binding of libwd_b200.so (declared in include/wd_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this module raises.
Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C workoutdetector_b200/csrc``.
"""
import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libwd_b200.so")
REPO_ROOT = os.path.dirname(_HERE)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--split-compile", "0",   # 0 = one optimisation thread per host core
]


class WdError(RuntimeError):
    """A libwd_b200 call returned a negative status."""

    def __init__(self, code, msg):
        super().__init__(f"libwd_b200 error {code}: {msg}")
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [
        ("arch", C.c_int32), ("num_class", C.c_int32), ("num_segments", C.c_int32), ("shift_div", C.c_int32),
        ("is_shift", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("max_clips", C.c_int32),
        ("mode", C.c_int32), ("device", C.c_int32),
    ]


class NamedTensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("numel", C.c_int64)]


ARCH_TSM_R50 = 0
ARCH_TDN_R50 = 1
MODE_BF16 = 0
MODE_FP32_VALIDATE = 1

# every symbol include/wd_b200.h declares: name -> (restype, argtypes)
_vp, _i, _f = C.c_void_p, C.c_int, C.c_float
SYMBOLS = {
    "wd_abi_version": (_i, []),
    "wd_last_error": (C.c_char_p, []),
    "wd_engine_create": (_i, [C.POINTER(ModelDesc), C.POINTER(_vp)]),
    "wd_engine_destroy": (_i, [_vp]),
    "wd_engine_load_weights": (_i, [_vp, C.POINTER(NamedTensor), _i]),
    "wd_engine_frame_bytes": (C.c_size_t, [_vp]),
    "wd_preprocess_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _f, _vp, _vp]),
    "wd_pack_nchw_f32": (_i, [_vp, _vp, _i, _vp, _vp]),
    "wd_engine_clip_bytes": (C.c_size_t, [_vp]),
    "wd_pack_tdn_f32": (_i, [_vp, _vp, _i, _vp, _vp]),
    "wd_preprocess_tdn_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _f, _vp, _vp]),
    "wd_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _f, _i, _vp]),
    "wd_forward_timed": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _f, _i, _vp, _vp]),
    "wd_count_reps": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "wd_scores_to_states": (_i, [_vp, _i, _i, _f, _i, _vp, _vp, _vp]),
    "wd_vote_states": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "wd_infer_u8_host": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp]),
    "wd_infer_u8_host_async": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _i, _vp, _vp, _vp]),
    "wd_infer_host_sync": (_i, [_vp]),
    "wd_engine_num_ops": (_i, [_vp]),
    "wd_engine_op_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "wd_engine_set_tap": (_i, [_vp, _i, _vp, C.c_int64]),
    "wd_engine_frame_geometry": (_i, [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "wd_engine_set_option": (_i, [_vp, C.c_char_p, _i]),
    "wd_engine_launch_count": (C.c_int64, [_vp]),
    "wd_debug_conv": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "wd_bench_conv": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                           C.POINTER(C.c_float)]),
}

_lib = None


def build(verbose=False):
    """Compile csrc/wd_engine.cu -> csrc/libwd_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(REPO_ROOT, "include", "wd_b200.h"))
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(CSRC, "wd_engine.cu"), "-lcudart"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


def load():
    """dlopen the library and attach prototypes. Raises if it is absent — there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` in the repo root.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.wd_abi_version() != 1:
        raise ImportError(f"libwd_b200 ABI version {lib.wd_abi_version()} != 1")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise WdError(rc, load().wd_last_error().decode("utf-8", "replace"))
