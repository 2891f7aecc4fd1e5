"""ctypes binding of ``libtriplane_b200.so`` (C ABI declared in ``include/triplane_b200.h``).

The library is built in-tree by ``triplaneturbo_b200/csrc/build.sh`` (``nvcc -gencode
arch=compute_100a,code=sm_100a``).  There is no fallback: if the shared object is missing, or no CUDA device
is usable, every compute entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TT_B200_LIB selects another build of the same library (e.g. the -DTT_MEMLOCK=0 build the sanitizer runs compare)
LIB_PATH = os.environ.get("TT_B200_LIB") or os.path.join(_HERE, "lib", "libtriplane_b200.so")

TT_OK = 0
fp = C.c_void_p      # device pointers travel as integers
i64 = C.c_int64


class TTConfig(C.Structure):
    """``tt_config`` of include/triplane_b200.h."""
    _fields_ = [("C", C.c_int32), ("R", C.c_int32), ("P", C.c_int32), ("rays_per_cache", C.c_int32),
                ("radius", C.c_float), ("sdf_bias_radius", C.c_float), ("inv_std", C.c_float),
                ("cos_anneal_ratio", C.c_float), ("near_plane", C.c_float), ("far_plane", C.c_float),
                ("render_step_size", C.c_float), ("flags", C.c_int32), ("image_h", C.c_int32), ("image_w", C.c_int32)]


_cfgp = C.POINTER(TTConfig)

# name -> (restype, argtypes); must list every symbol the header declares (tests/test_cabi.py checks it)
SIGNATURES = {
    "tt_version": (C.c_int, []),
    "tt_last_error": (C.c_char_p, []),
    "tt_device_ok": (C.c_int, []),
    "tt_set_impl": (C.c_int, [C.c_int]),
    "tt_get_impl": (C.c_int, []),
    "tt_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "tt_launch_count": (i64, []),
    "tt_profile_begin": (C.c_int, []),
    "tt_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
    "tt_wpack_floats": (C.c_size_t, [C.c_int]),
    "tt_wgrad_floats": (C.c_size_t, [C.c_int]),
    "tt_wgrad_offsets": (C.c_int, [C.c_int, C.POINTER(i64)]),
    "tt_pack_weights": (C.c_int, [fp] * 9 + [C.c_int, fp, fp]),
    "tt_repack_planes": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, fp, fp]),
    "tt_repack_planes_bwd": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, fp, fp]),
    "tt_repack_planes_bwd_split": (C.c_int, [fp] + [C.c_int] * 6 + [fp, fp]),
    "tt_geometry_fwd": (C.c_int, [fp, fp, _cfgp, fp, i64, C.c_int] + [fp] * 6 + [fp]),
    "tt_geometry_bwd_scratch_floats": (C.c_size_t, [_cfgp, i64]),
    "tt_geometry_bwd": (C.c_int, [fp, fp, _cfgp, fp, i64] + [fp] * 4 + [fp, fp, fp, fp]),
    "tt_wgrad_def_floats": (C.c_size_t, [C.c_int]),
    "tt_field_bwd": (C.c_int, [fp, fp, _cfgp, fp, i64, fp, fp, fp, fp, fp, fp, fp]),
    "tt_sample_scratch_floats": (C.c_size_t, [i64, C.c_int]),
    "tt_importance_sample": (C.c_int, [fp, fp, _cfgp, fp, fp, i64, C.c_int, C.c_int, fp, fp, fp, fp, fp]),
    "tt_render_fwd_scratch_floats": (C.c_size_t, [i64, C.c_int]),
    "tt_render_fwd": (C.c_int, [fp, fp, _cfgp, fp, fp, i64, fp, fp, i64, C.c_int] + [fp] * 8 + [fp, fp, fp]),
    "tt_render_bwd_scratch_floats": (C.c_size_t, [_cfgp, i64, C.c_int]),
    "tt_render_bwd": (C.c_int, [fp, fp, _cfgp, fp, fp, i64, fp, fp, i64, C.c_int] + [fp] * 6 + [fp] * 6 +
                      [C.c_float, fp, fp, fp, fp, fp]),
    "tt_to_channel_last": (C.c_int, [fp, i64, C.c_int, i64, fp, fp]),
    "tt_from_channel_last": (C.c_int, [fp, i64, C.c_int, i64, fp, fp]),
    "tt_sample_planes_fwd": (C.c_int, [fp] + [C.c_int] * 5 + [fp, i64, C.c_int, fp, fp]),
    "tt_sample_planes_bwd": (C.c_int, [fp] + [C.c_int] * 5 + [fp, i64, C.c_int, fp, fp, fp, fp]),
    "tt_sample_planes_bwdbwd": (C.c_int, [fp] + [C.c_int] * 5 + [fp, i64, C.c_int] + [fp] * 7),
    "tt_composite_fwd": (C.c_int, [fp, fp, i64, C.c_int, C.c_int, fp, fp, fp, fp]),
    "tt_composite_bwd": (C.c_int, [fp, fp, fp, fp, fp, i64, C.c_int, C.c_int, fp, fp, fp]),
}


def bind(lib):
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


class TTError(RuntimeError):
    pass


_lib = None


def load():
    """The one and only backend.  Fails loudly when the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TTError(f"{LIB_PATH} not found: build it with triplaneturbo_b200/csrc/build.sh "
                          "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def check(lib, code, what):
    if code != TT_OK:
        raise TTError(f"{what} failed ({code}): {lib.tt_last_error().decode()}")
