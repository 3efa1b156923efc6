"""ctypes binding of libsimulgen_b200.so (the C ABI declared in include/simulgen_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load, every kernel call
raises RuntimeError.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`python -m simulgen_vae_b200.build`.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIMULGEN_B200_LIB") or os.path.join(_HERE, "libsimulgen_b200.so")
LIB_PATH_FP16 = os.path.join(_HERE, "libsimulgen_b200_fp16.so")      # same sources, -DSG_OP16_HALF

_libs = {}
_errs = {}

c_void_p, c_int, c_float, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double
c_ll, c_ull = ctypes.c_longlong, ctypes.c_ulonglong

P, I, F, D, L, U = c_void_p, c_int, c_float, c_double, c_ll, c_ull

# name -> argtypes, in the order of include/simulgen_b200.h
SIGNATURES = {
    "sg_pack_input": [P, I, P, I, I, I, I, I, P],
    "sg_unpack_f32": [P, P, I, I, I, I, P],
    "sg_axpy_f32": [P, P, F, L, I, P],
    "sg_cast_f32": [P, P, L, I, P],
    "sg_sn_power_iter": [P, P, P, P, P, I, I, I, L, L, I, P],
    "sg_sn_pack_weight": [P, P, P, I, I, I, I, L, L, I, I, P],
    "sg_sn_weight_grad": [P, P, P, P, P, P, P, I, I, I, I, L, L, I, P],
    "sg_conv_fprop": [P, P, I, L, P, P, I, I, I, I, I, I, I, P],
    "sg_conv_fprop16": [P, P, I, L, P, P, I, I, I, I, I, I, P],
    "sg_conv_fprop_gn": [P, P, I, L, P, P, I, I, I, I, I, I, I, I, I, P, P, P, I, P],
    "sg_conv_dgrad": [P, P, I, L, P, I, I, I, I, I, I, I, I, P],
    "sg_conv_out16_ok": [I],
    "sg_set_sm_limit": [I],
    "sg_conv_wgrad": [P, I, L, P, I, L, P, I, I, I, I, I, I, P],
    "sg_gn_stats": [P, P, P, I, I, I, I, I, P],
    "sg_gn_act_fwd": [P, I, P, P, P, P, I, F, I, I, P, I, L, P, I, I, I, I, I, I, P],
    "sg_gn_act_bwd": [P, I, P, P, P, P, I, F, I, I, P, I, P, I, L, P, P, P, P, I, P, I, I, I, I, I, I, P],
    "sg_pack_static": [P, P, P, I, I, I, P],
    "sg_rows_compact16": [P, P, L, P],
    "sg_rows_expand_f32": [P, P, L, I, P],
    "sg_static_stats": [P, I, P, P, I, I, I, P],
    "sg_static_recon_fwd": [P, I, P, P, P, P, I, P, P, P, I, I, I, I, P],
    "sg_static_recon_bwd": [P, I, P, P, P, P, I, P, P, F, P, P, P, P, P, I, I, I, I, I, P],
    "sg_recon_fwd": [P, I, P, P, P, P, I, P, P, P, I, I, I, I, I, I, P],
    "sg_recon_bwd": [P, I, P, P, P, P, I, P, P, F, P, P, P, P, P, P, P, I, I, I, I, I, I, I, P],
    "sg_scale_f64_to_f32": [P, P, D, I, P],
    "sg_head_fwd": [P, P, P, P, P, I, I, I, I, I, P],
    "sg_head_bwd": [P, P, P, P, P, P, P, I, I, I, I, I, I, P],
    "sg_latent_fwd": [P, P, P, P, P, I, L, I, I, I, I, I, P],
    "sg_latent_bwd": [P, P, P, P, P, P, P, I, I, I, I, P],
    "sg_reparam_main_fwd": [P, P, P, P, I, I, P],
    "sg_reparam_main_bwd": [P, P, P, P, P, I, I, P],
    "sg_kl2_reparam_fwd": [P, P, P, P, F, P, I, L, P, P, I, I, I, I, I, P],
    "sg_kl2_reparam_bwd": [P, P, P, F, P, P, F, P, P, I, I, I, I, P],
    "sg_philox_normal": [P, I, L, U, U, L, P, P],
    "sg_counter_add": [P, L, P],
    "sg_adamw_step": [P, P, P, P, L, F, F, F, F, F, I, F, P, P],
    "sg_opt_step": [P, P, I, P, I, F, F, F, F, F, I, F, P, P, P, I, I, P],
    "sg_peer_reduce_dot": [P, P, I, P, I, I, I, I, P, P],
    "sg_sn_prepare": [P, P, I, P, L, I, I, P],
    "sg_assemble_batch": [P, I, P, P, P, P, P, I, I, I, I, U, U, I, P],
    "sg_minmax_fit": [P, I, P, L, L, P, L, P, P, I, P],
    "sg_minmax_transform": [P, I, L, L, P, P, P, P, L, P],
}


class SnLayer(ctypes.Structure):
    """sg_sn_layer of include/simulgen_b200.h"""
    _fields_ = [("w", c_void_p), ("u", c_void_p), ("v", c_void_p), ("sigma", c_void_p), ("wg", c_void_p),
                ("ws", c_void_p), ("so", c_ll), ("si", c_ll), ("H", c_int), ("Cin", c_int), ("k", c_int),
                ("Cin_p", c_int), ("flip", c_int), ("has_sn", c_int), ("has_wg", c_int), ("reserved", c_int)]


class OptItem(ctypes.Structure):
    """sg_opt_item of include/simulgen_b200.h"""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("u", c_void_p), ("vv", c_void_p),
                ("sigma", c_void_p), ("dot", c_void_p), ("n", c_ll), ("Cout", c_int), ("Cin", c_int), ("Cin_p", c_int),
                ("k", c_int), ("flip", c_int), ("reserved", c_int)]


class Peer(ctypes.Structure):
    """sg_peer of include/simulgen_b200.h"""
    _fields_ = [("world", c_int), ("rank", c_int), ("wbase", c_void_p * 8), ("vbase", c_void_p * 8), ("pbase", c_void_p * 8),
                ("wmc", c_void_p), ("vmc", c_void_p), ("pmc", c_void_p)]


def load(half=False):
    """Load a library variant once (bf16 operands by default, fp16 operands with half=True); raises RuntimeError (never
    falls back) when unavailable."""
    if half in _libs:
        return _libs[half]
    if half in _errs:
        raise RuntimeError(_errs[half])
    path = LIB_PATH_FP16 if half else LIB_PATH
    if not os.path.isfile(path):
        _errs[half] = ("simulgen_b200: %s not found - build the CUDA extension first "
                       "(python -m simulgen_vae_b200.build); there is no CPU fallback" % path)
        raise RuntimeError(_errs[half])
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:  # pragma: no cover
        _errs[half] = "simulgen_b200: cannot load %s: %s" % (path, e)
        raise RuntimeError(_errs[half])
    lib.sg_last_error.restype = ctypes.c_char_p
    lib.sg_last_error.argtypes = []
    lib.sg_version.restype = c_int
    lib.sg_device_supported.restype = c_int
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    _libs[half] = lib
    return lib


def call(name, *args, half=False):
    lib = load(half)
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError("simulgen_b200 %s failed (%d): %s" % (name, rc, lib.sg_last_error().decode()))
