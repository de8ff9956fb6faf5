"""ctypes binding of libpidnet_b200.so (the C ABI declared in include/pidnet_b200.h).

There is no CPU / PyTorch fallback: if the library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libpidnet_b200.so')


class Cfg(C.Structure):
    _fields_ = [('m', C.c_int), ('n', C.c_int), ('num_classes', C.c_int), ('planes', C.c_int),
                ('ppm_planes', C.c_int), ('head_planes', C.c_int), ('augment', C.c_int)]


class CriterionCfg(C.Structure):
    _fields_ = [('ignore_label', C.c_int64), ('ohem_keep', C.c_int64), ('ohem_thres', C.c_float),
                ('bd_threshold', C.c_float), ('balance_weight_aux', C.c_float), ('balance_weight_main', C.c_float),
                ('sb_weight', C.c_float), ('coeff_bce', C.c_float)]


_vp, _fp, _i, _i64p = C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/pidnet_b200.h one to one (tests check the list)
SIGNATURES = {
    'pidnet_last_error': (C.c_char_p, []),
    'pidnet_abi_version': (_i, []),
    'pidnet_debug_fastdiv': (C.c_uint, [C.c_uint, C.c_uint]),
    'pidnet_create': (_i, [C.POINTER(Cfg), C.POINTER(_vp)]),
    'pidnet_destroy': (_i, [_vp]),
    'pidnet_set_param': (_i, [_vp, C.c_char_p, _vp, _i64p, _i]),
    'pidnet_plan': (_i, [_vp, _i, _i, _i, C.POINTER(C.c_size_t)]),
    'pidnet_forward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i]),
    'pidnet_forward_u8': (_i, [_vp, _vp, _vp, C.POINTER(C.c_double), C.POINTER(C.c_double), _vp, _vp, _vp, _i]),
    'pidnet_num_launches': (_i, [_vp]),
    'pidnet_conv_flops': (C.c_double, [_vp]),
    'pidnet_profile': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    'pidnet_op_info': (_i, [_vp, _i, C.c_char_p, _i, C.c_char_p, _i, C.POINTER(C.c_double), C.POINTER(C.c_double),
                            C.POINTER(_i)]),
    'pidnet_set_option': (_i, [_vp, C.c_char_p, _i]),
    'pidnet_debug_tensor': (_i, [_vp, C.c_char_p, _vp, _i64p]),
    'pidnet_op_conv2d': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i]),
    'pidnet_op_stem': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    'pidnet_op_pag': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i]),
    'pidnet_op_upadd': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i]),
    'pidnet_op_pool': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i]),
    'pidnet_op_lightbag': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    'pidnet_criterion_workspace_bytes': (C.c_size_t, [_i, _i, _i]),
    'pidnet_criterion': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, C.POINTER(CriterionCfg), _vp,
                              C.c_size_t, _vp, _vp, _vp, _vp, _vp]),
    'pidnet_upsample_align_corners': (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i]),
    'pidnet_train_create': (_i, [C.POINTER(Cfg), C.POINTER(_vp)]),
    'pidnet_train_destroy': (_i, [_vp]),
    'pidnet_train_bind': (_i, [_vp, C.c_char_p, _vp, _vp, _i64p, _i]),
    'pidnet_train_plan': (_i, [_vp, _i, _i, _i, C.POINTER(C.c_size_t)]),
    'pidnet_train_step': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(CriterionCfg), _i, _vp, _vp, _vp, _vp, _vp]),
    'pidnet_train_backward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i]),
    'pidnet_train_num_segments': (_i, [_vp]),
    'pidnet_train_segment_ranges': (_i, [_vp, _i, _vp, _i64p, _i, C.POINTER(_i)]),
    'pidnet_train_profile': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(CriterionCfg), C.c_char_p, C.c_size_t, C.POINTER(C.c_float)]),
    'pidnet_train_debug_tensor': (_i, [_vp, C.c_char_p, _i, _vp, _i64p]),
    'pidnet_train_num_launches': (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    'pidnet_train_forward': (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    'pidnet_train_set_option': (_i, [_vp, C.c_char_p, _i]),
    'pidnet_postprocess': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, C.c_int64, _vp, _vp]),
    'pidnet_sgd_step': (_i, [_vp, _vp, _vp, _vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, _i, _i, C.c_float]),
    'pidnet_op_bag': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}

# hardware probes (include/pidnet_b200_probe.h): only in libpidnet_b200_probe.so, built on demand for tools/probe_*.py
PROBE_SIGNATURES = {
    'pidnet_probe_mn': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    'pidnet_probe_mma_rate': (_i, [_vp, _i, _i, _i, _i, _vp]),
    'pidnet_probe_pair': (_i, [_vp, _vp, _vp, _i, _vp]),
    'pidnet_probe_mma_rate_pair': (_i, [_vp, _i, _i, _i, _i, _vp]),
    'pidnet_probe_halo': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
}
PROBE_LIB_PATH = os.path.join(_HERE, 'lib', 'libpidnet_b200_probe.so')

_lib = None
_probe_lib = None


def load_probe():
    """The probe build (product sources + csrc/probe.cu with -DPIDNET_PROBES); `python -m pidnet_b200.build --probes`."""
    global _probe_lib
    if _probe_lib is None:
        if not os.path.exists(PROBE_LIB_PATH):
            raise RuntimeError(f'pidnet_b200: {PROBE_LIB_PATH} is missing -- build it with `python -m pidnet_b200.build --probes`')
        lib = C.CDLL(PROBE_LIB_PATH)
        for name, (res, args) in {**SIGNATURES, **PROBE_SIGNATURES}.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _probe_lib = lib
    return _probe_lib


def load():
    """Load the shared library (building it is `python -m pidnet_b200.build`); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'pidnet_b200: {LIB_PATH} is missing -- build it with `python -m pidnet_b200.build` '
            '(there is no CPU or PyTorch fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().pidnet_last_error()
        raise RuntimeError('pidnet_b200: ' + (msg.decode() if msg else f'error {rc}'))
