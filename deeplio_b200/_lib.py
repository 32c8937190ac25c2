"""ctypes binding of libdeeplio_b200.so (the C ABI declared in include/deeplio_b200.h).

There is no CPU fallback: if the shared library is missing, importing this module raises.  Build it with
``python -m deeplio_b200.build`` (nvcc, sm_100a).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdeeplio_b200.so")

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3, 4
GRAD_DIRECT, GRAD_POOL, GRAD_AVG = 0, 1, 2


class Tensor4(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("n", "h", "w", "c", "ph", "pw")]


class Conv(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("kh", "kw", "sh", "sw", "ph", "pw")]


class BnPool(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("relu", "res_mode", "pool_k", "pool_sh", "pool_sw", "c_off", "out_group", "zero_tail")]


class LossTerm(C.Structure):
    _fields_ = [("pred", C.c_void_p), ("gt", C.c_void_p), ("dpred", C.c_void_p)] + \
               [(k, C.c_int) for k in ("B", "G", "C", "pred_sb", "pred_ss", "gt_sb", "gt_ss", "d_sb", "d_ss")]


class DlioError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError("deeplio_b200: %s not found -- run `python -m deeplio_b200.build` (there is no CPU fallback)"
                      % LIB_PATH)

_lib = C.CDLL(LIB_PATH)

P, I, LL, F, SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
_PROTOS = {
    "dlio_abi_version": (I, []),
    "dlio_last_error": (C.c_char_p, []),
    "dlio_device_check": (I, [I]),
    "dlio_launch_count": (LL, []),
    "dlio_set_option": (I, [C.c_char_p, I]),
    "dlio_workspace_bytes": (SZ, [C.c_char_p, C.POINTER(C.c_longlong), I]),
    "dlio_profile_enable": (I, [I]),
    "dlio_profile_read": (I, [I, P, P]),
    "dlio_pack_input": (I, [P, LL, LL, LL, I, I, Tensor4, P, P, P]),
    "dlio_weight_to_s2d": (I, [P, I, I, I, I, I, P, P, P]),
    "dlio_weight_grad_from_s2d": (I, [P, I, I, I, I, I, P, P]),
    "dlio_fold_stats": (I, [P, I, I, P, P]),
    "dlio_conv2d_fwd": (I, [Tensor4, P, P, P, P, P, Conv, I, Tensor4, P, P, P]),
    "dlio_conv2d_bwd_data": (I, [Tensor4, P, P, P, P, P, P, Conv, Tensor4, P, I, P]),
    "dlio_conv2d_bwd_weight": (I, [Tensor4, P, P, Tensor4, P, P, Conv, P, P]),
    "dlio_weight_to_ohwi": (I, [P, I, I, I, I, I, P, P, P]),
    "dlio_weight_grad_to_oihw": (I, [P, I, I, I, I, I, P, P]),
    "dlio_weight_flip_transpose": (I, [P, I, I, I, I, P, P, P]),
    "dlio_weight_pack_f16": (I, [P, I, I, I, I, I, I, I, P, P, P]),
    "dlio_pack_f16": (I, [P, LL, I, P, P, P]),
    "dlio_absmax": (I, [P, LL, P, P]),
    "dlio_weight_to_s2d_f16": (I, [P, I, I, I, I, I, P, P, P]),
    "dlio_weight_grad_from_s2d_f16": (I, [P, I, I, I, I, I, P, P]),
    "dlio_conv2d_fwd_f16_folded": (I, [Tensor4, P, P, P, P, P, Conv, I, Tensor4, P, P, P]),
    "dlio_conv2d_bwd_weight_f16_folded": (I, [Tensor4, P, P, Tensor4, P, P, Conv, P, P]),
    "dlio_weight_pack_pair_f16": (I, [P, I, I, I, I, I, I, P, P, P]),
    "dlio_weight_grad_from_pair": (I, [P, I, I, I, I, P, P]),
    "dlio_conv2d_fwd_f16": (I, [Tensor4, P, P, P, P, P, Conv, I, Tensor4, P, P, P]),
    "dlio_conv2d_bwd_data_f16": (I, [Tensor4, P, P, P, P, Conv, Tensor4, P, I, P]),
    "dlio_conv2d_bwd_weight_f16": (I, [Tensor4, P, P, Tensor4, P, P, Conv, P, P]),
    "dlio_bn_finalize": (I, [P, LL, I, P, P, P, P, F, F, I, P, P, P, P, P, P, P, P]),
    "dlio_bn_act_pool_fwd": (I, [Tensor4, P, P, P, Tensor4, P, BnPool, Tensor4, P, P, P, P, P, P, P]),
    "dlio_pool_bwd_sums": (I, [Tensor4, P, I, I, P, P, P, I, P, P]),
    "dlio_bn_pool_bwd_apply": (I, [Tensor4, P, BnPool, Tensor4, P, P, P, LL, P, P, P, I, I, Tensor4, P, P, P, P, P, P, P, P]),
    "dlio_bn_act_pool_bwd_reduce": (I, [Tensor4, P, P, P, P, P, Tensor4, P, BnPool, I, Tensor4, P, I, P, P, P, I, I, P, I, P]),
    "dlio_bn_bwd_apply": (I, [Tensor4, P, P, P, LL, P, P, P, I, I, Tensor4, P, P, P, P, P, P, P, P, I, P]),
    "dlio_f64_to_f32": (I, [P, P, I, P]),
    "dlio_spatial_mean_fwd": (I, [Tensor4, P, P, P, I, P, I, I, P]),
    "dlio_spatial_dot": (I, [Tensor4, P, Tensor4, P, P, P]),
    "dlio_channel_scale_fwd": (I, [Tensor4, P, P, Tensor4, P, P, P]),
    "dlio_channel_scale_bwd": (I, [P, P, P, I, I, I, P, P]),
    "dlio_axpby": (I, [P, F, P, F, P, LL, P]),
    "dlio_sum_mid": (I, [P, P, LL, I, I, P]),
    "dlio_mul": (I, [P, P, P, LL, P]),
    "dlio_copy2d": (I, [P, LL, P, LL, LL, I, P]),
    "dlio_dropout_mask": (I, [P, LL, F, C.c_ulonglong, P, P]),
    "dlio_linear_fwd": (I, [P, I, P, P, I, I, I, I, P, I, P]),
    "dlio_linear_bwd": (I, [P, I, P, P, I, P, I, I, I, I, I, P, I, P, P, P, P]),
    "dlio_rnn_reserve_floats": (SZ, [I, I, I, I, I, I, I]),
    "dlio_rnn_bwd_scratch_floats": (SZ, [I, I, I, I, I, I, I]),
    "dlio_rnn_fwd": (I, [I, I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P]),
    "dlio_rnn_bwd": (I, [I, I, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P, SZ, P]),
    "dlio_hws_loss": (I, [P, P, P, P, I, F, F, P, P, P, P]),
    "dlio_adam_step": (I, [P, P, P, P, LL, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, I, C.c_double, P]),
    "dlio_pose_loss": (I, [LossTerm, LossTerm, LossTerm, LossTerm, P, P, I, F, P, P, P, P, P]),
    "dlio_se3_chain_fwd": (I, [P, P, I, I, P, P, P, P]),
    "dlio_se3_chain_bwd": (I, [P, P, I, I, P, P, P, P, P]),
    "dlio_gt_relative": (I, [P, I, I, P, I, P, P, P, P]),
    "dlio_finite_check": (I, [P, P, I, P, P]),
    "dlio_pair_gather": (I, [P, LL, LL, LL, I, I, P, I, I, I, Tensor4, P, P, P, P, P, I, P]),
    "dlio_scan_scratch_bytes": (SZ, [I, I]),
    "dlio_scan_project": (I, [P, I, I, I, F, F, F, F, P, I, P, P, P, P, P, P]),
    "dlio_imu_windows": (I, [P, P, I, P, I, I, P, P, P, P, P]),
}
EXPORTS = sorted(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(_lib, _name)  # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args

ABI_VERSION = 7
if _lib.dlio_abi_version() != ABI_VERSION:
    raise ImportError("deeplio_b200: ABI version mismatch (library %d, binding %d)" % (_lib.dlio_abi_version(), ABI_VERSION))


def last_error():
    return _lib.dlio_last_error().decode()


def launch_count():
    return int(_lib.dlio_launch_count())


def _checked(name):
    fn = getattr(_lib, name)

    def call(*args):
        rc = fn(*args)
        if rc != 0:
            raise DlioError("%s failed (%d): %s" % (name, rc, last_error()))
    call.__name__ = name
    return call


# status-returning entry points, wrapped to raise
for _name, (_res, _args) in _PROTOS.items():
    if _res is I and _name != "dlio_abi_version":
        globals()[_name[5:]] = _checked(_name)
scan_scratch_bytes = _lib.dlio_scan_scratch_bytes
rnn_reserve_floats = _lib.dlio_rnn_reserve_floats
rnn_bwd_scratch_floats = _lib.dlio_rnn_bwd_scratch_floats
abi_version = _lib.dlio_abi_version


def workspace_bytes(op, *dims):
    """dlio_workspace_bytes: caller-owned scratch of one call of `op` (see the header for the dims of each op)."""
    arr = (C.c_longlong * max(1, len(dims)))(*dims)
    n = _lib.dlio_workspace_bytes(op.encode(), arr, len(dims))
    if n == C.c_size_t(-1).value:
        raise DlioError("dlio_workspace_bytes failed: %s" % last_error())
    return n


PROF_KINDS = ("conv_fwd_simt", "conv_dgrad_simt", "conv_wgrad_simt", "conv_fwd_tc", "conv_dgrad_tc", "conv_wgrad_tc",
              "elementwise", "dense", "rnn", "optim")


def profile_read():
    """{kernel class: (total ms, launches)} since the last profile_enable(1)."""
    out = {}
    for k, name in enumerate(PROF_KINDS):
        ms, n = C.c_double(0.0), C.c_longlong(0)
        rc = _lib.dlio_profile_read(k, C.byref(ms), C.byref(n))
        if rc != 0:
            raise DlioError("dlio_profile_read failed (%d): %s" % (rc, last_error()))
        if n.value:
            out[name] = (ms.value, n.value)
    return out


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def ptr_array(tensors):
    """Host array of device pointers (float *const *)."""
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
    return arr
