"""ctypes binding of libmetabo_b200.so (C ABI declared in include/metabo_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MBO_LIB_PATH", os.path.join(_HERE, "libmetabo_b200.so"))   # override: A/B kernel builds

KERNEL_IDS = {"RBF": 0, "Matern32": 1, "Matern52": 2}
GP_FP64, GP_FP32 = 0, 1
MLP_FP32, MLP_TF32, MLP_BF16X3, MLP_FP16 = 0, 1, 2, 3
AF_EI, AF_UCB, AF_PI = 0, 1, 2
MLP_MODES = {"fp32": MLP_FP32, "tf32": MLP_TF32, "bf16x3": MLP_BF16X3, "fp16": MLP_FP16}
ENV_BRANIN, ENV_HARTMANN3 = 0, 1
REWARDS = {"none": 0, "neg_linear": 1, "neg_log10": 2}
ACT = {"relu": 0, "tanh": 1}
FEAT_MEAN, FEAT_STD, FEAT_X0, FEAT_INCUMBENT, FEAT_TPERC, FEAT_TIMESTEP, FEAT_BUDGET = 0, 1, 2, 10, 11, 12, 13


class Candidates(C.Structure):
    """mirror of `mbo_candidates`"""
    _fields_ = [("M", C.c_int), ("D", C.c_int), ("n_head", C.c_int), ("n_tail", C.c_int),
                ("local", C.c_int), ("n_cubes", C.c_int), ("head_is_f32", C.c_int), ("tail_is_f32", C.c_int),
                ("head", C.c_void_p), ("tail", C.c_void_p), ("cubes", C.c_void_p)]


class MboError(RuntimeError):
    pass


_lib = None

# every symbol include/metabo_b200.h declares: (name, restype, argtypes)
_vp, _i, _sz, _u64, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_double
SYMBOLS = [
    ("mbo_last_error", C.c_char_p, []),
    ("mbo_version", _i, []),
    ("mbo_gp_state_doubles", _sz, [_i, _i]),
    ("mbo_gp_fit", _i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    ("mbo_gp_posterior", _i, [_i, _i, _i, _i, _i, _vp, C.POINTER(Candidates), _vp, _vp, _i, _vp]),
    ("mbo_policy_create", _i, [C.POINTER(_vp)]),
    ("mbo_policy_destroy", None, [_vp]),
    ("mbo_policy_set_weights", _i, [_vp, _i, C.POINTER(C.c_int32), _i, C.POINTER(_vp), C.POINTER(_vp), _i, _vp]),
    ("mbo_af_logits", _i, [_vp, _i, _i, _i, C.POINTER(C.c_int32), C.POINTER(Candidates), _vp, _vp, _vp, _vp, _vp,
                           _sz, _vp]),
    ("mbo_af_workspace_bytes", _sz, [_i, _i, _i]),
    ("mbo_af_topk", _i, [_vp, _i, _i, _i, C.POINTER(C.c_int32), C.POINTER(Candidates), _vp, _vp, _vp, _i, _vp, _vp, _vp,
                         _vp, _sz, _vp]),
    ("mbo_af_logits_dense", _i, [_vp, _i, _i, _i, _i, _i, C.POINTER(C.c_int32), _vp, _vp, _vp, _sz, _vp]),
    ("mbo_af_eval_fused", _i, [_vp, _i, _i, _i, _i, _i, _vp, _i, C.POINTER(C.c_int32), C.POINTER(Candidates), _vp, _vp,
                               _vp, _vp, _vp, _sz, _vp]),
    ("mbo_value_net", _i, [_vp, _i, _vp, _vp, _vp]),
    ("mbo_af_closed_form", _i, [_i, _i, _i, _vp, _vp, _vp, _d, _d, _i, _vp, _vp]),
    ("mbo_logits_argmax", _i, [_i, _i, _vp, _vp, _vp, _vp]),
    ("mbo_logits_topk", _i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    ("mbo_logits_sample", _i, [_i, _i, _vp, _u64, _u64, _vp, _u64, _vp, _vp]),
    ("mbo_logits_logp_entropy", _i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    ("mbo_candidates_gather", _i, [_i, C.POINTER(Candidates), _i, _vp, _vp, _vp]),
    ("mbo_cubes_around", _i, [_i, _i, _i, _vp, _d, _vp, _vp]),
    ("mbo_state_pack", _i, [_i, C.POINTER(Candidates), _i, C.POINTER(C.c_int32), _vp, _vp, _vp, _vp, _vp]),
    ("mbo_env_step", _i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, C.c_float, C.c_float,
                          C.c_float, _vp]),
    ("mbo_ppo_workspace_bytes", _sz, [_i, _i]),
    ("mbo_ppo_debug_offsets", _i, [_i, _i, C.POINTER(_sz)]),
    ("mbo_ppo_policy_grad", _i, [_i, _i, _i, _i, C.POINTER(C.c_int32), _i, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp,
                                 _vp, C.c_float, C.c_float, C.c_float, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _sz, _vp]),
    ("mbo_ppo_value_workspace_bytes", _sz, [_i, _i]),
    ("mbo_ppo_value_grad", _i, [_i, _sz, _i, _i, _i, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp, C.c_float, C.POINTER(_vp),
                                C.POINTER(_vp), _vp, _vp, _vp, _sz, _vp]),
    ("mbo_ppo_prepare_minibatch", _i, [_i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("mbo_ppo_adv_stats", _i, [_i, _vp, _vp, _vp, _vp, _vp]),
    ("mbo_ppo_loss_sums", _i, [_i, _vp, _vp, C.c_float, C.c_float, C.c_float, _vp, _vp]),
    ("mbo_adam_step", _i, [_i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int64),
                           C.POINTER(_vp),
                           C.c_float, C.c_float, C.c_float, C.c_float, _vp]),
    ("mbo_ppo_simt_workspace_bytes", _sz, [C.c_longlong, _i, C.POINTER(C.c_int32), _i]),
    ("mbo_ppo_policy_grad_simt", _i, [_i, _i, _i, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i, _vp, C.POINTER(_vp),
                                      C.POINTER(_vp), _vp, _vp, _vp, C.c_float, C.c_float, C.c_float, C.POINTER(_vp),
                                      C.POINTER(_vp), _vp, _vp, _sz, _vp]),
    ("mbo_ppo_value_grad_simt", _i, [_i, _sz, _i, _i, _i, C.POINTER(C.c_int32), _i, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp,
                                     C.c_float, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _vp, _sz, _vp]),
    ("mbo_comm_create", _i, [_i, _i, _sz, C.POINTER(_vp)]),
    ("mbo_comm_export", _i, [_vp, _vp]),
    ("mbo_comm_connect", _i, [_vp, _vp]),
    ("mbo_comm_destroy", None, [_vp]),
    ("mbo_comm_allreduce", _i, [_vp, _i, _vp, _vp, _sz, _vp]),
    ("mbo_comm_allreduce_adam", _i, [_vp, _vp, _vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                     C.POINTER(C.c_int64), C.POINTER(_vp), C.c_float, C.c_float, C.c_float, C.c_float, _vp]),
    ("mbo_comm_status", _i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), _vp]),
]
COMM_HANDLE_BYTES = 64


def lib():
    """Loads the CUDA library; there is deliberately no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MboError("libmetabo_b200.so is missing (%s): build it with `python -c 'import __graft_entry__ as g; "
                           "g.build()'` or `make -C metabo_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            if "MBO_LIB_PATH" in os.environ and not hasattr(l, name):
                continue                              # A/B runs against older builds of the library
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        raise MboError("%s failed (%d): %s" % (what, rc, lib().mbo_last_error().decode()))
