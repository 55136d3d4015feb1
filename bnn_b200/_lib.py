"""ctypes binding of libbnn_b200.so (include/bnn_b200.h).

There is no CPU fallback: if the library is missing, importing an op raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BNN_B200_LIB selects another build of the same library (tools/tc_trace.py points it at the trace build)
LIB_PATH = os.environ.get("BNN_B200_LIB") or os.path.join(HERE, "libbnn_b200.so")

MAX_LINEAR = 8
MAX_WIDTH = 1024
MAX_FUSED_NOUT = 64

ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_IDENTITY = 0, 1, 2, 3
MASK_NONE, MASK_EXTERNAL, MASK_BERNOULLI, MASK_CONCRETE = 0, 1, 2, 3
PREC_FP32, PREC_TF32X3, PREC_TF32, PREC_TF32_BF16 = 0, 1, 2, 3
SCALE_BBB, SCALE_SVI = 0, 1


class MlpDesc(C.Structure):
    _fields_ = [("n_linear", C.c_int32), ("dims", C.c_int32 * (MAX_LINEAR + 1)), ("act", C.c_int32)]


class MaskSpec(C.Structure):
    _fields_ = [("mode", C.c_int32), ("layer_masked", C.c_int32 * MAX_LINEAR), ("mask", C.c_void_p),
                ("mask_stride", C.c_int64), ("p_drop", C.c_float * MAX_LINEAR), ("temperature", C.c_float),
                ("seed", C.c_uint64), ("sample_offset", C.c_int64)]


class ForwardArgs(C.Structure):
    _fields_ = [("desc", MlpDesc), ("X", C.c_void_p), ("N", C.c_int64), ("W", C.c_void_p), ("w_stride", C.c_int64),
                ("S", C.c_int64), ("mask", MaskSpec), ("precision", C.c_int32), ("pred", C.c_void_p),
                ("mean", C.c_void_p), ("var", C.c_void_p), ("part_mean", C.c_void_p), ("part_m2", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64)]


class SgldChain(C.Structure):
    _fields_ = [("desc", MlpDesc), ("batch", C.c_int32), ("theta", C.c_void_p), ("thetaT", C.c_void_p), ("V", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("X", C.c_void_p), ("y", C.c_void_p),
                ("perm", C.c_void_p), ("num_train", C.c_int64), ("lr_weight", C.c_float), ("lr_noise", C.c_float),
                ("precond_decay", C.c_float), ("diagonal_bias", C.c_float), ("alpha_n", C.c_float), ("beta_n", C.c_float),
                ("gain", C.c_float), ("precond_mode", C.c_int32), ("seed", C.c_uint64)]


class BbbTrain(C.Structure):
    _fields_ = [("desc", MlpDesc), ("batch", C.c_int32), ("mu", C.c_void_p), ("rho", C.c_void_p), ("adam", C.c_void_p),
                ("X", C.c_void_p), ("y", C.c_void_p), ("perm", C.c_void_p), ("num_train", C.c_int64), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("kl_factor", C.c_float),
                ("weight_std", C.c_float), ("seed", C.c_uint64)]


class DropoutTrain(C.Structure):
    _fields_ = [("desc", MlpDesc), ("batch", C.c_int32), ("W", C.c_void_p), ("adam", C.c_void_p), ("X", C.c_void_p),
                ("y", C.c_void_p), ("perm", C.c_void_p), ("num_train", C.c_int64), ("lr", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float), ("l2_reg", C.c_float), ("p_drop", C.c_float), ("seed", C.c_uint64)]


class CDropoutTrain(C.Structure):
    _fields_ = [("desc", MlpDesc), ("batch", C.c_int32), ("W", C.c_void_p), ("adam", C.c_void_p), ("p_logit", C.c_void_p),
                ("p_adam", C.c_void_p), ("noise_level", C.c_void_p), ("X", C.c_void_p), ("y", C.c_void_p), ("perm", C.c_void_p),
                ("num_train", C.c_int64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("lscale", C.c_float), ("dr", C.c_float), ("min_noise", C.c_float), ("seed", C.c_uint64)]


# name -> (restype, argtypes); mirrors include/bnn_b200.h one to one
PROTOTYPES = {
    "bnn_abi_version": (C.c_int32, []),
    "bnn_last_error": (C.c_char_p, []),
    "bnn_sm_count": (C.c_int32, []),
    "bnn_set_device": (C.c_int32, [C.c_int32]),
    "bnn_param_stride": (C.c_int64, [C.POINTER(MlpDesc)]),
    "bnn_layer_offset": (C.c_int64, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_layer_ld": (C.c_int32, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_mask_len": (C.c_int64, [C.POINTER(MlpDesc), C.POINTER(C.c_int32)]),
    "bnn_forward_workspace_bytes": (C.c_int64, [C.POINTER(ForwardArgs)]),
    "bnn_forward": (C.c_int32, [C.POINTER(ForwardArgs), C.c_void_p]),
    "bnn_forward_tc_supported": (C.c_int32, [C.POINTER(ForwardArgs)]),
    "bnn_tc_trace_dump": (C.c_int32, [C.c_void_p]),
    "bnn_tc_mma_bench": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bnn_tc_selftest": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "bnn_moments": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_int64, C.c_void_p]),
    "bnn_moments_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "bnn_moments_merge": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "bnn_moments_merge_partials": (C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_int64, C.c_void_p,
                                               C.c_void_p, C.c_void_p]),
    "bnn_nccl_load": (C.c_int32, [C.c_char_p]),
    "bnn_moments_allreduce_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int32]),
    "bnn_moments_allreduce": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "bnn_metrics": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                C.c_void_p, C.c_void_p]),
    "bnn_reparam_kl": (C.c_int32, [C.POINTER(MlpDesc), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                   C.c_int64, C.c_int64, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "bnn_dropout_masks": (C.c_int32, [C.POINTER(MlpDesc), C.POINTER(MaskSpec), C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p]),
    "bnn_apply_masks": (C.c_int32, [C.POINTER(MlpDesc), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_int64, C.c_void_p, C.c_void_p]),
    "bnn_psgld_step": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p]),
    "bnn_sgld_theta_len": (C.c_int64, [C.POINTER(MlpDesc)]),
    "bnn_sgld_theta_t_len": (C.c_int64, [C.POINTER(MlpDesc)]),
    "bnn_sgld_workspace_bytes": (C.c_int64, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_sgld_init": (C.c_int32, [C.POINTER(SgldChain), C.c_int32, C.c_void_p]),
    "bnn_sgld_run": (C.c_int32, [C.POINTER(SgldChain), C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "bnn_sgld_set_grid_cap": (C.c_int32, [C.c_int32]),
    "bnn_sgld_perm": (C.c_int32, [C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "bnn_sgld_trace": (C.c_int32, [C.c_int32, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "bnn_sgld_status": (C.c_int32, [C.POINTER(SgldChain), C.POINTER(C.c_int32), C.c_void_p]),
    "bnn_dropout_train_supported": (C.c_int32, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_dropout_train_mask_per_step": (C.c_int64, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_dropout_train_run": (C.c_int32, [C.POINTER(DropoutTrain), C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "bnn_cdropout_train_supported": (C.c_int32, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_cdropout_train_u_per_step": (C.c_int64, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_cdropout_train_run": (C.c_int32, [C.POINTER(CDropoutTrain), C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                           C.c_void_p, C.c_void_p]),
    "bnn_bbb_train_supported": (C.c_int32, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_bbb_train_eps_per_step": (C.c_int64, [C.POINTER(MlpDesc), C.c_int32]),
    "bnn_bbb_train_run": (C.c_int32, [C.POINTER(BbbTrain), C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "bnn_philox_words": (C.c_int32, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]),
    "bnn_philox_normals": (C.c_int32, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int64, C.c_void_p, C.c_void_p]),
    "bnn_measure_fp32_tflops": (C.c_double, []),
    "bnn_forward_host": (C.c_int32, [C.POINTER(ForwardArgs)]),
    "bnn_release_workspace": (None, []),
}

_lib = None


def lib():
    """Load the CUDA library (once).  Raises -- never falls back -- if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m bnn_b200.build` (nvcc, sm_100a). "
                "bnn_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)   # AttributeError if the .so does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        if handle.bnn_abi_version() != 2:
            raise RuntimeError("libbnn_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().bnn_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg}")
