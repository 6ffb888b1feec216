"""ctypes binding of libevrp_b200.so (the C ABI declared in include/evrp_b200.h).

There is NO fallback: if the shared library is missing or a kernel cannot be launched, the calling op
raises.  PyTorch only supplies device memory and the CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EVRP_B200_LIB") or os.path.join(_HERE, "libevrp_b200.so")   # override: instrumented debug builds (tools/)

EVRP_MAX_S = 128
ABI_VERSION = 22
EVRP_MAX_LAYERS = 8
ST_NAN_LOGIT, ST_INFEASIBLE_PICK, ST_STEP_CAP, ST_NONUNIFORM_DYNAMIC, ST_NO_FEASIBLE, ST_STUCK_AT_DEPOT = 1, 2, 4, 8, 16, 32
DECODE_GREEDY, DECODE_SAMPLE = 0, 1

_FP = C.c_void_p   # device pointers travel as void*


class Phys(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "aero", "kd", "kr", "g", "Cr", "velocity", "mc", "w", "max_load", "start_soc", "t_limit",
        "charging_time", "serve_time", "charge_plus_serve", "inv_velocity", "inv_3600")] + \
        [("div_mode", C.c_int32), ("objective", C.c_int32)]


class LayerWeights(C.Structure):
    _fields_ = [(n, _FP) for n in ("Wqkv", "Wo", "bn1_scale", "bn1_shift", "ff1_w", "ff1_b", "ff2_w", "ff2_b",
                                   "bn2_scale", "bn2_shift", "Wqkv_hi", "Wqkv_lo", "Wo_hi", "Wo_lo",
                                   "ff1_hi", "ff1_lo", "ff2_hi", "ff2_lo")]


class EncoderWeights(C.Structure):
    _fields_ = [(n, _FP) for n in ("depot_w", "depot_b", "station_w", "station_b", "cust_w", "cust_b")] + \
        [("n_layers", C.c_int32), ("layers", LayerWeights * EVRP_MAX_LAYERS),
         ("node_proj", _FP), ("fixed_ctx_w", _FP), ("node_proj_hi", _FP), ("node_proj_lo", _FP),
         ("step_proj", _FP), ("step_proj_hi", _FP), ("step_proj_lo", _FP),
         ("gemm_mode", C.c_int32), ("attn_mode", C.c_int32),
         ("fused_w", _FP), ("fused_params", _FP), ("fused_node_chunks", C.c_int32)]


class DecoderWeights(C.Structure):
    _fields_ = [(n, _FP) for n in ("w1_veh", "w1_max", "b1", "w2", "b2", "w1max_hi", "w1max_lo", "w2_hi", "w2_lo",
                                   "am_w_load")] + \
        [("w1_inv_scale", C.c_float), ("w2_inv_scale", C.c_float), ("ff_mode", C.c_int32), ("policy", C.c_int32),
         ("inv_scales", _FP), ("ff_stream", _FP)]


class RolloutCfg(C.Structure):
    _fields_ = [("decode_type", C.c_int32), ("temp", C.c_float), ("tanh_clipping", C.c_float),
                ("max_steps", C.c_int32), ("seed", C.c_uint64), ("n_replicas", C.c_int32),
                ("static_xy", _FP), ("elevations", _FP)]


# every symbol include/evrp_b200.h declares, with its ctypes signature
SIGNATURES = {
    "evrp_phys_default": (None, [C.POINTER(Phys), C.c_double, C.c_double, C.c_double, C.c_double]),
    "evrp_abi_version": (C.c_int, []),
    "evrp_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "evrp_update_dynamic": (C.c_int, [C.POINTER(Phys), _FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int,
                                      _FP, _FP, _FP, _FP]),
    "evrp_update_mask": (C.c_int, [C.POINTER(Phys), _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, _FP, _FP, _FP]),
    "evrp_pairwise_build": (C.c_int, [_FP, _FP, C.c_int, C.c_int, _FP, _FP, _FP]),
    "evrp_generate_instances": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                          _FP, _FP, _FP, _FP]),
    "evrp_encoder_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "evrp_encoder_forward": (C.c_int, [C.POINTER(EncoderWeights), _FP, _FP, C.c_int, C.c_int, C.c_int,
                                       _FP, _FP, _FP, _FP, _FP, C.c_size_t, _FP]),
    "evrp_encoder_fused_debug": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "evrp_rollout": (C.c_int, [C.POINTER(Phys), C.POINTER(DecoderWeights), C.POINTER(RolloutCfg),
                               _FP, _FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int,
                               _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP, C.c_size_t, _FP]),
    "evrp_rollout_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "evrp_sample_min": (C.c_int, [_FP, _FP, C.c_int, C.c_int, C.c_int, _FP, _FP, _FP, _FP]),
    "evrp_pointer_logp_saved_floats": (C.c_int, []),
    "evrp_pointer_logp_forward": (C.c_int, [_FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                            _FP, _FP, _FP]),
    "evrp_pointer_logp_backward": (C.c_int, [_FP, _FP, _FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_float,
                                             C.c_float, _FP, _FP, _FP, _FP]),
    "evrp_self_attention_forward": (C.c_int, [_FP, C.c_int, C.c_int, _FP, _FP, _FP]),
    "evrp_self_attention_backward": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int, C.c_int, _FP, _FP]),
    "evrp_gemm_3xtf32": (C.c_int, [C.c_int, _FP, C.c_int, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_int, _FP, _FP,
                                   C.c_int, _FP, _FP, _FP, _FP, _FP]),
    "evrp_split_tf32": (C.c_int, [_FP, C.c_int, C.c_int, C.c_longlong, C.c_longlong, _FP, _FP, _FP]),
    "evrp_wgrad_workspace_bytes": (C.c_size_t, []),
    "evrp_wgrad_3xtf32": (C.c_int, [_FP, C.c_int, C.c_int, _FP, C.c_int, C.c_int, C.c_int, _FP, C.c_int, _FP, _FP, C.c_int,
                                    _FP, C.c_size_t, _FP]),
    "evrp_bn_stats": (C.c_int, [_FP, C.c_int, _FP, _FP]),
    "evrp_bn_apply": (C.c_int, [_FP, C.c_int, _FP, _FP, _FP, C.c_float, C.c_float, _FP, _FP, _FP, _FP, _FP]),
    "evrp_bn_backward_reduce": (C.c_int, [_FP, _FP, C.c_int, _FP, _FP, _FP]),
    "evrp_bn_backward_apply": (C.c_int, [_FP, _FP, C.c_int, _FP, _FP, _FP, _FP, _FP]),
    "evrp_route_context_forward": (C.c_int, [_FP, _FP, _FP, C.c_int, C.c_int, C.c_int, _FP, _FP, _FP]),
    "evrp_route_context_backward": (C.c_int, [_FP, _FP, _FP, _FP, C.c_int, C.c_int, C.c_int, _FP, _FP, _FP]),
    "evrp_relu_mask": (C.c_int, [_FP, _FP, _FP, C.c_longlong, _FP]),
    "evrp_pack_decoder": (C.c_int, [_FP] * 14),
    "evrp_init_embed_forward": (C.c_int, [_FP, _FP, C.c_int, C.c_int, C.c_int, _FP, _FP, _FP, _FP, _FP, _FP, _FP, _FP]),
    "evrp_init_embed_backward_workspace_bytes": (C.c_size_t, []),
    "evrp_init_embed_backward": (C.c_int, [_FP, _FP, _FP, C.c_int, C.c_int, C.c_int, _FP, _FP, _FP, _FP, _FP, _FP, _FP, C.c_size_t,
                                           _FP]),
    "evrp_reinforce_loss": (C.c_int, [_FP, _FP, _FP, C.c_int, C.c_int, C.c_int, C.c_float, _FP, _FP, _FP, _FP, _FP, _FP, _FP]),
}

_lib = None


class EvrpLibraryError(RuntimeError):
    pass


def lib():
    """Load libevrp_b200.so once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EvrpLibraryError(
                f"{LIB_PATH} not found: build it with __graft_entry__.build() "
                "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU/PyTorch fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.evrp_abi_version() != ABI_VERSION:
            raise EvrpLibraryError("libevrp_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc, what):
    if rc == 0:
        return
    if rc <= -1000:
        names = {1000: "bad argument", 1001: f"more than {EVRP_MAX_S} nodes", 1002: "workspace too small"}
        raise EvrpLibraryError(f"{what}: {names.get(-rc, rc)}")
    raise EvrpLibraryError(f"{what}: CUDA error {-rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")


def ptr(t):
    """device pointer of a contiguous CUDA tensor (None -> NULL)"""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "evrp_b200 kernels need contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_phys(velocity=50., start_soc=80., t_limit=10., max_load=4, objective="EM", div_mode=0):
    """Folded constants evaluated with the reference's own Python expressions (problems/EVRP.py:29-41,247-252)."""
    Cd, A, Ad = 0.7, 6.66, 1.2041
    motor_d, motor_r, battery_d, battery_r = 1.18, 0.85, 1.11, 0.93
    charging_time, serve_time = 1, 0.33
    p = Phys()
    p.aero = 0.5 * Cd * A * Ad * (velocity / 3.6) ** 2
    p.kd = motor_d * battery_d
    p.kr = motor_r * battery_r
    p.g, p.Cr, p.velocity, p.mc, p.w, p.max_load = 9.81, 0.01, velocity, 4100, 1000, max_load
    p.start_soc, p.t_limit = start_soc, t_limit
    p.charging_time, p.serve_time = charging_time, serve_time
    p.charge_plus_serve = charging_time + serve_time
    p.inv_velocity, p.inv_3600 = 1.0 / velocity, 1.0 / 3600.
    p.div_mode = div_mode
    p.objective = 1 if objective == "DM" else 0
    return p
