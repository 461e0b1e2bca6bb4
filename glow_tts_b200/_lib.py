"""ctypes binding of libglowcore.so (include/glowcore.h).

The product path has no CPU fallback: if the shared library is missing or a
call fails, a ``GlowCoreError`` is raised -- loudly.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libglowcore.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "glowcore.h")

GLOW_F32, GLOW_I32, GLOW_BF16, GLOW_BF16_SIMT, GLOW_F32_TC = 0, 1, 2, 3, 4

_c = ctypes
_P, _I, _F, _Z, _U64, _U32 = _c.c_void_p, _c.c_int, _c.c_float, _c.c_size_t, _c.c_uint64, _c.c_uint32

class FlowConfig(ctypes.Structure):          # glow_flow_config
    _fields_ = [("blocks", _I), ("channels", _I), ("hidden", _I), ("layers", _I), ("kernel", _I),
                ("split", _I), ("spk_dim", _I), ("dropout", _F)]


class FlowCall(ctypes.Structure):            # glow_flow_call
    _fields_ = [("cfg", FlowConfig), ("precision", _I), ("batch", _I), ("t_max", _I), ("rows_pad", _I),
                ("training", _I), ("seed", _U64), ("step_dev", _P),
                ("row_utt", _P), ("row_t", _P), ("utt_off", _P), ("utt_len", _P),
                ("wpack", _P), ("wpack_tc", _P), ("spk", _P),
                ("ws_f32", _P), ("ws_act", _P), ("bw_f32", _P), ("bw_act", _P), ("stream", _P)]


class AttnCall(ctypes.Structure):            # glow_attn_call
    _fields_ = [("q", _P), ("k", _P), ("v", _P), ("wk", _P), ("wv", _P), ("lengths", _P), ("mask", _P),
                ("batch", _I), ("heads", _I), ("t", _I), ("head_dim", _I), ("window", _I),
                ("dropout", _F), ("seed", _U64), ("step_dev", _P), ("utt_off", _P), ("ld", _I), ("stream", _P)]


class RowsConvCall(ctypes.Structure):        # glow_rows_conv_call
    _fields_ = [("cin", _I), ("cout", _I), ("taps", _I), ("rows_pad", _I), ("row_utt", _P),
                ("relu", _I), ("p_out", _F), ("seed_out", _U64), ("step_dev", _P), ("stream", _P)]


class RowsNormCall(ctypes.Structure):        # glow_rows_norm_call
    _fields_ = [("rows_pad", _I), ("channels", _I), ("row_utt", _P), ("eps", _F),
                ("p_in", _F), ("seed_in", _U64), ("relu", _I), ("p_out", _F), ("seed_out", _U64),
                ("step_dev", _P), ("stream", _P)]


_PCFG, _PCALL, _PATTN = ctypes.POINTER(FlowConfig), ctypes.POINTER(FlowCall), ctypes.POINTER(AttnCall)
_PROWS = ctypes.POINTER(RowsConvCall)
_PNORM = ctypes.POINTER(RowsNormCall)

# name -> (restype, argtypes); kept in step with include/glowcore.h
# (tests/test_abi.py parses the header and checks every declared symbol is here
# and exported by the .so).
SIGNATURES = {
    "glow_abi_version": (_I, []),
    "glow_last_error": (_c.c_char_p, []),
    "glow_launch_count": (_U64, []),
    "glow_prof_enable": (_I, [_I]),
    "glow_prof_report": (_I, [_c.c_char_p, _Z]),
    "glow_mas_workspace_bytes": (_Z, [_I, _I, _I]),
    "glow_mas_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _I, _F, _P, _Z, _P]),
    "glow_mas_forward_host": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _I]),
    "glow_mas_align": (_I, [_P, _P, _P, _I, _I, _I, _P, _I, _F, _P, _P, _P]),
    "glow_align_logp": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "glow_align_expand_forward": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "glow_align_expand_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
    "glow_mle_loss_workspace_floats": (_Z, []),
    "glow_mle_loss_forward": (_I, [_P, _P, _P, _P, _P, _I, _Z, _I, _I, _P, _P, _P]),
    "glow_mle_loss_backward": (_I, [_P, _P, _P, _P, _P, _I, _Z, _P, _P, _P, _P, _P]),
    "glow_flow_param_slots": (_I, [_PCFG]),
    "glow_flow_wpack_floats": (_Z, [_PCFG]),
    "glow_flow_wpack_tc_elems": (_Z, [_PCFG]),
    "glow_flow_wpack_tc_elems_for": (_Z, [_PCFG, _I]),
    "glow_flow_wait_block_grads": (_I, [_P, _I]),
    "glow_flow_workspace_elems": (_I, [_PCFG, _I, _I, _I, ctypes.POINTER(_Z)]),
    "glow_flow_prepare": (_I, [_PCFG, _P, _P, _I, _P, _P, _P]),
    "glow_flow_forward": (_I, [_PCALL, _P, _P, _P]),
    "glow_flow_reverse": (_I, [_PCALL, _P, _P, _F]),
    "glow_flow_pack_rows": (_I, [_PCALL, _P, _P]),
    "glow_actnorm_stats": (_I, [_P, _P, _I, _I, _P, _P]),
    "glow_flow_block_forward": (_I, [_PCALL, _I, _P, _P]),
    "glow_flow_backward": (_I, [_PCALL, _P, _P, _P, _P, _P]),
    "glow_flow_backward_params": (_I, [_PCALL, _P, _P, _P, _P, _P, _P, _P, _P]),
    "glow_flow_param_grads": (_I, [_PCFG, _P, _P, _P, _P, _P, _P, _I, _P, _P]),
    "glow_conv_wgrad": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _I, _I, _P, _I, _c.c_longlong, _I, _I, _P]),
    "glow_rpr_attention_forward": (_I, [_PATTN, _P, _P, _P]),
    "glow_rpr_attention_backward": (_I, [_PATTN, _P, _P, _P, _P, _P, _P, _P, _P]),
    "glow_rows_conv_slab_elems": (_Z, [_I, _I, _I]),
    "glow_rows_conv_pack": (_I, [_PROWS, _P, _P, _P]),
    "glow_rows_conv_pack_multi": (_I, [_I, _P, _P, _P, _P, _P]),
    "glow_rows_conv_forward": (_I, [_PROWS, _P, _P, _P, _P]),
    "glow_rows_conv_backward_data": (_I, [_PROWS, _P, _P, _P]),
    "glow_rows_conv_backward_weight": (_I, [_PROWS, _P, _P, _P, _P]),
    "glow_rows_conv_backward_weight_accum": (_I, [_PROWS, _P, _P, _P, _P, _P]),
    "glow_side_join": (_I, [_P]),
    "glow_rows_act_backward": (_I, [_P, _I, _I, _I, _F, _U64, _P, _P, _P, _P, _P]),
    "glow_rows_norm_forward": (_I, [_PNORM, _P, _P, _P, _P, _P, _P, _P]),
    "glow_rows_norm_backward": (_I, [_PNORM, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "glow_sqnorm": (_I, [_P, _Z, _P, _P]),
    "glow_radam_step": (_I, [_P, _P, _P, _P, _Z, _F, _F, _F, _F, _F, _F, _I, _F, _F, _P, _P, _P]),
    "glow_radam_step_dev": (_I, [_P, _P, _P, _P, _Z, _P, _P, _P, _P]),
    "glow_selftest_umma_mn": (_I, [_P, _P, _P, _I, _I, _U32, _U32, _P]),
    "glow_selftest_umma": (_I, [_P, _P, _P, _I, _I, _I, _I, _U32, _U32, _U32, _U32, _I, _P]),
}


class GlowCoreError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libglowcore.so once; fail loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GlowCoreError(
                "libglowcore.so is not built (%s). Run `python -m glow_tts_b200.csrc.build` "
                "or __graft_entry__.build(); there is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().glow_last_error().decode("utf-8", "replace")
        raise GlowCoreError("%s failed with code %d: %s" % (what, rc, msg))


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise GlowCoreError(
            "%s must live on a CUDA device: glow_tts_b200 runs only on the sm_100a kernels "
            "in libglowcore.so (no CPU fallback)" % name)


# ---- device step counter (dropout under CUDA-graph replay) -------------------------------
# One int64 per device, owned by whoever drives training (train.TrainStep).  When registered,
# every flow / attention call hands its address to the kernels, which mix *counter into their
# dropout seed: a call captured in a CUDA graph then draws fresh masks on every replay.
_STEP_COUNTERS = {}


def set_step_counter(device, tensor):
    key = str(torch.device(device))
    if tensor is None:
        _STEP_COUNTERS.pop(key, None)
    else:
        assert tensor.dtype == torch.int64 and tensor.numel() == 1 and tensor.is_cuda
        _STEP_COUNTERS[key] = tensor


def step_counter_ptr(device):
    t = _STEP_COUNTERS.get(str(torch.device(device)))
    return None if t is None else t.data_ptr()


# ---- keeping cache-owned objects alive for captured CUDA graphs ----------------------------------
# A captured graph holds RAW POINTERS to whatever the step touched: workspaces, row maps, token rows, length
# tensors.  Those live in size-bounded caches (FlowPlan._ws, flow._ROWMAP_CACHE, rows._CACHE, _DEV_INTS below)
# that evict when other geometries come by -- after which a replay would read and write freed memory.  While a
# `capture_keepalive()` block is active every cache hands the objects it returns to the block's list as well; the
# owner of the graph keeps that list for as long as the graph lives.
_KEEPALIVE = []


class capture_keepalive:
    def __init__(self):
        self.objects = []

    def __enter__(self):
        _KEEPALIVE.append(self.objects)
        return self.objects

    def __exit__(self, *exc):
        _KEEPALIVE.remove(self.objects)
        return False


def keepalive(obj):
    for lst in _KEEPALIVE:
        lst.append(obj)
    return obj


# ---- small host lists -> cached device tensors ---------------------------------------------
# Lengths come from the host collater; turning them into device tensors is an H2D copy from
# pageable memory, which is a sync point and illegal inside CUDA-graph capture.  Steps that see
# the same lengths again (a captured step does, by construction) reuse the device copy.
_DEV_INTS = {}


def device_ints(values, dtype, device):
    key = (tuple(int(v) for v in values), dtype, str(torch.device(device)))
    t = _DEV_INTS.get(key)
    if t is None:
        if len(_DEV_INTS) > 512:
            _DEV_INTS.clear()
        t = _DEV_INTS[key] = torch.tensor(list(key[0]), dtype=dtype).to(device)
    return keepalive(t)


def side_join(device):
    """Make the current stream wait for everything libglowcore forked to its side stream (encoder wgrads)."""
    with torch.cuda.device(device):
        check(lib().glow_side_join(stream_ptr(device)), "glow_side_join")


def launch_count():
    return int(lib().glow_launch_count())


def prof_enable(on):
    check(lib().glow_prof_enable(int(bool(on))), "glow_prof_enable")


def prof_report():
    """{kernel family: (launches, total_ms)} since profiling was enabled; clears the log."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().glow_prof_report(buf, len(buf)), "glow_prof_report")
    out = {}
    for line in buf.value.decode().splitlines():
        name, n, ms = line.rsplit(" ", 2)
        out[name] = (int(n), float(ms))
    return out


def header_symbols():
    """Function names declared in include/glowcore.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(glow_[a-z0-9_]+)\s*\(", text)))
