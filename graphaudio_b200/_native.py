"""ctypes binding of include/graphaudio_cuda.h (libgraphaudio_cuda.so).

The library is the product: there is no Python or CPU fallback.  Loading fails loudly if the shared
object is missing, and every compute entry point fails with GAC_ERR_NO_DEVICE when no B200 is visible.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libgraphaudio_cuda.so")

GAC_OK = 0
GAC_ERR_INVALID_ARGUMENT = -1
GAC_ERR_OUT_OF_RANGE = -2
GAC_ERR_INVALID_OPERATION = -3
GAC_ERR_DISPOSED = -4
GAC_ERR_NO_DEVICE = -5
GAC_ERR_CUDA = -6
GAC_ERR_OUT_OF_MEMORY = -7
GAC_ERR_NCCL = -8
GAC_ERR_UNSUPPORTED = -9

GAC_EVENT_EPOCH = 4
GAC_OP_BIQUAD, GAC_OP_GAIN, GAC_OP_CONVOLVER, GAC_OP_DELAY, GAC_OP_PANNER, GAC_OP_CHANNEL, GAC_OP_GATE = 1, 2, 3, 4, 5, 6, 7
GAC_SAMPLE_S16, GAC_SAMPLE_S24, GAC_SAMPLE_S32, GAC_SAMPLE_F32 = 0, 1, 2, 3

fp = C.POINTER(C.c_float)
fpp = C.POINTER(fp)


class gac_context_desc(C.Structure):
    _fields_ = [("sample_rate", C.c_int), ("quantum", C.c_int), ("partition", C.c_int), ("device_id", C.c_int),
                ("mac_variant", C.c_int), ("tile_blocks", C.c_int), ("flags", C.c_int), ("reserved", C.c_int)]


GAC_FLAG_ASYNC_UPLOAD = 1
GAC_FLAG_UNIFORM_SEGMENTS = 2
GAC_FLAG_NO_FANIN_FUSION = 4


class gac_event(C.Structure):
    _fields_ = [("type", C.c_int32), ("value", C.c_float), ("target", C.c_float), ("time", C.c_double),
                ("time_constant", C.c_double)]


class gac_param(C.Structure):
    _fields_ = [("value", C.c_float), ("n_events", C.c_int32), ("events", C.POINTER(gac_event)), ("mod_bus", C.c_int32),
                ("min_value", C.c_float), ("max_value", C.c_float), ("reserved", C.c_int32)]


class gac_op_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("filter_type", C.c_int32), ("p0", gac_param), ("p1", gac_param), ("p2", gac_param),
                ("ir", C.c_void_p), ("aux", C.c_double)]


class gac_voice_desc(C.Structure):
    _fields_ = [("source", C.c_void_p), ("start_when", C.c_double), ("start_offset", C.c_double),
                ("start_duration", C.c_double), ("stop_when", C.c_double), ("playback_rate", C.c_float),
                ("n_ops", C.c_int32), ("ops", C.POINTER(gac_op_desc)), ("bus", C.c_int32), ("input", C.c_int32),
                ("loop", C.c_int32), ("source_kind", C.c_int32), ("loop_start", C.c_double), ("loop_end", C.c_double),
                ("source_param", gac_param), ("oscillator_type", C.c_int32), ("reserved", C.c_int32)]


class gac_bus_desc(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("ops", C.POINTER(gac_op_desc)), ("target", C.c_int32), ("n_inputs", C.c_int32),
                ("inputs", C.POINTER(C.c_int32)), ("flags", C.c_int32), ("reserved", C.c_int32), ("input_slots", C.POINTER(C.c_int32))]


GAC_BUS_MONO_INPUT = 1
GAC_SOURCE_BUFFER, GAC_SOURCE_CONSTANT, GAC_SOURCE_OSCILLATOR = 0, 1, 2


class gac_graph_desc(C.Structure):
    _fields_ = [("n_voices", C.c_int32), ("voices", C.POINTER(gac_voice_desc)), ("n_buses", C.c_int32),
                ("buses", C.POINTER(gac_bus_desc)), ("n_dest_inputs", C.c_int32), ("dest_inputs", C.POINTER(C.c_int32))]


class gac_stats(C.Structure):
    _fields_ = [("ms_total", C.c_double), ("ms_source", C.c_double), ("ms_automation", C.c_double),
                ("ms_biquad", C.c_double), ("ms_gain", C.c_double), ("ms_fft_fwd", C.c_double), ("ms_mac", C.c_double),
                ("ms_fft_inv", C.c_double), ("ms_mix", C.c_double), ("ms_d2h", C.c_double), ("conv_units", C.c_int64),
                ("algorithmic_bytes", C.c_double), ("mac_complex_macs", C.c_double), ("kernel_launches", C.c_int64),
                ("voices", C.c_int64), ("frames", C.c_int64), ("mac_flops", C.c_double), ("mac_bytes_moved", C.c_double),
                ("mac_variant_used", C.c_int32), ("mac_big_segments", C.c_int32), ("ms_delay", C.c_double), ("ms_panner", C.c_double),
                ("mac_h2_bytes_single", C.c_double), ("fanin_groups", C.c_int32), ("fanin_members", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/graphaudio_cuda.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gac_version": (C.c_int, []),
    "gac_last_error": (C.c_char_p, []),
    "gac_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gac_context_create": (C.c_int, [C.POINTER(gac_context_desc), C.POINTER(C.c_void_p)]),
    "gac_context_destroy": (C.c_int, [C.c_void_p]),
    "gac_synchronize": (C.c_int, [C.c_void_p]),
    "gac_buffer_create": (C.c_int, [C.c_void_p, fpp, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "gac_buffer_create_interleaved": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_void_p)]),
    "gac_buffer_destroy": (C.c_int, [C.c_void_p]),
    "gac_ir_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "gac_ir_destroy": (C.c_int, [C.c_void_p]),
    "gac_graph_create": (C.c_int, [C.c_void_p, C.POINTER(gac_graph_desc), C.POINTER(C.c_void_p)]),
    "gac_graph_destroy": (C.c_int, [C.c_void_p]),
    "gac_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, fpp, C.c_int, C.c_int64]),
    "gac_render_interleaved": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, fp, C.c_int, C.c_int64]),
    "gac_render_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_int]),
    "gac_render_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int64, fpp, C.c_int]),
    "gac_comm_unique_id": (C.c_int, [C.c_void_p]),
    "gac_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "gac_comm_destroy": (C.c_int, [C.c_void_p]),
    "gac_render_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, fpp, C.c_int]),
    "gac_group_create": (C.c_int, [C.POINTER(gac_context_desc), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "gac_group_destroy": (C.c_int, [C.c_void_p]),
    "gac_group_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "gac_group_context": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "gac_group_render": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int64, C.c_int64, fpp, C.c_int, C.c_int64]),
    "gac_get_stats": (C.c_int, [C.c_void_p, C.POINTER(gac_stats)]),
    "gac_rfft_fwd_batch": (C.c_int, [C.c_void_p, fp, C.c_int, C.c_int64, fp]),
    "gac_spectral_mac": (C.c_int, [C.c_void_p, fp, fp, C.c_int, C.c_int64, C.c_int, C.c_int, fp]),
    "gac_irfft_ola_batch": (C.c_int, [C.c_void_p, fp, C.c_int, C.c_int64, fp]),
    "gac_convolve_batch": (C.c_int, [C.c_void_p, fp, C.c_int, C.c_int64, fp, C.c_int64, C.c_int, fp]),
    "gac_automation_eval": (C.c_int, [C.c_void_p, C.POINTER(gac_param), C.c_int, C.c_int64, fp]),
    "gac_resample_cubic": (C.c_int, [C.c_void_p, fp, C.c_int64, C.c_double, C.c_int64, fp, C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int64)]),
    "gac_biquad_batch": (C.c_int, [C.c_void_p, fp, C.c_int, C.c_int64, C.POINTER(C.c_int), C.POINTER(gac_param), C.POINTER(gac_param),
                                   C.POINTER(gac_param), fp]),
    "gac_mix": (C.c_int, [C.c_void_p, fpp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), fp, C.c_int, C.c_int64, fp]),
    "gac_plan_segments": (C.c_int, [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gac_convolver_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "gac_convolver_destroy": (C.c_int, [C.c_void_p]),
    "gac_convolver_channels": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gac_convolver_reset": (C.c_int, [C.c_void_p]),
    "gac_convolver_process_block": (C.c_int, [C.c_void_p, fpp, C.c_int, fpp, C.c_int]),
    "gac_convolver_process": (C.c_int, [C.c_void_p, fpp, C.c_int, fpp, C.c_int, C.c_int64]),
}

_lib = None


def lib():
    """Loads libgraphaudio_cuda.so (no fallback: raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m graphaudio_b200.build` "
                "(graphaudio_b200 has no CPU or pure-Python fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().gac_last_error().decode("utf-8", "replace")
