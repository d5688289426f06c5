"""ctypes binding of libbpmf_b200.so (the C-ABI declared in include/bpmf_b200.h).

There is no CPU fallback: if the shared object is missing the import of any compute module raises,
and every compute entry point raises :class:`BpmfError` when the library reports a failure
(no CUDA device, bad shapes, out of memory).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BPMF_B200_LIB overrides the library path (A/B runs of kernel variants)
LIB_PATH = os.environ.get("BPMF_B200_LIB") or os.path.join(HERE, "lib", "libbpmf_b200.so")

OK = 0
MF_AUTO, MF_FP32, MF_TC_FP16, MF_TC_FP16X3, MF_TC_MIXED = 0, 1, 2, 3, 4
PRECISION_NAMES = {MF_FP32: "fp32", MF_TC_FP16: "fp16", MF_TC_FP16X3: "fp16x3",
                   MF_TC_MIXED: "fp16 / fp16x3 per template"}
PRECISIONS = {"auto": MF_AUTO, "fp32": MF_FP32, "fp16": MF_TC_FP16, "tc": MF_TC_FP16,
              "fp16x3": MF_TC_FP16X3, "tc3": MF_TC_FP16X3}


class BpmfError(RuntimeError):
    pass


fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)
z = C.c_size_t
vp = C.c_void_p

_SIGNATURES = {
    # name: (restype, argtypes)
    "bpmf_last_error": (C.c_char_p, []),
    "bpmf_last_status": (C.c_int, []),
    "bpmf_version": (C.c_char_p, []),
    "bpmf_device_count": (C.c_int, []),
    "bpmf_launch_count": (C.c_ulonglong, []),
    "bpmf_set_reserved_sms": (None, [C.c_int]),
    "bpmf_set_kernel_timing": (None, [C.c_int]),
    "bpmf_main_kernel_times": (C.c_int, [fp, C.c_int]),
    "caract_GPU": (None, []),
    "bpmf_mf_n_corr": (z, [z, z, z]),
    "bpmf_matched_filter": (C.c_int, [fp, ip, fp, fp, z, z, z, z, z, z, C.c_int, C.c_int, C.c_int, fp]),
    "bpmf_mf_workspace_bytes": (z, [z, z, z, z, z, z, C.c_int, C.c_int, C.c_int]),
    "bpmf_matched_filter_dev": (C.c_int, [vp, vp, vp, vp, z, z, z, z, z, z, C.c_int, C.c_int, C.c_int,
                                          vp, z, vp, z, vp]),
    "bpmf_mf_trace_workspace_bytes": (z, [z, z, z, z, z, C.c_int, C.c_int, C.c_int]),
    "bpmf_mf_template_workspace_bytes": (z, [z, z, z, z, z, z, C.c_int, C.c_int, C.c_int]),
    "bpmf_mf_prepare_traces_dev": (C.c_int, [vp, z, z, z, z, z, C.c_int, C.c_int, C.c_int, vp, z, vp]),
    "bpmf_mf_run_dev": (C.c_int, [vp, vp, vp, vp, z, z, z, z, z, z, C.c_int, C.c_int, C.c_int, vp, z, vp, z,
                                  vp, z, vp]),
    "bpmf_release_cache": (None, []),
    "bpmf_mf_hold_traces": (C.c_int, [fp, z, z, z]),
    "bpmf_mf_release_traces": (None, []),
    "bpmf_mf_resolve_precision": (C.c_int, [z, z, C.c_int, C.c_int, C.c_int]),
    "bpmf_mf_last_precision": (C.c_int, []),
    "matched_filter": (None, [fp, fp, ip, fp, fp, z, z, z, z, z, z, z, C.c_int, fp]),
    "matched_filter_no_sum": (None, [fp, fp, ip, fp, fp, z, z, z, z, z, z, z, C.c_int, fp]),
    "bpmf_beamform_max": (C.c_int, [fp, ip, fp, fp, z, z, z, z, z, C.c_int, C.c_int, fp, ip]),
    "bpmf_beamform_full": (C.c_int, [fp, ip, fp, fp, z, z, z, z, z, C.c_int, fp]),
    "bpmf_beam_workspace_bytes": (z, [z, z, z, z, z]),
    "bpmf_beamform_max_dev": (C.c_int, [vp, vp, vp, vp, z, z, z, z, z, C.c_int, C.c_int, vp, z, vp, vp, vp]),
    "bpmf_beamform_full_dev": (C.c_int, [vp, vp, vp, vp, z, z, z, z, z, C.c_int, vp, z, vp, vp]),
    "bpmf_beam_merge_dev": (C.c_int, [vp, vp, vp, vp, z, vp]),
    "bpmf_intertemplate_cc": (C.c_int, [fp, fp, C.c_char_p, z, z, z, z, z, fp]),
    "bpmf_intertemplate_workspace_bytes": (z, [z, z, z, z]),
    "bpmf_intertemplate_cc_dev": (C.c_int, [vp, vp, vp, z, z, z, z, z, vp, z, vp, vp]),
    "bpmf_moveouts_from_travel_times": (C.c_int, [vp, C.c_int, z, z, z, C.c_double, C.c_double, C.c_int, ip,
                                                  C.POINTER(C.c_double)]),
    "bpmf_moveouts_from_travel_times_dev": (C.c_int, [vp, C.c_int, z, z, z, C.c_double, C.c_double, C.c_int, vp, vp,
                                                      vp, vp]),
    "bpmf_detect_peaks_workspace_bytes": (z, [z]),
    "bpmf_detect_peaks": (C.c_int, [fp, fp, z, z, ip, z, C.POINTER(z)]),
    "bpmf_detect_peaks_dev": (C.c_int, [vp, vp, z, z, vp, z, vp, z, vp, vp]),
    "bpmf_normalize_traces": (C.c_int, [fp, z, z, fp]),
    "bpmf_normalize_traces_dev": (C.c_int, [vp, z, z, vp, vp, vp]),
    "kurtosis": (None, [fp, C.c_int, C.c_int, C.c_int, C.c_int, fp]),
    "find_similar_moveouts": (None, [fp, fp, fp, fp, fp, C.c_float, z, z, z, z, z, C.c_int, ip]),
    "find_similar_moveouts2": (None, [fp, fp, fp, fp, fp, C.c_float, z, z, z, z, z, C.c_int, ip]),
    "select_cc_indexes": (None, [fp, fp, z, z, ip]),
    "time_dependent_threshold": (None, [fp, fp, C.c_float, z, z, z, C.c_int, fp]),
    "swap": (None, [fp, fp]),
    "selectionSort": (None, [fp, C.c_int]),
    "sort_indexes": (None, [fp, C.POINTER(z), z]),
    "_compute_average": (C.c_float, [fp, z]),
    "_compute_std": (C.c_float, [fp, C.c_float, z]),
    "bpmf_threshold_workspace_bytes": (z, [z, z, z, z]),
    "bpmf_detect_dev": (C.c_int, [vp, z, z, vp, C.c_float, z, z, z, z, C.c_int, vp, z, vp, vp, vp, vp, vp, z, vp]),
    "bpmf_detect_gated_dev": (C.c_int, [vp, z, z, vp, C.c_float, z, z, z, z, C.c_int, vp, C.c_int, C.c_float, vp, vp, z,
                                        vp, vp, vp, vp, vp, z, vp]),
    "bpmf_time_dependent_threshold_dev": (C.c_int, [vp, z, z, vp, C.c_float, z, z, z, vp, z, vp, z, vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library once; raise loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BpmfError(
            "libbpmf_b200.so not found at %s -- build it with `python -m seismic_bpmf_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != OK:
        msg = load().bpmf_last_error()
        raise BpmfError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


def check_void(what):
    """After a reference-named (void) symbol: raise if the calling thread's last such call failed."""
    check(load().bpmf_last_status(), what)


def require_device():
    n = load().bpmf_device_count()
    if n <= 0:
        raise BpmfError("no CUDA device is visible: seismic_bpmf_b200 has no CPU fallback")
    return n
