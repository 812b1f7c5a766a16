"""ctypes binding of ``libdajin2_b200.so`` (the C ABI declared in ``include/dajin2_b200.h``).

There is no fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.
"""

from __future__ import annotations

import ctypes
from ctypes import c_double, c_int32, c_int64, c_uint32, c_void_p
from pathlib import Path

import os

# DAJIN2_B200_LIB points at an alternative build of the same library (kernel experiments under scripts/)
LIB_PATH = Path(os.environ.get("DAJIN2_B200_LIB") or Path(__file__).resolve().parent / "libdajin2_b200.so")

ABI_VERSION = 2
CODE_INVALID = 0x7F
CODE_INS = 0x80
FLAG_TOKEN_COUNT = 1
FLAG_BAD_TOKEN = 2
ENC_CHUNK = 1536
HIST_ROWS = 10
# rows of dj_hist (include/dajin2_b200.h: DJ_HIST_*)
HIST_MATCH, HIST_INS, HIST_DEL, HIST_SUB = 0, 1, 2, 3
HIST_SW_INS_P, HIST_SW_DEL_P, HIST_SW_SUB_P, HIST_SW_INS_M, HIST_SW_DEL_M, HIST_SW_SUB_M = 4, 5, 6, 7, 8, 9
MASK_INS, MASK_DEL, MASK_SUB = 1, 2, 4
KMER_AT, KMER_RAW = 256, 257


class KmerRecord(ctypes.Structure):
    _fields_ = [
        ("sel_index", ctypes.c_uint32),
        ("prev", ctypes.c_uint16),
        ("cur", ctypes.c_uint16),
        ("next", ctypes.c_uint16),
        ("reserved", ctypes.c_uint16),
        ("count_sample", ctypes.c_uint32),
        ("count_control", ctypes.c_uint32),
        ("example", ctypes.c_uint32),
        ("slot", ctypes.c_uint32),
    ]


class RowRecord(ctypes.Structure):
    _fields_ = [
        ("slot", ctypes.c_uint32),
        ("count", ctypes.c_uint32),
        ("count_plus", ctypes.c_uint32),
        ("first_row", ctypes.c_uint32),
        ("last_row", ctypes.c_uint32),
    ]


class ClusterStepOut(ctypes.Structure):
    _fields_ = [
        ("inertia", ctypes.c_double * 2),
        ("weight", ctypes.c_double * 2),
        ("silhouette", ctypes.c_double),
        ("iterations", ctypes.c_int32),
        ("control_clusters", ctypes.c_int32),
        ("sample_labels_distinct", ctypes.c_int32),
        ("seed_point", ctypes.c_int32 * 2),
    ]


# name -> (restype, argtypes); every name here must be exported by the library (tests/test_cabi.py)
SIGNATURES = {
    "dj_abi_version": (c_int32, []),
    "dj_last_error": (ctypes.c_char_p, []),
    "dj_device_info": (c_int32, [c_void_p]),
    "dj_code_pitch": (c_int32, [c_int32]),
    "dj_ckpt_pitch": (c_int32, [c_int32]),
    "dj_encode": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p, c_void_p]),
    "dj_encode_v1": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "dj_encode_ref": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dj_encode_ref_smem_bytes": (c_int64, [c_int32, c_int32]),
    "dj_hist": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "dj_window_cosine": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "dj_kmer_table_bytes": (c_int64, [c_uint32]),
    "dj_kmer_table_clear": (c_int32, [c_void_p, c_uint32, c_void_p]),
    "dj_kmer_count": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_uint32, c_void_p]),
    "dj_kmer_export": (c_int32, [c_void_p, c_uint32, c_void_p, c_uint32, c_void_p, c_void_p]),
    "dj_annotate": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                              c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_uint32,
                              c_void_p, c_void_p, c_void_p]),
    "dj_fetch_tokens": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32,
                                  c_void_p, c_void_p, c_int32, c_void_p]),
    "dj_row_table_bytes": (c_int64, [c_uint32]),
    "dj_row_table_clear": (c_int32, [c_void_p, c_uint32, c_void_p]),
    "dj_dedup_rows": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p,
                                c_void_p]),
    "dj_dedup_export": (c_int32, [c_void_p, c_uint32, c_void_p, c_uint32, c_void_p, c_void_p]),
    "dj_bisect": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_double,
                            c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dj_assign_rows": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "dj_silhouette": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "dj_cluster_step": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                  c_double, c_double, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dj_member_counts": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p]),
    "dj_member_locate": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "dj_gather_labels": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "dj_read_n_features": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "dj_loci_words": (c_int32, [c_int32]),
    "dj_sv_scan": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_uint32, c_void_p, c_void_p]),
    "dj_cs_midsv_measure": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "dj_cs_midsv_write": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
}

_NO_STATUS = {"dj_abi_version", "dj_last_error", "dj_code_pitch", "dj_ckpt_pitch", "dj_encode_ref_smem_bytes",
              "dj_kmer_table_bytes", "dj_row_table_bytes", "dj_loci_words"}

# kernels launched per call (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {
    "dj_encode": 1, "dj_encode_v1": 1, "dj_encode_ref": 1, "dj_hist": 2, "dj_window_cosine": 1, "dj_kmer_count": 1, "dj_kmer_export": 1, "dj_annotate": 1,
    "dj_fetch_tokens": 1, "dj_dedup_rows": 1, "dj_dedup_export": 1, "dj_bisect": 1, "dj_assign_rows": 1,
    "dj_silhouette": 2, "dj_cluster_step": 6, "dj_member_counts": 1, "dj_member_locate": 1, "dj_gather_labels": 1,
    "dj_read_n_features": 1, "dj_sv_scan": 1, "dj_cs_midsv_measure": 1, "dj_cs_midsv_write": 1,
}
launch_count = 0

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise if it has not been built (``python -m dajin2_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m dajin2_b200.build` (nvcc, sm_100a). "
            "dajin2_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dj_abi_version() != ABI_VERSION:
        raise RuntimeError("libdajin2_b200.so ABI version mismatch: rebuild the library")
    _lib = lib
    return lib


def call(name: str, *args):
    """Invoke an entry point that returns a status code; raise RuntimeError with the library's message on failure."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += KERNELS_PER_CALL.get(name, 0)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.dj_last_error().decode('utf-8', 'replace')}")
    return rc
