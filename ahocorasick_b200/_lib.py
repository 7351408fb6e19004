"""ctypes binding of libacgpu.so (C ABI: include/acgpu.h).

There is no fallback: if the CUDA library is missing or no device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACGPU_LIB") or os.path.join(_HERE, "libacgpu.so")  # ACGPU_LIB: A/B experiment builds only

OK, EINVAL, EILLEGALARG, ENODEVICE, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
NO_VALUE = 0xFFFFFFFF

AHOCORASICK, LONGEST, SHORTEST, WHOLEWORD, WHOLEWORDLONGEST = 0, 1, 2, 3, 4

# every symbol include/acgpu.h declares (tests check the .so exports all of them)
EXPORTS = [
    "acgpu_create_from_keywords", "acgpu_build_fingerprint", "acgpu_destroy", "acgpu_word_chars", "acgpu_info", "acgpu_char_classes",
    "acgpu_match_utf16", "acgpu_free_result", "acgpu_match_device", "acgpu_match_device_async",
    "acgpu_launches_per_match", "acgpu_stream_begin", "acgpu_stream_feed", "acgpu_stream_end",
    "acgpu_last_error", "acgpu_version", "acgpu_match_utf16_compact", "acgpu_free_matches", "acgpu_masks_to_records",
    "acgpu_chain_shard_layout", "acgpu_chain_shard_begin", "acgpu_chain_shard_finish", "acgpu_stream_set_values_only",
    "acgpu_create", "acgpu_desc_fingerprint",
]


class Result(C.Structure):
    _fields_ = [("n", C.c_int64), ("pos", C.POINTER(C.c_int32)), ("val", C.POINTER(C.c_uint32))]


class Matches(C.Structure):
    """acgpu_matches: records, or per-char hit masks for dense AhoCorasickSet streams (include/acgpu.h)."""
    _fields_ = [("n", C.c_int64), ("kind", C.c_int32), ("reserved", C.c_int32), ("pos", C.POINTER(C.c_int32)),
                ("val", C.POINTER(C.c_uint32)), ("masks", C.POINTER(C.c_uint16)), ("n_chars", C.c_int64)]


MATCHES_RECORDS, MATCHES_MASKS = 0, 1


class AcgpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("acgpu error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


_lib = None


def lib():
    """Load libacgpu.so; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libacgpu.so is missing (%s). Build it with `python -m ahocorasick_b200.build` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64
    L.acgpu_create_from_keywords.restype = C.c_int
    L.acgpu_create_from_keywords.argtypes = [C.c_int, vp, vp, vp, i64, i64, C.c_int, vp, C.c_int, C.POINTER(u64)]
    L.acgpu_create.restype = C.c_int
    L.acgpu_create.argtypes = [vp, C.POINTER(u64)]
    L.acgpu_desc_fingerprint.restype = C.c_int
    L.acgpu_desc_fingerprint.argtypes = [vp, C.POINTER(u64)]
    L.acgpu_build_fingerprint.restype = C.c_int
    L.acgpu_build_fingerprint.argtypes = [C.c_int, vp, vp, vp, i64, i64, C.c_int, vp, C.POINTER(u64), C.POINTER(C.c_double)]
    L.acgpu_destroy.restype = C.c_int
    L.acgpu_destroy.argtypes = [u64]
    L.acgpu_word_chars.restype = C.c_int
    L.acgpu_word_chars.argtypes = [C.c_int, vp, vp, i32, vp]
    L.acgpu_info.restype = C.c_int
    L.acgpu_info.argtypes = [u64, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    L.acgpu_char_classes.restype = C.c_int
    L.acgpu_char_classes.argtypes = [u64, vp, C.POINTER(i32)]
    L.acgpu_match_utf16.restype = C.c_int
    L.acgpu_match_utf16.argtypes = [u64, vp, i32, C.POINTER(Result)]
    L.acgpu_free_result.restype = None
    L.acgpu_free_result.argtypes = [C.POINTER(Result)]
    L.acgpu_match_device.restype = C.c_int
    L.acgpu_match_device.argtypes = [u64, vp, i64, i64, i64, vp, vp, i64, C.POINTER(i64), vp]
    L.acgpu_match_device_async.restype = C.c_int
    L.acgpu_match_device_async.argtypes = [u64, vp, i64, i64, i64, vp, vp, i64, vp, vp]
    L.acgpu_launches_per_match.restype = C.c_int
    L.acgpu_launches_per_match.argtypes = [u64]
    L.acgpu_stream_begin.restype = C.c_int
    L.acgpu_stream_begin.argtypes = [u64, C.POINTER(u64)]
    L.acgpu_stream_set_values_only.restype = C.c_int
    L.acgpu_stream_set_values_only.argtypes = [u64, C.c_int]
    L.acgpu_stream_feed.restype = C.c_int
    L.acgpu_stream_feed.argtypes = [u64, vp, i32, C.POINTER(Result)]
    L.acgpu_stream_end.restype = C.c_int
    L.acgpu_stream_end.argtypes = [u64, C.POINTER(Result)]
    L.acgpu_match_utf16_compact.restype = C.c_int
    L.acgpu_match_utf16_compact.argtypes = [u64, vp, i32, C.POINTER(Matches)]
    L.acgpu_free_matches.restype = None
    L.acgpu_free_matches.argtypes = [C.POINTER(Matches)]
    L.acgpu_masks_to_records.restype = i64
    L.acgpu_masks_to_records.argtypes = [vp, i64, i64, vp, i64]
    L.acgpu_chain_shard_layout.restype = C.c_int
    L.acgpu_chain_shard_layout.argtypes = [u64, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.acgpu_chain_shard_begin.restype = C.c_int
    L.acgpu_chain_shard_begin.argtypes = [u64, vp, i64, i64, vp, C.POINTER(u64), vp]
    L.acgpu_chain_shard_finish.restype = C.c_int
    L.acgpu_chain_shard_finish.argtypes = [u64, i32, i32, vp, vp, i64, vp, vp]
    L.acgpu_last_error.restype = C.c_char_p
    L.acgpu_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int):
    if rc != OK:
        msg = lib().acgpu_last_error().decode("utf-8", "replace")
        if rc == EILLEGALARG:
            raise IllegalArgumentException(msg)
        raise AcgpuError(rc, msg)


class IllegalArgumentException(ValueError):
    """Mirror of java.lang.IllegalArgumentException thrown by the WholeWord constructors
    (reference WholeWordMatchSet.java:149-153)."""
