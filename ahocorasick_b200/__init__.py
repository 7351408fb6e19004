"""ahocorasick_b200 — B200-native (sm_100a) drop-in for the matching hot path of RokLenarcic/AhoCorasick.

Host-side mirror of the reference's public Java API (package com.roklenarcic.util.strings):
same class names, constructor argument order, listener contract and error behaviour; the scan
itself runs in libacgpu.so (CUDA).  There is no CPU fallback.
"""
from ._lib import AcgpuError, IllegalArgumentException
from .matchers import (AhoCorasickMap, AhoCorasickSet, LongestMatchMap, LongestMatchSet, MapMatchListener,
                       RangeNodeThreshold, ReadableMatchListener, SetMatchListener, ShortestMatchMap,
                       ShortestMatchSet, StringMap, StringSet, Thresholder, WholeWordLongestMatchMap,
                       WholeWordLongestMatchSet, WholeWordMatchMap, WholeWordMatchSet, WordCharacters)

__all__ = [
    "AcgpuError", "IllegalArgumentException", "AhoCorasickMap", "AhoCorasickSet", "LongestMatchMap",
    "LongestMatchSet", "MapMatchListener", "RangeNodeThreshold", "ReadableMatchListener", "SetMatchListener",
    "ShortestMatchMap", "ShortestMatchSet", "StringMap", "StringSet", "Thresholder", "WholeWordLongestMatchMap",
    "WholeWordLongestMatchSet", "WholeWordMatchMap", "WholeWordMatchSet", "WordCharacters",
]
