package com.roklenarcic.util.strings.gpu;

/**
 * JNI veneer over libacgpu.so (C ABI: include/acgpu.h).  One native method per C entry point; no logic.
 * Source only in this repository: the build image has no JDK (SURVEY.md §0), so this file and
 * java/jni/acgpu_jni.c are exercised by a maintainer outside it.  Everything they call is tested
 * through the same C ABI from ctypes (tests/test_gpu_parity.py).
 */
final class AcGpuNative {
    static {
        System.loadLibrary("acgpu_jni"); // links against libacgpu.so
    }

    static final int AHOCORASICK = 0, LONGEST = 1, SHORTEST = 2, WHOLEWORD = 3, WHOLEWORDLONGEST = 4;

    private AcGpuNative() {
    }

    /** acgpu_create_from_keywords; nValues = -1 for a Set.  Throws IllegalArgumentException on ACGPU_EILLEGALARG. */
    static native long create(int family, char[] chars, long[] offsets, byte[] isNull, long nKeywords, long nValues,
            boolean caseSensitive, boolean[] wordChars, int device);

    /**
     * acgpu_word_chars: the boolean[65536] of WordCharacters.generateWordCharsFlags (WordCharacters.java:6-39; that class
     * is package-private in the reference, so the facade cannot call it).  mode 0 default, 1 only chars, 2 default + toggles.
     */
    static native boolean[] wordChars(int mode, char[] chars, boolean[] toggles);

    /** acgpu_destroy */
    static native void destroy(long handle);

    /**
     * acgpu_match_utf16 on the pinned char[] of the String (GetStringCritical).  Returns {pos, val}: pos = int[2*n]
     * (start,end) pairs in listener order, val = int[n] value indices (null for Sets).
     */
    static native Object[] match(long handle, String haystack);
    /** acgpu_match_utf16_compact: {int[] pos, int[] valueIdx, char[] masks} - masks != null for dense AhoCorasickSet streams. */
    static native Object[] matchCompact(long handle, String haystack);

    /** acgpu_stream_begin / feed / end; feed and end return {pos, val} like match (only val is used). */
    static native long streamBegin(long handle);
    static native void streamValuesOnly(long stream, boolean on);   // acgpu_stream_set_values_only

    static native Object[] streamFeed(long stream, char[] buf, int n);

    static native Object[] streamEnd(long stream);

    /** acgpu_info()[3]: charBufferSize (AhoCorasickMap.java:53) */
    static native int charBufferSize(long handle);
}
