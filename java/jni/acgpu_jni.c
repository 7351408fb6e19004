/*
 * JNI glue: com.roklenarcic.util.strings.gpu.AcGpuNative -> libacgpu.so (include/acgpu.h).
 * Build (outside this image; needs a JDK for jni.h):
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/jni/acgpu_jni.c \
 *       -Lahocorasick_b200 -lacgpu -o libacgpu_jni.so
 * No matching logic lives here: arrays are pinned, the C entry point is called, the records are copied into
 * Java int[]s.  Not compiled in this repository's image (no JDK) — see INTEGRATION.md.
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "acgpu.h"

static void throw_for(JNIEnv *env, int rc) {
    const char *cls = rc == ACGPU_EILLEGALARG ? "java/lang/IllegalArgumentException"
                    : rc == ACGPU_ENOMEM      ? "java/lang/OutOfMemoryError"
                                              : "java/lang/RuntimeException";
    (*env)->ThrowNew(env, (*env)->FindClass(env, cls), acgpu_last_error());
}

static jobjectArray wrap_result(JNIEnv *env, acgpu_result *r) {
    if (r->n > 0x3FFFFFFF) { /* 2 * n must fit a Java array length */
        acgpu_free_result(r);
        (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/OutOfMemoryError"), "more than 2^30 matches in one call");
        return NULL;
    }
    jclass obj = (*env)->FindClass(env, "java/lang/Object");
    jobjectArray out = obj ? (*env)->NewObjectArray(env, 3, obj, NULL) : NULL;
    jintArray pos = out ? (*env)->NewIntArray(env, (jsize)(2 * r->n)) : NULL;
    if (!pos) { /* OutOfMemoryError is pending */
        acgpu_free_result(r);
        return NULL;
    }
    if (r->n && r->pos) (*env)->SetIntArrayRegion(env, pos, 0, (jsize)(2 * r->n), (const jint *)r->pos); /* values-only streams: zeros */
    (*env)->SetObjectArrayElement(env, out, 0, pos);
    if (r->val) {
        jintArray val = (*env)->NewIntArray(env, (jsize)r->n);
        (*env)->SetIntArrayRegion(env, val, 0, (jsize)r->n, (const jint *)r->val);
        (*env)->SetObjectArrayElement(env, out, 1, val);
    } else if (r->n == 0) {
        (*env)->SetObjectArrayElement(env, out, 1, (*env)->NewIntArray(env, 0));
    }
    acgpu_free_result(r);
    return out;
}

JNIEXPORT jlong JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_create(
    JNIEnv *env, jclass c, jint family, jcharArray chars, jlongArray offsets, jbyteArray isNull, jlong nKeywords,
    jlong nValues, jboolean caseSensitive, jbooleanArray wordChars, jint device) {
    jchar *pc = (*env)->GetCharArrayElements(env, chars, NULL);
    jlong *po = (*env)->GetLongArrayElements(env, offsets, NULL);
    jbyte *pn = (*env)->GetByteArrayElements(env, isNull, NULL);
    jboolean *pw = wordChars ? (*env)->GetBooleanArrayElements(env, wordChars, NULL) : NULL;
    uint64_t h = 0;
    int rc = acgpu_create_from_keywords(family, (const uint16_t *)pc, (const int64_t *)po, (const uint8_t *)pn,
                                        nKeywords, nValues, caseSensitive ? 1 : 0, (const uint8_t *)pw, device, &h);
    (*env)->ReleaseCharArrayElements(env, chars, pc, JNI_ABORT);
    (*env)->ReleaseLongArrayElements(env, offsets, po, JNI_ABORT);
    (*env)->ReleaseByteArrayElements(env, isNull, pn, JNI_ABORT);
    if (pw) (*env)->ReleaseBooleanArrayElements(env, wordChars, pw, JNI_ABORT);
    if (rc != ACGPU_OK) throw_for(env, rc);
    return (jlong)h;
}

JNIEXPORT jbooleanArray JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_wordChars(JNIEnv *env, jclass c, jint mode,
                                                                                            jcharArray chars,
                                                                                            jbooleanArray toggles) {
    uint8_t local[65536];
    const jsize n = chars ? (*env)->GetArrayLength(env, chars) : 0;
    if (mode == 2 && (!toggles || (*env)->GetArrayLength(env, toggles) < n)) {
        (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/ArrayIndexOutOfBoundsException"), "toggleFlags shorter than wordCharacters");
        return NULL;
    }
    jchar *pc = n ? (*env)->GetCharArrayElements(env, chars, NULL) : NULL;
    jboolean *pt = (mode == 2 && n) ? (*env)->GetBooleanArrayElements(env, toggles, NULL) : NULL;
    int rc = acgpu_word_chars(mode, (const uint16_t *)pc, (const uint8_t *)pt, (int32_t)n, local);
    if (pc) (*env)->ReleaseCharArrayElements(env, chars, pc, JNI_ABORT);
    if (pt) (*env)->ReleaseBooleanArrayElements(env, toggles, pt, JNI_ABORT);
    if (rc != ACGPU_OK) {
        throw_for(env, rc);
        return NULL;
    }
    jbooleanArray out = (*env)->NewBooleanArray(env, 65536);
    (*env)->SetBooleanArrayRegion(env, out, 0, 65536, (const jboolean *)local);
    return out;
}

JNIEXPORT void JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_destroy(JNIEnv *env, jclass c, jlong h) {
    acgpu_destroy((uint64_t)h);
}

/* GetStringChars, not GetStringCritical: the call blocks on H2D copies, kernels and D2H copies for up to seconds, which a
 * JNI critical region must not do (it can stall the collector for every Java thread). */
JNIEXPORT jobjectArray JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_matchCompact(JNIEnv *env, jclass c, jlong h,
                                                                                              jstring haystack) {
    const jsize n = (*env)->GetStringLength(env, haystack);
    const jchar *p = (*env)->GetStringChars(env, haystack, NULL);
    if (!p) return NULL;
    acgpu_matches m;
    int rc = acgpu_match_utf16_compact((uint64_t)h, (const uint16_t *)p, (int32_t)n, &m);
    (*env)->ReleaseStringChars(env, haystack, p);
    if (rc != ACGPU_OK) {
        throw_for(env, rc);
        return NULL;
    }
    if (m.kind == ACGPU_MATCHES_RECORDS) {
        acgpu_result r = {m.n, m.pos, m.val};
        return wrap_result(env, &r);
    }
    jclass obj = (*env)->FindClass(env, "java/lang/Object");
    jobjectArray out = obj ? (*env)->NewObjectArray(env, 3, obj, NULL) : NULL;
    jcharArray masks = out ? (*env)->NewCharArray(env, (jsize)m.n_chars) : NULL;
    if (masks) {
        (*env)->SetCharArrayRegion(env, masks, 0, (jsize)m.n_chars, (const jchar *)m.masks);
        (*env)->SetObjectArrayElement(env, out, 2, masks);
    }
    acgpu_free_matches(&m);
    return masks ? out : NULL;
}

JNIEXPORT jobjectArray JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_match(JNIEnv *env, jclass c, jlong h,
                                                                                       jstring haystack) {
    const jsize n = (*env)->GetStringLength(env, haystack);
    const jchar *p = (*env)->GetStringChars(env, haystack, NULL);
    if (!p) return NULL;
    acgpu_result r;
    int rc = acgpu_match_utf16((uint64_t)h, (const uint16_t *)p, (int32_t)n, &r);
    (*env)->ReleaseStringChars(env, haystack, p);
    if (rc != ACGPU_OK) {
        throw_for(env, rc);
        return NULL;
    }
    return wrap_result(env, &r);
}

JNIEXPORT jlong JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_streamBegin(JNIEnv *env, jclass c, jlong h) {
    uint64_t s = 0;
    int rc = acgpu_stream_begin((uint64_t)h, &s);
    if (rc != ACGPU_OK) throw_for(env, rc);
    return (jlong)s;
}

JNIEXPORT void JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_streamValuesOnly(JNIEnv *env, jclass c, jlong s, jboolean on) {
    int rc = acgpu_stream_set_values_only((uint64_t)s, on ? 1 : 0);
    if (rc != ACGPU_OK) throw_for(env, rc);
}

JNIEXPORT jobjectArray JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_streamFeed(JNIEnv *env, jclass c,
                                                                                            jlong s, jcharArray buf,
                                                                                            jint n) {
    jchar *p = (*env)->GetCharArrayElements(env, buf, NULL); /* not a critical region: the feed blocks on the device */
    if (!p) return NULL;
    acgpu_result r;
    int rc = acgpu_stream_feed((uint64_t)s, (const uint16_t *)p, n, &r);
    (*env)->ReleaseCharArrayElements(env, buf, p, JNI_ABORT);
    if (rc != ACGPU_OK) {
        throw_for(env, rc);
        return NULL;
    }
    return wrap_result(env, &r);
}

JNIEXPORT jobjectArray JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_streamEnd(JNIEnv *env, jclass c,
                                                                                           jlong s) {
    acgpu_result r;
    int rc = acgpu_stream_end((uint64_t)s, &r);
    if (rc != ACGPU_OK) {
        throw_for(env, rc);
        return NULL;
    }
    return wrap_result(env, &r);
}

JNIEXPORT jint JNICALL Java_com_roklenarcic_util_strings_gpu_AcGpuNative_charBufferSize(JNIEnv *env, jclass c, jlong h) {
    int32_t cbs = 0;
    acgpu_info((uint64_t)h, NULL, NULL, NULL, &cbs, NULL);
    return cbs;
}
