/*
 * jtb_jni.c — JNI shim between jtb.Native (Clojure/Java) and the C ABI of libjtb_check.so.
 *
 * UNCOMPILED HERE: this image has no JDK (no jni.h).  Build on a host with a JDK:
 *     gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../include \
 *         jtb_jni.c -L../jepsen_tigerbeetle_b200 -ljtb_check -o libjtb_jni.so
 * Pure marshalling: pins the primitive arrays produced by jtb.checker/flatten-history, fills a
 * `jtb_history`, calls the C entry point, releases.  Mirrors jepsen_tigerbeetle_b200/native.py 1:1.
 * A non-zero status is turned into a RuntimeException so that jepsen's check-safe yields
 * {:valid? :unknown :error ...} (SURVEY §8(b) error convention).
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "jtb_check.h"

#define PIN(T, name, arr) T* name = (arr) ? (T*)(*env)->GetPrimitiveArrayCritical(env, (arr), NULL) : NULL
#define UNPIN(name, arr) if (arr) (*env)->ReleasePrimitiveArrayCritical(env, (arr), (void*)(name), JNI_ABORT)

static void throw_rt(JNIEnv* env, const char* msg) {
    jclass c = (*env)->FindClass(env, "java/lang/RuntimeException");
    if (c) (*env)->ThrowNew(env, c, msg);
}

JNIEXPORT jlong JNICALL Java_jtb_Native_create(JNIEnv* env, jclass cls, jint device) {
    (void)cls;
    jtb_opts o;
    memset(&o, 0, sizeof o);
    o.device = device;
    jtb_ctx* ctx = jtb_create(&o);
    if (!ctx) throw_rt(env, "jtb_create failed: no CUDA device (there is no CPU fallback)");
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT void JNICALL Java_jtb_Native_destroy(JNIEnv* env, jclass cls, jlong h) {
    (void)env; (void)cls;
    jtb_destroy((jtb_ctx*)(intptr_t)h);
}

/* long[] checkLinearizable0(long ctx, byte[] type, byte[] f, byte[] flags, int[] process, int[] index,
 *   long[] time, int[] a, int[] b, int[] c, long[] payloadOff, int[] payloadLen, int[] payload,
 *   long[] shardOff, long[] keyIds, int modelKind, int initValue, int[] accounts, boolean negOk)
 * returns 6 longs per shard: valid, witness, previous-ok, cause, configs, probes. */
JNIEXPORT jlongArray JNICALL Java_jtb_Native_checkLinearizable0(
    JNIEnv* env, jclass cls, jlong h, jbyteArray type, jbyteArray f, jbyteArray flags, jintArray process,
    jintArray index, jlongArray time, jintArray a, jintArray b, jintArray c, jlongArray poff, jintArray plen,
    jintArray payload, jlongArray shard_off, jlongArray key_ids, jint model_kind, jint init_value,
    jintArray accounts, jboolean neg_ok) {
    (void)cls;
    jtb_history hist;
    jtb_model m;
    memset(&hist, 0, sizeof hist);
    memset(&m, 0, sizeof m);
    hist.n_events = (*env)->GetArrayLength(env, type);
    hist.n_payload = (*env)->GetArrayLength(env, payload);
    hist.n_shards = (*env)->GetArrayLength(env, shard_off) - 1;
    m.kind = model_kind;
    m.init_value = init_value;
    m.negative_balances_ok = neg_ok ? 1 : 0;
    if (accounts) {
        jint n = (*env)->GetArrayLength(env, accounts);
        if (n > JTB_MAX_ACCOUNTS) { throw_rt(env, "at most 8 accounts"); return NULL; }
        m.n_accounts = n;
        (*env)->GetIntArrayRegion(env, accounts, 0, n, (jint*)m.account_ids);
    }
    jtb_lin_shard* shards = (jtb_lin_shard*)calloc((size_t)hist.n_shards, sizeof *shards);
    jtb_lin_result res;
    PIN(uint8_t, p_type, type); PIN(uint8_t, p_f, f); PIN(uint8_t, p_flags, flags);
    PIN(int32_t, p_proc, process); PIN(int32_t, p_index, index); PIN(int64_t, p_time, time);
    PIN(int32_t, p_a, a); PIN(int32_t, p_b, b); PIN(int32_t, p_c, c);
    PIN(int64_t, p_poff, poff); PIN(int32_t, p_plen, plen); PIN(int32_t, p_payload, payload);
    PIN(int64_t, p_soff, shard_off); PIN(int64_t, p_keys, key_ids);
    hist.type = p_type; hist.f = p_f; hist.flags = p_flags; hist.process = p_proc; hist.index = p_index;
    hist.time_ns = p_time; hist.a = p_a; hist.b = p_b; hist.c = p_c; hist.payload_off = p_poff;
    hist.payload_len = p_plen; hist.payload = p_payload; hist.shard_off = p_soff; hist.key_ids = p_keys;
    int rc = jtb_check_linearizable((jtb_ctx*)(intptr_t)h, &hist, &m, shards, &res);
    UNPIN(p_keys, key_ids); UNPIN(p_soff, shard_off); UNPIN(p_payload, payload); UNPIN(p_plen, plen);
    UNPIN(p_poff, poff); UNPIN(p_c, c); UNPIN(p_b, b); UNPIN(p_a, a); UNPIN(p_time, time);
    UNPIN(p_index, index); UNPIN(p_proc, process); UNPIN(p_flags, flags); UNPIN(p_f, f); UNPIN(p_type, type);
    if (rc != 0) {
        throw_rt(env, jtb_last_error((jtb_ctx*)(intptr_t)h));
        free(shards);
        return NULL;
    }
    jlongArray out = (*env)->NewLongArray(env, 6 * hist.n_shards);
    for (int s = 0; s < hist.n_shards; ++s) {
        jlong v[6] = {shards[s].valid, shards[s].witness_index, shards[s].previous_ok_index, shards[s].cause,
                      (jlong)shards[s].configs_explored, (jlong)shards[s].probes};
        (*env)->SetLongArrayRegion(env, out, 6 * s, 6, v);
    }
    free(shards);
    return out;
}
/* int[] finalConfigs0(long ctx, <the same history and model arguments as checkLinearizable0>, int shard, int cap)
 * knossos' :configs of an INVALID shard (jtb_final_configs): same pin -> fill jtb_history -> call -> unpin pattern;
 * call it directly after checkLinearizable0 on the same arrays.  Returns the total number of such configurations
 * followed by min(cap, total) records of sizeof(jtb_final_config)/4 = 140 ints each, in the struct's field order:
 * state, balances[8], n_pending, n_linearized_open, n_crashed_linearized, pending_index[64],
 * linearized_open_index[64]. */
static jintArray final_configs_result(JNIEnv* env, const jtb_final_config* buf, int32_t cap, int64_t total) {
    const jsize rec = (jsize)(sizeof(jtb_final_config) / 4);
    const jsize n = (jsize)(total < cap ? total : cap);
    jintArray out = (*env)->NewIntArray(env, 1 + n * rec);
    jint t = (jint)total;
    (*env)->SetIntArrayRegion(env, out, 0, 1, &t);
    if (n) (*env)->SetIntArrayRegion(env, out, 1, n * rec, (const jint*)buf);
    return out;
}

/* checkSetFull0 / checkBankTotals0 follow the same pin -> fill jtb_history -> call -> unpin pattern
 * around jtb_check_set_full / jtb_check_bank_totals and return their result structs as long[]. */
