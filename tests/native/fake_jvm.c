/*
 * fake_jvm.c — TEST INFRASTRUCTURE: a JNIEnv over plain C arrays (the function table of jni/stub/jni.h), so that the
 * JNI shim jni/jtb_jni.c can be compiled, linked against libjtb_check.so and driven end to end in an image without a
 * JDK.  tests/test_jni_shim.py calls the fj_* wrappers through ctypes and compares what comes out of the shim with the
 * direct ctypes binding of the same C ABI.
 */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct fake_arr {
    char kind; /* 'b' 'i' 'l' 'o' */
    int len;
    void* data;
} fake_arr;

static char g_exc[4096];
static int g_has_exc;
static int g_outstanding; /* Get*ArrayElements not yet released: must be 0 after every call */

static size_t elem_size(char k) { return k == 'b' ? 1 : k == 'i' ? 4 : 8; }

static jclass f_FindClass(JNIEnv* e, const char* n) { (void)e; (void)n; return (jclass)&g_exc; }
static jint f_ThrowNew(JNIEnv* e, jclass c, const char* m) {
    (void)e; (void)c;
    snprintf(g_exc, sizeof g_exc, "%s", m ? m : "");
    g_has_exc = 1;
    return 0;
}
static jsize f_GetArrayLength(JNIEnv* e, jarray a) { (void)e; return ((fake_arr*)a)->len; }
static jobject f_GetObjectArrayElement(JNIEnv* e, jobjectArray a, jsize i) { (void)e; return ((jobject*)((fake_arr*)a)->data)[i]; }
static jbyte* f_GetB(JNIEnv* e, jbyteArray a, jboolean* c) { (void)e; if (c) *c = 0; ++g_outstanding; return (jbyte*)((fake_arr*)a)->data; }
static jint* f_GetI(JNIEnv* e, jintArray a, jboolean* c) { (void)e; if (c) *c = 0; ++g_outstanding; return (jint*)((fake_arr*)a)->data; }
static jlong* f_GetL(JNIEnv* e, jlongArray a, jboolean* c) { (void)e; if (c) *c = 0; ++g_outstanding; return (jlong*)((fake_arr*)a)->data; }
static void f_RelB(JNIEnv* e, jbyteArray a, jbyte* p, jint m) { (void)e; (void)a; (void)p; (void)m; --g_outstanding; }
static void f_RelI(JNIEnv* e, jintArray a, jint* p, jint m) { (void)e; (void)a; (void)p; (void)m; --g_outstanding; }
static void f_RelL(JNIEnv* e, jlongArray a, jlong* p, jint m) { (void)e; (void)a; (void)p; (void)m; --g_outstanding; }
static void f_GetIntArrayRegion(JNIEnv* e, jintArray a, jsize s, jsize n, jint* out) {
    (void)e;
    memcpy(out, (jint*)((fake_arr*)a)->data + s, (size_t)n * 4);
}
static fake_arr* new_arr(char kind, int len) {
    fake_arr* a = (fake_arr*)calloc(1, sizeof *a);
    a->kind = kind;
    a->len = len;
    a->data = calloc((size_t)(len > 0 ? len : 1), kind == 'o' ? sizeof(void*) : elem_size(kind));
    return a;
}
static jlongArray f_NewLongArray(JNIEnv* e, jsize n) { (void)e; return (jlongArray)new_arr('l', n); }
static jintArray f_NewIntArray(JNIEnv* e, jsize n) { (void)e; return (jintArray)new_arr('i', n); }
static void f_SetLongArrayRegion(JNIEnv* e, jlongArray a, jsize s, jsize n, const jlong* v) {
    (void)e;
    memcpy((jlong*)((fake_arr*)a)->data + s, v, (size_t)n * 8);
}
static void f_SetIntArrayRegion(JNIEnv* e, jintArray a, jsize s, jsize n, const jint* v) {
    (void)e;
    memcpy((jint*)((fake_arr*)a)->data + s, v, (size_t)n * 4);
}

static const struct JNINativeInterface_ g_table = {
    f_FindClass, f_ThrowNew, f_GetArrayLength, f_GetObjectArrayElement, f_GetB, f_GetI, f_GetL, f_RelB, f_RelI, f_RelL,
    f_GetIntArrayRegion, f_NewLongArray, f_SetLongArrayRegion, f_NewIntArray, f_SetIntArrayRegion};
static JNIEnv g_env = &g_table;

/* ---- the shim's exported functions (jni/jtb_jni.c) ------------------------------------------------------------ */
jint Java_jtb_Native_deviceCount(JNIEnv*, jclass);
jlong Java_jtb_Native_create(JNIEnv*, jclass, jint, jint, jlong, jlong, jint);
void Java_jtb_Native_destroy(JNIEnv*, jclass, jlong);
jlong Java_jtb_Native_multiCreate(JNIEnv*, jclass, jint, jint, jlong, jlong, jint);
void Java_jtb_Native_multiDestroy(JNIEnv*, jclass, jlong);
jlongArray Java_jtb_Native_checkLinearizable(JNIEnv*, jclass, jlong, jboolean, jobjectArray, jint, jint, jintArray, jintArray, jboolean);
jintArray Java_jtb_Native_finalConfigs(JNIEnv*, jclass, jlong, jobjectArray, jint, jint, jintArray, jintArray, jboolean, jint, jint);
jlongArray Java_jtb_Native_checkSetFull(JNIEnv*, jclass, jlong, jboolean, jobjectArray, jboolean);
jlongArray Java_jtb_Native_checkBankTotals(JNIEnv*, jclass, jlong, jobjectArray, jintArray, jlong, jboolean);

/* ---- ctypes-facing helpers --------------------------------------------------------------------------------------- */
void* fj_new_array(char kind, int len, const void* data) {
    fake_arr* a = new_arr(kind, len);
    if (data && len > 0) memcpy(a->data, data, (size_t)len * (kind == 'o' ? sizeof(void*) : elem_size(kind)));
    return a;
}
void fj_free_array(void* a) {
    if (!a) return;
    free(((fake_arr*)a)->data);
    free(a);
}
int fj_array_len(void* a) { return a ? ((fake_arr*)a)->len : -1; }
void* fj_array_data(void* a) { return a ? ((fake_arr*)a)->data : NULL; }
const char* fj_exception(void) { return g_has_exc ? g_exc : NULL; }
void fj_clear_exception(void) { g_has_exc = 0; g_exc[0] = 0; }
int fj_outstanding(void) { return g_outstanding; }

int fj_device_count(void) { return Java_jtb_Native_deviceCount(&g_env, NULL); }
long long fj_create(int device, int flags, long long table_bytes, long long max_configs, int budget_ms) {
    return Java_jtb_Native_create(&g_env, NULL, device, flags, table_bytes, max_configs, budget_ms);
}
void fj_destroy(long long h) { Java_jtb_Native_destroy(&g_env, NULL, h); }
long long fj_multi_create(int n_gpus, int flags, long long table_bytes, long long max_configs, int budget_ms) {
    return Java_jtb_Native_multiCreate(&g_env, NULL, n_gpus, flags, table_bytes, max_configs, budget_ms);
}
void fj_multi_destroy(long long h) { Java_jtb_Native_multiDestroy(&g_env, NULL, h); }
void* fj_check_linearizable(long long h, int multi, void* hist, int kind, int init_value, void* accounts, void* init_bal,
                            int neg_ok) {
    return Java_jtb_Native_checkLinearizable(&g_env, NULL, h, (jboolean)multi, (jobjectArray)hist, kind, init_value,
                                             (jintArray)accounts, (jintArray)init_bal, (jboolean)neg_ok);
}
void* fj_final_configs(long long h, void* hist, int kind, int init_value, void* accounts, void* init_bal, int neg_ok,
                       int shard, int cap) {
    return Java_jtb_Native_finalConfigs(&g_env, NULL, h, (jobjectArray)hist, kind, init_value, (jintArray)accounts,
                                        (jintArray)init_bal, (jboolean)neg_ok, shard, cap);
}
void* fj_check_set_full(long long h, int multi, void* hist, int linearizable) {
    return Java_jtb_Native_checkSetFull(&g_env, NULL, h, (jboolean)multi, (jobjectArray)hist, (jboolean)linearizable);
}
void* fj_check_bank_totals(long long h, void* hist, void* accounts, long long total, int neg_ok) {
    return Java_jtb_Native_checkBankTotals(&g_env, NULL, h, (jobjectArray)hist, (jintArray)accounts, total, (jboolean)neg_ok);
}
