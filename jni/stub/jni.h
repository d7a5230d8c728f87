/*
 * jni/stub/jni.h — a MINIMAL stand-in for the JDK's <jni.h>, for images without a JDK (this one).
 *
 * It declares exactly the JNI types, macros and JNIEnv functions that jni/jtb_jni.c uses, with the JDK's names and
 * signatures, so that (1) CI can compile the shim (`gcc -Ijni/stub -Iinclude -c jni/jtb_jni.c`) and (2) the test
 * harness tests/native/fake_jvm.c can implement this function table over plain C arrays and drive the shim end to
 * end against libjtb_check.so.  The ORDER of the function table is NOT the JDK's: never mix objects compiled against
 * this header with a real JVM — on a host with a JDK build with -I$JAVA_HOME/include instead (jni/Makefile).
 */
#ifndef JTB_STUB_JNI_H
#define JTB_STUB_JNI_H
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_FALSE 0
#define JNI_TRUE 1

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jobjectArray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jlongArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv*, const char*);
    jint (*ThrowNew)(JNIEnv*, jclass, const char*);
    jsize (*GetArrayLength)(JNIEnv*, jarray);
    jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
    jbyte* (*GetByteArrayElements)(JNIEnv*, jbyteArray, jboolean*);
    jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
    jlong* (*GetLongArrayElements)(JNIEnv*, jlongArray, jboolean*);
    void (*ReleaseByteArrayElements)(JNIEnv*, jbyteArray, jbyte*, jint);
    void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
    void (*ReleaseLongArrayElements)(JNIEnv*, jlongArray, jlong*, jint);
    void (*GetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, jint*);
    jlongArray (*NewLongArray)(JNIEnv*, jsize);
    void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
    jintArray (*NewIntArray)(JNIEnv*, jsize);
    void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
};
#endif
