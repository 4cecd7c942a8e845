/*
 * jni.h — MOCK of the handful of JNI declarations jni/ranklib_b200_jni.c uses.  Test infrastructure only.
 *
 * There is no JDK in this image, so the shim could otherwise never be compiled, let alone run.  This header
 * declares, with the JNI specification's names and signatures, exactly the types and JNIEnv functions the shim
 * calls; mock_jvm.c implements them over plain C arrays.  The member ORDER of the function table is not the real
 * JVM's (the shim is source-compatible with the real <jni.h>, not binary-compatible with this mock) — build the
 * product shim against $JAVA_HOME/include as INTEGRATION.md says.
 */
#ifndef RLB_MOCK_JNI_H
#define RLB_MOCK_JNI_H

#include <stdarg.h>
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

typedef int32_t jint;
typedef int64_t jlong;
typedef float jfloat;
typedef double jdouble;
typedef uint8_t jboolean;
typedef int8_t jbyte;
typedef jint jsize;

struct mock_object;
typedef struct mock_object* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jfloatArray;
typedef jarray jdoubleArray;
typedef jarray jbyteArray;
struct mock_method;
typedef struct mock_method* jmethodID;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
    jclass (*FindClass)(JNIEnv* env, const char* name);
    jmethodID (*GetStaticMethodID)(JNIEnv* env, jclass cls, const char* name, const char* sig);
    jobject (*CallStaticObjectMethod)(JNIEnv* env, jclass cls, jmethodID m, ...);
    jstring (*NewStringUTF)(JNIEnv* env, const char* utf);
    jint (*Throw)(JNIEnv* env, jthrowable obj);
    jsize (*GetArrayLength)(JNIEnv* env, jarray a);
    void* (*GetPrimitiveArrayCritical)(JNIEnv* env, jarray a, jboolean* isCopy);
    void (*ReleasePrimitiveArrayCritical)(JNIEnv* env, jarray a, void* carray, jint mode);
    jint* (*GetIntArrayElements)(JNIEnv* env, jintArray a, jboolean* isCopy);
    void (*ReleaseIntArrayElements)(JNIEnv* env, jintArray a, jint* elems, jint mode);
    jfloat* (*GetFloatArrayElements)(JNIEnv* env, jfloatArray a, jboolean* isCopy);
    void (*ReleaseFloatArrayElements)(JNIEnv* env, jfloatArray a, jfloat* elems, jint mode);
    void (*SetIntArrayRegion)(JNIEnv* env, jintArray a, jsize start, jsize len, const jint* buf);
    const char* (*GetStringUTFChars)(JNIEnv* env, jstring s, jboolean* isCopy);
    void (*ReleaseStringUTFChars)(JNIEnv* env, jstring s, const char* utf);
    void (*SetByteArrayRegion)(JNIEnv* env, jbyteArray a, jsize start, jsize len, const jbyte* buf);
    void (*GetByteArrayRegion)(JNIEnv* env, jbyteArray a, jsize start, jsize len, jbyte* buf);
};

#endif
