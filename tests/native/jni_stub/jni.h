/* Minimal stand-in for <jni.h> — TEST INFRASTRUCTURE. Just enough of the JNI types for java/b200c_jni.c to compile without a JDK, so that
 * tests/test_java_shim.py can check the binding's C signatures against include/b200c.h. Never used to build a real JNI library. */
#ifndef JNI_STUB_H
#define JNI_STUB_H
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef int32_t jsize; typedef void* jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jintArray;
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
    jstring (*NewStringUTF)(JNIEnv*, const char*);
    void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
    jintArray (*NewIntArray)(JNIEnv*, jsize);
    void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
};
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#endif
