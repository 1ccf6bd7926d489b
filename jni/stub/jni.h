/* Minimal stand-in for the JDK's <jni.h>, ONLY so that `make -C jni check` can syntax-check b2_jni.c in an image without a
 * JDK (this one: java / javac are absent).  It declares just the JNI types and JNIEnv members the shim uses, with the
 * signatures of the JNI specification.  A real build (`make -C jni JAVA_HOME=...`) never sees this file. */
#ifndef B2_JNI_STUB_H
#define B2_JNI_STUB_H
#include <stdint.h>
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef jint jsize;
typedef struct _jobject* jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jarray; typedef jarray jobjectArray;
typedef jarray jlongArray; typedef jarray jintArray; typedef jobject jthrowable;
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_TRUE 1
#define JNI_FALSE 0
struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  jlong* (*GetLongArrayElements)(JNIEnv*, jlongArray, jboolean*);
  void (*ReleaseLongArrayElements)(JNIEnv*, jlongArray, jlong*, jint);
  jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
  void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
  jlongArray (*NewLongArray)(JNIEnv*, jsize);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
  void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
  jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  void (*DeleteLocalRef)(JNIEnv*, jobject);
};
#endif
