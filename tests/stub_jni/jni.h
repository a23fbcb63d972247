/* Minimal stand-in for <jni.h>: just enough of the JNI 1.8 C++ surface that jni/bs_jni.cpp uses, so the shim can be
 * SYNTAX-CHECKED in an image without a JDK (tests/test_jni_shim.py: g++ -fsyntax-only).  Never linked. */
#ifndef STUB_JNI_H
#define STUB_JNI_H
#include <cstdarg>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_VERSION_1_8 0x00010008
typedef int jint; typedef long long jlong; typedef signed char jbyte; typedef unsigned char jboolean;
typedef unsigned short jchar; typedef short jshort; typedef float jfloat; typedef double jdouble; typedef jint jsize;
class _jobject {}; class _jclass : public _jobject {}; class _jstring : public _jobject {}; class _jarray : public _jobject {};
class _jobjectArray : public _jarray {}; class _jlongArray : public _jarray {}; class _jintArray : public _jarray {};
class _jdoubleArray : public _jarray {}; class _jfloatArray : public _jarray {}; class _jshortArray : public _jarray {};
class _jbyteArray : public _jarray {}; class _jthrowable : public _jobject {};
typedef _jobject* jobject; typedef _jclass* jclass; typedef _jstring* jstring; typedef _jarray* jarray;
typedef _jobjectArray* jobjectArray; typedef _jlongArray* jlongArray; typedef _jintArray* jintArray;
typedef _jdoubleArray* jdoubleArray; typedef _jfloatArray* jfloatArray; typedef _jshortArray* jshortArray;
typedef _jbyteArray* jbyteArray; typedef _jthrowable* jthrowable;
struct JNIEnv {
    jclass FindClass(const char*);
    jint ThrowNew(jclass, const char*);
    jstring NewStringUTF(const char*);
    jsize GetArrayLength(jarray);
    jobject GetObjectArrayElement(jobjectArray, jsize);
    void* GetPrimitiveArrayCritical(jarray, jboolean*);
    void ReleasePrimitiveArrayCritical(jarray, void*, jint);
    void* GetDirectBufferAddress(jobject);
    jlong GetDirectBufferCapacity(jobject);
    jobject NewDirectByteBuffer(void*, jlong);
    jdoubleArray NewDoubleArray(jsize);
    jlongArray NewLongArray(jsize);
    jbyteArray NewByteArray(jsize);
    void SetByteArrayRegion(jbyteArray, jsize, jsize, const jbyte*);
    void GetByteArrayRegion(jbyteArray, jsize, jsize, jbyte*);
    void SetDoubleArrayRegion(jdoubleArray, jsize, jsize, const jdouble*);
    void SetLongArrayRegion(jlongArray, jsize, jsize, const jlong*);
    void GetLongArrayRegion(jlongArray, jsize, jsize, jlong*);
    void GetIntArrayRegion(jintArray, jsize, jsize, jint*);
    void GetDoubleArrayRegion(jdoubleArray, jsize, jsize, jdouble*);
    void GetFloatArrayRegion(jfloatArray, jsize, jsize, jfloat*);
    const char* GetStringUTFChars(jstring, jboolean*);
    void ReleaseStringUTFChars(jstring, const char*);
    jboolean IsInstanceOf(jobject, jclass);
    jboolean ExceptionCheck();
};
#endif
