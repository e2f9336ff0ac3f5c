// JNI shim: 1:1 wrappers from keystoneml.nodes.learning.gpu.KeystoneB200 (scala/.../KeystoneB200.scala) to the C ABI
// (include/keystone_b200.h).  Follows the reference's own native convention -- a Serializable Scala class whose
// constructor calls System.loadLibrary, @native methods taking only primitives / primitive arrays
// (/root/reference/src/main/scala/keystoneml/utils/external/VLFeat.scala:18-26, src/main/cpp/VLFeat.cxx:203-292) -- but
// errors become RuntimeExceptions instead of exit(-1) (src/main/cpp/EncEval.cxx:43-47).
//
// NOT COMPILED IN THIS IMAGE: there is no JDK (no jni.h).  Build where one exists:
//   g++ -O2 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude jni/keystone_b200_jni.cpp \
//       -Lkeystone_b200/lib -lkeystone_b200 -o lib/libkeystone_b200_jni.so
#include <jni.h>

#include <vector>

#include "keystone_b200.h"

#define JFN(ret, name) extern "C" JNIEXPORT ret JNICALL Java_keystoneml_nodes_learning_gpu_KeystoneB200_##name

static void check(JNIEnv* env, jlong ctx, int32_t rc) {
  if (rc == KS_OK) return;
  jclass ex = env->FindClass("java/lang/RuntimeException");
  env->ThrowNew(ex, ks_last_error(ctx));
}

JFN(jbyteArray, ncclUniqueId)(JNIEnv* env, jobject) {
  uint8_t id[KS_NCCL_ID_BYTES];
  check(env, 0, ks_nccl_unique_id(id));
  jbyteArray out = env->NewByteArray(KS_NCCL_ID_BYTES);
  env->SetByteArrayRegion(out, 0, KS_NCCL_ID_BYTES, reinterpret_cast<const jbyte*>(id));
  return out;
}

JFN(jlong, ctxCreate)(JNIEnv* env, jobject, jint device, jint rank, jint world, jbyteArray ncclId) {
  int64_t h = 0;
  jbyte* id = ncclId ? env->GetByteArrayElements(ncclId, nullptr) : nullptr;
  int32_t rc = ks_ctx_create(device, rank, world, reinterpret_cast<const uint8_t*>(id), &h);
  if (id) env->ReleaseByteArrayElements(ncclId, id, JNI_ABORT);
  check(env, 0, rc);
  return h;
}

JFN(void, ctxDestroy)(JNIEnv*, jobject, jlong ctx) { ks_ctx_destroy(ctx); }

// rows are passed as one flat row-major double array (MatrixUtils.rowsToMatrix without the column-major transpose)
JFN(jlong, matrixFromHost)(JNIEnv* env, jobject, jlong ctx, jdoubleArray rowMajor, jlong nRows, jlong nCols) {
  int64_t h = 0;
  void* p = env->GetPrimitiveArrayCritical(rowMajor, nullptr);  // no copy; the call only borrows the buffer
  int32_t rc = ks_matrix_from_host_f64(ctx, static_cast<const double*>(p), nRows, nCols, nCols, &h);
  env->ReleasePrimitiveArrayCritical(rowMajor, p, JNI_ABORT);
  check(env, ctx, rc);
  return h;
}

JFN(jdoubleArray, matrixToHost)(JNIEnv* env, jobject, jlong ctx, jlong m) {
  int64_t r = 0, c = 0;
  check(env, ctx, ks_matrix_shape(ctx, m, &r, &c));
  jdoubleArray out = env->NewDoubleArray(static_cast<jsize>(r * c));
  void* p = env->GetPrimitiveArrayCritical(out, nullptr);
  int32_t rc = ks_matrix_to_host_f64(ctx, m, static_cast<double*>(p), c);
  env->ReleasePrimitiveArrayCritical(out, p, 0);
  check(env, ctx, rc);
  return out;
}

JFN(void, matrixDestroy)(JNIEnv*, jobject, jlong ctx, jlong m) { ks_matrix_destroy(ctx, m); }

// W = DenseMatrix.data (column-major numOut x numIn), b = DenseVector.data
JFN(jlong, cosineRfCreate)(JNIEnv* env, jobject, jlong ctx, jdoubleArray W, jdoubleArray b, jlong nOut, jlong nIn) {
  int64_t h = 0;
  jdouble* w = env->GetDoubleArrayElements(W, nullptr);
  jdouble* bb = env->GetDoubleArrayElements(b, nullptr);
  int32_t rc = ks_cosine_rf_create(ctx, w, bb, nOut, nIn, &h);
  env->ReleaseDoubleArrayElements(W, w, JNI_ABORT);
  env->ReleaseDoubleArrayElements(b, bb, JNI_ABORT);
  check(env, ctx, rc);
  return h;
}

JFN(jlong, blockLsFit)(JNIEnv* env, jobject, jlong ctx, jlong features, jlong xIn, jlongArray rfs, jlong labels, jint blockSize,
                       jint numIter, jdouble lambda, jlong numFeaturesOr0, jint precisionMode) {
  int64_t h = 0;
  jsize n = rfs ? env->GetArrayLength(rfs) : 0;
  jlong* r = n ? env->GetLongArrayElements(rfs, nullptr) : nullptr;
  int32_t rc = ks_blockls_fit(ctx, features, xIn, reinterpret_cast<const int64_t*>(r), n, labels, blockSize, numIter, lambda,
                              numFeaturesOr0, precisionMode, &h);
  if (r) env->ReleaseLongArrayElements(rfs, r, JNI_ABORT);
  check(env, ctx, rc);
  return h;
}

JFN(jlong, blockWlsFit)(JNIEnv* env, jobject, jlong ctx, jlong features, jlong xIn, jlongArray rfs, jlong labels, jint blockSize,
                        jint numIter, jdouble lambda, jdouble mixtureWeight, jlong numFeaturesOr0) {
  int64_t h = 0;
  jsize n = rfs ? env->GetArrayLength(rfs) : 0;
  jlong* r = n ? env->GetLongArrayElements(rfs, nullptr) : nullptr;
  int32_t rc = ks_blockwls_fit(ctx, features, xIn, reinterpret_cast<const int64_t*>(r), n, labels, blockSize, numIter, lambda,
                               mixtureWeight, numFeaturesOr0, KS_PRECISION_TF32, &h);
  if (r) env->ReleaseLongArrayElements(rfs, r, JNI_ABORT);
  check(env, ctx, rc);
  return h;
}

JFN(jint, modelNumBlocks)(JNIEnv* env, jobject, jlong ctx, jlong model) {
  int32_t nb = 0, bs = 0;
  int64_t k = 0;
  check(env, ctx, ks_model_num_blocks(ctx, model, &nb, &k, &bs));
  return nb;
}

// returns W_j as DenseMatrix.data (column-major rows_j x k)
JFN(jdoubleArray, modelGetBlock)(JNIEnv* env, jobject, jlong ctx, jlong model, jint j) {
  int32_t nb = 0, bs = 0, has = 0;
  int64_t k = 0, rows = 0;
  check(env, ctx, ks_model_num_blocks(ctx, model, &nb, &k, &bs));
  check(env, ctx, ks_model_block_rows(ctx, model, j, &rows));
  jdoubleArray out = env->NewDoubleArray(static_cast<jsize>(rows * k));
  std::vector<double> w(static_cast<size_t>(rows * k));
  check(env, ctx, ks_model_get_block(ctx, model, j, w.data(), nullptr, &has));
  env->SetDoubleArrayRegion(out, 0, static_cast<jsize>(rows * k), w.data());
  return out;
}

JFN(jdoubleArray, modelGetBlockMean)(JNIEnv* env, jobject, jlong ctx, jlong model, jint j) {
  int32_t has = 0;
  int64_t rows = 0;
  check(env, ctx, ks_model_block_rows(ctx, model, j, &rows));
  std::vector<double> mu(static_cast<size_t>(rows));
  check(env, ctx, ks_model_get_block(ctx, model, j, nullptr, mu.data(), &has));
  if (!has) return nullptr;
  jdoubleArray out = env->NewDoubleArray(static_cast<jsize>(rows));
  env->SetDoubleArrayRegion(out, 0, static_cast<jsize>(rows), mu.data());
  return out;
}

JFN(jdoubleArray, modelGetIntercept)(JNIEnv* env, jobject, jlong ctx, jlong model) {
  int32_t nb = 0, bs = 0, has = 0;
  int64_t k = 0;
  check(env, ctx, ks_model_num_blocks(ctx, model, &nb, &k, &bs));
  std::vector<double> b(static_cast<size_t>(k));
  check(env, ctx, ks_model_get_intercept(ctx, model, b.data(), &has));
  if (!has) return nullptr;
  jdoubleArray out = env->NewDoubleArray(static_cast<jsize>(k));
  env->SetDoubleArrayRegion(out, 0, static_cast<jsize>(k), b.data());
  return out;
}

JFN(jlong, modelApply)(JNIEnv* env, jobject, jlong ctx, jlong model, jlong features, jlong xIn, jlongArray rfs) {
  int64_t h = 0;
  jsize n = rfs ? env->GetArrayLength(rfs) : 0;
  jlong* r = n ? env->GetLongArrayElements(rfs, nullptr) : nullptr;
  int32_t rc = ks_model_apply(ctx, model, features, xIn, reinterpret_cast<const int64_t*>(r), n, &h);
  if (r) env->ReleaseLongArrayElements(rfs, r, JNI_ABORT);
  check(env, ctx, rc);
  return h;
}

JFN(void, modelDestroy)(JNIEnv*, jobject, jlong ctx, jlong model) { ks_model_destroy(ctx, model); }
