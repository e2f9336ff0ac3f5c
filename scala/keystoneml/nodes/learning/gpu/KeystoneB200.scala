package keystoneml.nodes.learning.gpu

/**
 * JNI binding of libkeystone_b200 (include/keystone_b200.h via jni/keystone_b200_jni.cpp).
 * Same convention as the reference's only native nodes (keystoneml.utils.external.VLFeat / EncEval):
 * a Serializable class whose constructor loads the library, @native methods on primitives and primitive arrays.
 * Not compiled in the build image (no JVM); kept mechanical so it can be checked against the C header by eye.
 */
class KeystoneB200 extends Serializable {
  System.loadLibrary("keystone_b200_jni") // run-pipeline.sh passes -Djava.library.path=$FWDIR/lib

  @native def ncclUniqueId(): Array[Byte]
  @native def ctxCreate(device: Int, rank: Int, world: Int, ncclId: Array[Byte]): Long
  @native def ctxDestroy(ctx: Long): Unit
  @native def matrixFromHost(ctx: Long, rowMajor: Array[Double], nRows: Long, nCols: Long): Long
  @native def matrixToHost(ctx: Long, m: Long): Array[Double]
  @native def matrixDestroy(ctx: Long, m: Long): Unit
  @native def cosineRfCreate(ctx: Long, w: Array[Double], b: Array[Double], nOut: Long, nIn: Long): Long
  /** precisionMode: 0 = tf32 operands (KS_PRECISION_TF32), 1 = fp16 operands for generated cosine features (KS_PRECISION_F16). */
  @native def blockLsFit(ctx: Long, features: Long, xIn: Long, rfs: Array[Long], labels: Long,
      blockSize: Int, numIter: Int, lambda: Double, numFeaturesOr0: Long, precisionMode: Int): Long
  @native def blockWlsFit(ctx: Long, features: Long, xIn: Long, rfs: Array[Long], labels: Long,
      blockSize: Int, numIter: Int, lambda: Double, mixtureWeight: Double, numFeaturesOr0: Long): Long
  @native def modelNumBlocks(ctx: Long, model: Long): Int
  @native def modelGetBlock(ctx: Long, model: Long, j: Int): Array[Double]
  @native def modelGetBlockMean(ctx: Long, model: Long, j: Int): Array[Double]
  @native def modelGetIntercept(ctx: Long, model: Long): Array[Double]
  @native def modelApply(ctx: Long, model: Long, features: Long, xIn: Long, rfs: Array[Long]): Long
  @native def modelDestroy(ctx: Long, model: Long): Unit
}
