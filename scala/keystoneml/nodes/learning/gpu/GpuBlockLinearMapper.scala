package keystoneml.nodes.learning.gpu

import breeze.linalg._
import keystoneml.nodes.learning.BlockLinearMapper
import keystoneml.workflow.Transformer
import org.apache.spark.rdd.RDD

/**
 * BlockLinearMapper.apply / LinearMapper.apply on the GPU (BlockLinearMapper.scala:40-87, LinearMapper.scala:30-62): wraps a
 * fitted reference-side mapper (xs, blockSize, bOpt, featureScalersOpt), rebuilds the model on each executor's context once
 * (ks_model_from_host) and maps partitions through the fused apply GEMM; `applyArgmax` adds MaxClassifier on the device.
 * LinearMapper is the single-block case (xs = Seq(x), blockSize = x.rows).
 * Not compiled in the build image (no JVM).
 */
class GpuBlockLinearMapper(mapper: BlockLinearMapper, job: GpuJob) extends Transformer[DenseVector[Double], DenseVector[Double]] {
  private val xsData = mapper.xs.map(_.data).toArray                               // DenseMatrix.data: column-major rows_j x k
  private val k = mapper.xs.head.cols
  private val b = mapper.bOpt.map(_.data).orNull
  private val means = mapper.featureScalersOpt.map(_.map(_.mean.data).toArray).orNull
  private val blockSize = mapper.blockSize

  private def model(ctx: Long): Long = GpuExecutor.lib.modelFromHost(ctx, xsData, blockSize, k, b, means)

  private def withBatch[T](p: Int, rows: Array[DenseVector[Double]])(body: (KeystoneB200, Long, Long, Long) => T): T = {
    val lib = GpuExecutor.lib
    val rank = p % job.world
    val c = GpuExecutor.ctx(job.deviceOf(rank), rank, 1, null)
    val x = lib.matrixCreate(c, rows.length, rows(0).length)
    lib.matrixWriteRows(c, x, 0, GpuExecutor.flatten(rows), rows.length, rows(0).length)
    val m = model(c)
    try body(lib, c, m, x) finally { lib.modelDestroy(c, m); lib.matrixDestroy(c, x) }
  }

  override def apply(in: RDD[DenseVector[Double]]): RDD[DenseVector[Double]] = {
    val kk = k
    in.mapPartitionsWithIndex { case (p, it) =>
      val rows = it.toArray
      if (rows.isEmpty) Iterator.empty
      else withBatch(p, rows) { (lib, c, m, x) =>
        val y = lib.modelApply(c, m, x, 0L, null)
        val flat = lib.matrixToHost(c, y)
        lib.matrixDestroy(c, y)
        Iterator.tabulate(rows.length)(i => DenseVector(java.util.Arrays.copyOfRange(flat, i * kk, (i + 1) * kk)))
      }
    }
  }

  /** apply andThen MaxClassifier (K/nodes/util/MaxClassifier.scala:9-11), fused on the device. */
  def applyArgmax(in: RDD[DenseVector[Double]]): RDD[Int] = in.mapPartitionsWithIndex { case (p, it) =>
    val rows = it.toArray
    if (rows.isEmpty) Iterator.empty
    else withBatch(p, rows) { (lib, c, m, x) => lib.modelApplyArgmax(c, m, x, 0L, null, rows.length).iterator }
  }

  override def apply(in: DenseVector[Double]): DenseVector[Double] = mapper.apply(in)  // single datum: the JVM path
}
