package keystoneml.nodes.learning.gpu

import breeze.linalg._
import keystoneml.workflow.Transformer
import org.apache.spark.rdd.RDD

/**
 * Drop-in for keystoneml.nodes.stats.CosineRandomFeatures (CosineRandomFeatures.scala:19-44): same constructor (W, b), same
 * apply overloads.  apply(RDD) uploads each partition, runs the fused projection + cosine kernel and downloads the features;
 * in a pipeline that ends in GpuBlockLeastSquaresEstimator the features never need to leave the device: pass the feature-map
 * handles (`rfHandle`) with the raw input to blockLsFit (`xIn` + `rfs`) instead -- that is what the Python mirror's
 * LazyFeatures does.
 * Not compiled in the build image (no JVM).
 */
class GpuCosineRandomFeatures(val W: DenseMatrix[Double], val b: DenseVector[Double], job: GpuJob)
  extends Transformer[DenseVector[Double], DenseVector[Double]] {
  require(b.length == W.rows, "# of rows in W and size of b should match")        // CosineRandomFeatures.scala:24

  /** Feature-map handle on this executor's context (W is Breeze column-major numOut x numIn, exactly what the ABI takes). */
  def rfHandle(ctx: Long): Long = GpuExecutor.lib.cosineRfCreate(ctx, W.data, b.data, W.rows, W.cols)

  override def apply(in: RDD[DenseVector[Double]]): RDD[DenseVector[Double]] = {
    val (jb, self) = (job, this)
    in.mapPartitionsWithIndex { case (p, it) =>
      val rows = it.toArray
      if (rows.isEmpty) Iterator.empty
      else {
        val lib = GpuExecutor.lib
        val rank = p % jb.world
        val c = GpuExecutor.ctx(jb.deviceOf(rank), rank, 1, null)                 // apply is not collective
        val x = lib.matrixCreate(c, rows.length, rows(0).length)
        lib.matrixWriteRows(c, x, 0, GpuExecutor.flatten(rows), rows.length, rows(0).length)
        val rf = self.rfHandle(c)
        val f = lib.featureMapApply(c, rf, x)
        val flat = lib.matrixToHost(c, f)
        lib.matrixDestroy(c, f); lib.featureMapDestroy(c, rf); lib.matrixDestroy(c, x)
        val n = W.rows
        Iterator.tabulate(rows.length)(i => DenseVector(java.util.Arrays.copyOfRange(flat, i * n, (i + 1) * n)))
      }
    }
  }

  override def apply(in: DenseVector[Double]): DenseVector[Double] = {
    val features = W * in                                                          // single datum: stay on the JVM (:38-43)
    features :+= b
    breeze.numerics.cos.inPlace(features)
    features
  }
}
