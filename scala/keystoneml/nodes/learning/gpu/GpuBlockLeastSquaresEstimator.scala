package keystoneml.nodes.learning.gpu

import breeze.linalg._
import keystoneml.nodes.learning.BlockLinearMapper
import keystoneml.nodes.stats.StandardScalerModel
import keystoneml.workflow.{LabelEstimator, WeightedNode}
import org.apache.spark.rdd.RDD

/**
 * Drop-in for keystoneml.nodes.learning.BlockLeastSquaresEstimator (BlockLinearMapper.scala:199-257): same constructor
 * arguments, same fit signature, same returned BlockLinearMapper(xs, blockSize, Some(labelMean), Some(featureScalers)).
 * The body replaces VectorSplitter + StandardScaler passes + mlmatrix BlockCoordinateDescent by one ks_blockls_fit call.
 *
 * Deployment sketch (one Spark executor per GPU, `local[*]`-style single node in the target setup): every executor
 * uploads the rows of its partitions once (`matrixFromHost`) and all executors enter `blockLsFit` together -- the
 * all-reduce that replaces treeReduce happens inside the call.  With a single GPU the driver can do it directly,
 * which is what this reference implementation shows.
 * Not compiled in the build image (no JVM).
 */
class GpuBlockLeastSquaresEstimator(blockSize: Int, numIter: Int, lambda: Double = 0.0, numFeaturesOpt: Option[Int] = None,
    precisionMode: Int = 0)
  extends LabelEstimator[DenseVector[Double], DenseVector[Double], DenseVector[Double]] with WeightedNode {

  override val weight = (3 * numIter) + 1
  @transient private lazy val lib = new KeystoneB200()

  private def flatten(rows: Array[DenseVector[Double]]): Array[Double] = {
    val d = rows(0).length
    val out = new Array[Double](rows.length * d)
    var i = 0
    while (i < rows.length) { System.arraycopy(rows(i).toArray, 0, out, i * d, d); i += 1 }
    out
  }

  override def fit(trainingFeatures: RDD[DenseVector[Double]], trainingLabels: RDD[DenseVector[Double]]): BlockLinearMapper = {
    val feats = trainingFeatures.collect()
    val labels = trainingLabels.collect()
    val ctx = lib.ctxCreate(0, 0, 1, null)
    try {
      val f = lib.matrixFromHost(ctx, flatten(feats), feats.length, feats(0).length)
      val y = lib.matrixFromHost(ctx, flatten(labels), labels.length, labels(0).length)
      val m = lib.blockLsFit(ctx, f, 0L, null, y, blockSize, numIter, lambda, numFeaturesOpt.map(_.toLong).getOrElse(0L),
        precisionMode)
      val nb = lib.modelNumBlocks(ctx, m)
      val k = labels(0).length
      val xs = (0 until nb).map { j =>
        val w = lib.modelGetBlock(ctx, m, j)
        new DenseMatrix[Double](w.length / k, k, w) // column-major, as returned
      }
      val scalers = (0 until nb).map(j => new StandardScalerModel(DenseVector(lib.modelGetBlockMean(ctx, m, j)), None))
      val b = DenseVector(lib.modelGetIntercept(ctx, m))
      lib.modelDestroy(ctx, m); lib.matrixDestroy(ctx, f); lib.matrixDestroy(ctx, y)
      new BlockLinearMapper(xs, blockSize, Some(b), Some(scalers))
    } finally {
      lib.ctxDestroy(ctx)
    }
  }
}
