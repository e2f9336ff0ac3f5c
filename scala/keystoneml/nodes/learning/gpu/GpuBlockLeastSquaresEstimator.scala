package keystoneml.nodes.learning.gpu

import breeze.linalg._
import keystoneml.nodes.learning.BlockLinearMapper
import keystoneml.nodes.stats.StandardScalerModel
import keystoneml.workflow.{LabelEstimator, WeightedNode}
import org.apache.spark.rdd.RDD

/**
 * Drop-in for keystoneml.nodes.learning.BlockLeastSquaresEstimator (BlockLinearMapper.scala:199-257): same constructor
 * arguments, same fit signature, same returned BlockLinearMapper(xs, blockSize, Some(labelMean), Some(featureScalers)).
 * The body replaces VectorSplitter + the StandardScaler passes + mlmatrix BlockCoordinateDescent by one ks_blockls_fit call
 * per executor: features and labels are coalesced to one partition group per GPU, every executor uploads ITS rows
 * partition by partition (no collect() on the driver) and all executors enter the collective fit together (barrier stage);
 * the all-reduce that replaces treeReduce happens inside the call.  Rank 0 returns the model arrays.
 * Not compiled in the build image (no JVM).
 */
class GpuBlockLeastSquaresEstimator(blockSize: Int, numIter: Int, lambda: Double = 0.0, numFeaturesOpt: Option[Int] = None,
    job: GpuJob, precisionMode: Int = KeystoneB200.PrecisionDefault)
  extends LabelEstimator[DenseVector[Double], DenseVector[Double], DenseVector[Double]] with WeightedNode {

  override val weight = (3 * numIter) + 1

  override def fit(trainingFeatures: RDD[DenseVector[Double]], trainingLabels: RDD[DenseVector[Double]]): BlockLinearMapper = {
    val world = job.world
    val zipped = trainingFeatures.zip(trainingLabels).coalesce(world)          // one partition per GPU, rows stay where they are
    val (bs, ni, lam, nf, prec, jb) = (blockSize, numIter, lambda, numFeaturesOpt.map(_.toLong).getOrElse(0L), precisionMode, job)
    val models = zipped.barrier().mapPartitions { it =>
      val tc = org.apache.spark.BarrierTaskContext.get()
      val rank = tc.partitionId()
      val lib = GpuExecutor.lib
      val c = GpuExecutor.ctx(jb.deviceOf(rank), rank, jb.world, jb.ncclId)
      val rows = it.toArray
      val d = if (rows.isEmpty) 0 else rows(0)._1.length
      val k = if (rows.isEmpty) 0 else rows(0)._2.length
      val f = lib.matrixCreate(c, rows.length, d)
      val y = lib.matrixCreate(c, rows.length, k)
      val chunk = 65536                                                           // rows per upload: bounded JVM staging array
      var r0 = 0
      while (r0 < rows.length) {
        val r1 = math.min(rows.length, r0 + chunk)
        lib.matrixWriteRows(c, f, r0, GpuExecutor.flatten(rows.slice(r0, r1).map(_._1)), r1 - r0, d)
        lib.matrixWriteRows(c, y, r0, GpuExecutor.flatten(rows.slice(r0, r1).map(_._2)), r1 - r0, k)
        r0 = r1
      }
      tc.barrier()
      val m = lib.blockLsFit(c, f, 0L, null, y, bs, ni, lam, nf, prec)           // collective: NCCL all-reduce of G and C inside
      val out = if (rank == 0) {
        val nb = lib.modelNumBlocks(c, m)
        Iterator.single((
          (0 until nb).map(j => lib.modelGetBlock(c, m, j)).toArray,
          (0 until nb).map(j => lib.modelGetBlockMean(c, m, j)).toArray,
          lib.modelGetIntercept(c, m), k))
      } else Iterator.empty
      lib.modelDestroy(c, m); lib.matrixDestroy(c, f); lib.matrixDestroy(c, y)
      out
    }.collect()
    val (ws, mus, b, k) = models.head
    val xs = ws.map(w => new DenseMatrix[Double](w.length / k, k, w)).toSeq       // column-major, as returned
    val scalers = mus.map(mu => new StandardScalerModel(DenseVector(mu), None)).toSeq
    new BlockLinearMapper(xs, blockSize, Some(DenseVector(b)), Some(scalers))
  }
}
