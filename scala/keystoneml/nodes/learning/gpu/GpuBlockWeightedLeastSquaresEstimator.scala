package keystoneml.nodes.learning.gpu

import breeze.linalg._
import keystoneml.nodes.learning.BlockLinearMapper
import keystoneml.workflow.{LabelEstimator, WeightedNode}
import org.apache.spark.HashPartitioner
import org.apache.spark.rdd.RDD

/**
 * Drop-in for keystoneml.nodes.learning.BlockWeightedLeastSquaresEstimator (BlockWeightedLeastSquares.scala:36-84):
 * same constructor arguments and fit signatures; returns BlockLinearMapper(models, blockSize, Some(finalB)) without feature
 * scalers (:316-320).  The reference needs one class per partition (:111-131, groupByClasses :333-370); here the rows are
 * partitioned BY CLASS over the GPUs (class c -> executor c % world), the device regroups them inside a rank, and the library
 * checks that no class is split across ranks.  Collective; rank 0 returns the model.
 * Not compiled in the build image (no JVM).
 */
class GpuBlockWeightedLeastSquaresEstimator(blockSize: Int, numIter: Int, lambda: Double, mixtureWeight: Double,
    numFeaturesOpt: Option[Int] = None, job: GpuJob, precisionMode: Int = KeystoneB200.PrecisionDefault)
  extends LabelEstimator[DenseVector[Double], DenseVector[Double], DenseVector[Double]] with WeightedNode {

  override val weight = (3 * numIter) + 1

  override def fit(trainingFeatures: RDD[DenseVector[Double]], trainingLabels: RDD[DenseVector[Double]]): BlockLinearMapper = {
    val world = job.world
    val byClass = trainingFeatures.zip(trainingLabels)
      .map { case (x, y) => (argmax(y), (x, y)) }                                 // class of a row: argmax of the +-1 indicators (:133-139)
      .partitionBy(new HashPartitioner(world)).values                             // whole classes per executor
    val (bs, ni, lam, w, nf, prec, jb) =
      (blockSize, numIter, lambda, mixtureWeight, numFeaturesOpt.map(_.toLong).getOrElse(0L), precisionMode, job)
    val models = byClass.barrier().mapPartitions { it =>
      val tc = org.apache.spark.BarrierTaskContext.get()
      val rank = tc.partitionId()
      val lib = GpuExecutor.lib
      val c = GpuExecutor.ctx(jb.deviceOf(rank), rank, jb.world, jb.ncclId)
      val rows = it.toArray
      val d = rows(0)._1.length
      val k = rows(0)._2.length
      val f = lib.matrixCreate(c, rows.length, d)
      val y = lib.matrixCreate(c, rows.length, k)
      lib.matrixWriteRows(c, f, 0, GpuExecutor.flatten(rows.map(_._1)), rows.length, d)
      lib.matrixWriteRows(c, y, 0, GpuExecutor.flatten(rows.map(_._2)), rows.length, k)
      tc.barrier()
      val m = lib.blockWlsFit(c, f, 0L, null, y, bs, ni, lam, w, nf, prec)
      val out = if (rank == 0) {
        val nb = lib.modelNumBlocks(c, m)
        Iterator.single(((0 until nb).map(j => lib.modelGetBlock(c, m, j)).toArray, lib.modelGetIntercept(c, m), k))
      } else Iterator.empty
      lib.modelDestroy(c, m); lib.matrixDestroy(c, f); lib.matrixDestroy(c, y)
      out
    }.collect()
    val (ws, b, k) = models.head
    new BlockLinearMapper(ws.map(wj => new DenseMatrix[Double](wj.length / k, k, wj)).toSeq, blockSize, Some(DenseVector(b)))
  }
}
