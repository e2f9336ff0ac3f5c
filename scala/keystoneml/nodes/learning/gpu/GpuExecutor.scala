package keystoneml.nodes.learning.gpu

import breeze.linalg._
import org.apache.spark.TaskContext
import org.apache.spark.rdd.RDD

/**
 * One library context per executor JVM (one executor per GPU), created on first use and kept for the life of the JVM --
 * the counterpart of the `@transient lazy val` wrapper objects of the reference's native nodes
 * (keystoneml.nodes.images.external.SIFTExtractor:18).  `rank` / `world` / the NCCL id come from the job configuration
 * (spark.keystone.gpu.*), which the driver fills after calling `KeystoneB200.ncclUniqueId()` once.
 * Not compiled in the build image (no JVM).
 */
object GpuExecutor {
  @transient lazy val lib = new KeystoneB200()
  @volatile private var ctxHandle = 0L

  def ctx(device: Int, rank: Int, world: Int, ncclId: Array[Byte]): Long = synchronized {
    if (ctxHandle == 0L) ctxHandle = lib.ctxCreate(device, rank, world, if (world > 1) ncclId else null)
    ctxHandle
  }

  /** Rows of one partition as a flat row-major array (MatrixUtils.rowsToMatrix, K/utils/MatrixUtils.scala:48-93, minus the transpose). */
  def flatten(rows: Array[DenseVector[Double]]): Array[Double] = {
    if (rows.isEmpty) return Array.empty[Double]
    val d = rows(0).length
    val out = new Array[Double](rows.length * d)
    var i = 0
    while (i < rows.length) { System.arraycopy(rows(i).toArray, 0, out, i * d, d); i += 1 }
    out
  }

  /**
   * Uploads the rows this executor holds, one partition at a time, into ONE device matrix and returns its handle.
   * Runs inside a barrier stage (every executor takes part in the collective fit that follows); nothing is collected
   * on the driver.  `counts(p)` = rows of partition p (one cheap `mapPartitions(_.size)` pass by the caller).
   */
  def uploadPartitions(c: Long, parts: Iterator[(Int, Array[DenseVector[Double]])], myRows: Long, nCols: Int): Long = {
    val m = lib.matrixCreate(c, myRows, nCols)
    var row0 = 0L
    parts.foreach { case (_, rows) =>
      if (rows.nonEmpty) {
        lib.matrixWriteRows(c, m, row0, flatten(rows), rows.length, nCols)
        row0 += rows.length
      }
    }
    m
  }
}

/** Settings a job passes to its GPU nodes (one executor per GPU). */
case class GpuJob(world: Int, ncclId: Array[Byte], devicesPerHost: Int = 8) extends Serializable {
  def rankOf(tc: TaskContext): Int = tc.partitionId() % world
  def deviceOf(rank: Int): Int = rank % devicesPerHost
}
