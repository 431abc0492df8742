package com.stripe.rainier.cuda

import java.nio.ByteBuffer

/** JNI surface of librainier_jni.so (jni/rainier_jni.cpp), a 1:1 forward to the C ABI in include/rainier_cuda.h.
  * NOT COMPILED in the rainier_b200 repository (no JVM toolchain in its build image); see INTEGRATION.md. */
object Native {
  System.loadLibrary("rainier_jni") // which links librainier_cuda.so

  @native def create(rir: ByteBuffer, cols: Array[Array[Double]], device: Int): Long
  @native def nvars(handle: Long): Int
  @native def densityBatch(handle: Long, q: Array[Double], chains: Int, out: Array[Double]): Unit
  @native def sample(handle: Long,
                     config: ByteBuffer,
                     seeds: Array[Long],
                     samples: Array[Double],
                     mass: Array[Double],
                     stats: ByteBuffer): Unit
  /** page-locked host memory (rn_host_alloc) as a direct buffer; rn_sample DMAs straight into it */
  @native def hostAlloc(device: Int, bytes: Long): ByteBuffer
  @native def hostFree(device: Int, buf: ByteBuffer): Unit
  @native def sampleDirect(handle: Long,
                           config: ByteBuffer,
                           seeds: Array[Long],
                           samples: ByteBuffer,
                           mass: Array[Double],
                           stats: ByteBuffer): Unit
  @native def emitSource(handle: Long, config: ByteBuffer): String
  @native def configSize(): Int
  @native def statsSize(): Int
  @native def defaultConfig(config: ByteBuffer): Unit
  @native def destroy(handle: Long): Unit
  /** rn_function_*: Compiler.compile(inputs, outputs) + the CompiledFunction.output loop of Generator.prepare, batched */
  @native def functionCreate(rir: ByteBuffer, device: Int): Long
  @native def functionEval(handle: Long, draws: Array[Double], count: Long, out: Array[Double]): Unit
  @native def functionDestroy(handle: Long): Unit
  /** rn_optimize: Optimizer.lbfgs for a batch of starts; x0 == null: every start at 0 (the reference's start) */
  @native def optimize(handle: Long, x0: Array[Double], starts: Int, m: Int, eps: Double, maxEvals: Int,
                       x: Array[Double], info: Array[Int]): Unit
  @native def lastError(): String
}
