package com.stripe.rainier.compute

import java.nio.{ByteBuffer, ByteOrder}
import scala.collection.mutable
import com.stripe.rainier.ir._
import com.stripe.rainier.cuda.Native

/** Drop-in for `Compiler.compileTargets` (compute/Compiler.scala:14-20): runs the UNCHANGED `Translator`, flattens
  * the resulting `ir.Expr`s into the RIR container (include/rainier_rir.h) and hands it, with the data columns, to
  * `rn_model_create`.  Lives in package `compute` because `Translator`, `Target` internals and `Column.param` are
  * private[compute]/private[rainier].
  *
  * `withGradient = false` is the intended production setting: targets are built WITHOUT `Gradient.derive`
  * (compute/Target.scala:26-27), the CUDA emitter differentiates the primal DAG itself, and a Lookup over a large
  * table costs a scatter-add instead of one one-hot column per entry (compute/Gradient.scala:148-152).
  *
  * NOT COMPILED in the rainier_b200 repository (no JVM toolchain in its build image); see INTEGRATION.md.
  */
object CudaCompiler {
  final class CudaModel(val handle: Long, val nVars: Int) {
    def close(): Unit = Native.destroy(handle)
  }

  def compileTargets(group: TargetGroup, withGradient: Boolean = false, device: Int = 0): CudaModel = {
    val translator = new Translator
    val w = new RirWriter(group.parameters.map(_.param) ++ group.columns.map(_.param))
    // same output order as TargetGroup.outputs (compute/Target.scala:50-56); the Translator instance is shared so
    // CSE spans all outputs exactly as in Compiler.compile (compute/Compiler.scala:22-30)
    var firstInput = group.parameters.size
    val targets = group.targets.map { t =>
      val cols = t.columns ++ t.gradientColumns
      val outs = (t.real :: (if (withGradient) t.gradient else Nil)).map(r => w.node(translator.toExpr(r)))
      val meta = (firstInput, cols.size, cols.headOption.map(_.values.size).getOrElse(0), outs)
      firstInput += cols.size
      meta
    }
    val rir = w.finish(group.parameters.size, targets, withGradient)
    val data = group.columns.map(_.values).toArray
    val h = Native.create(rir, data, device)
    new CudaModel(h, group.parameters.size)
  }

  /** Drop-in for `Compiler.compile(inputs, outputs): CompiledFunction` (compute/Compiler.scala:22-30) as
    * `Generator.prepare` calls it (core/Generator.scala:72-76): one Translator over all outputs, serialised as a
    * RIR_FLAG_FUNCTION container (include/rainier_rir.h) for `Native.functionCreate`. */
  def compileFunction(inputs: Seq[Param], outputs: Seq[Real]): ByteBuffer = {
    val translator = new Translator
    val w = new RirWriter(inputs)
    val outs = outputs.map(r => w.node(translator.toExpr(r)))
    w.finish(inputs.size, Seq((inputs.size, 0, 0, outs)), flags = 2)
  }

  /** ir.Expr -> flat SSA.  VarDef(sym, rhs) becomes the node computing rhs, VarRef(sym) its index, SeqIR vanishes
    * (node order is evaluation order; defs precede refs by construction, compute/Translator.scala:178-179). */
  private final class RirWriter(inputs: Seq[Param]) {
    private val inputIndex = inputs.zipWithIndex.toMap
    private val nodes = mutable.ArrayBuffer.empty[Array[Byte]]
    private val lookupRefs = mutable.ArrayBuffer.empty[Int]
    private val ofSym = mutable.HashMap.empty[Sym, Int]
    private val ofParam = mutable.HashMap.empty[Param, Int]
    private val ofConst = mutable.HashMap.empty[Long, Int]

    private def rec(kind: Int, op: Int, a: Int, b: Int, c: Int, d: Int, value: Double): Int = {
      val bb = ByteBuffer.allocate(32).order(ByteOrder.LITTLE_ENDIAN)
      bb.put(kind.toByte).put(op.toByte).putShort(0).putInt(a).putInt(b).putInt(c).putInt(d).putInt(0).putDouble(value)
      nodes += bb.array
      nodes.size - 1
    }

    def node(e: Expr): Int = e match {
      case p: Param      => ofParam.getOrElseUpdate(p, rec(0, 0, inputIndex(p), 0, 0, 0, 0.0))
      case Const(v)      => ofConst.getOrElseUpdate(java.lang.Double.doubleToRawLongBits(v), rec(1, 0, 0, 0, 0, 0, v))
      case VarRef(sym)   => ofSym(sym)
      case VarDef(sym, rhs) =>
        val id = ir(rhs)
        ofSym(sym) = id
        id
    }

    private def ir(rhs: IR): Int = rhs match {
      case BinaryIR(l, r, op) =>
        val (a, b) = (node(l), node(r))
        rec(3, op match {
          case AddOp => 0; case MultiplyOp => 1; case SubtractOp => 2; case DivideOp => 3; case PowOp => 4; case CompareOp => 5
        }, a, b, 0, 0, 0.0)
      case UnaryIR(x, op) =>
        val a = node(x)
        rec(2, op match {
          case ExpOp => 0; case LogOp => 1; case AbsOp => 2; case NoOp => 3; case SinOp => 4; case CosOp => 5
          case TanOp => 6; case AsinOp => 7; case AcosOp => 8; case AtanOp => 9
        }, a, 0, 0, 0, 0.0)
      case LookupIR(index, table, low) =>
        val idx = node(index)
        val off = lookupRefs.size
        lookupRefs ++= table.map(node)
        rec(4, 0, idx, off, table.size, low, 0.0)
      case SeqIR(first, second) =>
        node(first) // evaluate-and-discard in the reference (ir/ExprMethodGenerator.scala:57-60)
        node(second)
      case MethodRef(_) => sys.error("MethodRef only exists after packing")
    }

    def finish(nParams: Int, targets: Seq[(Int, Int, Int, Seq[Int])], withGradient: Boolean): ByteBuffer =
      finish(nParams, targets, if (withGradient) 1 else 0)

    def finish(nParams: Int, targets: Seq[(Int, Int, Int, Seq[Int])], flags: Int): ByteBuffer = {
      def pad8(n: Int) = (n + 7) & ~7
      val size = 32 + nodes.size * 32 + pad8(lookupRefs.size * 4) + targets.map(t => 24 + pad8(t._4.size * 4)).sum
      val bb = ByteBuffer.allocateDirect(size).order(ByteOrder.LITTLE_ENDIAN)
      bb.putInt(0x31524952).putInt(1).putInt(nParams).putInt(inputs.size).putInt(nodes.size)
        .putInt(targets.size).putInt(lookupRefs.size).putInt(flags)
      nodes.foreach(bb.put)
      lookupRefs.foreach(bb.putInt)
      while (bb.position() % 8 != 0) bb.put(0.toByte)
      targets.foreach { case (first, nCols, nRows, outs) =>
        bb.putLong(nRows.toLong).putInt(first).putInt(nCols).putInt(outs.size).putInt(0)
        outs.foreach(bb.putInt)
        while (bb.position() % 8 != 0) bb.put(0.toByte)
      }
      bb
    }
  }
}
