// jni/rainier_jni.cpp -- the thin JNI shim between Rainier's Scala host code and the C ABI of librainier_cuda.so
// (include/rainier_cuda.h).  Binds the native methods of scala/com/stripe/rainier/cuda/Native.scala.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no JDK (no <jni.h>, no javac/scalac).  It is written
// against the JNI specification and guarded by __has_include so that `make -C jni` is a no-op where <jni.h> is
// missing.  Build where a JDK exists:
//   g++ -O2 -std=c++17 -fPIC -shared -I$JAVA_HOME/include -I$JAVA_HOME/include/linux jni/rainier_jni.cpp \
//       -Lrainier_b200 -lrainier_cuda -Wl,-rpath,'$ORIGIN' -o librainier_jni.so
//
// Every function is a 1:1 forward; arrays are pinned with Get/ReleasePrimitiveArrayCritical (no copies on HotSpot),
// the RIR container and the rn_config POD travel as direct ByteBuffers filled by the Scala side
// (scala/com/stripe/rainier/cuda/RIR.scala, CudaConfig.scala).  A non-zero return code becomes a RuntimeException
// carrying rn_last_error() -- mirroring the reference, where failures on this path are exceptions thrown from
// generated code (ir/MethodGenerator.scala:164-167).
#if defined(__has_include)
#if __has_include(<jni.h>)
#define RN_HAVE_JNI 1
#endif
#endif

#ifdef RN_HAVE_JNI
#include <jni.h>

#include <vector>

#include "../include/rainier_cuda.h"

namespace {
void throw_last(JNIEnv* env) {
  jclass c = env->FindClass("java/lang/RuntimeException");
  if (c) env->ThrowNew(c, rn_last_error());
}
struct Crit {  // RAII for Get/ReleasePrimitiveArrayCritical
  JNIEnv* env;
  jarray arr;
  void* p;
  jint mode;
  Crit(JNIEnv* e, jarray a, jint m = 0) : env(e), arr(a), p(a ? e->GetPrimitiveArrayCritical(a, nullptr) : nullptr), mode(m) {}
  ~Crit() {
    if (p) env->ReleasePrimitiveArrayCritical(arr, p, mode);
  }
};
}  // namespace

extern "C" {

// def create(rir: ByteBuffer, cols: Array[Array[Double]], device: Int): Long
JNIEXPORT jlong JNICALL Java_com_stripe_rainier_cuda_Native_create(JNIEnv* env, jclass, jobject rir, jobjectArray cols,
                                                                  jint device) {
  const void* rp = env->GetDirectBufferAddress(rir);
  const jlong rlen = env->GetDirectBufferCapacity(rir);
  const jsize n = cols ? env->GetArrayLength(cols) : 0;
  // columns are copied to the device inside rn_model_create, so pin them one at a time via Get<>ArrayElements
  std::vector<jdoubleArray> arrs(n);
  std::vector<jdouble*> ptrs(n);
  std::vector<int64_t> rows(n);
  for (jsize i = 0; i < n; i++) {
    arrs[i] = (jdoubleArray)env->GetObjectArrayElement(cols, i);
    rows[i] = env->GetArrayLength(arrs[i]);
    ptrs[i] = env->GetDoubleArrayElements(arrs[i], nullptr);
  }
  rn_model* m = nullptr;
  const int rc = rn_model_create(rp, (size_t)rlen, (const double* const*)ptrs.data(), rows.data(), (int)n, device, &m);
  for (jsize i = 0; i < n; i++) env->ReleaseDoubleArrayElements(arrs[i], ptrs[i], JNI_ABORT);
  if (rc != RN_OK) {
    throw_last(env);
    return 0;
  }
  return (jlong)(intptr_t)m;
}

JNIEXPORT jint JNICALL Java_com_stripe_rainier_cuda_Native_nvars(JNIEnv*, jclass, jlong h) {
  return rn_model_nvars((const rn_model*)(intptr_t)h);
}

// def densityBatch(h: Long, q: Array[Double], chains: Int, out: Array[Double]): Unit
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_densityBatch(JNIEnv* env, jclass, jlong h, jdoubleArray q,
                                                                       jint chains, jdoubleArray out) {
  int rc;
  {
    Crit cq(env, q, JNI_ABORT), co(env, out);
    rc = rn_density_batch((rn_model*)(intptr_t)h, (const double*)cq.p, chains, (double*)co.p);
  }
  if (rc != RN_OK) throw_last(env);
}

// def sample(h: Long, config: ByteBuffer, seeds: Array[Long], samples: Array[Double], mass: Array[Double],
//            stats: ByteBuffer): Unit          (config = rn_config bytes, stats = chains * sizeof(rn_chain_stats))
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_sample(JNIEnv* env, jclass, jlong h, jobject config,
                                                                 jlongArray seeds, jdoubleArray samples, jdoubleArray mass,
                                                                 jobject stats) {
  const rn_config* cfg = (const rn_config*)env->GetDirectBufferAddress(config);
  rn_chain_stats* st = stats ? (rn_chain_stats*)env->GetDirectBufferAddress(stats) : nullptr;
  const jint chains = env->GetArrayLength(seeds);
  int rc;
  {
    Crit cs(env, seeds, JNI_ABORT), co(env, samples), cm(env, mass);
    rc = rn_sample((rn_model*)(intptr_t)h, cfg, (const int64_t*)cs.p, chains, (double*)co.p, (double*)cm.p, st);
  }
  if (rc != RN_OK) throw_last(env);
}

// def hostAlloc(device: Int, bytes: Long): ByteBuffer   -- page-locked memory as a direct buffer (rn_host_alloc)
JNIEXPORT jobject JNICALL Java_com_stripe_rainier_cuda_Native_hostAlloc(JNIEnv* env, jclass, jint device, jlong bytes) {
  void* p = nullptr;
  if (rn_host_alloc(device, (size_t)bytes, &p) != RN_OK) {
    throw_last(env);
    return nullptr;
  }
  return env->NewDirectByteBuffer(p, bytes);
}
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_hostFree(JNIEnv* env, jclass, jint device, jobject buf) {
  if (rn_host_free(device, env->GetDirectBufferAddress(buf)) != RN_OK) throw_last(env);
}

// def sampleDirect(h: Long, config: ByteBuffer, seeds: Array[Long], samples: ByteBuffer, mass: Array[Double],
//                  stats: ByteBuffer): Unit    -- samples is a direct buffer (ideally from hostAlloc: the device->host
// copy of chains*iterations*n doubles then is one DMA into the buffer the JVM reads, with no staging copy)
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_sampleDirect(JNIEnv* env, jclass, jlong h, jobject config,
                                                                       jlongArray seeds, jobject samples, jdoubleArray mass,
                                                                       jobject stats) {
  const rn_config* cfg = (const rn_config*)env->GetDirectBufferAddress(config);
  rn_chain_stats* st = stats ? (rn_chain_stats*)env->GetDirectBufferAddress(stats) : nullptr;
  double* out = (double*)env->GetDirectBufferAddress(samples);
  const jint chains = env->GetArrayLength(seeds);
  int rc;
  {
    Crit cs(env, seeds, JNI_ABORT), cm(env, mass);
    rc = rn_sample((rn_model*)(intptr_t)h, cfg, (const int64_t*)cs.p, chains, out, (double*)cm.p, st);
  }
  if (rc != RN_OK) throw_last(env);
}

JNIEXPORT jstring JNICALL Java_com_stripe_rainier_cuda_Native_emitSource(JNIEnv* env, jclass, jlong h, jobject config) {
  const rn_config* cfg = config ? (const rn_config*)env->GetDirectBufferAddress(config) : nullptr;
  size_t need = 0;
  if (rn_emit_source((rn_model*)(intptr_t)h, cfg, nullptr, 0, &need) != RN_OK) {
    throw_last(env);
    return nullptr;
  }
  std::vector<char> buf(need);
  rn_emit_source((rn_model*)(intptr_t)h, cfg, buf.data(), buf.size(), &need);
  return env->NewStringUTF(buf.data());
}

JNIEXPORT jint JNICALL Java_com_stripe_rainier_cuda_Native_configSize(JNIEnv*, jclass) { return (jint)sizeof(rn_config); }
JNIEXPORT jint JNICALL Java_com_stripe_rainier_cuda_Native_statsSize(JNIEnv*, jclass) { return (jint)sizeof(rn_chain_stats); }
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_defaultConfig(JNIEnv* env, jclass, jobject config) {
  rn_config_default((rn_config*)env->GetDirectBufferAddress(config));
}
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_destroy(JNIEnv*, jclass, jlong h) {
  rn_model_destroy((rn_model*)(intptr_t)h);
}
// ---- compiled functions: Generator.prepare's Compiler.compile + CompiledFunction.output loop (core/Generator.scala:59-94) ----
// def functionCreate(rir: ByteBuffer, device: Int): Long
JNIEXPORT jlong JNICALL Java_com_stripe_rainier_cuda_Native_functionCreate(JNIEnv* env, jclass, jobject rir, jint device) {
  rn_function* f = nullptr;
  if (rn_function_create(env->GetDirectBufferAddress(rir), (size_t)env->GetDirectBufferCapacity(rir), device, RN_MATH_PARITY, &f) != RN_OK) {
    throw_last(env);
    return 0;
  }
  return (jlong)(intptr_t)f;
}
// def functionEval(handle: Long, draws: Array[Double], count: Long, out: Array[Double]): Unit   draws [count][n] -> out [count][m]
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_functionEval(JNIEnv* env, jclass, jlong h, jdoubleArray draws, jlong count,
                                                                        jdoubleArray out) {
  int rc;
  {
    Crit cx(env, draws, JNI_ABORT), co(env, out);
    rc = rn_function_eval((rn_function*)(intptr_t)h, (const double*)cx.p, (int64_t)count, (double*)co.p);
  }
  if (rc != RN_OK) throw_last(env);
}
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_functionDestroy(JNIEnv*, jclass, jlong h) {
  rn_function_destroy((rn_function*)(intptr_t)h);
}
// ---- Optimizer.lbfgs for a batch of starts (optimizer/Optimizer.scala:6-24) ----
// def optimize(handle: Long, x0: Array[Double] /* [starts][n] or null */, starts: Int, m: Int, eps: Double, maxEvals: Int,
//              x: Array[Double] /* [starts][n] */, info: Array[Int]): Unit
JNIEXPORT void JNICALL Java_com_stripe_rainier_cuda_Native_optimize(JNIEnv* env, jclass, jlong h, jdoubleArray x0, jint starts, jint m,
                                                                    jdouble eps, jint maxEvals, jdoubleArray x, jintArray info) {
  rn_optimize_config oc;
  rn_optimize_config_default(&oc);
  oc.history = m;
  oc.eps = eps;
  oc.max_evaluations = maxEvals;
  int rc;
  {
    Crit c0(env, x0, JNI_ABORT), cx(env, x), ci(env, info);
    rc = rn_optimize((rn_model*)(intptr_t)h, &oc, (const double*)c0.p, starts, (double*)cx.p, nullptr, (int32_t*)ci.p, nullptr);
  }
  if (rc != RN_OK) throw_last(env);
}
JNIEXPORT jstring JNICALL Java_com_stripe_rainier_cuda_Native_lastError(JNIEnv* env, jclass) {
  return env->NewStringUTF(rn_last_error());
}

}  // extern "C"
#endif  // RN_HAVE_JNI
