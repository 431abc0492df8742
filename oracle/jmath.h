/*
 * oracle/jmath.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).  Never linked into the product.
 *
 * Restates the JDK arithmetic the reference's hot path leans on but which is not under /root/reference:
 *
 *   java.util.Random (OpenJDK 11, pinned by .github/workflows/ci.yml:12-16): 48-bit LCG, nextDouble,
 *     polar-method nextGaussian with cached second variate.  Call sites: rainier-sampler/.../RNG.scala:20-26
 *     (scala.util.Random(seed) is a thin wrapper over java.util.Random(seed)).
 *   java.lang.StrictMath.log (fdlibm 5.3 e_log.c, as specified by the StrictMath javadoc) used inside
 *     nextGaussian; StrictMath.sqrt is IEEE correctly-rounded sqrt.
 *   java.lang.Math.pow corner cases that differ from C99 pow (javadoc of Math.pow); the remaining
 *     java.lang.Math functions are "within 1 ulp" intrinsics -> glibc libm here, parity by tolerance.
 *   JVM bytecode semantics the emitter relies on: DCMPL;I2D (ir/MethodGenerator.scala:62-65), D2I
 *     (ir/MethodGenerator.scala:130-132).
 */
#ifndef RAINIER_ORACLE_JMATH_H
#define RAINIER_ORACLE_JMATH_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace rno {

/* ---- fdlibm __ieee754_log (StrictMath.log) ---------------------------------------------------------- */
static inline int32_t hi_word(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return (int32_t)(u >> 32);
}
static inline uint32_t lo_word(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return (uint32_t)u;
}
static inline double with_hi_word(double x, int32_t hi) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  u = (u & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32);
  std::memcpy(&x, &u, 8);
  return x;
}

static inline double strict_log(double x) {
  static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                      two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                      Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                      Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                      Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
  const double zero = 0.0;
  double hfsq, f, s, z, R, w, t1, t2, dk;
  int32_t k, hx, i, j;
  uint32_t lx;
  hx = hi_word(x);
  lx = lo_word(x);
  k = 0;
  if (hx < 0x00100000) { /* x < 2**-1022 */
    if (((hx & 0x7fffffff) | lx) == 0) return -two54 / zero; /* log(+-0) = -inf */
    if (hx < 0) return (x - x) / zero;                       /* log(-#) = NaN */
    k -= 54;
    x *= two54; /* subnormal, scale up */
    hx = hi_word(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  i = (hx + 0x95f64) & 0x100000;
  x = with_hi_word(x, hx | (i ^ 0x3ff00000)); /* normalize x or x/2 */
  k += (i >> 20);
  f = x - 1.0;
  if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2**-20 */
    if (f == zero) {
      if (k == 0) return zero;
      dk = (double)k;
      return dk * ln2_hi + dk * ln2_lo;
    }
    R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    dk = (double)k;
    return dk * ln2_hi - ((R - dk * ln2_lo) - f);
  }
  s = f / (2.0 + f);
  dk = (double)k;
  z = s * s;
  i = hx - 0x6147a;
  w = z * z;
  j = 0x6b851 - hx;
  t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  i |= j;
  R = t2 + t1;
  if (i > 0) {
    hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
  } else {
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
  }
}

/* ---- java.util.Random ----------------------------------------------------------------------------------- */
struct JRandom {
  int64_t seed; /* scrambled 48-bit state */
  double next_next_gaussian;
  bool have_next_next_gaussian;

  static constexpr int64_t MULT = 0x5DEECE66DLL;
  static constexpr int64_t ADD = 0xBLL;
  static constexpr int64_t MASK = (1LL << 48) - 1;

  JRandom() : seed(0), next_next_gaussian(0.0), have_next_next_gaussian(false) {}
  explicit JRandom(int64_t s) : seed((s ^ MULT) & MASK), next_next_gaussian(0.0), have_next_next_gaussian(false) {}

  inline int32_t next(int bits) {
    seed = (int64_t)(((uint64_t)seed * (uint64_t)MULT + (uint64_t)ADD) & (uint64_t)MASK);
    return (int32_t)(seed >> (48 - bits));
  }
  inline double next_double() {
    return (double)(((int64_t)next(26) << 27) + (int64_t)next(27)) * 0x1.0p-53;
  }
  inline double next_gaussian() {
    if (have_next_next_gaussian) {
      have_next_next_gaussian = false;
      return next_next_gaussian;
    }
    double v1, v2, s;
    do {
      v1 = 2 * next_double() - 1;
      v2 = 2 * next_double() - 1;
      s = v1 * v1 + v2 * v2;
    } while (s >= 1 || s == 0);
    double multiplier = std::sqrt(-2 * strict_log(s) / s);
    next_next_gaussian = v2 * multiplier;
    have_next_next_gaussian = true;
    return v1 * multiplier;
  }
};

/* ---- java.lang.Math corner cases -------------------------------------------------------------------------- */
static inline double jpow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (std::isnan(y)) return std::numeric_limits<double>::quiet_NaN();
  if (std::isinf(y) && std::fabs(x) == 1.0) return std::numeric_limits<double>::quiet_NaN();
  return std::pow(x, y);
}
/* Math.min: NaN if either is NaN; -0.0 < +0.0 */
static inline double jmin(double a, double b) {
  if (a != a) return a;
  if (a == 0.0 && b == 0.0) return std::signbit(a) ? a : b;
  return (a <= b) ? a : b;
}
/* DCMPL ; I2D : -1 when less OR unordered */
static inline double jcompare(double a, double b) {
  if (a > b) return 1.0;
  if (a == b) return 0.0;
  return -1.0;
}
/* D2I : NaN -> 0, saturating truncation */
static inline int32_t jd2i(double v) {
  if (v != v) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (-2147483647 - 1);
  return (int32_t)v;
}

} /* namespace rno */
#endif
