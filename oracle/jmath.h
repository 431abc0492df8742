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

/*
 * The three functions below (log, exp, pow) are transcriptions of FDLIBM 5.3 (e_log.c, e_exp.c, e_pow.c), the
 * algorithm java.lang.StrictMath is specified to use.  FDLIBM's notice, preserved as its licence requires:
 *
 * ====================================================
 * Copyright (C) 1993, 2004 by Sun Microsystems, Inc. All rights reserved.
 *
 * Developed at SunSoft, a Sun Microsystems, Inc. business.
 * Permission to use, copy, modify, and distribute this
 * software is freely granted, provided that this notice
 * is preserved.
 * ====================================================
 */
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

static inline double with_lo_zero(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  u &= 0xffffffff00000000ull;
  std::memcpy(&x, &u, 8);
  return x;
}
static inline double from_words(int32_t hi, uint32_t lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | lo;
  double x;
  std::memcpy(&x, &u, 8);
  return x;
}

/* ---- fdlibm __ieee754_exp (StrictMath.exp) ---- */
static inline double strict_exp(double x) {
  const double one = 1.0, huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302,
               o_threshold = 7.09782712893383973096e+02, u_threshold = -7.45133219101941108420e+02,
               invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
               P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08,
               ln2HI0 = 6.93147180369123816490e-01, ln2LO0 = 1.90821492927058770002e-10;
  double y, hi = 0.0, lo = 0.0, c, t;
  int k = 0, xsb;
  unsigned hx;
  hx = (unsigned)hi_word(x);
  xsb = (int)((hx >> 31) & 1u);
  hx &= 0x7fffffffu;
  if (hx >= 0x40862E42u) {
    if (hx >= 0x7ff00000u) {
      if (((hx & 0xfffffu) | lo_word(x)) != 0) return x + x;
      return (xsb == 0) ? x : 0.0;
    }
    if (x > o_threshold) return huge * huge;
    if (x < u_threshold) return twom1000 * twom1000;
  }
  if (hx > 0x3fd62e42u) {
    if (hx < 0x3FF0A2B2u) {
      hi = x - (xsb ? -ln2HI0 : ln2HI0);
      lo = xsb ? -ln2LO0 : ln2LO0;
      k = 1 - xsb - xsb;
    } else {
      k = (int)(invln2 * x + (xsb ? -0.5 : 0.5));
      t = k;
      hi = x - t * ln2HI0;
      lo = t * ln2LO0;
    }
    x = hi - lo;
  } else if (hx < 0x3e300000u) {
    if (huge + x > one) return one + x;
  } else
    k = 0;
  t = x * x;
  c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return one - ((x * c) / (c - 2.0) - x);
  y = one - ((lo - (x * c) / (2.0 - c)) - hi);
  if (k >= -1021) return with_hi_word(y, hi_word(y) + (k << 20));
  y = with_hi_word(y, hi_word(y) + ((k + 1000) << 20));
  return y * twom1000;
}

/* ---- fdlibm __ieee754_pow (StrictMath.pow) ---- */
static inline double strict_pow(double x, double y) {
  const double zero = 0.0, one = 1.0, two = 2.0, two53 = 9007199254740992.0, huge = 1.0e300, tiny = 1.0e-300,
               L1 = 5.99999999999994648725e-01, L2 = 4.28571428578550184252e-01, L3 = 3.33333329818377432918e-01,
               L4 = 2.72728123808534006489e-01, L5 = 2.30660745775561754067e-01, L6 = 2.06975017800338417784e-01,
               P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08, lg2 = 6.93147180559945286227e-01,
               lg2_h = 6.93147182464599609375e-01, lg2_l = -1.90465429995776804525e-09,
               ovt = 8.0085662595372944372e-0017, cp = 9.61796693925975554329e-01, cp_h = 9.61796700954437255859e-01,
               cp_l = -7.02846165095275826516e-09, ivln2 = 1.44269504088896338700e+00,
               ivln2_h = 1.44269502162933349609e+00, ivln2_l = 1.92596299112661746887e-08;
  double z, ax, z_h, z_l, p_h, p_l;
  double y1, t1, t2, r, s, t, u, v, w;
  int i, j, k, yisint, n;
  int hx, hy, ix, iy;
  unsigned lx, ly;
  hx = hi_word(x);
  lx = lo_word(x);
  hy = hi_word(y);
  ly = lo_word(y);
  ix = hx & 0x7fffffff;
  iy = hy & 0x7fffffff;
  if ((iy | ly) == 0) return one;
  if (ix > 0x7ff00000 || ((ix == 0x7ff00000) && (lx != 0)) || iy > 0x7ff00000 || ((iy == 0x7ff00000) && (ly != 0)))
    return x + y;
  yisint = 0;
  if (hx < 0) {
    if (iy >= 0x43400000)
      yisint = 2;
    else if (iy >= 0x3ff00000) {
      k = (iy >> 20) - 0x3ff;
      if (k > 20) {
        j = (int)(ly >> (52 - k));
        if (((unsigned)j << (52 - k)) == ly) yisint = 2 - (j & 1);
      } else if (ly == 0) {
        j = iy >> (20 - k);
        if ((j << (20 - k)) == iy) yisint = 2 - (j & 1);
      }
    }
  }
  if (ly == 0) {
    if (iy == 0x7ff00000) {
      if (((ix - 0x3ff00000) | lx) == 0) return y - y;
      if (ix >= 0x3ff00000) return (hy >= 0) ? y : zero;
      return (hy < 0) ? -y : zero;
    }
    if (iy == 0x3ff00000) {
      if (hy < 0) return one / x;
      return x;
    }
    if (hy == 0x40000000) return x * x;
    if (hy == 0x3fe00000) {
      if (hx >= 0) return std::sqrt(x);
    }
  }
  ax = std::fabs(x);
  if (lx == 0) {
    if (ix == 0x7ff00000 || ix == 0 || ix == 0x3ff00000) {
      z = ax;
      if (hy < 0) z = one / z;
      if (hx < 0) {
        if (((ix - 0x3ff00000) | yisint) == 0) {
          z = (z - z) / (z - z);
        } else if (yisint == 1)
          z = -z;
      }
      return z;
    }
  }
  n = (hx < 0) ? 0 : 1; /* fdlibm: n = (hx>>31)+1 with an arithmetic shift */
  if ((n | yisint) == 0) return (x - x) / (x - x);
  s = one;
  if ((n | (yisint - 1)) == 0) s = -one;
  if (iy > 0x41e00000) {
    if (iy > 0x43f00000) {
      if (ix <= 0x3fefffff) return (hy < 0) ? huge * huge : tiny * tiny;
      if (ix >= 0x3ff00000) return (hy > 0) ? huge * huge : tiny * tiny;
    }
    if (ix < 0x3fefffff) return (hy < 0) ? s * huge * huge : s * tiny * tiny;
    if (ix > 0x3ff00000) return (hy > 0) ? s * huge * huge : s * tiny * tiny;
    t = ax - one;
    w = (t * t) * (0.5 - t * (0.3333333333333333333333 - t * 0.25));
    u = ivln2_h * t;
    v = t * ivln2_l - w * ivln2;
    t1 = u + v;
    t1 = with_lo_zero(t1);
    t2 = v - (t1 - u);
  } else {
    double ss, s2, s_h, s_l, t_h, t_l, bpk, dphk, dplk;
    n = 0;
    if (ix < 0x00100000) {
      ax *= two53;
      n -= 53;
      ix = hi_word(ax);
    }
    n += ((ix) >> 20) - 0x3ff;
    j = ix & 0x000fffff;
    ix = j | 0x3ff00000;
    if (j <= 0x3988E)
      k = 0;
    else if (j < 0xBB67A)
      k = 1;
    else {
      k = 0;
      n += 1;
      ix -= 0x00100000;
    }
    ax = with_hi_word(ax, ix);
    bpk = k ? 1.5 : 1.0;
    dphk = k ? 5.84962487220764160156e-01 : 0.0;
    dplk = k ? 1.35003920212974897128e-08 : 0.0;
    u = ax - bpk;
    v = one / (ax + bpk);
    ss = u * v;
    s_h = with_lo_zero(ss);
    t_h = from_words(((ix >> 1) | 0x20000000) + 0x00080000 + (k << 18), 0u);
    t_l = ax - (t_h - bpk);
    s_l = v * ((u - s_h * t_h) - s_h * t_l);
    s2 = ss * ss;
    r = s2 * s2 * (L1 + s2 * (L2 + s2 * (L3 + s2 * (L4 + s2 * (L5 + s2 * L6)))));
    r += s_l * (s_h + ss);
    s2 = s_h * s_h;
    t_h = 3.0 + s2 + r;
    t_h = with_lo_zero(t_h);
    t_l = r - ((t_h - 3.0) - s2);
    u = s_h * t_h;
    v = s_l * t_h + t_l * ss;
    p_h = u + v;
    p_h = with_lo_zero(p_h);
    p_l = v - (p_h - u);
    z_h = cp_h * p_h;
    z_l = cp_l * p_h + p_l * cp + dplk;
    t = (double)n;
    t1 = (((z_h + z_l) + dphk) + t);
    t1 = with_lo_zero(t1);
    t2 = z_l - (((t1 - t) - dphk) - z_h);
  }
  y1 = with_lo_zero(y);
  p_l = (y - y1) * t1 + y * t2;
  p_h = y1 * t1;
  z = p_l + p_h;
  j = hi_word(z);
  i = (int)lo_word(z);
  if (j >= 0x40900000) {
    if (((j - 0x40900000) | i) != 0) return s * huge * huge;
    if (p_l + ovt > z - p_h) return s * huge * huge;
  } else if ((j & 0x7fffffff) >= 0x4090cc00) {
    if (((j - (int)0xc090cc00) | i) != 0) return s * tiny * tiny;
    if (p_l <= z - p_h) return s * tiny * tiny;
  }
  i = j & 0x7fffffff;
  k = (i >> 20) - 0x3ff;
  n = 0;
  if (i > 0x3fe00000) {
    n = j + (0x00100000 >> (k + 1));
    k = ((n & 0x7fffffff) >> 20) - 0x3ff;
    t = from_words(n & ~(0x000fffff >> k), 0u);
    n = ((n & 0x000fffff) | 0x00100000) >> (20 - k);
    if (j < 0) n = -n;
    p_h -= t;
  }
  t = p_l + p_h;
  t = with_lo_zero(t);
  u = t * lg2_h;
  v = (p_l - (t - p_h)) * lg2 + t * lg2_l;
  z = u + v;
  w = v - (z - u);
  t = z * z;
  t1 = z - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  r = (z * t1) / (t1 - two) - (w + z * w);
  z = one - (r - z);
  j = hi_word(z);
  j += (n << 20);
  if ((j >> 20) <= 0)
    z = std::scalbn(z, n);
  else
    z = with_hi_word(z, hi_word(z) + (n << 20));
  return s * z;
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
/* java.lang.Math.{exp,log,pow} are specified only to 1 ulp and may delegate to StrictMath (they do when HotSpot's
 * intrinsics are off).  The oracle pins them to StrictMath = fdlibm, a bit-specified choice the JVM spec allows, so
 * that the CUDA path (which carries the same fdlibm restatement) can be compared bit for bit even where HMC
 * dynamics are chaotic (early warmup). */
static inline double jpow(double x, double y) { return strict_pow(x, y); }
static inline double jexp(double x) { return strict_exp(x); }
static inline double jlog(double x) { return strict_log(x); }
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
