"""
oracle/rainier_py/optimizer.py -- TEST INFRASTRUCTURE ONLY (oracle of SURVEY.md 8f-4).

Python restatement (IEEE doubles, no FMA -- like the JVM) of the reference's MAP optimizer:

  Optimizer.lbfgs(df)     rainier-sampler/src/main/scala/com/stripe/rainier/optimizer/Optimizer.scala:6-24
  class LBFGS             rainier-sampler/src/main/scala/com/stripe/rainier/optimizer/LBFGS.java
      constructor / work-array layout                 :42-60
      apply (reverse communication, two-loop update)  :62-190
      mcsrch (More-Thuente line search)               :240-383
      mcstep (safeguarded cubic/quadratic step)       :431-605
      ddot / daxpy (sequential order)                 :612-700
  Model.optimize          rainier-core/.../core/Model.scala:26-30

PARITY UNPINNED: the reference's own test (rainier-test/.../optimizer/OptimizerTest.scala:8-59) only compares LBFGS
against its Fortran-style ancestor on live values and holds no golden vector; docs/likelihoods.md:65-77 shows the call
without its output.  This restatement is checked by properties instead (tests/test_optimizer_host.py): the termination
test ||g|| <= eps*max(1,||x||) holds at the returned point, the Wolfe conditions hold at every accepted step, closed-form
optima are reached.
"""
import math

NaN = float("nan")


def jmin(a, b):
    """java.lang.Math.min(double,double): NaN if either is NaN, -0.0 < 0.0"""
    if a != a:
        return a
    if b != b:
        return b
    if a == 0.0 and b == 0.0:
        return a if math.copysign(1.0, a) < 0 else b
    return a if a <= b else b


def jmax(a, b):
    if a != a:
        return a
    if b != b:
        return b
    if a == 0.0 and b == 0.0:
        return b if math.copysign(1.0, a) < 0 else a
    return a if a >= b else b


def jsqrt(x):
    if x != x or x < 0:
        return NaN
    return math.sqrt(x)


def jdiv(a, b):
    """IEEE division (Python raises on x/0)"""
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return NaN
        return math.copysign(math.inf, a) * math.copysign(1.0, b)


def ddot(n, dx, ix0, dy, iy0):  # :650-700 with incx = incy = 1: a plain sequential sum
    dtemp = 0.0
    for i in range(n):
        dtemp = dtemp + dx[ix0 + i] * dy[iy0 + i]
    return dtemp


def daxpy(n, da, dx, ix0, dy, iy0):  # :612-648
    if n <= 0 or da == 0:
        return
    for i in range(n):
        dy[iy0 + i] = dy[iy0 + i] + da * dx[ix0 + i]


class LineSearchError(RuntimeError):
    """`throw new RuntimeException("dginit")`, LBFGS.java:253-254"""


class LBFGS:
    GTOL, STPMIN, STPMAX, XTOL, FTOL, MAXFEV, P5, P66, XTRAPF = 0.9, 1e-20, 1e20, 1e-16, 0.0001, 20, 0.5, 0.66, 4.0  # :193-205

    def __init__(self, x, m, eps):  # :42-60
        self.x, self.m, self.n, self.eps = x, m, len(x), eps
        n = self.n
        self.w = [0.0] * (n * (2 * m + 1) + 2 * m)
        self.iter = 0
        self.point = 0
        self.diag = [1.0] * n
        self.ispt = n + 2 * m
        self.iypt = self.ispt + n * m
        self.npt = 0
        self.info = 0
        self.stp = 0.0
        self.stp1 = 0.0
        self.nfev = 0
        self.bound = 0
        self.ys = self.yy = 0.0
        # mcsrch state :208-236
        self.dginit = self.dgtest = self.finit = 0.0
        self.stmin = self.stmax = self.width = self.width1 = 0.0
        self.stage1 = False
        self.infoc = 0
        self.brackt = False
        self.fx = self.dgx = self.fy = self.dgy = 0.0
        self.stx = self.sty = 0.0
        self.accepted = []  # test instrumentation: (stp, f, dg, finit, dginit) of every accepted line-search step

    def apply(self, f, g):  # :62-190
        n, m, w, diag = self.n, self.m, self.w, self.diag
        ispt, iypt = self.ispt, self.iypt
        whole = False
        if self.iter == 0:
            for i in range(n):
                w[ispt + i] = -g[i] * diag[i]
            gnorm = jsqrt(ddot(n, g, 0, g, 0))
            self.stp1 = jdiv(1.0, gnorm)
            whole = True
        while True:
            if whole:
                self.iter += 1
                self.info = 0
                self.bound = self.iter - 1
                if self.iter != 1:
                    if self.iter > m:
                        self.bound = m
                    self.ys = ddot(n, w, iypt + self.npt, w, ispt + self.npt)
                    self.yy = ddot(n, w, iypt + self.npt, w, iypt + self.npt)
                    d = jdiv(self.ys, self.yy)
                    for i in range(n):
                        diag[i] = d
                    cp = self.point
                    if self.point == 0:
                        cp = m
                    w[n + cp - 1] = jdiv(1.0, self.ys)
                    for i in range(n):
                        w[i] = -g[i]
                    cp = self.point
                    for _ in range(self.bound):
                        cp -= 1
                        if cp == -1:
                            cp = m - 1
                        sq = ddot(n, w, ispt + cp * n, w, 0)
                        inmc = n + m + cp
                        iycn = iypt + cp * n
                        w[inmc] = w[n + cp] * sq
                        daxpy(n, -w[inmc], w, iycn, w, 0)
                    for i in range(n):
                        w[i] = diag[i] * w[i]
                    for _ in range(self.bound):
                        yr = ddot(n, w, iypt + cp * n, w, 0)
                        beta = w[n + cp] * yr
                        inmc = n + m + cp
                        beta = w[inmc] - beta
                        iscn = ispt + cp * n
                        daxpy(n, beta, w, iscn, w, 0)
                        cp += 1
                        if cp == m:
                            cp = 0
                    for i in range(n):
                        w[ispt + self.point * n + i] = w[i]
                self.nfev = 0
                self.stp = 1.0
                if self.iter == 1:
                    self.stp = self.stp1
                for i in range(n):
                    w[i] = g[i]
            self.mcsrch(f, g)
            if self.info == -1:
                return False
            self.npt = self.point * n
            for i in range(n):
                w[ispt + self.npt + i] = self.stp * w[ispt + self.npt + i]
                w[iypt + self.npt + i] = g[i] - w[i]
            self.point += 1
            if self.point == m:
                self.point = 0
            gnorm = jsqrt(ddot(n, g, 0, g, 0))
            xnorm = jsqrt(ddot(n, self.x, 0, self.x, 0))
            xnorm = jmax(1.0, xnorm)
            if jdiv(gnorm, xnorm) <= self.eps:
                return True
            whole = True

    def mcsrch(self, f, g):  # :240-383
        n, w, diag, x = self.n, self.w, self.diag, self.x
        is0 = self.ispt + self.point * n
        if self.info != -1:
            self.infoc = 1
            self.dginit = 0.0
            for j in range(n):
                self.dginit = self.dginit + g[j] * w[is0 + j]
            if self.dginit >= 0:
                raise LineSearchError("dginit")
            self.brackt = False
            self.stage1 = True
            self.nfev = 0
            self.finit = f
            self.dgtest = self.FTOL * self.dginit
            self.width = self.STPMAX - self.STPMIN
            self.width1 = self.width / self.P5
            for j in range(n):
                diag[j] = x[j]
            self.stx = 0.0
            self.fx = self.finit
            self.dgx = self.dginit
            self.sty = 0.0
            self.fy = self.finit
            self.dgy = self.dginit
        while True:
            if self.info != -1:
                if self.brackt:
                    self.stmin = jmin(self.stx, self.sty)
                    self.stmax = jmax(self.stx, self.sty)
                else:
                    self.stmin = self.stx
                    self.stmax = self.stp + self.XTRAPF * (self.stp - self.stx)
                self.stp = jmax(self.stp, self.STPMIN)
                self.stp = jmin(self.stp, self.STPMAX)
                if ((self.brackt and (self.stp <= self.stmin or self.stp >= self.stmax)) or self.nfev >= self.MAXFEV - 1
                        or self.infoc == 0 or (self.brackt and self.stmax - self.stmin <= self.XTOL * self.stmax)):
                    self.stp = self.stx
                for j in range(n):
                    x[j] = diag[j] + self.stp * w[is0 + j]
                self.info = -1
                return
            self.info = 0
            self.nfev += 1
            dg = 0.0
            for j in range(n):
                dg = dg + g[j] * w[is0 + j]
            ftest1 = self.finit + self.stp * self.dgtest
            if (self.brackt and (self.stp <= self.stmin or self.stp >= self.stmax)) or self.infoc == 0:
                self.info = 6
            if self.stp == self.STPMAX and f <= ftest1 and dg <= self.dgtest:
                self.info = 5
            if self.stp == self.STPMIN and (f > ftest1 or dg >= self.dgtest):
                self.info = 4
            if self.nfev >= self.MAXFEV:
                self.info = 3
            if self.brackt and self.stmax - self.stmin <= self.XTOL * self.stmax:
                self.info = 2
            if f <= ftest1 and abs(dg) <= self.GTOL * (-self.dginit):
                self.info = 1
            if self.info != 0:
                self.accepted.append((self.info, self.stp, f, dg, self.finit, self.dginit))
                return
            if self.stage1 and f <= ftest1 and dg >= jmin(self.FTOL, self.GTOL) * self.dginit:
                self.stage1 = False
            if self.stage1 and f <= self.fx and f > ftest1:
                fm = f - self.stp * self.dgtest
                fxm = self.fx - self.stx * self.dgtest
                fym = self.fy - self.sty * self.dgtest
                dgm = dg - self.dgtest
                dgxm = self.dgx - self.dgtest
                dgym = self.dgy - self.dgtest
                fxm, dgxm, fym, dgym = self.mcstep(fxm, dgxm, fym, dgym, fm, dgm)
                self.fx = fxm + self.stx * self.dgtest
                self.fy = fym + self.sty * self.dgtest
                self.dgx = dgxm + self.dgtest
                self.dgy = dgym + self.dgtest
            else:
                self.fx, self.dgx, self.fy, self.dgy = self.mcstep(self.fx, self.dgx, self.fy, self.dgy, f, dg)
            if self.brackt:
                if abs(self.sty - self.stx) >= self.P66 * self.width1:
                    self.stp = self.stx + self.P5 * (self.sty - self.stx)
                self.width1 = self.width
                self.width = abs(self.sty - self.stx)

    def mcstep(self, fx, dx, fy, dy, fp, dp):  # :431-605 ; the 1-element arrays become return values
        stp, stx, sty = self.stp, self.stx, self.sty
        self.infoc = 0
        if ((self.brackt and (stp <= jmin(stx, sty) or stp >= jmax(stx, sty))) or dx * (stp - stx) >= 0.0
                or self.stmax < self.stmin):
            return fx, dx, fy, dy
        sgnd = dp * jdiv(dx, abs(dx))

        def max3(a, b, c):  # :703
            return (c if b < c else b) if a < b else (c if a < c else a)

        def sqr(v):
            return v * v

        if fp > fx:
            self.infoc = 1
            bound = True
            theta = jdiv(3 * (fx - fp), stp - stx) + dx + dp
            s = max3(abs(theta), abs(dx), abs(dp))
            gamma = s * jsqrt(sqr(jdiv(theta, s)) - jdiv(dx, s) * jdiv(dp, s))
            if stp < stx:
                gamma = -gamma
            p = (gamma - dx) + theta
            q = ((gamma - dx) + gamma) + dp
            r = jdiv(p, q)
            stpc = stx + r * (stp - stx)
            stpq = stx + jdiv(jdiv(dx, jdiv(fx - fp, stp - stx) + dx), 2) * (stp - stx)
            if abs(stpc - stx) < abs(stpq - stx):
                stpf = stpc
            else:
                stpf = stpc + (stpq - stpc) / 2
            self.brackt = True
        elif sgnd < 0.0:
            self.infoc = 2
            bound = False
            theta = jdiv(3 * (fx - fp), stp - stx) + dx + dp
            s = max3(abs(theta), abs(dx), abs(dp))
            gamma = s * jsqrt(sqr(jdiv(theta, s)) - jdiv(dx, s) * jdiv(dp, s))
            if stp > stx:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dx
            r = jdiv(p, q)
            stpc = stp + r * (stx - stp)
            stpq = stp + jdiv(dp, dp - dx) * (stx - stp)
            if abs(stpc - stp) > abs(stpq - stp):
                stpf = stpc
            else:
                stpf = stpq
            self.brackt = True
        elif abs(dp) < abs(dx):
            self.infoc = 3
            bound = True
            theta = jdiv(3 * (fx - fp), stp - stx) + dx + dp
            s = max3(abs(theta), abs(dx), abs(dp))
            gamma = s * jsqrt(jmax(0.0, sqr(jdiv(theta, s)) - jdiv(dx, s) * jdiv(dp, s)))
            if stp > stx:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = (gamma + (dx - dp)) + gamma
            r = jdiv(p, q)
            if r < 0.0 and gamma != 0.0:
                stpc = stp + r * (stx - stp)
            elif stp > stx:
                stpc = self.stmax
            else:
                stpc = self.stmin
            stpq = stp + jdiv(dp, dp - dx) * (stx - stp)
            if self.brackt:
                stpf = stpc if abs(stp - stpc) < abs(stp - stpq) else stpq
            else:
                stpf = stpc if abs(stp - stpc) > abs(stp - stpq) else stpq
        else:
            self.infoc = 4
            bound = False
            if self.brackt:
                theta = jdiv(3 * (fp - fy), sty - stp) + dy + dp
                s = max3(abs(theta), abs(dy), abs(dp))
                gamma = s * jsqrt(sqr(jdiv(theta, s)) - jdiv(dy, s) * jdiv(dp, s))
                if stp > sty:
                    gamma = -gamma
                p = (gamma - dp) + theta
                q = ((gamma - dp) + gamma) + dy
                r = jdiv(p, q)
                stpc = stp + r * (sty - stp)
                stpf = stpc
            elif stp > stx:
                stpf = self.stmax
            else:
                stpf = self.stmin
        if fp > fx:
            sty = stp
            fy = fp
            dy = dp
        else:
            if sgnd < 0.0:
                sty = stx
                fy = fx
                dy = dx
            stx = stp
            fx = fp
            dx = dp
        stpf = jmin(self.stmax, stpf)
        stpf = jmax(self.stmin, stpf)
        stp = stpf
        if self.brackt and bound:
            if sty > stx:
                stp = jmin(stx + 0.66 * (sty - stx), stp)
            else:
                stp = jmax(stx + 0.66 * (sty - stx), stp)
        self.stp, self.stx, self.sty = stp, stx, sty
        return fx, dx, fy, dy


def lbfgs(density_batch, n, x0=None, m=5, eps=0.1, max_evals=None):
    """Optimizer.lbfgs (Optimizer.scala:6-24): x starts at 0 (or x0, the multi-start extension), m = 5, eps = 0.1;
    f = -density, g = -gradient; loops until LBFGS.apply returns true.  density_batch(q[1][n]) -> [1][n+1] (density then
    gradient), e.g. OracleModel.density_batch.  max_evals: the reference has no cap; the batched kernel needs one.
    Returns dict(x, evals, info, f, lb): info 0 = converged, 1 = max_evals reached, 2 = "dginit" RuntimeException."""
    x = [0.0] * n if x0 is None else [float(v) for v in x0]
    lb = LBFGS(x, m, eps)
    evals, info, f = 0, 0, NaN
    while True:
        if max_evals is not None and evals >= max_evals:
            info = 1
            break
        out = density_batch([list(x)])[0]
        evals += 1
        f = float(out[0]) * -1
        g = [float(v) * -1 for v in out[1:]]
        try:
            if lb.apply(f, g):
                break
        except LineSearchError:
            info = 2
            break
    return {"x": list(x), "evals": evals, "info": info, "f": f, "lb": lb}
