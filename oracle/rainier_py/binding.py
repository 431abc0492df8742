"""
oracle/rainier_py/binding.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/librainier_oracle.so
(the `rno_*` mirror of include/rainier_cuda.h) plus a Python-side java.util.Random that shares its arithmetic.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rainier_b200.abi import ChainStats, Config, RngState

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.dirname(_HERE)
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _ORACLE_DIR], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ORACLE_DIR, "librainier_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.rno_last_error.restype = C.c_char_p
        L.rno_strict_log.restype = C.c_double
        L.rno_strict_log.argtypes = [C.c_double]
        L.rno_jpow.restype = C.c_double
        L.rno_jpow.argtypes = [C.c_double, C.c_double]
        L.rno_model_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int,
                                       C.c_int, C.POINTER(C.c_void_p)]
        L.rno_model_destroy.argtypes = [C.c_void_p]
        L.rno_model_nvars.argtypes = [C.c_void_p]
        L.rno_density_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.rno_model_emit_cpp.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.rno_model_use_compiled.argtypes = [C.c_void_p, C.c_char_p]
        L.rno_sample_traced.argtypes = [C.c_void_p, C.POINTER(Config), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        L.rno_config_default.argtypes = [C.POINTER(Config)]
        L.rno_function_create.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.rno_function_ninputs.argtypes = [C.c_void_p]
        L.rno_function_noutputs.argtypes = [C.c_void_p]
        L.rno_function_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.rno_function_destroy.argtypes = [C.c_void_p]
        L.rno_jr_seed.argtypes = [C.POINTER(RngState), C.c_int64]
        L.rno_jr_doubles.argtypes = [C.POINTER(RngState), C.c_void_p, C.c_int64]
        L.rno_jr_gaussians.argtypes = [C.POINTER(RngState), C.c_void_p, C.c_int64]
        _LIB = L
    return _LIB


class OracleError(RuntimeError):
    pass


def default_config():
    c = Config()
    lib().rno_config_default(C.byref(c))
    return c


class JRandom:
    """java.util.Random, bit-exact (scalar draws in Python; StrictMath.log via the oracle library)."""
    MULT = 0x5DEECE66D
    MASK = (1 << 48) - 1

    def __init__(self, seed=None):
        self.seed = 0 if seed is None else ((seed ^ self.MULT) & self.MASK)
        self.next_next = 0.0
        self.have_next = False
        self._slog = lib().rno_strict_log

    def next(self, bits):
        self.seed = (self.seed * self.MULT + 0xB) & self.MASK
        v = self.seed >> (48 - bits)
        if v >= 1 << (bits - 1) and bits == 32:
            v -= 1 << 32
        return v

    def nextDouble(self):
        return ((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))

    def nextGaussian(self):
        if self.have_next:
            self.have_next = False
            return self.next_next
        while True:
            v1 = 2 * self.nextDouble() - 1
            v2 = 2 * self.nextDouble() - 1
            s = v1 * v1 + v2 * v2
            if not (s >= 1 or s == 0):
                break
        multiplier = float(np.sqrt(np.float64(-2 * self._slog(s) / s)))
        self.next_next = v2 * multiplier
        self.have_next = True
        return v1 * multiplier

    # bulk draws through the C implementation (same stream)
    def state(self):
        return RngState(self.seed, self.next_next, 1 if self.have_next else 0, 0)

    def set_state(self, st):
        self.seed, self.next_next, self.have_next = st.seed48, st.next_gaussian, bool(st.have_next)

    def gaussians(self, n):
        st = self.state()
        out = np.empty(n, dtype=np.float64)
        lib().rno_jr_gaussians(C.byref(st), out.ctypes.data, n)
        self.set_state(st)
        return out

    def doubles(self, n):
        st = self.state()
        out = np.empty(n, dtype=np.float64)
        lib().rno_jr_doubles(C.byref(st), out.ctypes.data, n)
        self.set_state(st)
        return out


class ScalaRNG:
    """rainier-sampler/.../sampler/RNG.scala:6-26"""

    def __init__(self, seed):
        self.rand = JRandom(seed)

    def standardUniform(self):
        return self.rand.nextDouble()

    def standardNormal(self):
        return self.rand.nextGaussian()

    def int(self, until):
        return min(int(self.standardUniform() * until), until - 1)


class OracleModel:
    def __init__(self, rir, cols):
        L = lib()
        self._cols = [np.ascontiguousarray(c, dtype=np.float64) for c in cols]
        n = len(self._cols)
        ptrs = (C.c_void_p * max(n, 1))(*[c.ctypes.data for c in self._cols])
        rows = (C.c_int64 * max(n, 1))(*[len(c) for c in self._cols])
        h = C.c_void_p()
        self._rir = bytes(rir)
        rc = L.rno_model_create(self._rir, len(self._rir), ptrs, rows, n, 0, C.byref(h))
        if rc != 0:
            raise OracleError(L.rno_last_error().decode())
        self.h = h
        self.n = L.rno_model_nvars(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().rno_model_destroy(self.h)
            self.h = None

    def emit_cpp(self):
        """C++ source of this model's DataFunction as straight-line code (rno_model_emit_cpp)"""
        L = lib()
        need = C.c_size_t()
        L.rno_model_emit_cpp(self.h, None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value)
        L.rno_model_emit_cpp(self.h, buf, need.value, C.byref(need))
        return buf.value.decode()

    def compile_density(self, enable=True):
        """Switch the density to its COMPILED form: the node list as straight-line C++ built with the oracle's own flags
        (g++ -O2 -ffp-contract=off) -- the stand-in for the JVM executing rainier-compute's generated bytecode after JIT
        compilation, where the default form is a switch-per-node interpreter.  Bit-identical to the interpreter
        (tests/test_oracle_compiled.py)."""
        import hashlib
        import tempfile
        L = lib()
        if not enable:
            L.rno_model_use_compiled(self.h, None)
            return self
        src = self.emit_cpp()
        d = os.path.join(tempfile.gettempdir(), "rno_compiled")
        os.makedirs(d, exist_ok=True)
        key = hashlib.sha1(src.encode()).hexdigest()[:16]
        so = os.path.join(d, key + ".so")
        if not os.path.exists(so):
            cpp = os.path.join(d, key + ".cpp")
            with open(cpp, "w") as f:
                f.write(src)
            tmp = so + ".tmp%d" % os.getpid()
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-I", _ORACLE_DIR,
                            cpp, "-o", tmp], check=True)
            os.replace(tmp, so)
        if L.rno_model_use_compiled(self.h, so.encode()) != 0:
            raise OracleError(L.rno_last_error().decode())
        return self

    def density_batch(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.n)
        out = np.empty((q.shape[0], self.n + 1), dtype=np.float64)
        rc = lib().rno_density_batch(self.h, q.ctypes.data, q.shape[0], out.ctypes.data)
        if rc != 0:
            raise OracleError(lib().rno_last_error().decode())
        return out

    def sample(self, cfg, seeds=None, rng_states=None, trace=False, dense_mass=False):
        """returns dict(samples[chains][iterations][n], mass, stats, trace)"""
        L = lib()
        if rng_states is not None:
            chains = len(rng_states)
            arr = (RngState * chains)(*rng_states)
            cfg.rng_states = C.cast(arr, C.POINTER(RngState))
            seeds_a = np.zeros(chains, dtype=np.int64)
        else:
            seeds_a = np.ascontiguousarray(seeds, dtype=np.int64)
            chains = len(seeds_a)
            cfg.rng_states = None
        n = self.n
        samples = np.empty((chains, cfg.iterations, n), dtype=np.float64)
        mass = np.empty((chains, n * n if dense_mass else n), dtype=np.float64)
        stats = (ChainStats * chains)()
        tr = np.zeros((chains, cfg.warmup_iterations + cfg.iterations, 4), dtype=np.float64) if trace else None
        rc = L.rno_sample_traced(self.h, C.byref(cfg), seeds_a.ctypes.data, chains, samples.ctypes.data, mass.ctypes.data,
                                 C.cast(stats, C.c_void_p), tr.ctypes.data if trace else None)
        if rc != 0:
            raise OracleError(L.rno_last_error().decode())
        return {"samples": samples, "mass": mass, "stats": stats, "trace": tr}


class OracleFunction:
    """rno_function_*: Compiler.compile(inputs, outputs) + the CompiledFunction.output loop of Generator.prepare."""

    def __init__(self, rir):
        L = lib()
        self._rir = bytes(rir)
        h = C.c_void_p()
        if L.rno_function_create(self._rir, len(self._rir), 0, 0, C.byref(h)) != 0:
            raise OracleError(L.rno_last_error().decode())
        self.h = h
        self.nInputs = L.rno_function_ninputs(h)
        self.nOutputs = L.rno_function_noutputs(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().rno_function_destroy(self.h)
            self.h = None

    def __call__(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, max(self.nInputs, 1))[:, : self.nInputs]
        x = np.ascontiguousarray(x)
        out = np.empty((x.shape[0], self.nOutputs), dtype=np.float64)
        if lib().rno_function_eval(self.h, x.ctypes.data, x.shape[0], out.ctypes.data) != 0:
            raise OracleError(lib().rno_last_error().decode())
        return out
