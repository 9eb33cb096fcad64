"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Import the UNMODIFIED reference modules from /root/reference/MC/tools inside this
(build) container by providing stand-ins for the third-party packages the reference
needs and this image lacks (scimath, numpy-quaternion, dask, memory_profiler).

This file does not restate any reference algorithm; the only arithmetic in it is the
third-party arithmetic that is not vendored in /root/reference:

  * numpy-quaternion (no version pin in the reference; README.md:9-14 names it):
    Hamilton product and ``rotate_vectors`` (v' = q v q^-1, norm-safe matrix form),
    used at MC/tools/rotation.py:44-53.
  * scimath unit objects: every unit expression on the path collapses to a plain
    SI factor (stoppingPowers.py:139 -> 1e-14, scattering.py:35 -> 1e14,
    material.py:30 -> 1e6, crossSections.py:281 -> 1e4, crossSections.py:55 -> 1e-3).

/root/reference does not exist on the GPU box: only ``oracle/gen_golden.py`` (run here,
outputs committed under tests/golden/) and the not-gpu validation tests that are skipped
when the reference is absent may call :func:`load_reference`.
"""
import os
import sys
import types

import numpy as np
import numpy.ma as ma

REFERENCE_ROOT = os.environ.get("MCD_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MC", "tools"))


class _Quat:
    """numpy-quaternion stand-in: (w, x, y, z) with Hamilton product."""

    def __init__(self, w, x, y, z):
        self.w = float(w)
        self.v = np.array([x, y, z], dtype=float)

    real = property(lambda s: s.w)
    imag = property(lambda s: s.v)

    def __mul__(a, b):
        w = a.w * b.w - a.v @ b.v
        v = a.w * b.v + b.w * a.v + np.cross(a.v, b.v)
        return _Quat(w, *v)

    def conj(self):
        return _Quat(self.w, *(-self.v))


def _rotate_vectors(q, vs):
    """v' = q v q^-1 for each row of vs (numpy-quaternion's norm-safe matrix form)."""
    vs = np.asarray(vs, dtype=float)
    w, (x, y, z) = q.w, q.v
    n = w * w + x * x + y * y + z * z
    m = np.array([
        [1 - 2 * (y * y + z * z) / n, 2 * (x * y - z * w) / n, 2 * (x * z + y * w) / n],
        [2 * (x * y + z * w) / n, 1 - 2 * (x * x + z * z) / n, 2 * (y * z - x * w) / n],
        [2 * (x * z - y * w) / n, 2 * (y * z + x * w) / n, 1 - 2 * (x * x + y * y) / n],
    ])
    return vs @ m.T


class _DaskArray:
    def __init__(self, a, chunks):
        self.a, self.c = a, int(chunks)

    def map_blocks(self, func, dtype=None):
        outer = self

        class _R:
            def compute(_):
                return ma.concatenate(
                    [func(outer.a[i:i + outer.c]) for i in range(0, len(outer.a), outer.c)], axis=0)
        return _R()


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _mod(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


_loaded = None


def load_reference(workdir="/tmp/mcd_oracle_work"):
    """Install the shims, chdir into a scratch MC/ dir and import the reference's tools package.

    Returns the ``tools`` package (attributes: material, crossSections, stoppingPowers,
    scattering, electron, probTables, singleScatter, multiScatter, rotation, fileTools,
    extFunctions).
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    if not hasattr(np, "NaN"):            # scattering.py:179 (removed in numpy 2)
        np.NaN = np.nan
    if not hasattr(np, "trapz"):          # probTables.py:181,216,220,353,454
        np.trapz = np.trapezoid

    def has_units(f=None, **kw):
        return f if f is not None else (lambda g: g)

    class UnitScalar(float):
        def __new__(cls, v, units=None):
            return float.__new__(cls, v)

    _mod("scimath")
    _mod("scimath.units")
    _mod("scimath.units.api", has_units=has_units, UnitScalar=UnitScalar,
         UnitArray=lambda v, units=None: np.asarray(v))
    _mod("scimath.units.length", angstrom=1e-10, cm=1e-2, m=1.0, nm=1e-9)
    _mod("scimath.units.energy", eV=1.602176634e-19, KeV=1.602176634e-16)
    _mod("scimath.units.mass", kg=1.0)
    _mod("scimath.units.time", s=1.0)
    _mod("scimath.units.dimensionless", dim=1.0)

    np.quaternion = _Quat
    _mod("quaternion", rotate_vectors=_rotate_vectors, quaternion=_Quat)

    _mod("dask", array=_mod("dask.array", from_array=lambda a, chunks: _DaskArray(a, chunks)),
         compute=lambda *a: a)
    _mod("dask.diagnostics", ProgressBar=_NullCtx)
    _mod("memory_profiler", profile=lambda f: f)

    os.makedirs(os.path.join(workdir, "MC"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "logs"), exist_ok=True)          # probTables.py:627
    os.makedirs(os.path.join(workdir, "tables", "integration"), exist_ok=True)  # probTables.py:562
    os.chdir(os.path.join(workdir, "MC"))
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "MC"))
    import tools                                                         # noqa: F401
    from tools import (material, crossSections, stoppingPowers, scattering, electron,  # noqa: F401
                       probTables, singleScatter, multiScatter, rotation, fileTools, extFunctions)
    _loaded = tools
    return tools


class injected_uniforms:
    """Context manager: route the reference's ``random.random`` to a given sequence.

    Every draw on the path goes through ``random.random`` (scattering.py:56,89,147,556,676,807,831).
    """

    def __init__(self, us):
        self.it = iter(us)
        self.count = 0

    def _next(self):
        self.count += 1
        return float(next(self.it))

    def __enter__(self):
        import random
        self._random = random
        self._saved = random.random
        random.random = self._next
        return self

    def __exit__(self, *a):
        self._random.random = self._saved
        return False
