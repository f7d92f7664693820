"""ctypes binding of the host-parallel text reader / writer (include/sph_textio.h -> libsph_textio.so).

Same surface as `io.read_data_from_file` / `io.make_save` (SUMMER_SPH.f90:594-738 | Variable.f90:729-942), for the
16M-64M particle files where a scalar parser dominates a run.  Pure host code (no GPU needed); `io.py` remains the
readable statement of the format and is what the tests compare this against.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from ._abi import SphParams, MODE_VARIABLE_H
from .state import Bodies, Sinks, GAS_FIELDS

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsph_textio.so")
_LIB = None


def build(force=False):
    src = os.path.join(_HERE, "..", "host", "sph_textio.cpp")
    hdr = os.path.join(_HERE, "..", "include", "sph_textio.h")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "..", "host"), "-s", "textio"])
    return LIB_PATH


def load_library():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not found: build it with `make -C host textio`")
        lib = C.CDLL(LIB_PATH)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        lib.sph_textio_last_error.restype = C.c_char_p
        lib.sph_ics_open.argtypes = [C.c_char_p, i32, dbl, dbl, i32, C.POINTER(vp)]
        lib.sph_ics_sizes.argtypes = [vp, C.POINTER(i64), C.POINTER(i32)]
        lib.sph_ics_fetch.argtypes = [vp] + [vp] * 18
        lib.sph_ics_close.argtypes = [vp]
        lib.sph_save_write.argtypes = [C.c_char_p, i32, i64] + [vp] * 10 + [i32] + [vp] * 7 + [i32]
        _LIB = lib
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _err(lib):
    return (lib.sph_textio_last_error() or b"").decode()


def read_data_from_file(filename, params: SphParams, threads=0, log=print):
    """Native twin of io.read_data_from_file: same rows, same values (bit for bit), same error texts."""
    lib = load_library()
    h = C.c_void_p()
    variable = 1 if (params.mode & MODE_VARIABLE_H) else 0
    rc = lib.sph_ics_open(os.fsencode(filename), variable, params.h_fixed, params.sink_radius, int(threads), C.byref(h))
    if rc == -2:
        raise FileNotFoundError(_err(lib))
    if rc:
        raise ValueError(_err(lib))
    try:
        n, ns = C.c_int64(), C.c_int32()
        lib.sph_ics_sizes(h, C.byref(n), C.byref(ns))
        bodies = Bodies.empty(n.value)
        sinks = Sinks.empty(ns.value)
        lib.sph_ics_fetch(h, *[_p(getattr(bodies, k)) for k in GAS_FIELDS],
                          *[_p(getattr(sinks, k)) for k in ("x", "y", "z", "vx", "vy", "vz", "m", "radius")])
    finally:
        lib.sph_ics_close(h)
    if log:
        log(f" Successfully read {len(bodies)} bodies and {len(sinks)} sinks from {filename}.")   # F:714
    return bodies, sinks


def make_save(bodies: Bodies, sinks: Sinks, number, params: SphParams, directory=".", threads=0):
    """Native twin of io.make_save: byte-identical `save<number>.txt`; an existing file is an error (F:728)."""
    lib = load_library()
    path = os.path.join(directory, f"save{number}.txt")
    variable = 1 if (params.mode & MODE_VARIABLE_H) else 0
    cols = [np.ascontiguousarray(getattr(bodies, k), dtype=np.float64) for k in GAS_FIELDS]
    scols = [np.ascontiguousarray(getattr(sinks, k), dtype=np.float64) for k in ("x", "y", "z", "vx", "vy", "vz", "m")]
    rc = lib.sph_save_write(os.fsencode(path), variable, len(bodies), *[_p(c) for c in cols], len(sinks), *[_p(c) for c in scols], int(threads))
    if rc == -5:
        raise FileExistsError(_err(lib))
    if rc:
        raise OSError(_err(lib))
    return path
