"""ctypes wrapper of the CPU oracle (oracle/sph_oracle.cpp).  TEST INFRASTRUCTURE — see that file's header.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from summersph_b200._abi import SphParams, SphCounts, EVAL_ALL, CONSERVED, conserved_dict
from summersph_b200.state import Bodies, Sinks

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libsph_oracle.so")
    src = os.path.join(_HERE, "sph_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "sph_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_create.restype = C.c_void_p
        _LIB.orc_create.argtypes = [C.POINTER(SphParams), C.c_int]
        _LIB.orc_lookup_grav_kernel.restype = C.c_double
        _LIB.orc_lookup_grav_kernel.argtypes = [C.c_void_p, C.c_double, C.c_double]
        _LIB.orc_lookup_kernel.argtypes = [C.c_void_p, C.c_double, C.c_double, _dp, _dp]
        _LIB.orc_G.restype = C.c_double
        _LIB.orc_download_neighbours.restype = C.c_int64
        _LIB.orc_kick.argtypes = [C.c_void_p, C.c_double]
        _LIB.orc_drift.argtypes = [C.c_void_p, C.c_double]
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self, params: SphParams, threads=1):
        self.params = params
        self._l = lib()
        self._c = C.c_void_p(self._l.orc_create(C.byref(params), int(threads)))

    def close(self):
        if self._c:
            self._l.orc_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- tables / lookups (SUMMER_SPH.f90:55-146) -------------------------------------------------
    def tables(self):
        nq = self.params.nq
        w, dw, g = (np.zeros(nq + 1) for _ in range(3))
        self._l.orc_tables(self._c, _p(w), _p(dw), _p(g))
        return w, dw, g

    def lookup_kernel(self, r, h):
        W, dW = C.c_double(), C.c_double()
        self._l.orc_lookup_kernel(self._c, float(r), float(h), C.byref(W), C.byref(dW))
        return W.value, dW.value

    def lookup_grav_kernel(self, r, h):
        return self._l.orc_lookup_grav_kernel(self._c, float(r), float(h))

    @property
    def G(self):
        return self._l.orc_G()

    # -- state ---------------------------------------------------------------------------------------
    def upload(self, b: Bodies, s: Sinks):
        rad = np.where(np.isnan(s.radius), self.params.sink_radius, s.radius) if len(s) else s.radius
        self._keep = (b, s, rad)
        self._l.orc_upload(self._c, C.c_int64(len(b)), _p(b.x), _p(b.y), _p(b.z), _p(b.vx), _p(b.vy), _p(b.vz),
                           _p(b.u), _p(b.m), _p(b.alpha), _p(b.h), C.c_int32(len(s)),
                           _p(s.x), _p(s.y), _p(s.z), _p(s.vx), _p(s.vy), _p(s.vz), _p(s.m), _p(rad))

    def sizes(self):
        n, ns = C.c_int64(), C.c_int32()
        self._l.orc_sizes(self._c, C.byref(n), C.byref(ns))
        return n.value, ns.value

    def record_neighbours(self, on=True):
        self._l.orc_record_neighbours(self._c, int(on))

    def evaluate(self, mask=EVAL_ALL):
        self._l.orc_evaluate(self._c, int(mask))

    def step(self, dt, t):
        cdt, ct = C.c_double(dt), C.c_double(t)
        self._l.orc_step(self._c, C.byref(cdt), C.byref(ct))
        return cdt.value, ct.value

    def calc_smoothing(self):
        self._l.orc_calc_smoothing(self._c)

    def kick(self, dt):
        self._l.orc_kick(self._c, float(dt))

    def drift(self, dt):
        self._l.orc_drift(self._c, float(dt))

    def next_timestep(self, dt):
        c = C.c_double(dt)
        self._l.orc_next_timestep(self._c, C.byref(c))
        return c.value

    def accrete(self):
        self._l.orc_accrete(self._c)

    def check_bounds(self):
        self._l.orc_check_bounds(self._c)

    def check_sink_creation(self):
        self._l.orc_check_sink_creation(self._c)

    def download(self):
        n, ns = self.sizes()
        b, s = Bodies.empty(n), Sinks.empty(ns)
        self._l.orc_download(self._c, _p(b.x), _p(b.y), _p(b.z), _p(b.vx), _p(b.vy), _p(b.vz), _p(b.u), _p(b.m),
                             _p(b.alpha), _p(b.h), _p(s.x), _p(s.y), _p(s.z), _p(s.vx), _p(s.vy), _p(s.vz),
                             _p(s.m), _p(s.radius))
        return b, s

    def diag(self):
        n, ns = self.sizes()
        d = {k: np.zeros(n) for k in ("rho", "omega", "P", "c", "ax", "ay", "az", "udot", "alphadot")}
        d.update({k: np.zeros(ns) for k in ("sink_ax", "sink_ay", "sink_az")})
        self._l.orc_download_diag(self._c, *[_p(d[k]) for k in ("rho", "omega", "P", "c", "ax", "ay", "az", "udot",
                                                                "alphadot", "sink_ax", "sink_ay", "sink_az")])
        return d

    def accel_parts(self):
        n, _ = self.sizes()
        a = [np.zeros(n) for _ in range(6)]
        self._l.orc_download_accel_parts(self._c, *[_p(v) for v in a])
        return {"grav": np.stack(a[:3], 1), "grav_sink": np.stack(a[3:], 1)}

    def tree(self):
        n, _ = self.sizes()
        t = {"order": np.zeros(n, np.int32), "level": np.zeros(n, np.int32), "cx": np.zeros(n), "cy": np.zeros(n),
             "cz": np.zeros(n), "size": np.zeros(n), "n_in_leaf": np.zeros(n, np.int32)}
        self._l.orc_download_tree(self._c, *[_p(t[k]) for k in ("order", "level", "cx", "cy", "cz", "size", "n_in_leaf")])
        ctr = np.zeros(3); sz = C.c_double()
        self._l.orc_root(self._c, _p(ctr), C.byref(sz))
        t["root_center"], t["root_size"] = ctr, sz.value
        return t

    def neighbours(self, with_list=True):
        n, _ = self.sizes()
        count = np.zeros(n, np.int32); hsh = np.zeros(n, np.uint64); off = np.zeros(n + 1, np.int64)
        tot = self._l.orc_download_neighbours(self._c, _p(count), _p(hsh), _p(off), None, C.c_int64(0))
        lst = None
        if with_list:
            lst = np.zeros(max(tot, 1), np.int32)
            self._l.orc_download_neighbours(self._c, _p(count), _p(hsh), _p(off), _p(lst), C.c_int64(tot))
            lst = lst[:tot]
        return count, hsh, off, lst

    def sample_eval(self, targets):
        """One evaluation's results for a few target particles only (0-based numbers); the tree is built over the
        whole set.  For checking the engine at sizes the full serial pair loop cannot reach (sph_oracle.cpp: sample_eval)."""
        tg = np.ascontiguousarray(targets, dtype=np.int32)
        nt = len(tg)
        d = {k: np.zeros(nt) for k in ("rho", "omega", "P", "c", "ax", "ay", "az", "udot", "alphadot")}
        d["count"] = np.zeros(nt, np.int32); d["hash"] = np.zeros(nt, np.uint64); d["pairs"] = np.zeros(nt, np.int64)
        self._l.orc_sample_eval(self._c, C.c_int32(nt), _p(tg), *[_p(d[k]) for k in ("rho", "omega", "P", "c", "ax", "ay", "az", "udot",
                                                                                   "alphadot", "count", "hash", "pairs")])
        return d

    def check_sink_merger(self):
        self._l.orc_check_sink_merger(self._c)

    def sink_spin(self):
        _, ns = self.sizes()
        sp = [np.zeros(ns) for _ in range(3)]
        self._l.orc_download_sink_spin(self._c, *[_p(v) for v in sp])
        return np.stack(sp, 1)

    def conserved(self):
        out = np.zeros(len(CONSERVED))
        self._l.orc_conserved(self._c, _p(out))
        return conserved_dict(out)

    def counters(self):
        c = SphCounts()
        self._l.orc_counters(self._c, C.byref(c))
        return c.as_dict()
