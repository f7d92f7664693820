"""ctypes mirror of include/sph_b200.h (struct layouts and constants only; no library loading).

Shared by the engine binding (engine.py) and by test infrastructure.
"""
import ctypes as C

MODE_FIXED_H = 0        # SUMMER_SPH.f90
MODE_VARIABLE_H = 1     # "SUMMER_SPH - Variable.f90"
FLAG_SOFT_USES_HI = 2   # "(test new)" softening 0.001*h_i (T:298)
FLAG_SINK_MERGE_SPIN = 4   # opt-in, not the reference: sink spin bookkeeping (F:33,509) + the merger stub V:1067-1073

EVAL_TREE, EVAL_DENSITY, EVAL_GRAVITY, EVAL_SINKS, EVAL_SPH = 1, 2, 4, 8, 16
EVAL_ALL = 31

ERRORS = {0: "ok", -1: "bad argument", -2: "no CUDA device", -3: "CUDA error", -4: "out of memory",
          -5: "bad state", -6: "key depth exceeded", -7: "communicator error"}


# slots of sph_conserved() (include/sph_b200.h)
CONSERVED = ("e_kin", "e_int", "e_pot", "px", "py", "pz", "lx", "ly", "lz", "mass", "e_pot_gas", "e_pot_sink")


def conserved_dict(out):
    d = {k: float(v) for k, v in zip(CONSERVED, out)}
    d["e_total"] = d["e_kin"] + d["e_int"] + d["e_pot"]
    return d


def drift_report(first, last):
    """Drift of the conserved sums between two `conserved()` dicts: energy relative to |E_total| of the first,
    momentum relative to sqrt(2 E_kin M) (the momentum the system would carry if it all moved one way; the larger
    E_kin of the two states, so that a start from rest still has a scale), angular momentum relative to the larger
    |L| of the two.  Absolute changes are given as well; a relative value is None when its scale is zero.
    Tree gravity (Barnes-Hut monopoles) is not symmetric, so momentum is not conserved by the reference's scheme
    (SURVEY.md Appendix D): this is a report, not a test."""
    import math
    norm = lambda d, ks: math.sqrt(sum(d[k] ** 2 for k in ks))   # noqa: E731
    p_scale = math.sqrt(2.0 * max(abs(first["e_kin"]), abs(last["e_kin"])) * max(first["mass"], last["mass"]))
    l_scale = max(norm(first, ("lx", "ly", "lz")), norm(last, ("lx", "ly", "lz")))
    dp = math.sqrt(sum((last[k] - first[k]) ** 2 for k in ("px", "py", "pz")))
    dl = math.sqrt(sum((last[k] - first[k]) ** 2 for k in ("lx", "ly", "lz")))
    de = last["e_total"] - first["e_total"]
    return {
        "energy_rel": de / abs(first["e_total"]) if first["e_total"] else None,
        "momentum_rel": dp / p_scale if p_scale else None,
        "angular_momentum_rel": dl / l_scale if l_scale else None,
        "mass_rel": (last["mass"] - first["mass"]) / first["mass"] if first["mass"] else None,
        "energy_abs": de, "momentum_abs": dp, "angular_momentum_abs": dl,
        "e_total_first": first["e_total"], "e_total_last": last["e_total"],
    }


class SphParams(C.Structure):
    """`sph_params` — V's type(param) (Variable.f90:54-64) + F's compile-time constants (SUMMER_SPH.f90:7-11)."""
    _fields_ = [
        ("mode", C.c_int32), ("max_depth", C.c_int32), ("nq", C.c_int32), ("n_ranks", C.c_int32),
        ("h_fixed", C.c_double), ("bounding_size", C.c_double), ("theta", C.c_double), ("gamma", C.c_double),
        ("eta", C.c_double), ("convergence_criteria", C.c_double), ("max_length", C.c_double),
        ("timestep_scale", C.c_double), ("end_time", C.c_double), ("sink_radius", C.c_double),
        ("theta_override", C.c_int32), ("decomposition", C.c_int32),
    ]

    @property
    def variable_h(self):
        return bool(self.mode & MODE_VARIABLE_H)

    def copy(self, **kw):
        q = SphParams.from_buffer_copy(bytes(self))
        for k, v in kw.items():
            setattr(q, k, v)
        return q


class SphCounts(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "n_gas", "n_nodes", "density_candidates", "density_contributing", "sph_pairs",
        "grav_opened", "grav_accepted", "h_iterations")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def default_params(mode=MODE_FIXED_H, **kw):
    """Reference defaults. Fixed h: the compile-time constants of SUMMER_SPH.f90:7-11,465-466,694,851,873.
    Variable h: nq/sink radius of Variable.f90:8,830 and the parameters.txt values SURVEY.md §8(d) uses."""
    p = SphParams()
    p.mode = mode
    p.n_ranks = 1
    p.h_fixed = 2.5
    p.theta = 0.5
    p.theta_override = 0
    if mode & MODE_VARIABLE_H:
        p.max_depth, p.nq = 1000, 2500
        p.bounding_size, p.gamma, p.eta = 1500.0, 1.4, 1.2
        p.convergence_criteria, p.max_length = 1e-3, 50.0
        p.timestep_scale, p.end_time, p.sink_radius = 0.25, 0.1, 5.0
    else:
        p.max_depth, p.nq = 1000, 5000
        p.bounding_size, p.gamma, p.eta = 1500.0, 1.4, 1.2
        p.convergence_criteria, p.max_length = 1e-3, 50.0
        p.timestep_scale, p.end_time, p.sink_radius = 0.25, 1000.0, 3.5
    for k, v in kw.items():
        setattr(p, k, v)
    return p
