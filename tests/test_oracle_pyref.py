"""The two CPU restatements of the reference held to each other.

oracle/sph_oracle.cpp (index ranges, snapshots, the checker of the CUDA engine) and oracle/pyref.py (a literal,
routine-by-routine Python reading of the Fortran with the reference's own AoS records and pointer octree of particle
copies) were written independently from the same source.  Neither can be run against the reference itself (no Fortran
compiler, no golden vectors: "parity unpinned"), but a misreading would have to be made twice to pass here.
Everything is compared after one evaluation and after whole loop bodies, in both programs, including accretion,
bounds removal, the h iteration and sink creation.  Both perform the reference's operations in the reference's order in
IEEE double precision without contraction (sqrt is the only library function), so every compared number must be
bit-identical: the tolerance is zero; dt, t, particle counts and iteration counts must be equal."""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, FLAG_SOFT_USES_HI, ics, Sinks
from oracle.oracle import Oracle
from oracle.pyref import Program
from conftest import relerr



def program_for(p):
    variable = bool(p.mode & MODE_VARIABLE_H)
    return Program(variable, max_depth=p.max_depth, bounding_size=p.bounding_size, gamma=p.gamma, eta=p.eta,
                   convergence_criteria=p.convergence_criteria, max_length=p.max_length,
                   timestep_scale=p.timestep_scale, sink_radius=p.sink_radius, soft_uses_hi=bool(p.mode & FLAG_SOFT_USES_HI))


def col(items, attr, k=None):
    return np.array([getattr(b, attr) if k is None else getattr(b, attr)[k] for b in items])


def compare_rates(py, o):
    d = o.diag()
    pairs = {"rho": col(py.bodies, "density"), "P": col(py.bodies, "pressure"), "c": col(py.bodies, "sound_speed"),
             "ax": col(py.bodies, "acceleration", 0), "ay": col(py.bodies, "acceleration", 1), "az": col(py.bodies, "acceleration", 2),
             "udot": col(py.bodies, "internal_energy_rate"), "alphadot": col(py.bodies, "alpha_rate"),
             "sink_ax": col(py.sinks, "acceleration", 0), "sink_ay": col(py.sinks, "acceleration", 1), "sink_az": col(py.sinks, "acceleration", 2)}
    if py.variable:
        pairs["omega"] = col(py.bodies, "omega")
    for k, v in pairs.items():
        assert np.array_equal(d[k], v), (k, relerr(d[k], v))


def compare_state(py, o):
    b, s = o.download()
    assert len(b) == len(py.bodies) and len(s) == len(py.sinks)
    for name, attr, k in (("x", "position", 0), ("y", "position", 1), ("z", "position", 2), ("vx", "velocity", 0),
                          ("vy", "velocity", 1), ("vz", "velocity", 2), ("u", "internal_energy", None), ("m", "mass", None),
                          ("alpha", "alpha", None)) + ((("h", "s_length", None),) if py.variable else ()):
        assert np.array_equal(getattr(b, name), col(py.bodies, attr, k)), name
    for name, attr, k in (("x", "position", 0), ("y", "position", 1), ("z", "position", 2), ("vx", "velocity", 0),
                          ("vy", "velocity", 1), ("vz", "velocity", 2), ("m", "mass", None), ("radius", "radius", None)):
        assert np.array_equal(getattr(s, name), col(py.sinks, attr, k)), "sink " + name


def leaves_in_dfs_order(node, level=0, out=None):
    """(number, level, centre, size, n_particles) of every childless node in the reference's recursion order (F:240-244)."""
    out = [] if out is None else out
    if node.children is None:
        for q in node.particles:
            out.append((q.number, level, tuple(node.center), node.size, node.n_particles))
        return out
    for ch in node.children:
        if ch.n_particles > 0:
            leaves_in_dfs_order(ch, level + 1, out)
    return out


def compare_tree(py, o):
    """Depth-first leaf order (= the Morton order the engine sorts into), leaf level, cell centre and size."""
    t = o.tree()
    lv = leaves_in_dfs_order(py.root)
    order = np.array([q[0] - 1 for q in lv], np.int32)
    assert np.array_equal(order, t["order"])
    assert np.array_equal(py.root.center, t["root_center"]) and py.root.size == t["root_size"]
    for (num, level, ctr, size, npart) in lv:
        i = num - 1
        assert level == t["level"][i] and size == t["size"][i] and npart == t["n_in_leaf"][i], num
        assert ctr == (t["cx"][i], t["cy"][i], t["cz"][i]), num


def test_tables_and_literals_agree():
    for mode in (MODE_FIXED_H, MODE_VARIABLE_H):
        p = default_params(mode)
        o = Oracle(p); py = program_for(p)
        w, dw, g = o.tables()
        assert np.array_equal(w, py.w_table) and np.array_equal(dw, py.dw_table) and np.array_equal(g, py.grav_table)
        assert o.G == py.G
        for r, h in ((0.3, 1.1), (1.7, 0.9), (2.5, 1.0), (0.0, 2.5)):
            assert o.lookup_kernel(r, h) == py.lookup_kernel(r, h)
            assert o.lookup_grav_kernel(r, h) == py.lookup_grav_kernel(r, h)


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H, MODE_VARIABLE_H | FLAG_SOFT_USES_HI])
def test_one_evaluation(mode):
    """create_tree + get_density + EOS + find_forces (F:894-898) with two sinks and viscosity on."""
    p = default_params(mode)
    b, _ = ics.keplerian_disc(260, seed=31)
    if mode == MODE_FIXED_H:
        p = p.copy(h_fixed=6.0)                 # the module constant `smoothing`: enough neighbours at this N
    s = Sinks([0.0, 35.0], [0.0, 4.0], [0.0, 0.5], [0.0, 0.1], [0.0, 5.0], [0.0, 0.0], [1.0, 0.02], [12.0, 5.0])
    o = Oracle(p); o.upload(b, s); o.evaluate()
    py = program_for(p); py.smoothing = p.h_fixed; py.load(b, s)
    for i, q in enumerate(py.bodies):
        q.number = i + 1
    py.evaluate()
    compare_rates(py, o)
    compare_tree(py, o)
    co = o.counters()
    assert {k: co[k] for k in py.cnt} == py.cnt          # the interaction counts the engine's parity tests compare with
    assert max(abs(q.alpha_rate) for q in py.bodies) > 0 and max(abs(q.internal_energy_rate) for q in py.bodies) > 0


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_loop_bodies_with_accretion_and_bounds(mode):
    """Three loop bodies (F:886-928 | V:1120-1162): accreting sinks, a tight bounding cube, the dt ladder, and in the
    variable-h program the Newton-Raphson h update."""
    p = default_params(mode, bounding_size=85.0)
    if mode == MODE_FIXED_H:
        p = p.copy(h_fixed=6.0)
    b, _ = ics.keplerian_disc(300, seed=9)
    s = Sinks([0.0, 40.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 6.0], [0.0, 0.0], [1.0, 0.01], [22.0, 9.0])
    o = Oracle(p); o.upload(b, s)
    py = program_for(p); py.smoothing = p.h_fixed; py.load(b, s)
    dto = dtp = 0.01; to = tp = 0.0
    for k in range(3):
        dto, to = o.step(dto, to)
        dtp, tp = py.step(dtp, tp)
        assert (dto, to) == (dtp, tp), k
        assert o.sizes() == (len(py.bodies), len(py.sinks)), k
        compare_state(py, o)
    assert len(py.bodies) < 300                  # something was accreted or left the cube
    assert dtp != 0.01                           # the ladder moved


def test_h_iteration_counts_and_sink_creation():
    """calc_smoothing's inner re-walks (V:529-539) and check_sink_creation (V:549-597) followed by the new sink
    accreting its seed particle in the same loop body."""
    p = default_params(MODE_VARIABLE_H, max_length=0.21)
    b, s = ics.keplerian_disc(280, seed=21)
    b.m[123] = 5e-3; b.h[123] = 0.2
    o = Oracle(p); o.upload(b, s)
    py = program_for(p); py.load(b, s)
    # the loop body by hand on the Python side, to count the h iterations like the oracle's counter
    for i, q in enumerate(py.bodies):
        q.number = i + 1
    py.evaluate(); py.kick(0.01); py.drift(0.01); py.evaluate(); py.kick(0.01)
    dtp = py.get_next_timestep(0.01)
    iters = py.calc_smoothing()
    py.check_sink_creation()
    assert len(py.sinks) == 2 and py.sinks[1].mass == 1e-11
    py.initiate_sink_accretion(); py.check_bounds()
    dto, to = o.step(0.01, 0.0)
    assert dto == dtp and o.counters()["h_iterations"] == iters
    assert o.sizes() == (len(py.bodies), len(py.sinks)) == (279, 2)
    compare_state(py, o)


def test_no_sink_file_dummy_sink():
    p = default_params(MODE_VARIABLE_H)
    b, _ = ics.uniform_sphere(200)
    o = Oracle(p); o.upload(b, Sinks.empty(0)); o.evaluate()
    py = program_for(p); py.load(b, Sinks.empty(0))
    for i, q in enumerate(py.bodies):
        q.number = i + 1
    py.evaluate()
    assert len(py.sinks) == 1 and py.sinks[0].mass == 0.0
    compare_rates(py, o)


@pytest.mark.parametrize("mode", [MODE_FIXED_H, MODE_VARIABLE_H])
def test_depth_limited_tree_and_coincident_particles(mode):
    """max_depth below the natural leaf depth and two particles at the same place: multi-particle childless nodes are
    skipped by the density / SPH walks and taken whole by gravity (F:182,431,443)."""
    p = default_params(mode, max_depth=4)
    if mode == MODE_FIXED_H:
        p = p.copy(h_fixed=6.0)
    b, s = ics.keplerian_disc(240, seed=4)
    b.x[17], b.y[17], b.z[17] = b.x[16], b.y[16], b.z[16]
    o = Oracle(p); o.upload(b, s); o.evaluate()
    py = program_for(p); py.smoothing = p.h_fixed; py.load(b, s)
    for i, q in enumerate(py.bodies):
        q.number = i + 1
    py.evaluate()
    # nobody gathers from a particle inside a childless multi-particle node (neither branch of F:431/443 fires)
    d = o.diag()
    rho = col(py.bodies, "density")
    for k, v in (("ax", col(py.bodies, "acceleration", 0)), ("udot", col(py.bodies, "internal_energy_rate")), ("c", col(py.bodies, "sound_speed"))):
        assert np.array_equal(d[k], v, equal_nan=True), k
    assert np.array_equal(d["rho"], rho)
    compare_tree(py, o)
    assert max(q[4] for q in leaves_in_dfs_order(py.root)) > 1            # there are multi-particle childless nodes


# ------------------------------------------------------------------------------------------------------
# The committed golden fixtures (tests/golden/*.npz) were generated from the C++ oracle; the literal Python
# restatement reproduces them bit for bit as well, so the small-case targets of the CUDA engine
# (tests/test_golden.py::test_engine_matches_golden) are what a second, independent reading of the Fortran computes.
def _mix64(z):
    z = (z + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


@pytest.mark.parametrize("name", ["disc600_fixed", "disc600_fixed_accrete", "disc600_variable", "disc600_variable_accrete", "sod_variable"])
def test_pyref_reproduces_golden(name):
    from _cases import load_golden
    z, p, b, s = load_golden(name)
    py = program_for(p); py.smoothing = p.h_fixed; py.record_ngb = True
    py.load(b, s)
    for i, q in enumerate(py.bodies):
        q.number = i + 1
    py.evaluate()
    ev = {"rho": col(py.bodies, "density"), "P": col(py.bodies, "pressure"), "c": col(py.bodies, "sound_speed"),
          "ax": col(py.bodies, "acceleration", 0), "ay": col(py.bodies, "acceleration", 1), "az": col(py.bodies, "acceleration", 2),
          "udot": col(py.bodies, "internal_energy_rate"), "alphadot": col(py.bodies, "alpha_rate"),
          "sink_ax": col(py.sinks, "acceleration", 0), "sink_ay": col(py.sinks, "acceleration", 1), "sink_az": col(py.sinks, "acceleration", 2)}
    if py.variable:
        ev["omega"] = col(py.bodies, "omega")
    for k, v in ev.items():
        assert np.array_equal(v, z["ev_" + k], equal_nan=True), k
    lv = leaves_in_dfs_order(py.root)
    assert np.array_equal(np.array([q[0] - 1 for q in lv], np.int32), z["tree_order"])
    for (num, level, ctr, size, _) in lv:
        i = num - 1
        assert level == z["tree_level"][i] and size == z["tree_size"][i] and ctr == (z["tree_cx"][i], z["tree_cy"][i], z["tree_cz"][i])
    assert np.array_equal(np.array([len(v) for v in py.ngb], np.int32), z["ngb_count"])
    hsh = np.array([sum(_mix64(j) for j in v) & 0xFFFFFFFFFFFFFFFF for v in py.ngb], np.uint64)
    assert np.array_equal(hsh, z["ngb_hash"])
    # the loop bodies
    py.record_ngb = False
    py.load(b, s)
    dt, t = 0.01, 0.0
    for _ in range(int(z["steps"][0])):
        dt, t = py.step(dt, t)
    assert np.array_equal(np.array([dt, t]), z["st_dt_t"])
    for name_, attr, k in (("x", "position", 0), ("y", "position", 1), ("z", "position", 2), ("vx", "velocity", 0), ("vy", "velocity", 1),
                           ("vz", "velocity", 2), ("u", "internal_energy", None), ("m", "mass", None), ("alpha", "alpha", None)):
        assert np.array_equal(col(py.bodies, attr, k), z["st_" + name_]), name_
    if py.variable:
        assert np.array_equal(col(py.bodies, "s_length"), z["st_h"])
    for name_, attr, k in (("x", "position", 0), ("vx", "velocity", 0), ("m", "mass", None), ("radius", "radius", None)):
        assert np.array_equal(col(py.sinks, attr, k), z["st_sink_" + name_]), "sink " + name_


# ------------------------------------------------------------------------------------------------------
# Randomised cases: program, parameters (tight / wide bounding cube, max_depth 3 / 6 / 1000, gamma, eta, tolerance,
# timestep scale, max_length), geometry, viscosity and 0-3 sinks (some massless) drawn per seed; two loop bodies each.
# Shallow depth limits put most particles into childless multi-particle nodes, so NaNs (rho = 0) run through kicks,
# the bounding box (MAXVAL / MINVAL skip them), the dt minimum and the accretion sums - the regime where the two
# formulations are least alike.  Everything must still be bit-identical, NaN patterns included.
def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    mode = [MODE_FIXED_H, MODE_VARIABLE_H, MODE_VARIABLE_H | FLAG_SOFT_USES_HI][seed % 3]
    kw = dict(bounding_size=float(rng.choice([60.0, 85.0, 1500.0])), max_depth=int(rng.choice([3, 6, 1000])),
              gamma=float(rng.choice([1.4, 5 / 3])), eta=float(rng.uniform(1.0, 1.4)),
              convergence_criteria=float(rng.choice([1e-2, 1e-3, 1e-4])), timestep_scale=float(rng.choice([0.1, 0.25, 0.5])),
              max_length=float(rng.choice([50.0, 3.0])))
    p = default_params(mode, **kw)
    if mode == MODE_FIXED_H:
        p = p.copy(h_fixed=float(rng.uniform(4, 8)))
    n = int(rng.integers(60, 220))
    kind = seed % 4
    if kind == 0:
        b, _ = ics.keplerian_disc(n, seed=seed)
    elif kind == 1:
        b, _ = ics.thin_ring(n, seed=seed)
    elif kind == 2:
        b, _ = ics.uniform_sphere(n, seed=seed); b.vx[:] = rng.normal(0, 0.3, n); b.vy[:] = rng.normal(0, 0.3, n)
    else:
        b, _ = ics.keplerian_disc(n, seed=seed); b.alpha[:] = rng.uniform(0, 1, n)
    ns = int(rng.integers(0, 4))
    s = Sinks(rng.uniform(-40, 40, ns), rng.uniform(-40, 40, ns), rng.uniform(-2, 2, ns), rng.normal(0, 1, ns), rng.normal(0, 1, ns),
              rng.normal(0, .1, ns), rng.choice([0.0, 0.01, 1.0], ns), rng.uniform(2, 25, ns)) if ns else Sinks.empty(0)
    return p, b, s


@pytest.mark.parametrize("seed", range(12))
def test_random_cases_two_loop_bodies(seed):
    p, b, s = _random_case(seed)
    o = Oracle(p); o.upload(b, s)
    py = program_for(p); py.smoothing = p.h_fixed; py.load(b, s)
    dto = dtp = 0.01; to = tp = 0.0
    for k in range(2):
        dto, to = o.step(dto, to)
        dtp, tp = py.step(dtp, tp)
        assert (dto, to) == (dtp, tp), k
        assert o.sizes() == (len(py.bodies), len(py.sinks)), k
        if len(py.bodies) < 2:
            break
        bo, so = o.download()
        for name, attr, kk in (("x", "position", 0), ("y", "position", 1), ("z", "position", 2), ("vx", "velocity", 0), ("vy", "velocity", 1),
                               ("vz", "velocity", 2), ("u", "internal_energy", None), ("alpha", "alpha", None), ("h", "s_length", None)):
            if name == "h" and not py.variable:
                continue
            assert np.array_equal(getattr(bo, name), col(py.bodies, attr, kk), equal_nan=True), (k, name)
        for name, attr, kk in (("x", "position", 0), ("vy", "velocity", 1), ("m", "mass", None), ("radius", "radius", None)):
            assert np.array_equal(getattr(so, name), col(py.sinks, attr, kk), equal_nan=True), (k, "sink " + name)
