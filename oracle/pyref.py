"""pyref — a second, literal restatement of the reference's loop body in pure Python.  TEST INFRASTRUCTURE.

Why it exists: the reference ships no tests or golden vectors and there is no Fortran compiler in the image, so the C++
oracle (oracle/sph_oracle.cpp) cannot be run against the reference itself ("parity unpinned").  This file is an independent
second reading of the same Fortran, written routine by routine from the source with the reference's own data model —
AoS particle records, a pointer octree whose nodes own *copies* of their particles, recursive walks — instead of the
oracle's index ranges and snapshots.  tests/test_oracle_pyref.py holds the two restatements to each other, bit for bit, on
small cases; a misreading of the Fortran would have to be made twice, in two different formulations, to go unnoticed.
Pure-Python loops: a few hundred particles only.  Only tests/ may import it.

Citations: F = /root/reference/SUMMER_SPH.f90, V = "/root/reference/SUMMER_SPH - Variable.f90".
Arithmetic kept from the source: FP64 Python floats (no fused multiply-add), `x**n` expanded the way libgcc's __powidf2
does it, default-real literals (0.01, 0.15, 0.1, 0.0001, pi and G in V) rounded through single precision, left-to-right
array sums starting from zero (gfortran's inline SUM / DOT_PRODUCT), sequential loops (the OpenMP directives are ignored:
the serial order is the parity order).
"""
import math
import struct
import sys

sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))   # build_tree recurses to max_depth (1000) when particles coincide or are NaN


def r4(x):
    """A default-real (single precision) literal as the double the compiler promotes it to."""
    return struct.unpack("f", struct.pack("f", x))[0]


def ipow(x, n):
    """x**n for integer n >= 1 in libgcc __powidf2 order (gfortran without -ffast-math)."""
    y = x if (n % 2) else 1.0
    n >>= 1
    while n:
        x = x * x
        if n % 2:
            y = y * x
        n >>= 1
    return y


def div(a, b):
    """IEEE-754 division (Fortran gives Inf / NaN where Python raises): only needed where a density can be zero,
    i.e. for particles inside a depth-limited multi-particle node, which the density walk skips (F:431,443)."""
    if b != 0.0:
        return a / b
    if a != a or a == 0.0:
        return math.nan
    return math.copysign(math.inf, a) * math.copysign(1.0, b)


def fmax(a, b):
    """Fortran MAX on reals as gfortran compiles it (a NaN operand yields the other one, like C fmax)."""
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def maxval(v):
    """gfortran MAXVAL / MINVAL on reals skip NaNs (NaN only when every element is one)."""
    r = None
    for a in v:
        if a == a and (r is None or a > r):
            r = a
    return math.nan if r is None else r


def minval(v):
    r = None
    for a in v:
        if a == a and (r is None or a < r):
            r = a
    return math.nan if r is None else r


def vsum(v):
    s = 0.0
    for a in v:
        s = s + a
    return s


class Particle:                                   # F:14-27 | V:14-29
    __slots__ = ("number", "mass", "density", "internal_energy", "pressure", "sound_speed", "internal_energy_rate",
                 "alpha", "alpha_rate", "s_length", "omega", "position", "velocity", "acceleration")

    def copy(self):
        q = Particle()
        for k in Particle.__slots__:
            v = getattr(self, k)
            setattr(q, k, list(v) if isinstance(v, list) else v)
        return q


class Sink:                                       # F:30-37
    __slots__ = ("mass", "radius", "spin", "position", "velocity", "acceleration")


class Branch:                                     # F:40-49 | V:42-52
    __slots__ = ("center", "size", "n_particles", "particles", "mass_total", "max_len", "mass_center", "children")

    def __init__(self):
        self.children = None                      # allocated(node%children) == (children is not None)
        self.particles = None
        self.n_particles = 0


class Program:
    """Module-level state of one of the two programs: constants, kernel tables, bodies(:), sinks(:)."""

    def __init__(self, variable, max_depth=1000, bounding_size=1500.0, gamma=1.4, eta=1.2, convergence_criteria=1e-3,
                 max_length=50.0, timestep_scale=0.25, sink_radius=None, soft_uses_hi=False):
        self.variable = variable
        self.soft_uses_hi = soft_uses_hi                          # "SUMMER_SPH - Variable (test new)).f90":298
        self.smoothing = 2.5                                      # F:11 | V:11
        self.nq = 2500 if variable else 5000                      # V:8 | F:8
        self.dq = 2.0 / self.nq                                   # F:10
        self.G = r4(39.47841760435743)                            # F:7 | V:7 (default-real literal)
        self.pi = r4(3.1415926535897932) if variable else 3.14159265359   # V:7 | F:125
        self.max_depth, self.bounding_size = max_depth, bounding_size
        self.gamma, self.eta, self.convergence_criteria = gamma, eta, convergence_criteria
        self.max_length, self.timestep_scale = max_length, timestep_scale
        self.sink_radius = sink_radius if sink_radius is not None else (5.0 if variable else 3.5)   # V:830 | F:694
        self.bodies, self.sinks, self.root = [], [], None
        self._ngb_open = False
        # interaction counts of the most recent evaluation (not in the reference: what include/sph_b200.h's sph_counts reports)
        self.record_ngb, self.ngb = False, None               # per particle: numbers (0-based) of the leaves whose box test passes (F:443 | V:479)
        self.cnt = {"density_candidates": 0, "density_contributing": 0, "sph_pairs": 0, "grav_opened": 0, "grav_accepted": 0}
        self.init_kernel_table(); self.init_grav_kernel_table()

    # ---- F:55-101 | V:69-115 ----------------------------------------------------------------------------
    def init_kernel_table(self):
        self.w_table, self.dw_table = [0.0] * (self.nq + 1), [0.0] * (self.nq + 1)
        for i in range(self.nq + 1):
            q = i * self.dq
            if 0.0 <= q <= 1.0:
                self.w_table[i] = 1.0 - 1.5 * ipow(q, 2) + 0.75 * ipow(q, 3)
                self.dw_table[i] = -3.0 * q + 2.25 * ipow(q, 2)
            elif 1.0 < q <= 2.0:
                self.w_table[i] = 0.25 * ipow(2.0 - q, 3)
                self.dw_table[i] = -0.75 * ipow(2.0 - q, 2)

    def init_grav_kernel_table(self):
        self.grav_table = [0.0] * (self.nq + 1)
        for i in range(self.nq + 1):
            q = i * self.dq
            if 0.0 <= q <= 1.0:
                self.grav_table[i] = ((40.0 * ipow(q, 3)) - (36.0 * ipow(q, 5)) + (15.0 * ipow(q, 6))) / 30.0
            elif 1.0 < q <= 2.0:
                self.grav_table[i] = ((80.0 * ipow(q, 3)) - (90.0 * ipow(q, 4)) + (36.0 * ipow(q, 5)) - (5 * ipow(q, 6)) - 2) / 30.0
            else:
                self.grav_table[i] = 1.0

    # ---- F:105-146 | V:119-160 --------------------------------------------------------------------------
    def lookup_kernel(self, r, hi):
        qi = r / hi
        if 0.0 <= qi <= 2.0:
            i = min(int(qi / self.dq), self.nq - 1)
            alpha = (qi - i * self.dq) / self.dq
            Wi = (1.0 - alpha) * self.w_table[i] + alpha * self.w_table[i + 1]
            dWi = (1.0 - alpha) * self.dw_table[i] + alpha * self.dw_table[i + 1]
        else:
            Wi, dWi = 0.0, 0.0
        hn = hi if self.variable else self.smoothing              # F:125-126 normalises with the module constant
        return Wi / (self.pi * ipow(hn, 3)), dWi / (self.pi * ipow(hn, 4))

    def lookup_grav_kernel(self, r, hi):
        qi = r / hi
        if 0.0 <= qi <= 2.0:
            i = min(int(qi / self.dq), self.nq - 1)
            alpha = (qi - i * self.dq) / self.dq
            return (1.0 - alpha) * self.grav_table[i] + alpha * self.grav_table[i + 1]
        return 1.0

    # ---- F:149-246 | V:163-267 --------------------------------------------------------------------------
    def build_tree(self, node, depth, max_particles):
        node.mass_total = 0.0
        node.mass_center = [0.0, 0.0, 0.0]
        lens = []
        for p in node.particles:
            node.mass_total = node.mass_total + p.mass
            node.mass_center = [node.mass_center[k] + p.mass * p.position[k] for k in range(3)]
            lens.append(p.s_length)
        node.max_len = maxval(lens)                               # V:192
        if node.mass_total > 0.0:
            node.mass_center = [c / node.mass_total for c in node.mass_center]
        else:
            node.mass_center = list(node.center)
        if len(node.particles) <= max_particles or depth == 0:
            return
        node.children = [Branch() for _ in range(8)]
        for i in range(1, 9):
            ch = node.children[i - 1]
            ch.size = node.size * 0.5
            ch.n_particles = 0
            offset = [0.25 * node.size if ((i - 1) >> b) & 1 else -0.25 * node.size for b in range(3)]
            ch.center = [node.center[k] + offset[k] for k in range(3)]
        which_child, counts = [], [0] * 9
        for p in node.particles:
            child_index = 1
            for j in range(1, 4):
                if p.position[j - 1] > node.center[j - 1]:
                    child_index = ((child_index - 1) | (1 << (j - 1))) + 1
            which_child.append(child_index)
            counts[child_index] += 1
        for i in range(1, 9):
            node.children[i - 1].n_particles = counts[i]
            if counts[i] > 0:
                node.children[i - 1].particles = []
        for p, ci in zip(node.particles, which_child):
            node.children[ci - 1].particles.append(p.copy())     # node%children(..)%particles(..) = node%particles(p): a copy
        for i in range(1, 9):
            if node.children[i - 1].n_particles > 0:
                self.build_tree(node.children[i - 1], depth - 1, max_particles)

    def create_tree(self):                                        # F:795-816 | V:999-1020
        root = Branch()
        mx = [maxval(b.position[k] for b in self.bodies) for k in range(3)]
        mn = [minval(b.position[k] for b in self.bodies) for k in range(3)]
        root.center = [(mx[k] + mn[k]) / 2.0 for k in range(3)]
        root.size = maxval(mx[k] - mn[k] for k in range(3))
        root.n_particles = len(self.bodies)
        root.particles = [b.copy() for b in self.bodies]
        self.build_tree(root, self.max_depth, 1)
        self.root = root

    # ---- gravity F:264-290 | V:285-311 ------------------------------------------------------------------
    def particle_gravforce_one(self, node, p, theta):
        direction = [p.position[k] - node.mass_center[k] for k in range(3)]
        d2 = vsum([d * d for d in direction]) + (0.001 * p.s_length if self.soft_uses_hi else 0.001 * self.smoothing)   # T:298 | F:275, V:296
        dist = math.sqrt(d2)
        if (node.size / dist) < theta or node.children is None:
            self.cnt["grav_accepted"] += 1
            if node.mass_total > 0.0 and dist > 0.0:
                W = self.lookup_grav_kernel(dist, p.s_length if self.variable else self.smoothing)   # V:301 | F:280
                d3 = ipow(dist, 3)
                p.acceleration = [p.acceleration[k] - (self.G * node.mass_total * W * direction[k] / d3) for k in range(3)]
        else:
            self.cnt["grav_opened"] += 1
            for ch in node.children:
                if ch.n_particles > 0:
                    self.particle_gravforce_one(ch, p, theta)

    def sink_gravforces(self):                                    # F:559-591 | V:691-726
        for s in self.sinks:
            for b in self.bodies:
                vect_dr = [b.position[k] - s.position[k] for k in range(3)]
                dr = math.sqrt(vsum([v * v for v in vect_dr]))
                w = [self.G * v / (dr * dr * dr) for v in vect_dr]
                s.acceleration = [s.acceleration[k] + (b.mass * w[k]) for k in range(3)]
                b.acceleration = [b.acceleration[k] - (s.mass * w[k]) for k in range(3)]
        if len(self.sinks) < 2:
            return
        for i in range(len(self.sinks)):
            for j in range(i):
                si, sj = self.sinks[i], self.sinks[j]
                vect_dr = [sj.position[k] - si.position[k] for k in range(3)]
                dr = math.sqrt(vsum([v * v for v in vect_dr]))
                w = [self.G * v / (dr * dr * dr) for v in vect_dr]
                si.acceleration = [si.acceleration[k] + (sj.mass * w[k]) for k in range(3)]
                sj.acceleration = [sj.acceleration[k] - (si.mass * w[k]) for k in range(3)]

    # ---- density F:398-457 | V:440-496 ------------------------------------------------------------------
    def reach(self, node):
        if node.n_particles == 0:
            return 0.0        # an empty child reached from get_SPH's loop over all eight (F:307): max_len is undefined, neither branch can fire
        return 2.0 * node.max_len if self.variable else 2.0 * self.smoothing      # V:471 | F:431

    def density_tree_search(self, node, body):
        over_dr = [body.position[k] - node.center[k] for k in range(3)]
        lim = self.reach(node) + node.size / 2.0
        inside = all(abs(o) < lim for o in over_dr)
        if node.n_particles > 1 and inside and node.children is not None:
            for ch in node.children:
                if ch.n_particles > 0:
                    self.density_tree_search(ch, body)
        elif node.n_particles == 1 and inside:
            other = node.particles[0]
            nr = [body.position[k] - other.position[k] for k in range(3)]
            dr = math.sqrt(vsum([a * a for a in nr]))
            hb = body.s_length if self.variable else self.smoothing
            Wj, dWj_mag = self.lookup_kernel(dr, hb)
            self.cnt["density_candidates"] += 1
            if self.record_ngb and self._ngb_open:
                self.ngb[body.number - 1].append(other.number - 1)
            if dr / hb <= 2.0:
                self.cnt["density_contributing"] += 1
            body.density = body.density + other.mass * Wj
            if self.variable:
                W_h = -(dr * dWj_mag - 3 * Wj) / body.s_length   # V:487
                body.omega = body.omega + other.mass * W_h

    def get_density(self):
        self._ngb_open = True
        if self.record_ngb:
            self.ngb = [[] for _ in self.bodies]
        for b in self.bodies:
            b.density = 0.0
            b.omega = 0.0
            self.density_tree_search(self.root, b)
            if self.variable:
                b.omega = 1.0 + div(b.s_length, 3 * b.density) * b.omega          # V:455

    def get_pressure_and_sound_speed(self):                       # F:459-468 | V:502-512
        gm1 = (self.gamma - 1.0) if self.variable else 0.4
        gam = self.gamma if self.variable else 1.4
        for b in self.bodies:
            b.pressure = gm1 * b.internal_energy * b.density
            b.sound_speed = math.sqrt(div(gam * b.pressure, b.density))

    # ---- SPH pairs F:295-395 | V:324-432 ----------------------------------------------------------------
    def SPH_tree_search(self, node, body):
        bodies = self.bodies
        over_dr = [body.position[k] - node.center[k] for k in range(3)]
        lim = self.reach(node) + node.size / 2.0
        inside = all(abs(o) < lim for o in over_dr)
        if node.n_particles > 1 and inside and node.children is not None:
            for ch in node.children:
                if ch.n_particles > 0:
                    self.SPH_tree_search(ch, body)
        elif node.n_particles == 1 and inside:
            leafp = node.particles[0]
            if leafp.number >= body.number:
                return
            other = bodies[leafp.number - 1]
            self.cnt["sph_pairs"] += 1
            nr = [body.position[k] - other.position[k] for k in range(3)]
            dr = math.sqrt(vsum([a * a for a in nr]))
            vij = [body.velocity[k] - other.velocity[k] for k in range(3)]
            vdotr = vsum([vij[k] * nr[k] for k in range(3)])
            if vdotr >= 0:
                vdotr = 0.0
            nr = [a / dr for a in nr]
            lit001 = r4(0.01)
            if self.variable:
                Wj, dWj_mag = self.lookup_kernel(dr, body.s_length)
                Wi, dWi_mag = self.lookup_kernel(dr, leafp.s_length)                  # the tree copy's h (V:396)
                dWj = [a * dWj_mag for a in nr]
                dWi = [a * dWi_mag for a in nr]
                vdotgradW = (vsum([dWj[k] * vij[k] for k in range(3)]) + vsum([dWi[k] * vij[k] for k in range(3)])) / 2
                avg_len = (body.s_length + leafp.s_length) / 2
                vis_nu = (avg_len * vdotr) / (dr * dr + lit001 * avg_len * avg_len)
            else:
                Wj, dWj_mag = self.lookup_kernel(dr, self.smoothing)
                dWj = [a * dWj_mag for a in nr]
                vdotgradW = vsum([dWj[k] * vij[k] for k in range(3)])
                vis_nu = (self.smoothing * vdotr) / (dr * dr + lit001 * self.smoothing * self.smoothing)
            avg_sound_speed = 0.5 * (body.sound_speed + other.sound_speed)
            avg_alpha = 0.5 * (body.alpha + other.alpha)
            viscous_cont = div(-avg_alpha * avg_sound_speed * vis_nu + 2 * avg_alpha * vis_nu * vis_nu, 0.5 * (body.density + other.density))
            if self.variable:
                pi_ = div(body.pressure, body.omega * body.density * body.density)
                pj_ = div(other.pressure, other.omega * other.density * other.density)
                acc_contrib = [pi_ * dWj[k] + pj_ * dWi[k] + viscous_cont * (dWi[k] + dWj[k]) / 2 for k in range(3)]
            else:
                pi_ = div(body.pressure, body.density * body.density)
                pj_ = div(other.pressure, other.density * other.density)
                acc_contrib = [(pi_ + pj_ + viscous_cont) * dWj[k] for k in range(3)]
            body.acceleration = [body.acceleration[k] - other.mass * acc_contrib[k] for k in range(3)]
            other.acceleration = [other.acceleration[k] + body.mass * acc_contrib[k] for k in range(3)]
            body.internal_energy_rate = body.internal_energy_rate + other.mass * vdotgradW * (pi_ + 0.5 * viscous_cont)
            other.internal_energy_rate = other.internal_energy_rate + body.mass * vdotgradW * (pj_ + 0.5 * viscous_cont)
            body.alpha_rate = body.alpha_rate + other.mass * vdotgradW
            other.alpha_rate = other.alpha_rate + body.mass * vdotgradW

    def get_SPH(self):
        for b in self.bodies:
            if self.root.children is not None:
                for ch in self.root.children:                     # all eight, empty ones return at once (F:307)
                    self.SPH_tree_search(ch, b)
        lit015, lit01 = r4(0.15), 0.1
        for b in self.bodies:
            hh = b.s_length if self.variable else self.smoothing
            b.alpha_rate = fmax(div(b.alpha_rate, b.density), 0.0) + lit015 * ((lit01 - b.alpha) * b.sound_speed / hh)   # F:317 | V:346

    def zero_rates(self):                                         # F:779-793
        for b in self.bodies:
            b.acceleration = [0.0, 0.0, 0.0]; b.internal_energy_rate = 0.0; b.alpha_rate = 0.0
        for s in self.sinks:
            s.acceleration = [0.0, 0.0, 0.0]

    def find_forces(self):                                        # F:818-829 | V:1022-1033
        self.zero_rates()
        for b in self.bodies:
            self.particle_gravforce_one(self.root, b, 0.5)
        self.sink_gravforces()
        self.get_SPH()

    # ---- integrator F:742-776, timestep F:831-860 | V:1035-1065 -----------------------------------------
    def kick(self, dt):
        for b in self.bodies:
            b.velocity = [b.velocity[k] + 0.5 * b.acceleration[k] * dt for k in range(3)]
        for s in self.sinks:
            s.velocity = [s.velocity[k] + 0.5 * s.acceleration[k] * dt for k in range(3)]
        for b in self.bodies:
            b.internal_energy = b.internal_energy + 0.5 * b.internal_energy_rate * dt
            b.alpha = b.alpha + b.alpha_rate * dt * 0.5

    def drift(self, dt):
        for b in self.bodies:
            b.position = [b.position[k] + b.velocity[k] * dt for k in range(3)]
        for s in self.sinks:
            s.position = [s.position[k] + s.velocity[k] * dt for k in range(3)]

    def get_next_timestep(self, dt):
        cands = []
        for b in self.bodies:
            hh = b.s_length if self.variable else self.smoothing
            vv = vsum([v * v for v in b.velocity]); aa = vsum([a * a for a in b.acceleration])
            cands.append(math.sqrt(vv / aa) if aa != 0.0 else math.inf)
            cands.append(b.internal_energy / abs(b.internal_energy_rate) if b.internal_energy_rate != 0.0 else math.inf)
            cands.append(hh / math.sqrt(vv) if vv != 0.0 else math.inf)
            cands.append(hh / (b.sound_speed + 1.2 * b.sound_speed))
        dt_candidate = minval(cands) * (self.timestep_scale if self.variable else 0.25)   # V:1056 | F:851
        if dt_candidate > 2 * dt and 1.5 * dt < r4(0.1):
            dt = 1.5 * dt
        elif dt_candidate < 0.5 * dt and dt * 0.5 > r4(0.0001):
            dt = 0.5 * dt
        return dt

    # ---- V:515-546 --------------------------------------------------------------------------------------
    def calc_smoothing(self):
        self._ngb_open = False                    # its re-walks are not part of the evaluation's neighbour sets
        eta = self.eta
        iterations = 0
        for b in self.bodies:
            old_len = b.s_length
            b.s_length = b.s_length * (1 + div(div(b.mass * ipow(eta / b.s_length, 3), b.density) - 1, 3 * b.omega))
            if b.s_length < self.max_length and b.s_length > r4(0.01):
                while ((b.s_length - old_len) / old_len) > self.convergence_criteria and b.s_length < 10.0:
                    old_len = b.s_length
                    b.density = 0.0
                    b.omega = 0.0
                    self.density_tree_search(self.root, b)
                    b.omega = 1.0 + div(b.s_length, 3 * b.density) * b.omega
                    b.s_length = b.s_length * (1 + div(div(b.mass * ipow(eta / b.s_length, 3), b.density) - 1, 3 * b.omega))
                    iterations += 1
            else:
                b.s_length = old_len
        return iterations

    # ---- V:549-597 --------------------------------------------------------------------------------------
    def check_sink_creation(self):
        for b in self.bodies:
            if b.mass * ipow(self.eta / b.s_length, 3) > 0.5:
                for s in self.sinks:
                    dr = math.sqrt(vsum([(s.position[k] - b.position[k]) * (s.position[k] - b.position[k]) for k in range(3)]))
                    if dr < s.radius + 2 * b.s_length:
                        return
                ns = Sink()
                ns.position = list(b.position); ns.velocity = list(b.velocity)
                ns.acceleration = [0.0, 0.0, 0.0]; ns.spin = [0.0, 0.0, 0.0]
                ns.mass = 0.00000000001; ns.radius = 2 * b.s_length
                self.sinks.append(ns)
                return

    # ---- accretion F:484-556 | V:616-688 ----------------------------------------------------------------
    def sink2gasdists(self, sink_i, node, mask):
        over_dr = [node.center[k] - sink_i.position[k] for k in range(3)]
        if node.n_particles > 1 and all(abs(o) < (sink_i.radius + node.size / 2.0) for o in over_dr) and node.children is not None:
            for ch in node.children:
                if ch.n_particles > 0:
                    self.sink2gasdists(sink_i, ch, mask)
            return
        if self.variable:
            if node.n_particles == 1 and all(abs(o) < (sink_i.radius + node.size / 2.0) for o in over_dr):   # V:668
                dr = vsum([math.sqrt((node.particles[0].position[k] - sink_i.position[k]) * (node.particles[0].position[k] - sink_i.position[k]))
                           for k in range(3)])                                                                 # V:669
                if dr < sink_i.radius:
                    mask[node.particles[0].number - 1] = False
        else:
            if node.n_particles == 1 and all(abs(o) < (2 * sink_i.radius + node.size / 2.0) for o in over_dr):   # F:536
                terms = []
                for k in range(3):
                    a = node.center[k] * node.center[k] - sink_i.position[k] * sink_i.position[k]
                    terms.append(math.sqrt(a) if a >= 0.0 else math.nan)                                       # F:537
                dr = vsum(terms)
                if dr < sink_i.radius:
                    mask[node.particles[0].number - 1] = False

    def initiate_sink_accretion(self):
        n = len(self.bodies)
        masks = []
        for s in self.sinks:
            mask = [True] * n
            self.sink2gasdists(s, self.root, mask)
            acc = [b for b, keep in zip(self.bodies, mask) if not keep]
            new_mass = s.mass + vsum([b.mass for b in acc])
            s.position = [div(s.mass * s.position[k] + vsum([b.mass * b.position[k] for b in acc]), new_mass) for k in range(3)]   # 0/0 for a massless sink that accretes nothing
            s.velocity = [div(s.mass * s.velocity[k] + vsum([b.mass * b.velocity[k] for b in acc]), new_mass) for k in range(3)]
            s.mass = s.mass + vsum([b.mass for b in acc])
            masks.append(mask)
        keep_all = [all(m[j] for m in masks) for j in range(n)]   # pack_sinks F:546-556
        self.bodies = [b for b, k in zip(self.bodies, keep_all) if k]

    def check_bounds(self):                                       # F:471-482 | V:599-614
        B = self.bounding_size
        self.bodies = [b for b in self.bodies if all(abs(x) <= B for x in b.position)]
        if self.variable:
            self.sinks = [s for s in self.sinks if all(abs(x) <= B for x in s.position)]

    # ---- one body of simulate's loop F:886-928 | V:1120-1162 --------------------------------------------
    def evaluate(self):
        for k in self.cnt:
            self.cnt[k] = 0
        self.create_tree()
        self.get_density()
        self.get_pressure_and_sound_speed()
        self.find_forces()

    def step(self, dt, t):
        for i, b in enumerate(self.bodies):
            b.number = i + 1                                      # F:886-888
        self.evaluate()
        self.kick(dt)
        self.drift(dt)
        self.evaluate()
        self.kick(dt)
        t = t + dt
        dt = self.get_next_timestep(dt)
        if self.variable:
            self.calc_smoothing()
            self.check_sink_creation()
        if any(s.mass > 0.0 for s in self.sinks):
            self.initiate_sink_accretion()
        self.check_bounds()
        return dt, t

    # ---- host-side hand-off (what read_data_from_file leaves in bodies(:) / sinks(:), F:594-716 | V:729-852) ----
    def load(self, bodies, sinks):
        """bodies / sinks: summersph_b200.state containers (SoA numpy)."""
        self.bodies, self.sinks = [], []
        for i in range(len(bodies)):
            p = Particle()
            p.number = i + 1
            p.mass = float(bodies.m[i]); p.internal_energy = float(bodies.u[i])
            p.alpha = float(bodies.alpha[i])
            p.s_length = float(bodies.h[i]) if self.variable else self.smoothing
            p.density = 0.0; p.pressure = 0.0; p.sound_speed = 0.0; p.internal_energy_rate = 0.0; p.alpha_rate = 0.0; p.omega = 1.0
            p.position = [float(bodies.x[i]), float(bodies.y[i]), float(bodies.z[i])]
            p.velocity = [float(bodies.vx[i]), float(bodies.vy[i]), float(bodies.vz[i])]
            p.acceleration = [0.0, 0.0, 0.0]
            self.bodies.append(p)
        for i in range(len(sinks)):
            s = Sink()
            s.mass = float(sinks.m[i])
            rad = float(sinks.radius[i])
            s.radius = self.sink_radius if rad != rad else rad
            s.spin = [0.0, 0.0, 0.0]
            s.position = [float(sinks.x[i]), float(sinks.y[i]), float(sinks.z[i])]
            s.velocity = [float(sinks.vx[i]), float(sinks.vy[i]), float(sinks.vz[i])]
            s.acceleration = [0.0, 0.0, 0.0]
            self.sinks.append(s)
        if not self.sinks:                                        # dummy sink F:698-707
            s = Sink()
            s.mass = 0.0; s.radius = 0.0; s.spin = [0.0] * 3
            s.position = [0.0] * 3; s.velocity = [0.0] * 3; s.acceleration = [0.0] * 3
            self.sinks.append(s)
