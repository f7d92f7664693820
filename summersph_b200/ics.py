"""Seeded synthetic initial conditions; the reference ships only a broken sketch
(Disc_ICs.py) and none of the IC files it names, so the generators are this repo's own.

All generators return (Bodies, Sinks) in the reference's units: AU, M_sun, yr, G = 39.478416442871094
(the real(4) literal of SUMMER_SPH.f90:7 as compiled).
"""
import numpy as np
from .state import Bodies, Sinks

G_EFF = float(np.float32(39.47841760435743))   # SUMMER_SPH.f90:7


def keplerian_disc(n, seed=20251018, r_in=10.0, r_out=100.0, aspect=0.05, m_star=1.0, m_disc=0.01,
                   u=0.25, alpha=0.1, eta=1.2, with_sink=True):
    """Uniform-surface-density Keplerian disc + central sink (configs 1, 4, 5).
    u = 0.25 as Disc_ICs.py:26; column 10 (h) is eta*(m/rho)^(1/3) from the analytic density."""
    rng = np.random.default_rng(seed)
    r = np.sqrt(rng.uniform(r_in * r_in, r_out * r_out, n))
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    H = aspect * r
    zeta = np.clip(rng.standard_normal(n), -3.0, 3.0)
    z = zeta * H
    x = r * np.cos(phi); y = r * np.sin(phi)
    vphi = np.sqrt(G_EFF * m_star / r)
    vx = -vphi * np.sin(phi); vy = vphi * np.cos(phi); vz = np.zeros(n)
    m = np.full(n, m_disc / n)
    sigma = m_disc / (np.pi * (r_out ** 2 - r_in ** 2))
    rho = sigma / (np.sqrt(2.0 * np.pi) * H) * np.exp(-0.5 * zeta ** 2)
    h = eta * (m / rho) ** (1.0 / 3.0)
    bodies = Bodies(x, y, z, vx, vy, vz, np.full(n, u), m, np.full(n, alpha), h)
    if with_sink:
        sinks = Sinks(*[np.zeros(1) for _ in range(6)], np.array([m_star]), np.array([0.0]))
        sinks.radius[:] = np.nan      # filled in by the reader / caller with params.sink_radius
    else:
        sinks = Sinks.empty(0)
    return bodies, sinks


def thin_ring(n, seed=20251019, r0=50.0, sigma_r=2.5, sigma_z=0.5, m_star=1.0, m_ring=1e-3,
              aspect=0.02, alpha=0.1, eta=1.2):
    """Thin ring around a 1 M_sun sink (config 3); u chosen so that c/v_kepler ~ H/R = aspect."""
    rng = np.random.default_rng(seed)
    r = r0 + sigma_r * np.clip(rng.standard_normal(n), -4.0, 4.0)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    zeta = np.clip(rng.standard_normal(n), -4.0, 4.0)
    z = sigma_z * zeta
    x = r * np.cos(phi); y = r * np.sin(phi)
    vphi = np.sqrt(G_EFF * m_star / r)
    vx = -vphi * np.sin(phi); vy = vphi * np.cos(phi); vz = np.zeros(n)
    m = np.full(n, m_ring / n)
    gamma = 1.4
    cs2 = (aspect ** 2) * G_EFF * m_star / r0
    u = np.full(n, cs2 / (gamma * (gamma - 1.0)))
    rho = (m_ring / (2.0 * np.pi * r0)) / (2.0 * np.pi * sigma_r * sigma_z) * np.exp(-0.5 * ((r - r0) / sigma_r) ** 2 - 0.5 * zeta ** 2)
    h = np.minimum(eta * (m / rho) ** (1.0 / 3.0), 5.0)
    bodies = Bodies(x, y, z, vx, vy, vz, u, m, np.full(n, alpha), h)
    sinks = Sinks(*[np.zeros(1) for _ in range(6)], np.array([m_star]), np.array([np.nan]))
    return bodies, sinks


def sod_tube(n_target=100_000, rho_scale=1e-9, gamma=1.4, eta=1.2, alpha=1.0, width=8):
    """Sod shock tube (config 2): cubic lattices, 8x denser on the left (spacing ratio 1:2), equal-mass
    particles, rho_L:rho_R = 1:0.125, P_L:P_R = 1:0.1, v = 0, x in [-0.5, 0.5].  Densities are scaled by
    `rho_scale` so the always-on self gravity is negligible (Euler equations are density-scale invariant).
    No particle sits at the origin (the dummy sink would NaN it, SUMMER_SPH.f90:572).  No sinks."""
    # choose the right-hand spacing so that the total count is close to n_target
    # N = (0.5/dr)*w^2 + (0.5/dl)*(2w)^2 with dl = dr/2  ->  N = 0.5*w^2/dr * (1 + 8)
    w = width
    dr = 4.5 * w * w / n_target
    dl = dr / 2.0
    nxr = int(round(0.5 / dr)); nxl = int(round(0.5 / dl))
    dr = 0.5 / nxr; dl = dr / 2.0; nxl = 2 * nxr

    def lattice(nx, ny, d, x0):
        i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(ny), indexing="ij")
        return (x0 + (i.ravel() + 0.5) * d, (j.ravel() + 0.5) * d, (k.ravel() + 0.5) * d)

    xl, yl, zl = lattice(nxl, 2 * w, dl, -0.5)
    xr, yr, zr = lattice(nxr, w, dr, 0.0)
    x = np.concatenate([xl, xr]); y = np.concatenate([yl, yr]); z = np.concatenate([zl, zr])
    n = x.size
    rho_l, rho_r = 1.0 * rho_scale, 0.125 * rho_scale
    m = np.full(n, rho_l * dl ** 3)             # = rho_r * dr^3
    p_l, p_r = 1.0 * rho_scale, 0.1 * rho_scale
    u = np.concatenate([np.full(xl.size, p_l / ((gamma - 1.0) * rho_l)), np.full(xr.size, p_r / ((gamma - 1.0) * rho_r))])
    h = np.concatenate([np.full(xl.size, eta * dl), np.full(xr.size, eta * dr)])
    bodies = Bodies(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), u, m, np.full(n, alpha), h)
    return bodies, Sinks.empty(0)


def sod_box(n_target=100_000, t_end=0.2, rho_scale=1e-9, gamma=1.4, eta=1.2, alpha=1.0, core_cells=4, omega=1.0):
    """Sod shock tube sized for a comparison with the exact solution at time `t_end` (config 2).

    The reference has no periodic or wall boundaries, so a lattice tube expands into vacuum from every face.
    The box is therefore made just large enough that, at `t_end`, a core of `core_cells` right-hand lattice
    cells around the axis and the whole wave pattern (rarefaction head at -c_L t, shock at s t) have not
    been reached by the rarefactions coming in from the side walls (speed <= c_L) and the two ends, with one
    kernel radius (2 h_right) to spare.  Same states as `sod_tube`: rho 1 | 0.125, P 1 | 0.1 (times `rho_scale`,
    which makes the always-on self gravity negligible), equal-mass particles on cubic lattices of spacing
    dl | 2 dl, h = eta * spacing, no particle at the origin, no sinks.

    `omega`: the variable-h program divides its pressure terms by Omega ~ 3 (see analytic.effective_gas), which
    slows every wave; pass the lattice value (analytic.OMEGA_LATTICE) to size the box for those waves instead.
    Returns (Bodies, Sinks, geom) with geom = {len_l, len_r, width, dl, dr, h_left, h_right, t_end, omega}."""
    from .analytic import riemann_star, effective_gas, SOD
    g_eff, st = effective_gas(gamma, omega)
    c_l = np.sqrt(g_eff * st["p_l"] / st["rho_l"]); c_r = np.sqrt(g_eff * st["p_r"] / st["rho_r"])
    ps, _ = riemann_star(gamma=g_eff, **st)
    s_shock = c_r * np.sqrt((g_eff + 1) / (2 * g_eff) * ps / st["p_r"] + (g_eff - 1) / (2 * g_eff))

    def geometry(dl):
        dr = 2.0 * dl; pad = 2.0 * eta * dr
        nw = int(np.ceil((2.0 * (c_l * t_end + pad) + core_cells * dr) / dr))          # width in right-hand cells
        nl = int(np.ceil((2.0 * c_l * t_end + pad + dr) / dl))                          # left length in dl
        nr = int(np.ceil(((s_shock + c_r) * t_end + pad + dr) / dr))                    # right length in dr
        return nl, nr, nw, nl * (2 * nw) ** 2 + nr * nw ** 2

    lo, hi = 1e-4, 1.0                                # particle count falls monotonically with dl
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        if geometry(mid)[3] > n_target:
            lo = mid
        else:
            hi = mid
    dl = hi; dr = 2.0 * dl
    nl, nr, nw, _ = geometry(dl)

    def lattice(nx, ny, d, x0):
        i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(ny), indexing="ij")
        return (x0 + (i.ravel() + 0.5) * d, (j.ravel() + 0.5) * d, (k.ravel() + 0.5) * d)

    xl, yl, zl = lattice(nl, 2 * nw, dl, -nl * dl)
    xr, yr, zr = lattice(nr, nw, dr, 0.0)
    x = np.concatenate([xl, xr]); y = np.concatenate([yl, yr]); z = np.concatenate([zl, zr])
    n = x.size
    rho_l, p_l, p_r, rho_r = rho_scale, rho_scale, 0.1 * rho_scale, 0.125 * rho_scale
    m = np.full(n, rho_l * dl ** 3)
    u = np.concatenate([np.full(xl.size, p_l / ((gamma - 1.0) * rho_l)), np.full(xr.size, p_r / ((gamma - 1.0) * rho_r))])
    h = np.concatenate([np.full(xl.size, eta * dl), np.full(xr.size, eta * dr)])
    geom = {"len_l": nl * dl, "len_r": nr * dr, "width": nw * dr, "dl": dl, "dr": dr, "h_left": eta * dl,
            "h_right": eta * dr, "t_end": t_end, "omega": omega}
    bodies = Bodies(x, y, z, np.zeros(n), np.zeros(n), np.zeros(n), u, m, np.full(n, alpha), h)
    return bodies, Sinks.empty(0), geom


def uniform_sphere(n, seed=7, radius=100.0, m_total=5.0, u=0.25, alpha=0.1, eta=1.2):
    """The 'Collapse' geometry Disc_ICs.py sketches (uniform sphere R<=100 AU, v=0 here)."""
    rng = np.random.default_rng(seed)
    r = radius * rng.uniform(0, 1, n) ** (1.0 / 3.0)
    ct = rng.uniform(-1, 1, n); st = np.sqrt(1 - ct * ct); ph = rng.uniform(0, 2 * np.pi, n)
    x, y, z = r * st * np.cos(ph), r * st * np.sin(ph), r * ct
    m = np.full(n, m_total / n)
    rho = m_total / (4.0 / 3.0 * np.pi * radius ** 3)
    h = np.full(n, eta * (m[0] / rho) ** (1.0 / 3.0))
    zero = np.zeros(n)
    return Bodies(x, y, z, zero, zero, zero, np.full(n, u), m, np.full(n, alpha), h), Sinks.empty(0)
