"""Analytic solutions the runs are compared with (BASELINE.json configs[1]: "Sod shock tube ... compared against
the analytic solution and the reference").  Host-side analysis only: nothing here is on the step path.

The reference ships one un-numbered Sod plot (README.md:16-18) and no numbers; the comparison defined here is
this build's own: the exact solution of the Riemann problem for the ideal-gas Euler equations (Toro, "Riemann
Solvers and Numerical Methods for Fluid Dynamics", ch. 4) sampled at the particle positions, and the L1 error
of the SPH density / velocity / pressure over the particles of the tube's undisturbed core.
"""
import numpy as np


def _f_side(p, rho_k, p_k, gamma):
    """Toro eq. 4.6-4.7: velocity jump function of one side and its derivative."""
    c_k = np.sqrt(gamma * p_k / rho_k)
    if p > p_k:                                   # shock
        a = 2.0 / ((gamma + 1.0) * rho_k); b = (gamma - 1.0) / (gamma + 1.0) * p_k
        s = np.sqrt(a / (p + b))
        return (p - p_k) * s, s * (1.0 - 0.5 * (p - p_k) / (p + b))
    e = (gamma - 1.0) / (2.0 * gamma)             # rarefaction
    return 2.0 * c_k / (gamma - 1.0) * ((p / p_k) ** e - 1.0), (p / p_k) ** (-(gamma + 1.0) / (2.0 * gamma)) / (rho_k * c_k)


def riemann_star(rho_l, v_l, p_l, rho_r, v_r, p_r, gamma=1.4):
    """Pressure and velocity between the two nonlinear waves (Newton-Raphson on Toro eq. 4.5)."""
    p = max(1e-12, 0.5 * (p_l + p_r))
    for _ in range(100):
        fl, dl = _f_side(p, rho_l, p_l, gamma); fr, dr = _f_side(p, rho_r, p_r, gamma)
        p_new = max(1e-14, p - (fl + fr + (v_r - v_l)) / (dl + dr))
        if abs(p_new - p) <= 1e-15 * (p_new + p):
            p = p_new
            break
        p = p_new
    fl, _ = _f_side(p, rho_l, p_l, gamma); fr, _ = _f_side(p, rho_r, p_r, gamma)
    return p, 0.5 * (v_l + v_r) + 0.5 * (fr - fl)


def riemann_exact(xi, rho_l, v_l, p_l, rho_r, v_r, p_r, gamma=1.4):
    """Exact (rho, v, P) of the Riemann problem at the similarity coordinates xi = (x - x0) / t (Toro §4.5)."""
    xi = np.asarray(xi, dtype=float)
    g = gamma; gm, gp = g - 1.0, g + 1.0
    ps, vs = riemann_star(rho_l, v_l, p_l, rho_r, v_r, p_r, g)
    c_l, c_r = np.sqrt(g * p_l / rho_l), np.sqrt(g * p_r / rho_r)
    rho = np.empty_like(xi); v = np.empty_like(xi); p = np.empty_like(xi)
    left = xi <= vs
    # ---- left of the contact
    if ps > p_l:                                                     # left shock
        rs = rho_l * ((ps / p_l + gm / gp) / (gm / gp * ps / p_l + 1.0))
        s = v_l - c_l * np.sqrt(gp / (2 * g) * ps / p_l + gm / (2 * g))
        pre = left & (xi < s)
        rho[left], v[left], p[left] = rs, vs, ps
        rho[pre], v[pre], p[pre] = rho_l, v_l, p_l
    else:                                                            # left rarefaction
        rs = rho_l * (ps / p_l) ** (1.0 / g); cs = c_l * (ps / p_l) ** (gm / (2 * g))
        head, tail = v_l - c_l, vs - cs
        rho[left], v[left], p[left] = rs, vs, ps
        pre = left & (xi < head)
        rho[pre], v[pre], p[pre] = rho_l, v_l, p_l
        fan = left & (xi >= head) & (xi <= tail)
        c = 2.0 / gp * (c_l + 0.5 * gm * (v_l - xi[fan]))
        v[fan] = 2.0 / gp * (c_l + 0.5 * gm * v_l + xi[fan])
        rho[fan] = rho_l * (c / c_l) ** (2.0 / gm); p[fan] = p_l * (c / c_l) ** (2.0 * g / gm)
    # ---- right of the contact
    right = ~left
    if ps > p_r:                                                     # right shock
        rs = rho_r * ((ps / p_r + gm / gp) / (gm / gp * ps / p_r + 1.0))
        s = v_r + c_r * np.sqrt(gp / (2 * g) * ps / p_r + gm / (2 * g))
        rho[right], v[right], p[right] = rs, vs, ps
        post = right & (xi > s)
        rho[post], v[post], p[post] = rho_r, v_r, p_r
    else:                                                            # right rarefaction
        rs = rho_r * (ps / p_r) ** (1.0 / g); cs = c_r * (ps / p_r) ** (gm / (2 * g))
        head, tail = v_r + c_r, vs + cs
        rho[right], v[right], p[right] = rs, vs, ps
        post = right & (xi > head)
        rho[post], v[post], p[post] = rho_r, v_r, p_r
        fan = right & (xi <= head) & (xi >= tail)
        c = 2.0 / gp * (c_r - 0.5 * gm * (v_r - xi[fan]))
        v[fan] = 2.0 / gp * (-c_r + 0.5 * gm * v_r + xi[fan])
        rho[fan] = rho_r * (c / c_r) ** (2.0 / gm); p[fan] = p_r * (c / c_r) ** (2.0 * g / gm)
    return rho, v, p


SOD = dict(rho_l=1.0, v_l=0.0, p_l=1.0, rho_r=0.125, v_r=0.0, p_r=0.1)


def sod_exact(x, t, gamma=1.4, rho_scale=1.0):
    """Sod's problem (rho 1 | 0.125, P 1 | 0.1, membrane at x = 0) at time t; densities and pressures times
    `rho_scale` (the Euler equations are invariant under a common scale of rho and P)."""
    rho, v, p = riemann_exact(np.asarray(x, dtype=float) / t, gamma=gamma, **SOD)
    return rho * rho_scale, v, p * rho_scale


# Omega of the variable-h program on a uniform cubic lattice with h = 1.2 spacings ("SUMMER_SPH - Variable.f90":455,487
# as coded: 1 + h/(3 rho) sum m (3 W - r dW/dr)/h; tests/test_widen_shock_tube.py measures it with the oracle).
OMEGA_LATTICE = 2.98


def effective_gas(gamma=1.4, omega=1.0):
    """The gas the variable-h program actually evolves.  Its momentum and energy equations both carry P/Omega
    (V:413-425) and its Omega is ~3 where the textbook grad-h factor is ~1 (V:487 has -(r dW - 3W)/h where
    dW/dh = -(3W + r dW/dr)/h; SURVEY.md §8(a) #9 keeps it as coded).  With a uniform Omega the particles follow
    an ideal gas of pressure P/Omega and adiabatic index 1 + (gamma-1)/Omega at unchanged rho and u.
    Returns (gamma_eff, Sod states with the effective pressures); omega = 1 is the fixed-h program / the textbook."""
    st = dict(SOD)
    st["p_l"] = SOD["p_l"] / omega; st["p_r"] = SOD["p_r"] / omega
    return 1.0 + (gamma - 1.0) / omega, st


def sod_core_mask(x, y, z, t, geom, gamma=1.4):
    """Particles whose history is still one-dimensional at time t in a tube with free (vacuum) boundaries:
    the lateral rarefactions move in from the side walls at the sound speed of the left state, the end
    rarefactions from x = -len_l and x = +len_r; one kernel radius (2 h_right) is kept clear on top, which is
    what `ics.sod_box` sizes the tube for (wave speeds of the effective gas for geom['omega'])."""
    g_eff, st = effective_gas(gamma, geom.get("omega", 1.0))
    c_l = np.sqrt(g_eff * st["p_l"] / st["rho_l"]); c_r = np.sqrt(g_eff * st["p_r"] / st["rho_r"])
    pad = 2.0 * geom["h_right"]
    half = 0.5 * geom["width"] - c_l * t - pad
    if half <= 0:
        raise ValueError("the lateral rarefactions have reached the axis: tube too narrow for this time")
    yc = zc = 0.5 * geom["width"]
    return ((np.abs(y - yc) < half) & (np.abs(z - zc) < half) &
            (x > -geom["len_l"] + c_l * t + pad) & (x < geom["len_r"] - c_r * t - pad))
