"""Independent cross-checks of the oracle's tree code: an O(N^2) numpy brute
force with the same table interpolation validates the geometry-independent fixed-h definitions, and an
independent derivation of leaf cells (sort by descent key, leaf level from the sorted neighbours'
common prefix) validates the recursive build and the variable-h neighbour criterion."""
import numpy as np
import pytest

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, ics, EVAL_TREE, EVAL_DENSITY, EVAL_SPH, EVAL_GRAVITY
from oracle.oracle import Oracle
from conftest import relerr

N = 1500


def np_lookup(tab, q, nq):
    dq = 2.0 / nq
    i = np.minimum((q / dq).astype(np.int64), nq - 1)
    a = (q - i * dq) / dq
    out = (1.0 - a) * tab[i] + a * tab[i + 1]
    return np.where(q <= 2.0, out, 0.0)


@pytest.fixture(scope="module")
def disc():
    return ics.keplerian_disc(N, seed=11)


def test_fixed_h_density_and_forces_vs_bruteforce(disc):
    b, s = disc
    b = b.copy(); b.alpha[:] = 0.0          # what the fixed-h reader does (F:681): no viscosity
    p = default_params(MODE_FIXED_H)
    o = Oracle(p)
    o.upload(b, s); o.evaluate(EVAL_TREE | EVAL_DENSITY | EVAL_SPH)
    d = o.diag()
    w, dw, _ = o.tables()
    h, pi = p.h_fixed, 3.14159265359
    X = np.stack([b.x, b.y, b.z], 1); V = np.stack([b.vx, b.vy, b.vz], 1)
    D = X[:, None, :] - X[None, :, :]
    r = np.sqrt(np.sum(D * D, -1))
    qq = np.minimum(r / h, 2.5)
    W = np_lookup(w, qq, p.nq) / (pi * h ** 3)
    dW = np_lookup(dw, qq, p.nq) / (pi * h ** 4)
    rho = W @ b.m
    assert relerr(d["rho"], rho) < 1e-12
    P = 0.4 * b.u * rho; c = np.sqrt(1.4 * P / rho)
    assert relerr(d["P"], P) < 1e-12 and relerr(d["c"], c) < 1e-12
    # pair forces, gather form (Appendix B)
    np.fill_diagonal(r, 1.0)
    nhat = D / r[:, :, None]
    np.fill_diagonal(dW, 0.0)
    por2 = P / rho ** 2
    scal = (por2[:, None] + por2[None, :]) * dW
    acc = -np.sum((b.m[None, :] * scal)[:, :, None] * nhat, 1)
    assert relerr(d["ax"], acc[:, 0]) < 1e-11 and relerr(d["ay"], acc[:, 1]) < 1e-11 and relerr(d["az"], acc[:, 2]) < 1e-11
    vij = V[:, None, :] - V[None, :, :]
    vdg = np.sum(nhat * vij, -1) * dW
    udot = np.sum(b.m[None, :] * vdg, 1) * por2
    assert relerr(d["udot"], udot) < 1e-11


def test_gravity_bounded_by_direct_sum(disc):
    b, s = disc
    p = default_params(MODE_FIXED_H)
    o = Oracle(p)
    o.upload(b, s); o.evaluate(EVAL_TREE | EVAL_GRAVITY)
    d = o.diag()
    _, _, g = o.tables()
    X = np.stack([b.x, b.y, b.z], 1)
    D = X[:, None, :] - X[None, :, :]
    d2 = np.sum(D * D, -1) + 0.001 * p.h_fixed
    dist = np.sqrt(d2)
    qq = dist / p.h_fixed
    dq = 2.0 / p.nq
    i = np.minimum((np.minimum(qq, 2.0) / dq).astype(np.int64), p.nq - 1)
    a = (np.minimum(qq, 2.0) - i * dq) / dq
    Wg = np.where(qq <= 2.0, (1 - a) * g[i] + a * g[i + 1], 1.0)
    f = o.G * b.m[None, :] * Wg / dist ** 3
    np.fill_diagonal(f, 0.0)
    acc = -np.sum(f[:, :, None] * D, 1)
    tree = np.stack([d["ax"], d["ay"], d["az"]], 1)
    err = np.linalg.norm(tree - acc, axis=1) / np.linalg.norm(acc, axis=1)
    assert np.median(err) < 0.02 and np.max(err) < 0.3      # theta = 0.5 monopole


def descent(X, centre, size, levels):
    """numpy replay of the centre descent (F:190-214): keys + per-level centres."""
    n = X.shape[0]
    c = np.tile(centre, (n, 1)); s = size
    keys = np.zeros(n, np.uint64)
    centres = [c.copy()]; sizes = [s]
    for _ in range(levels):
        bits = X > c
        dig = bits[:, 0].astype(np.uint64) | (bits[:, 1].astype(np.uint64) << np.uint64(1)) | (bits[:, 2].astype(np.uint64) << np.uint64(2))
        keys = (keys << np.uint64(3)) | dig
        c = c + np.where(bits, 0.25 * s, -0.25 * s)
        s = s * 0.5
        centres.append(c.copy()); sizes.append(s)
    return keys, centres, sizes


def test_leaf_cells_independent_derivation(disc):
    b, s = disc
    p = default_params(MODE_VARIABLE_H)
    o = Oracle(p)
    o.record_neighbours(True)
    o.upload(b, s); o.evaluate(EVAL_TREE | EVAL_DENSITY)
    t = o.tree()
    X = np.stack([b.x, b.y, b.z], 1)
    mn, mx = X.min(0), X.max(0)
    centre = (mx + mn) / 2.0; size = float(np.max(mx - mn))
    assert np.array_equal(centre, t["root_center"]) and size == t["root_size"]
    L = 21
    keys, centres, sizes = descent(X, centre, size, L)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(order.astype(np.int32), t["order"])
    ks = keys[order]
    x = ks[1:] ^ ks[:-1]
    assert np.all(x != 0)
    lcp = np.array([(63 - int(v).bit_length()) // 3 for v in x])
    lev_sorted = 1 + np.maximum(np.concatenate([[-1], lcp]), np.concatenate([lcp, [-1]]))
    lev = np.empty(N, np.int64); lev[order] = lev_sorted
    assert np.array_equal(lev, t["level"])
    cs = np.stack(centres, 0)                                # (L+1, n, 3)
    leafc = cs[lev, np.arange(N)]
    assert np.array_equal(leafc[:, 0], t["cx"]) and np.array_equal(leafc[:, 1], t["cy"]) and np.array_equal(leafc[:, 2], t["cz"])
    leafs = np.array(sizes)[lev]
    assert np.array_equal(leafs, t["size"])
    # variable-h candidate set: j is a candidate of i iff all |x_i - c_leaf(j)| < 2 h_j + size_j/2   (V:479)
    R = 2.0 * b.h + leafs / 2.0
    inbox = np.all(np.abs(X[:, None, :] - leafc[None, :, :]) < R[None, :, None], -1)     # [i, j]
    count, hsh, off, lst = o.neighbours()
    assert np.array_equal(inbox.sum(1).astype(np.int32), count)
    ii, jj = np.nonzero(inbox)
    assert np.array_equal(jj.astype(np.int32), lst)
    # density from that set with the gatherer's h (V:486)
    w, dw, _ = o.tables()
    pi = float(np.float32(3.1415926535897932))
    r = np.sqrt(np.sum((X[:, None, :] - X[None, :, :]) ** 2, -1))
    q = np.minimum(r / b.h[:, None], 2.5)
    W = np_lookup(w, q, p.nq) / (pi * b.h[:, None] ** 3)
    rho = np.sum(np.where(inbox, W * b.m[None, :], 0.0), 1)
    assert relerr(o.diag()["rho"], rho) < 1e-12
