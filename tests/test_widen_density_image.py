"""Column-density image (`sph_column_density`, SURVEY.md §8(f) item 3: the Density_Image.py counterpart).

The checker is a dense numpy statement of the same definition — Sigma = sum_j m_j F(|d|/h_j)/(pi h_j^2) with the
line-of-sight integral F of the M4 shape taken by scipy quadrature — evaluated at every pixel centre.  CPU tests
pin the checker (normalisation, agreement with a brute z-sum like Density_Image.py:120-145) and the PGM writer;
GPU tests compare the CUDA image with it: 1e-9 of the image maximum when the checker uses the engine's 1024-sample
table definition, 2e-6 against the exact integral (the table's linear-interpolation error)."""
import numpy as np
import pytest
from scipy.integrate import quad

from summersph_b200 import default_params, MODE_FIXED_H, MODE_VARIABLE_H, ics
from summersph_b200.density_image import to_gray, save_pgm

IMG_TABLE = 1024


def w_shape(q):
    q = np.asarray(q, dtype=float)
    return np.where(q <= 1.0, 1.0 - 1.5 * q ** 2 + 0.75 * q ** 3, np.where(q <= 2.0, 0.25 * (2.0 - q) ** 3, 0.0))


def F_exact(qb):
    if qb >= 2.0:
        return 0.0
    smax = np.sqrt(4.0 - qb * qb)
    pts = [np.sqrt(1.0 - qb * qb)] if qb < 1.0 else None
    return 2.0 * quad(lambda s: float(w_shape(np.sqrt(qb * qb + s * s))), 0.0, smax, points=pts, epsabs=1e-13, epsrel=1e-13)[0]


_TABLE = None


def F_table(q):
    """The engine's definition: F at IMG_TABLE + 1 nodes on [0, 2], linear interpolation, zero from q = 2."""
    global _TABLE
    if _TABLE is None:
        _TABLE = np.array([F_exact(2.0 * i / IMG_TABLE) for i in range(IMG_TABLE)] + [0.0, 0.0])
    x = np.asarray(q, dtype=float) * (IMG_TABLE / 2.0)
    j = np.minimum(x.astype(int), IMG_TABLE - 1)
    f = x - j
    return np.where(np.asarray(q) < 2.0, (1.0 - f) * _TABLE[j] + f * _TABLE[j + 1], 0.0)


_FINE = None


def F_fine(q, nodes=16384):
    """F by quadrature on a 16x finer grid than the engine's table (interpolation error ~4e-9 instead of ~1e-6):
    stands in for the exact integral where millions of pixel-particle pairs are needed."""
    global _FINE
    if _FINE is None:
        _FINE = np.array([F_exact(2.0 * i / nodes) for i in range(nodes)] + [0.0, 0.0])
    x = np.asarray(q, dtype=float) * (nodes / 2.0)
    j = np.minimum(x.astype(int), nodes - 1)
    f = x - j
    return np.where(np.asarray(q) < 2.0, (1.0 - f) * _FINE[j] + f * _FINE[j + 1], 0.0)


def image_ref(a, b, m, h, extent, shape, F=F_table):
    """Dense numpy image: every particle against every pixel centre (small cases only)."""
    u0, u1, v0, v1 = extent; nv, nu = shape
    du, dv = (u1 - u0) / nu, (v1 - v0) / nv
    uc = u0 + (np.arange(nu) + 0.5) * du; vc = v0 + (np.arange(nv) + 0.5) * dv
    h = np.maximum(h, 0.5 * max(du, dv))
    img = np.zeros((nv, nu))
    for j in range(len(a)):
        d = np.sqrt((uc[None, :] - a[j]) ** 2 + (vc[:, None] - b[j]) ** 2)
        img += m[j] * F(d / h[j]) / (np.pi * h[j] ** 2)
    return img


def test_fine_table_is_the_exact_integral():
    q = np.random.default_rng(1).uniform(0.0, 2.0, 200)
    assert np.max(np.abs(F_fine(q) - np.array([F_exact(v) for v in q]))) < 2e-8
    assert np.max(np.abs(F_table(q) - np.array([F_exact(v) for v in q]))) < 2e-6      # the engine's 1024 samples


def test_projected_kernel_is_normalised():
    """int F(q) 2 pi q dq = int w d^3q = pi (the M4 kernel integrates to one, F:125)."""
    q = np.linspace(0.0, 2.0, 20001)
    val = np.trapezoid(F_table(q) * 2.0 * np.pi * q, q)
    assert val == pytest.approx(np.pi, rel=1e-6)
    assert F_exact(0.0) == pytest.approx(2.0 * (1.0 - 0.5 + 0.1875) + 2.0 * 0.25 * 0.25, rel=1e-12)   # 2 int_0^2 w(s) ds = 1.5


def test_checker_matches_a_grid_sum_along_z():
    """Density_Image.py sums rho on a z grid (:120-145); with a fine grid, sum * dz tends to the column density."""
    rng = np.random.default_rng(3)
    n = 40
    x, y, z = rng.uniform(-3, 3, n), rng.uniform(-3, 3, n), rng.uniform(-1, 1, n)
    m = rng.uniform(0.5, 1.5, n); h = rng.uniform(0.8, 1.6, n)
    extent, shape = (-4.0, 4.0, -4.0, 4.0), (16, 16)
    ref = image_ref(x, y, m, h, extent, shape, F=np.vectorize(F_exact))
    zi = np.linspace(-5.0, 5.0, 2001); dz = zi[1] - zi[0]
    uc = -4.0 + (np.arange(16) + 0.5) * 0.5
    img = np.zeros((16, 16))
    for j in range(n):
        r = np.sqrt((uc[None, :, None] - x[j]) ** 2 + (uc[:, None, None] - y[j]) ** 2 + (zi[None, None, :] - z[j]) ** 2)
        img += m[j] * np.sum(w_shape(r / h[j]), axis=2) / (np.pi * h[j] ** 3) * dz
    assert np.max(np.abs(img - ref)) < 1e-5 * ref.max()


def test_gray_levels_and_pgm(tmp_path):
    img = np.array([[0.0, 1e-4, 1e-2], [1.0, np.nan, 1e-9]])
    g = to_gray(img)
    assert g.dtype == np.uint8 and g[1, 0] == 255 and g[0, 0] == 0 and g[0, 1] == 0 and g[0, 2] == 128 and g[1, 2] == 0 and g[1, 1] == 0
    assert np.array_equal(to_gray(np.zeros((2, 2))), np.zeros((2, 2), np.uint8))
    f = tmp_path / "a.pgm"
    save_pgm(str(f), img)
    raw = f.read_bytes()
    assert raw.startswith(b"P5\n3 2\n255\n") and raw[-6:] == bytes(g[::-1].ravel())


# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def E(built_engine):
    from summersph_b200.engine import Engine
    return Engine


AXES = {"x": ("y", "z"), "y": ("z", "x"), "z": ("x", "y")}


@pytest.mark.gpu
@pytest.mark.parametrize("axis", ["z", "x", "y"])
def test_gpu_image_matches_checker(axis, E):
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(1500, seed=12)
    extent, shape = (-110.0, 90.0, -60.0, 120.0), (48, 40)         # off-centre frame, non-square pixels, clipped disc
    with E(p) as e:
        e.upload(b, s)
        img = e.column_density(axis, extent, shape)
        e.evaluate()                                               # the state is re-ordered by the tree build ...
        img2 = e.column_density(axis, extent, shape)               # ... the image does not care
    ua, va = AXES[axis]
    ref = image_ref(getattr(b, ua), getattr(b, va), b.m, b.h, extent, shape)
    assert img.shape == shape
    assert np.max(np.abs(img - ref)) < 1e-9 * ref.max()
    assert np.max(np.abs(img2 - img)) < 1e-12 * ref.max()         # atomics: summation order only
    exact = image_ref(getattr(b, ua), getattr(b, va), b.m, b.h, extent, shape, F=F_fine)
    assert np.max(np.abs(img - exact)) < 2e-6 * exact.max()


@pytest.mark.gpu
def test_gpu_image_conserves_mass_and_widens_subpixel_particles(E):
    """Fixed-h mode uses `smoothing` for every particle; h below half a pixel is widened so no particle is lost."""
    b, s = ics.keplerian_disc(4000, seed=2)
    for h_fixed, pixels in ((2.5, 256), (0.05, 64)):
        p = default_params(MODE_FIXED_H, h_fixed=h_fixed)
        with E(p) as e:
            e.upload(b, s)
            img = e.column_density("z", (-120.0, 120.0, -120.0, 120.0), (pixels, pixels))
        px = (240.0 / pixels) ** 2
        assert img.sum() * px == pytest.approx(b.m.sum(), rel=2e-2 if h_fixed < 1 else 2e-3)
        assert img.min() >= 0.0 and np.isfinite(img).all()
    ref = image_ref(b.x, b.y, b.m, np.full(len(b), 0.05), (-120.0, 120.0, -120.0, 120.0), (64, 64))
    assert np.max(np.abs(img - ref)) < 1e-9 * ref.max()


@pytest.mark.gpu
def test_gpu_image_frame_without_particles_and_bad_arguments(E):
    from summersph_b200.engine import SphError
    p = default_params(MODE_VARIABLE_H)
    b, s = ics.keplerian_disc(500, seed=1)
    with E(p) as e:
        e.upload(b, s)
        assert not e.column_density("z", (1000.0, 1100.0, 1000.0, 1100.0), (8, 8)).any()
        with pytest.raises(SphError):
            e.column_density("z", (1.0, -1.0, 0.0, 1.0), (8, 8))
        with pytest.raises(SphError):
            e.column_density(3, (-1.0, 1.0, -1.0, 1.0), (8, 8))
