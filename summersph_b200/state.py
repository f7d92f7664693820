"""Host-side particle containers (SoA numpy, FP64) mirroring the reference's `bodies(:)` / `sinks(:)` arrays
(type particle SUMMER_SPH.f90:14-27 | Variable.f90:14-29, type sink SUMMER_SPH.f90:30-37)."""
from dataclasses import dataclass, field
import numpy as np

GAS_FIELDS = ("x", "y", "z", "vx", "vy", "vz", "u", "m", "alpha", "h")
SINK_FIELDS = ("x", "y", "z", "vx", "vy", "vz", "m", "radius")


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.shape != (n,):
        raise ValueError(f"expected shape ({n},), got {a.shape}")
    return a


@dataclass
class Bodies:
    x: np.ndarray; y: np.ndarray; z: np.ndarray
    vx: np.ndarray; vy: np.ndarray; vz: np.ndarray
    u: np.ndarray; m: np.ndarray; alpha: np.ndarray; h: np.ndarray

    def __post_init__(self):
        n = len(np.atleast_1d(self.x))
        for k in GAS_FIELDS:
            setattr(self, k, _f64(getattr(self, k), n))

    def __len__(self):
        return self.x.shape[0]

    @staticmethod
    def empty(n):
        return Bodies(*[np.zeros(n) for _ in GAS_FIELDS])

    def take(self, idx):
        return Bodies(*[getattr(self, k)[idx] for k in GAS_FIELDS])

    def copy(self):
        return Bodies(*[getattr(self, k).copy() for k in GAS_FIELDS])


@dataclass
class Sinks:
    x: np.ndarray; y: np.ndarray; z: np.ndarray
    vx: np.ndarray; vy: np.ndarray; vz: np.ndarray
    m: np.ndarray; radius: np.ndarray

    def __post_init__(self):
        n = len(np.atleast_1d(self.x))
        for k in SINK_FIELDS:
            setattr(self, k, _f64(np.atleast_1d(getattr(self, k)), n))

    def __len__(self):
        return self.x.shape[0]

    @staticmethod
    def empty(n=0):
        return Sinks(*[np.zeros(n) for _ in SINK_FIELDS])

    @staticmethod
    def dummy():
        """The zero-mass, zero-radius sink the reference creates when the IC file has none (SUMMER_SPH.f90:698-707)."""
        return Sinks.empty(1)

    def copy(self):
        return Sinks(*[getattr(self, k).copy() for k in SINK_FIELDS])
