"""Text I/O with the reference's surface: same names, same file formats.

read_data_from_file  — SUMMER_SPH.f90:594-716 | "SUMMER_SPH - Variable.f90":729-852
read_params_from_file — Variable.f90:854-919
make_save            — SUMMER_SPH.f90:719-738 | Variable.f90:921-942
"""
import os
import numpy as np
import pandas as pd

from ._abi import SphParams, default_params, MODE_VARIABLE_H
from .state import Bodies, Sinks


def read_data_from_file(filename, params: SphParams):
    """Header line skipped; whitespace-separated rows. Fixed h reads the first 8 columns
    `x y z vx vy vz u m` (extras ignored) and sets alpha := 0 (F:647,681); variable h reads 10
    (`... alpha h`, V:782). A row with u == 0.0 exactly is a sink (F:658-659); sink radius is
    params.sink_radius (F:694 | V:830). No sink rows -> one dummy zero sink (F:698-707).
    Deviation (documented): a variable-h sink row may carry only 8 columns anywhere in the file; the
    Fortran list-directed read would only tolerate that for trailing rows."""
    if not os.path.exists(filename):
        raise FileNotFoundError(f"Error opening file: {filename}")          # F:612-615
    variable = bool(params.mode & MODE_VARIABLE_H)
    df = pd.read_csv(filename, sep=r"\s+", header=None, skiprows=1, names=list(range(16)), engine="c",
                     dtype=np.float64, on_bad_lines="error", float_precision="round_trip")
    a = df.to_numpy(dtype=np.float64)[:, :10]
    if a.shape[0] == 0:
        raise ValueError(f"No data found in file: {filename}")              # F:625-628
    if np.isnan(a[:, :8]).any():
        bad = int(np.nonzero(np.isnan(a[:, :8]).any(axis=1))[0][0]) + 1
        raise ValueError(f"Error reading line {bad}")                       # F:648-651
    is_sink = a[:, 6] == 0.0
    gas = a[~is_sink]
    if variable and np.isnan(gas[:, 8:10]).any():
        bad = int(np.nonzero(np.isnan(a[:, 8:10]).any(axis=1) & ~is_sink)[0][0]) + 1
        raise ValueError(f"Error reading line {bad}")
    alpha = gas[:, 8].copy() if variable else np.zeros(gas.shape[0])        # F:681
    h = gas[:, 9].copy() if variable else np.full(gas.shape[0], params.h_fixed)
    bodies = Bodies(gas[:, 0], gas[:, 1], gas[:, 2], gas[:, 3], gas[:, 4], gas[:, 5], gas[:, 6], gas[:, 7], alpha, h)
    sk = a[is_sink]
    if sk.shape[0] > 0:
        sinks = Sinks(sk[:, 0], sk[:, 1], sk[:, 2], sk[:, 3], sk[:, 4], sk[:, 5], sk[:, 7], np.full(sk.shape[0], params.sink_radius))
    else:
        sinks = Sinks.dummy()
    print(f" Successfully read {len(bodies)} bodies and {len(sinks)} sinks from {filename}.")   # F:714
    return bodies, sinks


def read_params_from_file(filename, base: SphParams = None):
    """parameters.txt: header, then rows of 9 values (last row wins, V:898-904):
    bounding_size max_depth theta gamma eta convergence_criteria max_length timestep_scale end_time."""
    if not os.path.exists(filename):
        raise FileNotFoundError(f"Error opening file: {filename}")
    p = base.copy() if base is not None else default_params(MODE_VARIABLE_H)
    rows = []
    with open(filename) as f:
        f.readline()
        for line in f:
            tok = line.replace(",", " ").split()
            if not tok:
                break
            if len(tok) < 9:
                raise ValueError(f"Error reading line {len(rows) + 1}")
            rows.append(tok[:9])
    if not rows:
        raise ValueError(f"No data found in file: {filename}")
    r = rows[-1]
    p.bounding_size = float(r[0]); p.max_depth = int(float(r[1])); p.theta = float(r[2]); p.gamma = float(r[3])
    p.eta = float(r[4]); p.convergence_criteria = float(r[5]); p.max_length = float(r[6])
    p.timestep_scale = float(r[7]); p.end_time = float(r[8])
    print(f" Successfully read parameters from{filename}.")                 # V:917
    return p


def write_ics(filename, bodies: Bodies, sinks: Sinks = None, columns=10):
    """IC file in the reference's format: header + `%.15e` rows (Disc_ICs.py:40), sink rows (u = 0) last."""
    names = ["x", "y", "z", "vx", "vy", "vz", "energy", "mass", "alpha", "smoothing"][:columns]
    cols = [bodies.x, bodies.y, bodies.z, bodies.vx, bodies.vy, bodies.vz, bodies.u, bodies.m, bodies.alpha, bodies.h][:columns]
    with open(filename, "w") as f:
        f.write(" ".join(names) + "\n")
        np.savetxt(f, np.stack(cols, 1), fmt="%.15e")
        if sinks is not None and len(sinks):
            srows = np.stack([sinks.x, sinks.y, sinks.z, sinks.vx, sinks.vy, sinks.vz, np.zeros(len(sinks)), sinks.m], 1)
            np.savetxt(f, srows, fmt="%.15e")


def make_save(bodies: Bodies, sinks: Sinks, number, params: SphParams, directory="."):
    """`save<number>.txt`; like the reference's status="new" (F:728) an existing file is an error.
    Gas rows: x y z vx vy vz energy mass alpha [smoothing]; then sink rows x y z vx vy vz 0.0 m.
    Values are written with 17 significant digits (lossless, like gfortran's list-directed output)."""
    variable = bool(params.mode & MODE_VARIABLE_H)
    path = os.path.join(directory, f"save{number}.txt")
    with open(path, "x") as f:                                              # status="new"
        hdr = " x  y  z  vx  vy vz energy mass  alpha  " + ("smoothing" if variable else "")
        f.write(hdr.rstrip() + "\n")
        cols = [bodies.x, bodies.y, bodies.z, bodies.vx, bodies.vy, bodies.vz, bodies.u, bodies.m, bodies.alpha]
        if variable:
            cols.append(bodies.h)
        if len(bodies):
            np.savetxt(f, np.stack(cols, 1), fmt="%25.17E")
        if len(sinks):
            srows = np.stack([sinks.x, sinks.y, sinks.z, sinks.vx, sinks.vy, sinks.vz, np.zeros(len(sinks)), sinks.m], 1)
            np.savetxt(f, srows, fmt="%25.17E")
    return path
