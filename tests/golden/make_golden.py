"""Generates the committed golden fixtures of tests/golden/ from the CPU oracle (run in the build container):
    python tests/golden/make_golden.py
* notebook_example3.json -- the reference's only known-answer numbers (notebooks/Example_simulations.ipynb cells 21/23/25),
  copied from the notebook outputs, plus the oracle's reproduction of them.
* tm_small.npz / te_small.npz / mod_small.npz -- small seeded problems: inputs, CSR of the system matrix and the solved
  (Nx,Ny,3) fields, so the GPU tests can also be checked against committed vectors.
* eig_ring.json -- eigenfrequency() of the notebook's ring resonator (cell 31: Cylinder R=1.0 eps 12.25 with an R=0.8 air core,
  4 x 4 um, Npml 15) at 200^2 and 400^2, TM and TE, 6 modes :LM.  The 400^2 values are the regression values SURVEY.md §8c
  records from its probe of the same restatement; they are NOT reference-published numbers (the reference has none for this path)."""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import fdfd_oracle as O  # noqa: E402

W = 2 * math.pi * 200e12


def small(pol):
    g = O.Grid2D(0.05, [6, 5], [0, 2.0], [0, 1.5])
    d = O.Device(g, [W])
    rng = np.random.default_rng(7)
    d.eps_r = (1 + 11 * rng.random(g.size())).round(2) + 0j
    d.src[10, 12] = 1j
    A, b, _ = O.system_matrix(d, W, pol)
    A.sort_indices()
    f = O.solve(d, pol)
    return dict(dh=0.05, npml=[6, 5], xr=[0, 2.0], yr=[0, 1.5], omega=W, eps_r=d.eps_r, src=d.src, indptr=A.indptr.astype(np.int64),
                indices=A.indices.astype(np.int64), data=A.data, fields=f["data"])


def modulated_small():
    g = O.Grid2D(0.04, [8, 8], [0, 3.0], [-1.0, 1.0])
    d = O.ModulatedDevice(g, [2 * math.pi * 1.939e14], Omega=4.541e14, nsidebands=1)
    a = 0.2202
    O.mask_values(d.eps_r, g, lambda x, y: -a / 2 <= y <= a / 2, 12.25)
    O.mask_values(d.deps_r, g, lambda x, y: (0.6 <= x <= 2.4) and (-a / 2 <= y <= 0), lambda x, y: np.exp(1j * 2.9263 * x))
    d.src[12, :] = 1j
    f = O.solve_modulated(d)[0]
    return dict(dh=0.04, npml=[8, 8], xr=[0, 3.0], yr=[-1.0, 1.0], omega=d.omega[0], Omega=d.Omega, eps_r=d.eps_r, deps_r=d.deps_r,
                src=d.src, fields=np.stack([x["data"] for x in f], axis=3))


def eig_ring():
    out = {"source": "oracle (SciPy ARPACK shift-invert, eigen.jl:69-115 restated); 400^2 values = SURVEY.md section 8c probe values",
           "geometry": "Grid(4/n, [15,15], [-2,2], [-2,2]); Cylinder((0,0),0.8,eps 1) over Cylinder((0,0),1.0,eps 12.25); omega0 = 2 pi 200e12; nev 6, which LM"}
    for n in (200, 400):
        g = O.Grid2D(4.0 / n, [15, 15], [-2.0, 2.0], [-2.0, 2.0])
        d = O.Device(g, [W])
        xs, ys = O.xc(g)[:, None], O.yc(g)[None, :]
        r2 = xs ** 2 + ys ** 2
        d.eps_r[r2 <= 1.0] = 12.25
        d.eps_r[r2 <= 0.64] = 1.0          # the air core is listed first in the notebook: first shape wins (device.jl:47-61)
        for pol, name in ((O.TM, "TM"), (O.TE, "TE")):
            om, _ = O.eigenfrequency(d, pol, 6, which="LM", v0=np.ones(len(g), dtype=complex))
            out[f"{name}_{n}"] = [[float(z.real), float(z.imag)] for z in om]
    return out


def phc_cavity_eps(g, a=0.5, r=0.1, eps_rod=12.25, nx=11, ny=9, missing=((-1, 0), (0, 0), (1, 0))):
    """L3-type cavity: nx x ny square lattice of rods (pitch a, radius r) centred on the origin, three rods in a row removed"""
    xs, ys = O.xc(g)[:, None], O.yc(g)[None, :]
    eps = np.ones(g.size(), dtype=complex)
    for i in range(-(nx // 2), nx // 2 + 1):
        for j in range(-(ny // 2), ny // 2 + 1):
            if (i, j) in missing:
                continue
            eps[(xs - i * a) ** 2 + (ys - j * a) ** 2 <= r * r] = eps_rod
    return eps


def eig_config4():
    """BASELINE config 4: eigenfrequency() of the notebook ring (cell 31) at 400^2 and of a photonic-crystal cavity, 10 nearest modes.
    12 modes are stored so that a test asking for 10 can match each of its values to a stored one even when the 10th / 11th are a
    degenerate pair."""
    out = {"source": "oracle (SciPy ARPACK shift-invert, eigen.jl:69-115 restated); regression values, not reference-published numbers",
           "ring": "Grid(0.01, [15,15], [-2,2], [-2,2]); Cylinder((0,0),0.8,eps 1) over Cylinder((0,0),1.0,eps 12.25); omega0 = 2 pi 200e12; nev 12, which LM",
           "phc": "Grid(0.025, [15,15], [-3.2,3.2], [-2.8,2.8]); 11 x 9 square lattice of rods (pitch 0.5, radius 0.1, eps 12.25), rods (-1,0),(0,0),(1,0) removed; TM; omega0 = 2 pi 200e12; nev 12, which LM"}
    g = O.Grid2D(0.01, [15, 15], [-2.0, 2.0], [-2.0, 2.0])
    d = O.Device(g, [W])
    xs, ys = O.xc(g)[:, None], O.yc(g)[None, :]
    r2 = xs ** 2 + ys ** 2
    d.eps_r[r2 <= 1.0] = 12.25
    d.eps_r[r2 <= 0.64] = 1.0
    for pol, name in ((O.TM, "TM"), (O.TE, "TE")):
        om, _ = O.eigenfrequency(d, pol, 12, which="LM", v0=np.ones(len(g), dtype=complex))
        out[f"ring_{name}_400"] = [[float(z.real), float(z.imag)] for z in om]
    g = O.Grid2D(0.025, [15, 15], [-3.2, 3.2], [-2.8, 2.8])
    d = O.Device(g, [W])
    d.eps_r = phc_cavity_eps(g)
    om, _ = O.eigenfrequency(d, O.TM, 12, which="LM", v0=np.ones(len(g), dtype=complex))
    out["phc_TM"] = [[float(z.real), float(z.imag)] for z in om]
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "config4":
        json.dump(eig_config4(), open(os.path.join(HERE, "eig_config4.json"), "w"), indent=1)
        sys.exit(0)
    json.dump(eig_config4(), open(os.path.join(HERE, "eig_config4.json"), "w"), indent=1)
    json.dump(eig_ring(), open(os.path.join(HERE, "eig_ring.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "tm_small.npz"), **small(O.TM))
    np.savez_compressed(os.path.join(HERE, "te_small.npz"), **small(O.TE))
    np.savez_compressed(os.path.join(HERE, "mod_small.npz"), **modulated_small())
    nb = {"source": "reference notebooks/Example_simulations.ipynb, Example 3 (cells 17-25)",
          "Nin_cell21": 1.3625216010889075e-20, "Nout_cell23": 1.3618731014650896e-20, "ratio_cell25": 0.9995240445191477,
          "oracle_Nin": 1.3625216010888627e-20, "oracle_Nout": 1.3618731014650806e-20,
          "oracle_rel_err": [-3.3e-14, -6.7e-15]}
    json.dump(nb, open(os.path.join(HERE, "notebook_example3.json"), "w"), indent=1)
    print("wrote", os.listdir(HERE))
