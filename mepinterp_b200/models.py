"""In-memory tight-binding model generators (no files): the synthetic large model of BASELINE config 5.

SURVEY.md 8(d) C5 recipe: R = all integer triples with |R|^2 <= rmax2 in lexicographic order (R1 slowest) like the
reference's writers (scripts/tb_input_writer.py); H_R = (g1 + i g2) exp(-|R| decay) eV with g ~ N(0,1) from
numpy.random.default_rng(seed), symmetrised H_R <- (H_R + H_{-R}^dagger)/2, onsite diag += linspace(-span, span, nw) eV;
orthorhombic 5 x 6 x 7 Bohr cell; Wannier FT convention unless tb_centers.
"""
from __future__ import annotations

import numpy as np

from .api import make_model

AU_TO_EV = 27.211385


def r_ball(rmax2: int) -> np.ndarray:
    n = int(np.sqrt(rmax2)) + 1
    out = [(a, b, c) for a in range(-n, n + 1) for b in range(-n, n + 1) for c in range(-n, n + 1) if a * a + b * b + c * c <= rmax2]
    return np.array(out, dtype=np.float64).T


def synthetic_tb(nw: int = 64, rmax2: int = 61, seed: int = 20261017, tb_centers: bool = False, decay: float = 0.5, a_latt=None,
                 onsite_span_ev: float = 3.0):
    rng = np.random.default_rng(seed)
    R = r_ball(rmax2)
    nR = R.shape[1]
    H = np.zeros((nw, nw, nR), dtype=np.complex128)
    for i in range(nR):
        g = rng.normal(size=(nw, nw)) + 1j * rng.normal(size=(nw, nw))
        H[:, :, i] = g * np.exp(-np.linalg.norm(R[:, i]) * decay)
    index = {tuple(R[:, i]): i for i in range(nR)}
    Hs = np.zeros_like(H)
    for i in range(nR):
        j = index[tuple(-R[:, i])]
        Hs[:, :, i] = 0.5 * (H[:, :, i] + H[:, :, j].conj().T)
    Hs[:, :, index[(0.0, 0.0, 0.0)]] += np.diag(np.linspace(-onsite_span_ev, onsite_span_ev, nw))
    Hs /= AU_TO_EV
    if a_latt is None:
        a_latt = np.diag([5.0, 6.0, 7.0])
    wfc = rng.uniform(size=(3, nw)) if tb_centers else None
    return make_model(R, Hs, np.asarray(a_latt, dtype=np.float64), wfc)
