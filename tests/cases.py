"""Shared test inputs: the golden run directories and seeded synthetic tight-binding models."""
from __future__ import annotations

import copy
import os

import numpy as np

from oracle import mep_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name: str, **overrides):
    """(Model, Params) of tests/golden/<name>, parsed by the oracle's restatement of the reference readers."""
    model, params = mo.load_model(os.path.join(GOLDEN, name))
    params = copy.copy(params)
    for k, v in overrides.items():
        setattr(params, k, v)
    return model, params


def r_ball(rmax2: int):
    """All integer R with |R|^2 <= rmax2 in lexicographic order (R1 slowest), like the reference's writers."""
    n = int(np.sqrt(rmax2)) + 1
    out = [(a, b, c) for a in range(-n, n + 1) for b in range(-n, n + 1) for c in range(-n, n + 1)
           if a * a + b * b + c * c <= rmax2]
    return np.array(out, dtype=np.float64).T


def synthetic_model(nw: int, rmax2: int, seed: int = 20261017, tb_centers: bool = False, decay: float = 0.5,
                    a_latt=None, onsite_span_ev: float = 3.0):
    """SURVEY.md 8(d) C5 recipe at any size: H_R = (g1 + i g2) exp(-|R| decay) eV, symmetrised
    H_R <- (H_R + H_{-R}^dagger)/2, onsite diag += linspace(-span, span, nw) eV; orthorhombic 5x6x7 Bohr cell."""
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
    i0 = index[(0.0, 0.0, 0.0)]
    Hs[:, :, i0] += np.diag(np.linspace(-onsite_span_ev, onsite_span_ev, nw))
    Hs /= mo.AU_TO_EV
    if a_latt is None:
        a_latt = np.diag([5.0, 6.0, 7.0])
    wfc = rng.uniform(size=(3, nw)) if tb_centers else None
    return mo.Model(nw=nw, nR=nR, R=R, H=Hs, a_latt=np.asarray(a_latt, dtype=np.float64), wf_centers=wfc)


def ev(x):
    return x / mo.AU_TO_EV


def rel_err(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    s = np.abs(b).max() if scale is None else scale
    return float(np.abs(a - b).max() / max(s, 1e-300))
