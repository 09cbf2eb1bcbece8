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


# ------------------------------------------------------------------------------------------------------------------
# Cubic 8-band model of New J. Phys. 12, 053032 (2010) as the reference's generator parametrises it
# (scripts/souza_tb_model.py:394-442: get_souza_tb(phi_para)).  Restated as a bond table:
# per axis four bonds, each with an intra-cell hopping H(0)(i,j) = t[a] and an inter-cell hopping H(+-e)(i',j') = t[a'] (the SAME
# value at +e and -e, as the reference writes it), plus the Hermitian partners.  t[axis][a] = 27.21139 eV * exp(i pi phi[axis][a]).
# (i, j, a) are 1-based like the _hr.dat columns "m n".
_SOUZA_BONDS = (
    ((1, 2, 1, 2, 1, 2), (4, 3, 4, 3, 4, 3), (5, 6, 5, 6, 5, 6), (8, 7, 8, 7, 8, 7)),      # x
    ((1, 4, 1, 4, 1, 4), (2, 3, 2, 3, 2, 3), (5, 8, 5, 8, 5, 8), (6, 7, 6, 7, 6, 7)),      # y
    ((1, 5, 1, 5, 1, 5), (2, 6, 2, 6, 2, 6), (3, 7, 3, 7, 3, 7), (4, 8, 4, 8, 4, 8)),      # z
)
_SOUZA_ONSITE = (-6.5, 0.9, 1.4, 1.2, -6.0, 1.5, 0.8, 1.2)
_SOUZA_PHI = ((None, 1.3, 0.8, 0.3, 1.4, 0.6, 0.8, 1.9), (0.5, 0.2, 1.4, 1.9, 0.8, 1.7, 0.6, 0.3), (1.7, 0.5, 0.6, 1.0, 0.3, 0.7, 1.2, 1.4))
_PY_AU_TO_EV = 27.21139          # the generator's constant; the Fortran reader divides by 27.211385 (keep both)


def souza_tb(phi: float = 0.0, as_file: bool = True):
    """In-memory model of the Souza toy system for the phase parameter `phi` (the knob scripts/probe_phi.py sweeps).

    as_file=True reproduces what the reference pipeline hands to core_worker bit for bit: the generator's eV values rounded to
    the 8 decimals of tb_input_writer.py ("{:16.8f}"), the reader's |x| < 1e-7 eV -> 0 rule and its division by 27.211385
    (file_io.f90:732-741).  R points in the writer's (sorted) order; cubic cell a = 2 Bohr; Wannier FT convention (N_wf = 0)."""
    R = np.array([(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0)], dtype=np.float64).T
    index = {tuple(int(x) for x in R[:, i]): i for i in range(7)}
    H = np.zeros((8, 8, 7), dtype=np.complex128)
    for n, e in enumerate(_SOUZA_ONSITE):
        H[n, n, index[(0, 0, 0)]] = e * _PY_AU_TO_EV
    for ax in range(3):
        t = [_PY_AU_TO_EV * np.exp(1j * np.pi * (phi if p is None else p)) for p in _SOUZA_PHI[ax]]
        e = [0, 0, 0]
        e[ax] = 1
        for (i0, j0, a0, i1, j1, a1) in _SOUZA_BONDS[ax]:
            H[i0 - 1, j0 - 1, index[(0, 0, 0)]] = t[a0 - 1]
            H[j0 - 1, i0 - 1, index[(0, 0, 0)]] = np.conj(t[a0 - 1])
            for sgn in (-1, 1):
                r = index[tuple(sgn * x for x in e)]
                H[i1 - 1, j1 - 1, r] = t[a1 - 1]
                H[j1 - 1, i1 - 1, r] = np.conj(t[a1 - 1])
    if as_file:
        re = np.vectorize(lambda v: float("{:.8f}".format(v)))(H.real)
        im = np.vectorize(lambda v: float("{:.8f}".format(v)))(H.imag)
        re[np.abs(re) < 1e-7] = 0.0
        im[np.abs(im) < 1e-7] = 0.0
        H = re + 1j * im
    H = H / AU_TO_EV
    return make_model(R, H, np.diag([2.0, 2.0, 2.0]), None)


# ------------------------------------------------------------------------------------------------------------------
# FeMn 3q / spin-spiral tight-binding model with an in-plane Rashba term, as the reference's generator parametrises it
# (scripts/FeMn_WX_rashba.py:76-428 write_3q_HR, scripts/FeMn_latt_OPT.py, scripts/FeMn_get_spin_config.py:60-78; SURVEY.md 8(d) C4,
# 8(f) row 3).  Restated, not copied: four Fe/Mn sites of the fcc cell in the rhombohedral setting, two spin blocks (orbital index =
# 4 spin + site), nearest-neighbour hopping within (111) planes (t) and between them (t / strain^2), local exchange -J m_i.sigma,
# Rashba-like complex hopping on the bonds the generator's index test selects (its `i <= natoms` / `j > natoms` comparisons are
# 0-based against natoms = 4, so only the orbital pairs (4, 5..7) and (5..7, 4) receive it — reproduced literally), lattice strained
# along (111) at constant cell volume (the same bounded scalar minimisation the reference runs, scipy.optimize.minimize_scalar).
_FEMN_LATT0 = ((5.60092859, 0.00000000, 3.96045459), (-2.80046430, 4.85054644, 3.96045459), (-2.80046430, -4.85054644, 3.96045459))
_FEMN_SITES = ((0.0, 0.0, 0.0), (0.5, 0.5, 0.0), (0.5, 0.0, 0.5), (0.0, 0.5, 0.5))
_FEMN_PHI_DEG = (45.0, -135.0, 135.0, -45.0)


def femn_strained_lattice(strain: float) -> np.ndarray:
    """Rows a_1, a_2, a_3 (Bohr) of the rhombohedral cell stretched by `strain` along (111) at the volume of the unstrained cell."""
    from scipy.optimize import minimize_scalar
    latt0 = np.array(_FEMN_LATT0)
    vol0 = np.dot(latt0[0, :], np.cross(latt0[1, :], latt0[2, :]))
    z1 = latt0[0, 2] * strain

    def cell(x):
        return np.array([[2.0 * x, 0.0, z1], [-x, x * np.sqrt(3), z1], [-x, -x * np.sqrt(3), z1]])

    def miss(x):
        c = cell(x)
        return np.abs(np.dot(c[0, :], np.cross(c[1, :], c[2, :])) - vol0)

    opt = minimize_scalar(miss, bounds=(0, 6.8597), method="bounded", tol=None, options={"xatol": 1e-7, "maxiter": 5000, "disp": 0})
    return cell(opt.x)


def femn_rashba_tb(theta_deg: float = 54.7, t: float = -1.2, strain: float = 1.2, J_ex: float = -3.0, tso: float = -0.3,
                   as_file: bool = True):
    """In-memory FeMn spin-spiral + Rashba model for the canting angle `theta_deg` (the knob scripts/FeMn_WX_rashba.py:loop_spiral
    sweeps); defaults = BASELINE config 4.  as_file=True reproduces what the reference pipeline hands to core_worker bit for bit:
    the generator's eV values rounded to the 6 decimals of its `{:12.6f}` writer, the reader's |x| < 1e-7 eV -> 0 rule and its
    division by 27.211385 (file_io.f90:732-741).  Tight-binding FT convention (wf_centers = the site positions, twice)."""
    sx, sy, sz = np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]])
    nat, norb = 4, 8
    t_intra, t_inter = t, t / float(strain) ** 2
    nns = 2 if abs(strain - 1.0) > 1e-3 else 1
    latt = femn_strained_lattice(strain)
    sites = np.array(_FEMN_SITES)
    sites_cart = sites @ latt
    th = np.pi / 180.0 * theta_deg
    theta = np.array([th, th, np.pi - th, np.pi - th])
    phi = np.pi / 180.0 * np.array(_FEMN_PHI_DEG)
    mag = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=1)
    # neighbour shells: distances from the four sites to every site of the 3 x 3 x 3 block, distinct values in ascending order
    cells = [(x, y, z) for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)]
    block = np.array([x * latt[0] + y * latt[1] + z * latt[2] + sites_cart[i] for (x, y, z) in cells for i in range(nat)])
    dist = np.unique(np.array([np.linalg.norm(sites_cart[i] - block[j]) for i in range(nat) for j in range(len(block))]))
    shells = [d for d in dist if abs(d) > 1e-3]
    for i in range(len(shells)):
        if shells[i] > 0:
            for j in range(i + 1, len(shells)):
                if abs(shells[j] - shells[i]) < 1e-5:
                    shells[j] = 0
    shells = [d for d in shells if abs(d) > 1e-3]
    cutoff = shells[nns + 1]
    # exchange term: -J sum_i (m_i . sigma) (x) |i><i|
    H1 = np.zeros((norb, norb), dtype=np.complex128)
    for i in range(nat):
        proj = np.zeros((nat, nat))
        proj[i, i] = 1
        H1 = H1 + np.kron(mag[i, 0] * sx + mag[i, 1] * sy + mag[i, 2] * sz, proj)
    H1 = -J_ex * H1
    R = np.array(cells, dtype=np.float64).T
    H = np.zeros((norb, norb, len(cells)), dtype=np.complex128)
    for ir, (x, y, z) in enumerate(cells):
        Rf = np.array([x, y, z], dtype=np.float64)
        Rc = x * latt[0] + y * latt[1] + z * latt[2]
        for i in range(norb):
            for j in range(norb):
                hh = 0
                ii, jj = i % nat, j % nat
                if ii != jj and ((i < nat) == (j < nat)):
                    d_frac = Rf - sites[jj] - sites[ii]
                    d_cart = Rc + sites_cart[jj] - sites_cart[ii]
                    if np.linalg.norm(d_cart) < cutoff:
                        if abs(d_cart[2]) < 1e-5:
                            hh = t_intra
                            if i <= nat and j > nat:
                                hh = hh + 1j * tso * (d_frac[1] + 1j * d_frac[0])
                            elif i > nat and j <= nat:
                                hh = hh + 1j * tso * (d_frac[1] - 1j * d_frac[0])
                        else:
                            hh = t_inter
                if x == 0 and y == 0 and z == 0:
                    hh = hh + H1[i, j]
                H[i, j, ir] = hh
    if as_file:
        re = np.vectorize(lambda v: float("{:.6f}".format(v)))(H.real)
        im = np.vectorize(lambda v: float("{:.6f}".format(v)))(H.imag)
        re[np.abs(re) < 1e-7] = 0.0
        im[np.abs(im) < 1e-7] = 0.0
        H = re + 1j * im
    H = H / AU_TO_EV
    wfc = np.ascontiguousarray(np.vstack([sites, sites]).T)
    return make_model(R, H, latt, wfc)
