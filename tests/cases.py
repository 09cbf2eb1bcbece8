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
    from mepinterp_b200.models import r_ball as rb
    return rb(rmax2)


def synthetic_model(nw: int, rmax2: int, seed: int = 20261017, tb_centers: bool = False, decay: float = 0.5,
                    a_latt=None, onsite_span_ev: float = 3.0):
    """SURVEY.md 8(d) C5 recipe at any size (generator: mepinterp_b200/models.py), as an oracle Model."""
    from mepinterp_b200.models import synthetic_tb
    m = synthetic_tb(nw, rmax2, seed, tb_centers, decay, a_latt, onsite_span_ev)
    return mo.Model(nw=m.nw, nR=m.nR, R=m.R, H=m.H, a_latt=m.a_latt, wf_centers=m.wf_centers)


def ev(x):
    return x / mo.AU_TO_EV


def rel_err(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    s = np.abs(b).max() if scale is None else scale
    return float(np.abs(a - b).max() / max(s, 1e-300))


_FAMILIES = [("mep_ic", "mep_lc", "mep_cs", "mep_tens", "mep_bands"),
             ("kubo_mep_ic", "kubo_mep_lc", "kubo_mep_cs", "kubo_mep_tens")]


def family_scale(ref: dict, key: str) -> float:
    """Magnitude a tensor's error is measured against.  The MEP tensors are sums over bands / terms that cancel by many
    orders (mep_ic ~ 1e-8 from band terms ~ 1e-2 for FeMn), so their floating-point reassociation error scales with the
    un-cancelled band-resolved magnitude: use the largest |.| within the tensor's family (SURVEY.md 7 'compare with
    tolerance relative to sum |term|')."""
    keys = (key,)
    for fam in _FAMILIES:
        if key in fam:
            keys = fam
    s = 0.0
    for k in keys:
        if k in ref and isinstance(ref[k], np.ndarray) and ref[k].size:
            s = max(s, float(np.abs(ref[k]).max()))
    return max(s, 1e-300)


def assert_results_close(out: dict, ref: dict, tol: float, tols: dict = None, skip=()):
    assert out["n_k"] == ref["n_k"], (out["n_k"], ref["n_k"])
    assert out["eF_idx"] == ref["eF_idx"]
    worst = {}
    for k, v in ref.items():
        if not isinstance(v, np.ndarray) or k in skip:
            continue
        assert k in out, k
        err = float(np.abs(np.asarray(out[k]) - v).max()) / family_scale(ref, k)
        worst[k] = err
        t = (tols or {}).get(k, tol)
        assert err <= t, f"{k}: rel err {err:.3e} > {t:.1e}"
    return worst


def perturbed(model, seed: int, eps: float = 4.4e-16):
    """The same model with every H(R) entry changed by ~2 ulp (Hermitian closure preserved)."""
    rng = np.random.default_rng(seed)
    H = model.H * (1.0 + eps * rng.normal(size=model.H.shape))
    idx = {tuple(model.R[:, i]): i for i in range(model.nR)}
    Hs = H.copy()
    for i in range(model.nR):
        j = idx.get(tuple(-model.R[:, i]))
        if j is not None and np.abs(model.H[:, :, i] - model.H[:, :, j].conj().T).max() == 0.0:
            Hs[:, :, i] = 0.5 * (H[:, :, i] + H[:, :, j].conj().T)
    m2 = copy.copy(model)
    m2.H = Hs
    return m2


def oracle_reference(model, params, nthreads=None, n_perturb=2):
    """Oracle result + the two yardsticks the parity bar is measured with:
      scale[k] = max over the tensor of  sum_k |X(k)| / norm   (the un-cancelled magnitude of the k-sum)
      sens[k]  = max |oracle(H) - oracle(H (1 + 2 ulp))|        (how far the REFERENCE ALGORITHM itself moves under a
                 rounding-level change of its input: unguarded 1/dE^2, 1/dE^3 factors at near-degenerate bands,
                 kubo.f90:65, kubo_mep.f90:73,126, make some tensors ill-conditioned, e.g. kubo_mep on FeMn)."""
    from oracle import cpu_ref
    nthreads = nthreads or os.cpu_count()
    h = cpu_ref.Handle(model, params)
    g = params.mp_grid
    nk = g[0] * g[1] * g[2]
    with np.errstate(all="ignore"):
        ref = h.run_mesh(nthreads)
        sc = h.finalize(h.accumulate_abs(0, nk, nthreads))
        scale, sens = {}, {}
        for k, v in ref.items():
            if isinstance(v, np.ndarray):
                scale[k] = max(float(np.abs(sc[k]).max()) if k in sc and np.size(sc[k]) else 0.0, float(np.abs(v).max()) if v.size else 0.0)
                sens[k] = 0.0
        for s in range(n_perturb):
            r2 = cpu_ref.Handle(perturbed(model, 1000 + s), params).run_mesh(nthreads)
            for k in sens:
                sens[k] = max(sens[k], float(np.nanmax(np.abs(r2[k] - ref[k]))) if ref[k].size else 0.0)
    for fam in _FAMILIES:     # derived sums (mep_tens = ic + lc + cs) inherit the family's yardsticks
        ks = [k for k in fam if k in scale]
        if ks:
            smax, tmax = max(scale[k] for k in ks), max(sens[k] for k in ks)
            for k in ks:
                scale[k], sens[k] = smax, max(sens[k], tmax if k.endswith("tens") else sens[k])
    return ref, scale, sens


def assert_parity(out, ref, scale, sens, tol=1e-8, k_sens=20.0, label=""):
    """|gpu - oracle| <= tol * sum|term|  +  k_sens * (oracle's own movement under a 2-ulp input change)."""
    assert out["n_k"] == ref["n_k"], (out["n_k"], ref["n_k"])
    assert out["eF_idx"] == ref["eF_idx"]
    assert out["n_degenerate"] == ref["n_degenerate"]
    report = {}
    for k, v in ref.items():
        if not isinstance(v, np.ndarray):
            continue
        assert k in out, k
        err = float(np.abs(np.asarray(out[k]) - v).max()) if v.size else 0.0
        allowed = tol * scale[k] + k_sens * sens[k]
        report[k] = (err, scale[k], sens[k])
        assert err <= allowed, f"{label} {k}: |gpu-oracle| {err:.3e} > {allowed:.3e} (scale {scale[k]:.3e}, sens {sens[k]:.3e})"
    return report


def packed_slices(params, nw=None):
    """name -> slice of the packed partial-sum buffer (include/mepb.h; identical in oracle/cpu_ref.c and mep_oracle.pack_partials)."""
    p = params
    o, sl = 2, {"n_k": slice(0, 1), "n_degenerate": slice(1, 2)}

    def take(name, n):
        nonlocal o
        sl[name] = slice(o, o + n)
        o += n

    take("Ne", p.n_eF)
    if p.do_mep:
        take("mep_bands_ic", 9 * p.valence_bands); take("mep_bands_lc", 9 * p.valence_bands)
    if p.do_kubo:
        take("kubo_mep_ic", 9 * p.n_eF); take("kubo_mep_lc", 9 * p.n_eF)
    if p.do_ahc:
        take("kubo_ahc", 9 * p.n_eF); take("velo_ahc", 18 * p.n_hw * p.n_eF); take("kubo_ohc", 18 * p.n_hw * p.n_eF)
    if p.do_opt:
        take("photo2", 54 * p.n_hw * p.n_eF)
    return sl, o


def oracle_eigenpairs(model, params, kpts):
    """(e (nw,nk), V (3,nw,nw,nk)) of the C oracle (zheevd + zgemm) on a k list."""
    from oracle import cpu_ref
    e, V, _, _ = cpu_ref.Handle(model, params).interp(kpts, want_V=True)
    return e, V


def band_sum_reference(model, params, e, V):
    """Oracle band sums on supplied eigenpairs + the per-tensor yardstick sum |term| (numpy oracle, mag=True)."""
    from oracle import cpu_ref
    ref = cpu_ref.Handle(model, params).band_sums(e, V)
    with np.errstate(all="ignore"):
        mg = mo.pack_partials(params, mo.band_sums(params, np.ascontiguousarray(e.T), np.ascontiguousarray(np.moveaxis(V, 3, 0)), mag=True),
                              model.nw)
    sl, n = packed_slices(params)
    assert n == ref.size == mg.size
    return ref, {k: float(np.nanmax(mg[v])) if mg[v].size else 0.0 for k, v in sl.items()}, sl
