"""Host-side mirror of the reference's file-level contract for the hot path (Python plumbing around libmepb.so).

    ./input.cfg + ./w90files/<seed>_hr.dat  ->  core_worker on the GPU  ->  ./out/*.npy (+ out/occ.dat)

mirrors  src/input_paras.f90:79-293 (init_parameters), src/file_io.f90:663-746 (read_hr_file), src/core.f90:79-298
(core_worker) and src/file_io.f90:186-398 (write_*_tensors): same keys, same units (eV in the file, Hartree inside), same
output file names, shapes and (Fortran) index order.  The C++ `mepInterp` driver (mepinterp_b200/csrc/driver) does the same
without Python.  Nothing here computes physics; all of it happens in the CUDA kernels behind the C ABI.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

from . import api

AU_TO_EV = 27.211385          # src/constants.f90:18  (must be this literal: inputs are eV, kernels work in Hartree)


# ------------------------------------------------------------------------------------------------ input.cfg
def _logical(tok: str) -> bool:
    t = tok.strip().strip(".").lower()
    if t[:1] == "t":
        return True
    if t[:1] == "f":
        return False
    raise ValueError(f"input.cfg: cannot read logical from {tok!r}")


def read_input_cfg(path: str):
    """Parse the reference's INI dialect (thirdparty/config_fortran/m_config.f90:226-311): `[section]`, `key = v1 v2 ..`,
    `#`/`;` comments, separators blank , ' " tab.  Returns (params namespace in atomic units, a_latt (3,3) rows = a_i*a0,
    wf_centers (3,N_wf) or None, meta dict with seed_name / plot_bands / R_vect_float)."""
    kv = {}
    sec = ""
    with open(path) as fh:
        for raw in fh:
            line = raw
            for ch in "#;":
                pos = line.find(ch)
                if pos >= 0:
                    line = line[:pos]
            line = line.strip()
            if not line:
                continue
            if line[0] == "[" and "]" in line:
                sec = line[1:line.index("]")].strip()
                continue
            if "=" in line:
                k, v = line.split("=", 1)
                for ch in ",'\"\t":
                    v = v.replace(ch, " ")
                kv[(sec, k.strip())] = v.split()

    def val(sec, key, conv, default=None, n=None):
        toks = kv.get((sec, key))
        if toks is None:
            if default is None:
                raise KeyError(f"input.cfg: [{sec}] {key} is required")
            return default
        if n is None:
            return conv(toks[0])
        if len(toks) < n:
            raise ValueError(f"input.cfg: [{sec}] {key} needs {n} values, found {len(toks)}")
        return [conv(t) for t in toks[:n]]

    a0 = val("unitCell", "a0", float, 1.0)
    a_latt = a0 * np.array([val("unitCell", "a%d" % i, float, n=3) for i in (1, 2, 3)])      # input_paras.f90:165-168
    n_wf = val("wannBase", "N_wf", int, 0)
    wfc = None
    if n_wf > 0:                                                                             # input_paras.f90:127-132
        wfc = np.array([val("wannBase", "wf_centers_" + ax, float, n=n_wf) for ax in "xyz"])
    ev = 1.0 / AU_TO_EV
    p = api.make_params(
        mp_grid=tuple(val("wannInterp", "mp_grid", int, n=3)),
        do_gauge_trafo=val("wannInterp", "doGaugeTrafo", _logical, True),
        use_cart_velo=val("wannInterp", "use_cart_velo", _logical, False),
        do_mep=val("jobs", "do_mep", _logical, False), do_kubo=val("jobs", "do_kubo", _logical, False),
        do_ahc=val("jobs", "do_ahc", _logical, False), do_opt=val("jobs", "do_opt", _logical, False),
        do_write_mep_bands=val("MEP", "do_write_mep_bands", _logical, False),
        valence_bands=val("MEP", "valence_bands", int, 1),
        n_hw=max(1, val("Laser", "N_hw", int, 1)),                                           # input_paras.f90:174
        n_eF=val("Fermi", "N_eF", int, 1),
        hw_min=val("Laser", "hw_min", float, 0.0) * ev, hw_max=val("Laser", "hw_max", float, 0.0) * ev,   # :171-172
        eF_min=val("Fermi", "eF_min", float, 0.0) * ev, eF_max=val("Fermi", "eF_max", float, 0.0) * ev,   # :176-178
        T_kelvin=val("Fermi", "Tkelvin", float, 0.0),
        eta=val("Fermi", "eta_smearing", float, 0.0) * ev,
        kubo_tol=val("Fermi", "kuboTol", float, 1e-5))
    meta = dict(seed_name=val("wannBase", "seed_name", str, "wf1"), plot_bands=val("jobs", "plot_bands", _logical, False),
                R_vect_float=val("jobs", "R_vect_float", _logical, False), do_gyro=val("jobs", "do_gyro", _logical, False))
    return p, a_latt, wfc, meta


# ------------------------------------------------------------------------------------------------ <seed>_hr.dat
def read_hr_dat(path: str):
    """wannier90 `_hr.dat` as the reference reads it (src/file_io.f90:663-746): comment; num_wann; nrpts; nrpts degeneracy
    integers (read, NOT applied: :685-709); then nrpts*num_wann^2 rows `R1 R2 R3 m n Re Im` in eV.  R of a block is taken
    from its first row (:723-729); |Re|,|Im| < 1e-7 eV are zeroed (:732-733); H(m,n,R); result in Hartree."""
    with open(path) as fh:
        fh.readline()
        nw = int(fh.readline().split()[0])
        nR = int(fh.readline().split()[0])
        ndeg = 0
        while ndeg < nR:
            ndeg += len(fh.readline().split())
        if ndeg != nR:
            raise ValueError("[read_hr_file]: Wigner-Seitz degeneracy block has the wrong length")
        body = np.loadtxt(fh, dtype=np.float64, ndmin=2)
    if body.shape != (nR * nw * nw, 7):
        raise ValueError(f"[read_hr_file]: expected {nR * nw * nw} rows of 7 columns, found {body.shape}")
    blocks = body.reshape(nR, nw * nw, 7)
    R = np.ascontiguousarray(blocks[:, 0, 0:3].T)
    re, im = blocks[:, :, 5].copy(), blocks[:, :, 6].copy()
    re[np.abs(re) < 1e-7] = 0.0
    im[np.abs(im) < 1e-7] = 0.0
    H = np.zeros((nw, nw, nR), dtype=np.complex128)
    m = blocks[:, :, 3].astype(np.int64) - 1
    n = blocks[:, :, 4].astype(np.int64) - 1
    sc = np.broadcast_to(np.arange(nR)[:, None], m.shape)
    H[m, n, sc] = re + 1j * im
    return R, H / AU_TO_EV


AU_TO_ANGSTRM = 0.52917721092  # src/constants.f90:19


def read_r_dat(path: str, R: np.ndarray, nw: int) -> np.ndarray:
    """`<seed>_r.dat` as read_r_file reads it (src/file_io.f90:750-795): comment; num_wann; nR*nw^2 rows
    `R1 R2 R3 m n x_re x_im y_re y_im z_re z_im` in Angstrom, R blocks in the order of the _hr.dat (checked, :772-776).
    Returns r (3,nw,nw,nR) in Bohr.  (The reference appends `_r.dat` twice to the path, :838 vs :760, and can never open the
    file; the file itself is well defined — scripts/tb_input_writer.py:75-97 writes it.)"""
    with open(path) as fh:
        fh.readline()
        if int(fh.readline().split()[0]) != nw:
            raise ValueError("[read_tb_basis]: r_mat and H_mat have different basis")
        body = np.loadtxt(fh, dtype=np.float64, ndmin=2)
    nR = R.shape[1]
    if body.shape != (nR * nw * nw, 11):
        raise ValueError(f"[read_r_file]: expected {nR * nw * nw} rows of 11 columns, found {body.shape}")
    blocks = body.reshape(nR, nw * nw, 11)
    if np.abs(blocks[:, :, 0:3] - R.T[:, None, :]).max() > 1e-8:
        raise ValueError("[read_r_file]: different R_vect order in _hr.dat and _r.dat file")
    m = blocks[:, :, 3].astype(np.int64) - 1
    n = blocks[:, :, 4].astype(np.int64) - 1
    sc = np.broadcast_to(np.arange(nR)[:, None], m.shape)
    r = np.zeros((3, nw, nw, nR), dtype=np.complex128)
    for a in range(3):
        r[a][m, n, sc] = blocks[:, :, 5 + 2 * a] + 1j * blocks[:, :, 6 + 2 * a]
    return r / AU_TO_ANGSTRM


def read_tb_dat(path: str):
    """wannier90 `<seed>_tb.dat` as read_tb_file reads it (src/file_io.f90:555-658): comment; three lattice rows in Angstrom (they
    OVERRIDE a_latt, :575-580); num_wann; nrpts; degeneracies (skipped); per R a blank line, `R1 R2 R3`, nw^2 rows `m n Re Im` (eV,
    |x| < 1e-7 -> 0); then per R a blank line, `R1 R2 R3`, nw^2 rows `m n x_re x_im y_re y_im z_re z_im` (Angstrom).
    Returns (a_latt rows in Bohr, R (3,nR) — the integer triples of the file, SURVEY Q9 —, H Hartree, r Bohr)."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    cell = np.array([[float(t) for t in lines[1 + i].split()[:3]] for i in range(3)]) / AU_TO_ANGSTRM
    nw, nR = int(lines[4].split()[0]), int(lines[5].split()[0])
    pos = 6 + -(-nR // 15)
    R = np.zeros((3, nR))
    H = np.zeros((nw, nw, nR), dtype=np.complex128)
    r = np.zeros((3, nw, nw, nR), dtype=np.complex128)
    for sc in range(nR):
        pos += 1                                                       # blank line
        R[:, sc] = [float(t) for t in lines[pos].split()[:3]]
        blk = np.array([ln.split() for ln in lines[pos + 1: pos + 1 + nw * nw]], dtype=np.float64)
        re, im = blk[:, 2].copy(), blk[:, 3].copy()
        re[np.abs(re) < 1e-7] = 0.0
        im[np.abs(im) < 1e-7] = 0.0
        H[blk[:, 0].astype(int) - 1, blk[:, 1].astype(int) - 1, sc] = re + 1j * im
        pos += 1 + nw * nw
    for sc in range(nR):
        pos += 1
        r3 = np.array([float(t) for t in lines[pos].split()[:3]])
        if np.abs(r3 - R[:, sc]).max() > 1e-8:
            raise ValueError("[read_tb_basis]: rHopp has different sc order than tHopp")
        blk = np.array([ln.split() for ln in lines[pos + 1: pos + 1 + nw * nw]], dtype=np.float64)
        for a in range(3):
            r[a, blk[:, 0].astype(int) - 1, blk[:, 1].astype(int) - 1, sc] = blk[:, 2 + 2 * a] + 1j * blk[:, 3 + 2 * a]
        pos += 1 + nw * nw
    return cell, R, H / AU_TO_EV, r / AU_TO_ANGSTRM


def load_run_dir(run_dir: str):
    """(model, params, meta) of a reference run directory (main.f90:24-45, core.f90:144; read_tb_basis file_io.f90:800-860:
    `<seed>_tb.dat` wins, else `<seed>_hr.dat` plus `<seed>_r.dat` when present — a position operator switches on the Berry
    connection branch)."""
    p, a_latt, wfc, meta = read_input_cfg(os.path.join(run_dir, "input.cfg"))
    seed = os.path.join(run_dir, "w90files", meta["seed_name"])
    r = None
    if os.path.exists(seed + "_tb.dat"):
        a_latt, R, H, r = read_tb_dat(seed + "_tb.dat")
    else:
        R, H = read_hr_dat(seed + "_hr.dat")
        if os.path.exists(seed + "_r.dat"):
            r = read_r_dat(seed + "_r.dat", R, H.shape[0])
    if wfc is not None and wfc.shape[1] != H.shape[0]:
        raise ValueError("input.cfg: N_wf differs from num_wann of the _hr.dat")
    p.do_gyro = bool(meta["do_gyro"])
    return api.make_model(R, H, a_latt, wfc, r), p, meta


# ------------------------------------------------------------------------------------------------ outputs
def hw_list(params) -> np.ndarray:
    if params.n_hw > 1:
        return params.hw_min + np.arange(params.n_hw) * ((params.hw_max - params.hw_min) / (params.n_hw - 1))
    return np.array([params.hw_min])


def write_outputs(run_dir: str, params, res: dict):
    """out/*.npy exactly as write_*_tensors do (src/file_io.f90:186-398): Fortran-ordered <f8 / <c16, atomic units;
    hw_lst in eV when n_hw > 1 (:194-210); out/occ.dat (core.f90:378)."""
    out = os.path.join(run_dir, "out")
    os.makedirs(out, exist_ok=True)

    def save(name, arr):
        np.save(os.path.join(out, name + ".npy"), np.asfortranarray(arr))

    hw = hw_list(params)
    save("hw_lst", hw * AU_TO_EV if params.n_hw > 1 else hw)
    for k in ("mep_ic", "mep_lc", "mep_cs", "mep_tens", "mep_bands", "kubo_mep_ic", "kubo_mep_lc", "kubo_mep_cs",
              "kubo_mep_tens", "ahc_tens", "ahcVELO", "ohcVELO", "opt_Ssymm", "opt_Asymm", "photoC_2nd", "gyro_C", "gyro_D", "gyro_Dw"):
        if k in res:
            save(k, res[k])
    ef = params.eF_min + (np.arange(params.n_eF) * ((params.eF_max - params.eF_min) / (params.n_eF - 1)) if params.n_eF > 1 else 0.0)
    with open(os.path.join(out, "occ.dat"), "w") as fh:
        for e, n in zip(np.atleast_1d(ef), res["Ne"]):          # '(f16.8,a,f16.8)' with a tab, core.f90:378
            fh.write(f"{e * AU_TO_EV:16.8f}\t{n:16.8f}\n")
    return out


def core_worker(run_dir: str = ".", device: int = 0, write: bool = True) -> dict:
    """The reference's response run (main.f90 -> core_worker) on one B200: read inputs, run the mesh, write out/."""
    model, params, meta = load_run_dir(run_dir)
    if meta["plot_bands"]:
        raise api.MepbError(api.ERR_UNSUPPORTED, "plot_bands=T selects band mode; use mepinterp_b200.host.band_worker")
    ctx = api.Context(model, params, device)
    try:
        res = ctx.run_mesh()
    finally:
        ctx.close()
    if write:
        write_outputs(run_dir, params, res)
    return res


def band_worker(run_dir: str, kpts: np.ndarray, device: int = 0) -> np.ndarray:
    """Band mode (src/band_calc.f90:30-92): eigenvalues (nk, nw) in Hartree on an explicit list of fractional k-points."""
    model, params, _ = load_run_dir(run_dir)
    ctx = api.Context(model, params, device)
    try:
        e, _ = ctx.interp(np.asarray(kpts, dtype=np.float64), want_V=False)
    finally:
        ctx.close()
    return np.ascontiguousarray(e.T)


def probe_phi(phis, params, device: int = 0) -> list:
    """The parameter sweep of scripts/probe_phi.py / mep_worker.py without processes or files: the reference writes a run
    directory and spawns `mpirun mepInterp` per phi (mep_worker.py:142); here ONE live context takes the regenerated H(R)
    through mepb_update_model (same R set, same FT convention) and reruns the mesh.  Returns the result dict per phi."""
    from . import models
    phis = [float(x) for x in phis]
    ctx = api.Context(models.souza_tb(phis[0]), params, device)
    out = []
    try:
        for i, phi in enumerate(phis):
            if i:
                ctx.update_model(models.souza_tb(phi).H)
            out.append(ctx.run_mesh())
    finally:
        ctx.close()
    return out


def probe_hw(model, params, sweeps, device: int = 0) -> list:
    """scripts/probe_hw.py:167-208 as a library call: the reference starts `mpirun mepInterp` in a fresh run directory per
    (hw window, eta, E_F, T) setting; here ONE live context (model resident on the GPU) takes each setting through
    mepb_update_params and reruns the mesh.  `sweeps`: iterable of dicts with any of hw_min, hw_max, eF_min, eF_max, T_kelvin,
    eta, kubo_tol (atomic units).  Returns the result dict per setting."""
    ctx = api.Context(model, params, device)
    out = []
    try:
        for kw in sweeps:
            ctx.update_params(**kw)
            out.append(ctx.run_mesh())
    finally:
        ctx.close()
    return out


def probe_spiral(thetas, params, device: int = 0, **model_kw) -> list:
    """scripts/FeMn_WX_rashba.py:loop_spiral as a library call: the FeMn spin-spiral model regenerated in memory per canting angle
    (models.femn_rashba_tb) and re-uploaded into ONE live context (mepb_update_model); returns the result dict per angle."""
    from . import models
    thetas = [float(x) for x in thetas]
    ctx = api.Context(models.femn_rashba_tb(thetas[0], **model_kw), params, device)
    out = []
    try:
        for i, th in enumerate(thetas):
            if i:
                ctx.update_model(models.femn_rashba_tb(th, **model_kw).H)
            out.append(ctx.run_mesh())
    finally:
        ctx.close()
    return out
