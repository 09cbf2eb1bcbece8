"""GPU parity tests (`-m gpu`): libmepb's hand-written sm_100a path, called through the C ABI (ctypes, include/mepb.h),
against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): eigenvalues 1e-10 relative, response tensors 1e-8 relative in FP64, gauge-invariant
quantities only.  "Relative" is measured against the un-cancelled magnitude of each k-sum, plus the oracle's own
movement under a 2-ulp perturbation of H(R) for the tensors the reference leaves unguarded at near-degenerate bands
(tests/cases.py::oracle_reference explains both yardsticks).  Fourier sums are gauge independent and are compared
element-wise at 1e-13.  Nothing here reads /root/reference.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import cpu_ref
from oracle import mep_oracle as mo
from tests.cases import (GOLDEN, assert_parity, band_sum_reference, ev, load_case, oracle_eigenpairs, oracle_reference, rel_err,
                         synthetic_model)

pytestmark = pytest.mark.gpu

EIG_TOL = 1e-10
TENS_TOL = 1e-8


def _cases_small():
    all_jobs = dict(do_mep=True, do_kubo=True, do_ahc=True, do_opt=True, n_hw=7, hw_min=ev(0.03), hw_max=ev(6.0), eta=ev(0.1))
    yield "C1_souza_ahc_mep_32^3_full", load_case("souza_phi0", mp_grid=(32, 32, 32))
    yield "C1_souza_kubo_16^3", load_case("souza_phi0", mp_grid=(16, 16, 16), do_kubo=True)
    yield "C2_femn_ahc_nhw121_16^3", load_case("femn_d0.900", mp_grid=(16, 16, 16))
    yield "C2_femn_d1.1_ahc_12^3", load_case("femn_d1.100", mp_grid=(12, 12, 12))
    yield "C3_femn_mep_kubo_16^3", load_case("femn_d0.900", mp_grid=(16, 16, 16), do_ahc=False, do_mep=True, do_kubo=True, n_hw=1, hw_max=0.0)
    yield "C4_rashba_opt_ahc_nhw200_8^3", load_case("femn_rashba", mp_grid=(8, 8, 8))
    yield "femn_all_jobs_T300_neF3_ragged", load_case("femn_d0.900", mp_grid=(11, 12, 13), n_eF=3, eF_min=ev(-0.5), eF_max=ev(0.5),
                                                       T_kelvin=300.0, **all_jobs)
    yield "rashba_all_jobs_T0", load_case("femn_rashba", mp_grid=(7, 6, 5), **dict(all_jobs, n_hw=9))
    sm = synthetic_model(8, 5, tb_centers=True)
    yield "syn_nw8_tb_cart_velo", (sm, mo.Params(mp_grid=(10, 11, 12), valence_bands=3, use_cart_velo=True, **all_jobs))
    yield "syn_nw8_no_gauge_trafo", (sm, mo.Params(mp_grid=(6, 6, 6), valence_bands=3, do_gauge_trafo=False, T_kelvin=120.0, **all_jobs))
    for nw in (4, 16):
        yield f"syn_nw{nw}_wannier_T200", (synthetic_model(nw, 3), mo.Params(mp_grid=(8, 8, 8), valence_bands=nw // 2, T_kelvin=200.0,
                                                                              **dict(all_jobs, n_hw=5)))
    yield "syn_nw8_nval6_T0_all_jobs", (sm, mo.Params(mp_grid=(7, 7, 6), valence_bands=6, **all_jobs))       # mep_sums_kernel helper lanes
    yield "syn_nw4_nval1_T0_all_jobs", (synthetic_model(4, 3), mo.Params(mp_grid=(9, 8, 7), valence_bands=1, **dict(all_jobs, n_hw=5)))
    yield "syn_nw16_tb_neF2", (synthetic_model(16, 2, tb_centers=True),
                               mo.Params(mp_grid=(5, 5, 5), valence_bands=7, n_eF=2, eF_min=ev(-0.3), eF_max=ev(0.2), **all_jobs))


CASES = dict(_cases_small())


def _cases_large():
    """General path (mepb_large.cu): num_wann up to 64.  C5 = synthetic nw=64, nR=2007 on a reduced mesh."""
    all_jobs = dict(do_mep=True, do_kubo=True, do_ahc=True, do_opt=True, n_hw=5, hw_min=ev(0.05), hw_max=ev(4.0), eta=ev(0.1))
    yield "C5_syn_nw64_nR2007_ahc_T300_4x3x5", (synthetic_model(64, 61), mo.Params(mp_grid=(4, 3, 5), do_ahc=True, T_kelvin=300.0))
    yield "syn_nw64_nR123_all_jobs", (synthetic_model(64, 9), mo.Params(mp_grid=(3, 3, 3), valence_bands=30, T_kelvin=300.0, **all_jobs))
    yield "syn_nw32_tb_centers_T0", (synthetic_model(32, 4, tb_centers=True), mo.Params(mp_grid=(4, 4, 3), valence_bands=11, **all_jobs))
    yield "syn_nw13_odd_cart_velo", (synthetic_model(13, 5, tb_centers=True),
                                     mo.Params(mp_grid=(5, 4, 3), valence_bands=6, use_cart_velo=True, T_kelvin=150.0, **all_jobs))
    yield "syn_nw24_no_gauge_neF2", (synthetic_model(24, 3), mo.Params(mp_grid=(3, 4, 5), valence_bands=9, do_gauge_trafo=False, n_eF=2,
                                                                      eF_min=ev(-0.2), eF_max=ev(0.3), T_kelvin=200.0, **all_jobs))


CASES_LARGE = dict(_cases_large())
FORCED_LARGE = ["C1_souza_kubo_16^3", "C2_femn_ahc_nhw121_16^3", "C4_rashba_opt_ahc_nhw200_8^3", "femn_all_jobs_T300_neF3_ragged",
                "syn_nw8_tb_cart_velo", "syn_nw4_wannier_T200", "syn_nw16_tb_neF2"]


class force_path:
    """MEPB_PATH=large routes a small-num_wann model through the general kernels (testing knob of mepb_create)."""

    def __init__(self, which):
        self.which = which

    def __enter__(self):
        self.old = os.environ.get("MEPB_PATH")
        if self.which:
            os.environ["MEPB_PATH"] = self.which

    def __exit__(self, *a):
        os.environ.pop("MEPB_PATH", None)
        if self.old is not None:
            os.environ["MEPB_PATH"] = self.old


@pytest.mark.parametrize("name", list(CASES))
def test_fourier_sums_elementwise(libmepb, name):
    """FT_R_to_k (wann_interp.f90:151-281): H(k), dH/dk_a(k) are gauge independent -> element-wise."""
    model, params = CASES[name]
    kp = np.random.default_rng(5).uniform(-0.5, 0.5, size=(301, 3))
    kp[0] = 0.0                                         # Gamma
    kp[1] = (0.5, -0.5, 0.25)                           # zone boundary
    _, _, Hk_r, Hka_r = cpu_ref.Handle(model, params).interp(kp, want_V=False, want_H=True)
    ctx = libmepb.Context(model, params)
    Hk, Hka = ctx.fourier(kp)
    ctx.close()
    assert rel_err(Hk, Hk_r) < 1e-13
    assert rel_err(Hka, Hka_r) < 1e-13


@pytest.mark.parametrize("name", list(CASES))
def test_eigenvalues_and_gauge_invariants(libmepb, name):
    """zheevd + rotate_gauge (matrix_math.f90:192-226, wann_interp.f90:502-550): eigenvalues ascending, 1e-10 relative;
    sum rules tr H = sum E, ||H_upper-Hermitian||_F^2 = sum E^2; V diagonal (band velocities) and the gauge-invariant
    Frobenius norms sum_nm |V_a(n,m)|^2 = ||dH_a||_F^2."""
    model, params = CASES[name]
    kp = np.random.default_rng(6).uniform(-0.5, 0.5, size=(257, 3))
    h = cpu_ref.Handle(model, params)
    e_r, V_r, Hk_r, Hka_r = h.interp(kp, want_V=True, want_H=True)
    ctx = libmepb.Context(model, params)
    e, V = ctx.interp(kp)
    ctx.close()
    hnorm = np.abs(e_r).max()
    assert np.all(np.diff(e, axis=0) >= 0.0)
    assert np.abs(e - e_r).max() <= EIG_TOL * hnorm
    assert np.abs(e.sum(axis=0) - np.einsum("nnk->k", Hk_r).real).max() <= 1e-12 * hnorm * model.nw
    if params.do_gauge_trafo:
        fro_V = (np.abs(V) ** 2).sum(axis=(1, 2))
        fro_H = (np.abs(Hka_r) ** 2).sum(axis=(1, 2))
        assert np.abs(fro_V - fro_H).max() <= 1e-11 * fro_H.max()
        # band velocities: invariant wherever the band is isolated
        gap = np.minimum(np.diff(e_r, axis=0, prepend=-np.inf), np.diff(e_r, axis=0, append=np.inf)[..., :])
        iso = np.abs(gap) > 1e-4
        dV = np.einsum("annk->ank", V) - np.einsum("annk->ank", V_r)
        if iso.any():                                   # FeMn bands come in near-degenerate pairs: nothing isolated there
            assert np.abs(dV[:, iso]).max() <= 1e-9 * np.abs(V_r).max()
    else:
        assert rel_err(V, V_r) < 1e-13                  # W gauge: V = dH


@pytest.mark.parametrize("name", list(CASES))
def test_run_mesh_parity(libmepb, name):
    """core_worker (core.f90:79-298) end to end: every tensor the reference writes to out/*.npy."""
    model, params = CASES[name]
    ref, scale, sens = oracle_reference(model, params)
    ctx = libmepb.Context(model, params)
    out = ctx.run_mesh()
    launches = ctx.kernel_launches
    ctx.close()
    assert launches > 0
    assert_parity(out, ref, scale, sens, tol=TENS_TOL, label=name)
    if params.do_ahc and "rashba" not in name:       # (the generator's Rashba H(R) is not Hermitian-closed: no antisymmetry)
        a = out["ahc_tens"]
        # sigma_ij = -sigma_ji up to the same yardsticks (pairs of occupied near-degenerate bands cancel only to rounding x 1/dE^2)
        assert np.abs(a + a.T).max() <= TENS_TOL * scale["ahc_tens"] + 20.0 * sens["ahc_tens"]


def _case(name):
    if name in CASES_LARGE:
        return CASES_LARGE[name], None
    return CASES[name], "large"


@pytest.mark.parametrize("name", list(CASES_LARGE) + FORCED_LARGE)
def test_general_path_fourier_and_eigen(libmepb, name):
    """Dense Fourier contraction (TMA-staged H(R), phase tables) element-wise at 1e-13; CTA-per-k Jacobi eigenvalues 1e-10."""
    (model, params), force = _case(name)
    kp = np.random.default_rng(11).uniform(-0.5, 0.5, size=(70, 3))
    kp[0] = 0.0
    e_r, V_r, Hk_r, Hka_r = cpu_ref.Handle(model, params).interp(kp, want_V=True, want_H=True)
    with force_path(force):
        ctx = libmepb.Context(model, params)
    Hk, Hka = ctx.fourier(kp)
    e, V = ctx.interp(kp)
    ctx.close()
    assert rel_err(Hk, Hk_r) < 1e-13
    assert rel_err(Hka, Hka_r) < 1e-13
    hnorm = np.abs(e_r).max()
    assert np.all(np.diff(e, axis=0) >= 0.0)
    assert np.abs(e - e_r).max() <= EIG_TOL * hnorm
    assert np.abs((e ** 2).sum(axis=0) - (e_r ** 2).sum(axis=0)).max() <= 1e-11 * hnorm ** 2 * model.nw
    if params.do_gauge_trafo:
        fro_V = (np.abs(V) ** 2).sum(axis=(1, 2))
        fro_H = (np.abs(Hka_r) ** 2).sum(axis=(1, 2))
        assert np.abs(fro_V - fro_H).max() <= 1e-11 * fro_H.max()
    else:
        assert rel_err(V, V_r) < 1e-13


@pytest.mark.parametrize("name", list(CASES_LARGE) + ["femn_all_jobs_T300_neF3_ragged", "syn_nw8_tb_cart_velo"])
def test_general_path_mesh_separable_fourier(libmepb, name):
    """Mesh mode of the general path: the axis-by-axis (separable) Fourier transform on Monkhorst-Pack points equals the
    oracle's nR-term sums element-wise (1e-13), on a ragged range that starts and ends inside mesh lines."""
    (model, params), force = _case(name)
    N1, N2, N3 = params.mp_grid
    nk_mesh = N1 * N2 * N3
    b0, b1 = min(2, nk_mesh - 1), nk_mesh - min(1, nk_mesh - 2)
    q = np.arange(b0, b1)
    kp = np.stack([(2.0 * (q % N1 + 1) - N1 - 1) / (2.0 * N1), (2.0 * ((q // N1) % N2 + 1) - N2 - 1) / (2.0 * N2),
                   (2.0 * (q // (N1 * N2) + 1) - N3 - 1) / (2.0 * N3)], axis=1)                  # k_space.f90:351-375
    _, _, Hk_r, Hka_r = cpu_ref.Handle(model, params).interp(kp, want_V=True, want_H=True)
    with force_path(force):
        ctx = libmepb.Context(model, params)
    Hk, Hka = ctx.fourier_mesh(b0, b1 - b0)
    ctx.close()
    assert rel_err(Hk, Hk_r) < 1e-13
    assert rel_err(Hka, Hka_r) < 1e-13


class chunk_log2:
    """MEPB_CHUNK_LOG2=n makes mepb_create schedule the mesh in waves of 2^n k-points (tuning knob): with n = 10 even the small test
    meshes run as many waves alternating over the two CUDA streams, like the BASELINE meshes do with the default 2^18."""

    def __init__(self, n):
        self.n = n

    def __enter__(self):
        self.old = os.environ.get("MEPB_CHUNK_LOG2")
        os.environ["MEPB_CHUNK_LOG2"] = str(self.n)

    def __exit__(self, *a):
        os.environ.pop("MEPB_CHUNK_LOG2", None)
        if self.old is not None:
            os.environ["MEPB_CHUNK_LOG2"] = self.old


def _mesh_points(grid, b0, b1):
    N1, N2, N3 = grid
    q = np.arange(b0, b1)
    return np.stack([(2.0 * (q % N1 + 1) - N1 - 1) / (2.0 * N1), (2.0 * ((q // N1) % N2 + 1) - N2 - 1) / (2.0 * N2),
                     (2.0 * (q // (N1 * N2) + 1) - N3 - 1) / (2.0 * N3)], axis=1)                  # k_space.f90:351-375


@pytest.mark.parametrize("grid,span", [((200, 2, 2), (3, 797)), ((72, 5, 3), (70, 1011)), ((65, 3, 2), (0, 390)), ((130, 1, 3), (64, 325))])
def test_mesh_separable_fourier_long_lines_c5(libmepb, grid, span):
    """BASELINE config 5's model (nw = 64, nR = 2007) on meshes whose lines are LONGER than one 64-k CTA tile (N1 = 65, 72, 130,
    200 — C5 itself is 200^3): second and third tiles of a line (q1_0 > 0), partial last tiles, ranges that start and end inside a
    line, against the oracle's 2007-term sums element-wise (wann_interp.f90:254-279)."""
    model = synthetic_model(64, 61)
    params = mo.Params(mp_grid=grid, do_ahc=True, T_kelvin=300.0)
    b0, b1 = span
    kp = _mesh_points(grid, b0, b1)
    _, _, Hk_r, Hka_r = cpu_ref.Handle(model, params).interp(kp, want_V=False, want_H=True)
    ctx = libmepb.Context(model, params)
    Hk, Hka = ctx.fourier_mesh(b0, b1 - b0)
    ctx.close()
    assert rel_err(Hk, Hk_r) < 1e-13
    assert rel_err(Hka, Hka_r) < 1e-13


def test_c5_run_mesh_2160_kpoints_three_waves(libmepb):
    """BASELINE config 5 (synthetic nw = 64, nR = 2007, AHC, T = 300 K) against the oracle on 72 x 5 x 6 = 2160 k-points, scheduled
    in waves of 1024 (MEPB_CHUNK_LOG2=10): mesh lines of two CTA tiles, waves that start and end inside lines and alternate over
    both streams — the schedule the 200^3 run uses (core.f90:165-270 on the whole mesh)."""
    model = synthetic_model(64, 61)
    params = mo.Params(mp_grid=(72, 5, 6), do_ahc=True, T_kelvin=300.0)
    from tests.cases import oracle_reference as oref
    ref, scale, sens = oref(model, params, n_perturb=1)
    with chunk_log2(10):
        ctx = libmepb.Context(model, params)
    out = ctx.run_mesh()
    launches = ctx.kernel_launches
    ctx.close()
    assert launches >= 3 * 7                              # three waves of the seven-kernel general path
    rep = assert_parity(out, ref, scale, sens, tol=TENS_TOL, label="C5 72x5x6")
    # the allowance actually needed with the conditioning term switched off (sigma = 0): printed for the record
    print("[C5 2160 k] error / (1e-8 sum|term|) with sigma = 0: " +
          " ".join(f"{k}={e / max(1e-8 * s, 1e-300):.2e}" for k, (e, s, _) in rep.items()))


MULTIWAVE = ["C2_femn_ahc_nhw121_16^3", "C3_femn_mep_kubo_16^3", "C4_rashba_opt_ahc_nhw200_8^3", "femn_all_jobs_T300_neF3_ragged",
             "syn_nw16_tb_neF2", "C1_souza_kubo_16^3"]


@pytest.mark.parametrize("force", [None, "large"])
@pytest.mark.parametrize("name", MULTIWAVE)
def test_run_mesh_parity_many_waves(libmepb, name, force):
    """The same oracle comparison as test_run_mesh_parity with the mesh cut into waves of 1024 k-points (4 waves for 16^3, ragged
    last waves for 11x12x13 and 8^3) on both paths (thread-per-k + band-sum kernels, and the general kernels): the two-stream
    wave schedule of mepb_accumulate, which the BASELINE-size runs use, against the oracle."""
    model, params = CASES[name]
    ref, scale, sens = oracle_reference(model, params)
    with force_path(force), chunk_log2(10):
        ctx = libmepb.Context(model, params)
    out = ctx.run_mesh()
    ctx.close()
    assert_parity(out, ref, scale, sens, tol=TENS_TOL, label=f"{name} waves of 1024 ({force or 'small'} path)")


@pytest.mark.parametrize("name", list(CASES_LARGE) + FORCED_LARGE)
def test_general_path_run_mesh_parity(libmepb, name):
    (model, params), force = _case(name)
    ref, scale, sens = oracle_reference(model, params)
    with force_path(force):
        ctx = libmepb.Context(model, params)
    out = ctx.run_mesh()
    prof_launches = ctx.kernel_launches
    ctx.close()
    assert prof_launches > 0
    assert_parity(out, ref, scale, sens, tol=TENS_TOL, label=name)


def test_golden_ahc_anchor_on_gpu(libmepb):
    """The reference's golden FeMn sigma_xy (config/FeMn_TB/delta0.900 via tests/golden) at the test's own 64^3 mesh and at
    BASELINE config 2's 128^3: mesh-convergence-level pin (1.1e-3 and 7.4e-4 from the golden)."""
    gold = np.load(os.path.join(GOLDEN, "femn_d0.900", "ahc_tens.npy"))
    gold_w = np.load(os.path.join(GOLDEN, "femn_d0.900", "ahcVELO.npy"))
    for mesh, tol, tol_w in ((64, 2e-3, 2.5e-4), (128, 1.5e-3, 5e-5)):
        m, p = load_case("femn_d0.900", mp_grid=(mesh,) * 3)
        ctx = libmepb.Context(m, p)
        out = ctx.run_mesh()
        ctx.close()
        assert out["n_k"] == mesh ** 3
        assert rel_err(out["ahc_tens"], gold) < tol
        assert np.abs(out["ahcVELO"] - gold_w).max() < tol_w
        assert np.abs(out["ahc_tens"] + out["ahc_tens"].T).max() < 1e-7 * np.abs(out["ahc_tens"]).max()   # rounding x 1/dE^2 of occupied pairs
    m, p = load_case("femn_d1.000", mp_grid=(32, 32, 32))
    ctx = libmepb.Context(m, p)
    out = ctx.run_mesh()
    ctx.close()
    assert np.abs(out["ahc_tens"]).max() < 1e-10         # zero by symmetry, like the golden


def test_full_size_c2_shard_linearity(libmepb):
    """BASELINE config 2 at full size (128^3, n_hw=121): size-independent properties — the sum of ragged slab partials
    equals the whole-mesh partials (linearity / 'sum over ranks'), and a second context fed the summed buffer through
    mepb_set_partials finalizes to the same tensors."""
    m, p = load_case("femn_d0.900", mp_grid=(128, 128, 128))
    nk = 128 ** 3
    ctx = libmepb.Context(m, p)
    ctx.reset(); ctx.accumulate(0, nk)
    whole = ctx.get_partials().copy()
    res_whole = ctx.finalize()
    cuts = [0, 1, 777_777, 777_777, 1_500_001, nk]          # ragged, one empty slab
    tot = np.zeros_like(whole)
    for a, b in zip(cuts[:-1], cuts[1:]):
        ctx.reset(); ctx.accumulate(a, b)
        tot += ctx.get_partials()
    assert tot[0] == whole[0] == nk
    mag = np.abs(whole).max()
    assert np.abs(tot - whole).max() <= 1e-11 * mag
    ctx2 = libmepb.Context(m, p)
    ctx2.set_partials(tot)
    res2 = ctx2.finalize()
    ctx2.close(); ctx.close()
    for k in ("ahc_tens", "ahcVELO", "ohcVELO"):
        assert rel_err(res2[k], res_whole[k]) < 1e-10


def test_sharded_two_contexts_equal_oracle(libmepb):
    """The MPI-host variant of the ABI (accumulate / get_partials / set_partials / finalize) with 3 'ranks'."""
    from mepinterp_b200 import dist as mdist
    m, p = load_case("souza_phi0", mp_grid=(9, 8, 7), do_kubo=True)
    nk = 9 * 8 * 7
    ref, scale, sens = oracle_reference(m, p)
    bufs = []
    for r in range(3):
        c = libmepb.Context(m, p)
        c.reset()
        c.accumulate(*mdist.slab(r, 3, nk))
        bufs.append(c.get_partials().copy())
        c.close()
    c = libmepb.Context(m, p)
    c.set_partials(mdist.sum_partials_host(bufs))
    out = c.finalize()
    c.close()
    assert_parity(out, ref, scale, sens, tol=TENS_TOL)
    # raw packed layout is the oracle's layout (include/mepb.h): compare un-normalised sums entry by entry
    raw = cpu_ref.Handle(m, p).accumulate(0, nk)
    tot = mdist.sum_partials_host(bufs)
    assert tot.shape == raw.shape and tot[0] == raw[0]


def test_edge_meshes_and_ranges(libmepb):
    m, p = load_case("souza_phi0", mp_grid=(1, 1, 1), do_kubo=True)
    ref, scale, sens = oracle_reference(m, p)
    ctx = libmepb.Context(m, p)
    assert_parity(ctx.run_mesh(), ref, scale, sens, tol=TENS_TOL)
    # empty range is legal and adds nothing; finalize with nothing accumulated is a state error
    ctx.reset(); ctx.accumulate(0, 0)
    with pytest.raises(libmepb.MepbError) as ei:
        ctx.finalize()
    assert ei.value.code == libmepb.api.ERR_STATE
    with pytest.raises(libmepb.MepbError) as ei:
        ctx.accumulate(0, 2)                             # outside the 1-point mesh
    assert ei.value.code == libmepb.api.ERR_ARG
    ctx.close()
    for grid in ((1, 1, 37), (3, 1, 2), (2, 5, 1)):
        m, p = load_case("femn_d0.900", mp_grid=grid, n_hw=3)
        ref, scale, sens = oracle_reference(m, p)
        ctx = libmepb.Context(m, p)
        assert_parity(ctx.run_mesh(), ref, scale, sens, tol=TENS_TOL, label=str(grid))
        ctx.close()


def test_call_order_and_argument_errors(libmepb):
    m, p = load_case("souza_phi0", mp_grid=(2, 2, 2))
    ctx = libmepb.Context(m, p)
    with pytest.raises(libmepb.MepbError) as ei:
        ctx.finalize()
    assert ei.value.code == libmepb.api.ERR_STATE
    ctx.set_partials(np.zeros(ctx.partials_size()))
    with pytest.raises(libmepb.MepbError) as ei:
        ctx.accumulate(0, 8)                             # accumulate after set_partials needs a reset
    assert ei.value.code == libmepb.api.ERR_STATE
    ctx.reset(); ctx.accumulate(0, 8)
    assert ctx.finalize()["n_k"] == 8
    ctx.close()
    cm = ctx.cm
    h = C.c_void_p()
    assert libmepb.lib().mepb_create(C.byref(cm), C.byref(ctx.cp), 99, C.byref(h)) == libmepb.api.ERR_ARG
    bad = mo.Params(mp_grid=(2, 2, 0))
    with pytest.raises(libmepb.MepbError):
        libmepb.Context(m, bad)


def test_exact_degeneracy_matches_reference_semantics(libmepb):
    """Q6: no guard in kubo_ahc_tens / kubo_ohc_tens / kubo_mep (kubo.f90:65,186; kubo_mep.f90:73,126) -> inf/NaN propagate,
    guarded mep_niu (mep_niu.f90:73,132) and the eta-regularised velo_ahc stay finite; offending pairs are counted."""
    nw = 4
    H = np.zeros((nw, nw, 1), dtype=np.complex128)
    H[:, :, 0] = np.diag([-1.0, -1.0, 1.0, 1.0])
    m = mo.Model(nw=nw, nR=1, R=np.zeros((3, 1)), H=H, a_latt=np.eye(3) * 3.0)
    p = mo.Params(mp_grid=(2, 2, 2), do_ahc=True, do_mep=True, do_kubo=True, valence_bands=2, n_hw=2, hw_max=ev(1.0), eta=ev(0.1))
    with np.errstate(all="ignore"):
        ref = cpu_ref.Handle(m, p).run_mesh()
    ctx = libmepb.Context(m, p)
    out = ctx.run_mesh()
    ctx.close()
    assert out["n_degenerate"] == ref["n_degenerate"] == 16
    for k in ("ahc_tens", "kubo_mep_lc"):
        assert np.array_equal(np.isfinite(out[k]), np.isfinite(ref[k])), k
    for k in ("ahcVELO", "mep_ic", "mep_lc"):
        assert np.all(np.isfinite(out[k])) and np.all(np.isfinite(ref[k])), k
        assert np.abs(out[k] - ref[k]).max() < 1e-12


def test_repeatability_bitwise(libmepb):
    """Fixed launch shapes + fixed-order tree reductions: two runs of the same mesh give bit-identical tensors."""
    m, p = load_case("femn_d0.900", mp_grid=(24, 24, 24), n_hw=11)
    ctx = libmepb.Context(m, p)
    a = ctx.run_mesh()
    b = ctx.run_mesh()
    ctx.close()
    for k in ("ahc_tens", "ahcVELO", "ohcVELO"):
        assert np.array_equal(a[k], b[k]), k


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_nccl_sharded_if_multi_gpu(libmepb):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 8)), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "NCCL_SHARDED_OK" in r.stdout


def test_multi_context_all_devices(libmepb):
    """mepb_create_multi / mepb_multi_run_mesh (one process, every GPU of the box; round-robin blocks, in-library NCCL all-reduce
    when there is more than one device) against the oracle, and bit-identical to itself on a second run."""
    n = max(1, min(_n_gpus(), 8))
    m, p = load_case("femn_d0.900", mp_grid=(20, 18, 16), n_hw=9, do_mep=True, do_kubo=True)
    ref, scale, sens = oracle_reference(m, p)
    with chunk_log2(10):
        mc = libmepb.MultiContext(m, p, list(range(n)))
    out = mc.run_mesh()
    out2 = mc.run_mesh()
    tm = mc.timings()
    mc.close()
    assert_parity(out, ref, scale, sens, tol=TENS_TOL, label=f"multi x{n}")
    for k in ("ahc_tens", "ahcVELO", "mep_bands"):
        assert np.array_equal(out[k], out2[k]), k
    assert len(tm) == n and all(t["accumulate_ms"] > 0.0 for t in tm)
    if n > 1:
        assert all(t["allreduce_ms"] > 0.0 for t in tm)


def test_strided_blocks_equal_whole_mesh(libmepb):
    """mepb_accumulate_strided: the round-robin blocks of 3 'ranks' (ragged last block, a rank with fewer blocks) sum to the
    partials of the whole mesh (the reference's round-robin ownership, mpi_comm.f90:80-99, at block granularity)."""
    m, p = load_case("femn_rashba", mp_grid=(13, 11, 9), n_hw=17)
    nk = 13 * 11 * 9
    with chunk_log2(10):
        ctx = libmepb.Context(m, p)
    assert ctx.chunk_size == 1024
    ctx.reset(); ctx.accumulate(0, nk)
    whole = ctx.get_partials().copy()
    tot = np.zeros_like(whole)
    for block in (ctx.balanced_block(nk, 3), 100):
        tot[:] = 0.0
        for r in range(3):
            ctx.reset(); ctx.accumulate_strided(0, nk, block, r, 3)
            tot += ctx.get_partials()
        assert tot[0] == whole[0] == nk
        assert np.abs(tot - whole).max() <= 1e-11 * np.abs(whole).max()
    with pytest.raises(libmepb.MepbError):
        ctx.accumulate_strided(0, nk, 10, 3, 3)
    ctx.close()


def _run_variant(tmp_path, tag, case, over, env):
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / (tag + ".npz"))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "run_case_worker.py"), case, json.dumps(over), out],
                       capture_output=True, text=True, timeout=900, cwd=root, env=e)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return dict(np.load(out))


def test_full_size_c3_independent_eigensolvers_agree(libmepb, tmp_path):
    """BASELINE config 3 at full size (FeMn MEP, 256^3 = 16.8 M k-points): the Householder + QL eigensolver and the cyclic
    Jacobi one (MEPB_SMALL_EIG=jacobi) are independent implementations of zheevd's contract; the guarded MEP tensors
    (mep_niu, band resolved before the cancellation between bands) must agree.  kubo_mep is left out: it has no
    degeneracy guard (kubo_mep.f90:73,126) and moves by 1e-2 under a 2-ulp change of the input on FeMn."""
    over = dict(mp_grid=[256, 256, 256], do_ahc=False, do_mep=True, do_kubo=True, n_hw=1, hw_max=0.0, valence_bands=4,
                do_write_mep_bands=True)
    a = _run_variant(tmp_path, "ql", "femn_d0.900", over, {})
    b = _run_variant(tmp_path, "jacobi", "femn_d0.900", over, {"MEPB_SMALL_EIG": "jacobi"})
    assert int(a["n_k"]) == int(b["n_k"]) == 256 ** 3 and int(a["eF_idx"]) == int(b["eF_idx"])
    assert abs(float(a["Ne"][0]) - float(b["Ne"][0])) <= 1e-9 * float(a["Ne"][0])
    s = np.abs(a["mep_bands"]).max()
    for k in ("mep_bands", "mep_ic", "mep_lc"):
        assert np.abs(a[k] - b[k]).max() <= 1e-8 * s, k


def test_full_size_c4_dmma_and_fma_sweeps_agree(libmepb, tmp_path):
    """BASELINE config 4 at full size (FeMn Rashba, photo2 + AHC, n_hw = 200, 128^3): the tensor-core (DMMA) sweeps against
    the thread-per-frequency FMA sweeps (MEPB_SWEEP_FMA=1) on identical items."""
    over = dict(mp_grid=[128, 128, 128])
    a = _run_variant(tmp_path, "dmma", "femn_rashba", over, {})
    b = _run_variant(tmp_path, "fma", "femn_rashba", over, {"MEPB_SWEEP_FMA": "1"})
    assert int(a["n_k"]) == int(b["n_k"]) == 128 ** 3
    for k in ("ahc_tens", "ahcVELO", "ohcVELO", "photoC_2nd"):
        assert a[k].shape == b[k].shape
        assert np.abs(a[k] - b[k]).max() <= 1e-10 * np.abs(b[k]).max(), k


def test_probe_phi_live_context_equals_fresh_contexts(libmepb):
    """scripts/probe_phi.py as a library call: H(R) regenerated per phi and re-uploaded into ONE context
    gives bit-identical tensors to a fresh context per phi, and phi = 0 matches the oracle on the reference-generated fixture."""
    from mepinterp_b200 import host, models
    m0, p = load_case("souza_phi0", mp_grid=(12, 12, 12), do_kubo=True)
    phis = [0.0, 0.3, 0.85]
    sweep = host.probe_phi(phis, p)
    for phi, res in zip(phis, sweep):
        ctx = libmepb.Context(models.souza_tb(phi), p)
        fresh = ctx.run_mesh()
        ctx.close()
        for k, v in fresh.items():
            assert np.array_equal(np.asarray(v), np.asarray(res[k])), (phi, k)
    ref, scale, sens = oracle_reference(m0, p)
    assert_parity(sweep[0], ref, scale, sens, tol=TENS_TOL, label="souza phi=0")
    assert np.abs(sweep[1]["mep_tens"] - sweep[0]["mep_tens"]).max() > 1e-6 * np.abs(sweep[0]["mep_bands"]).max()   # phi matters


@pytest.mark.parametrize("n_hw", [16, 17, 40, 64, 100, 130, 200, 300])
@pytest.mark.parametrize("name", ["femn_d0.900", "femn_rashba"])
def test_dmma_sweep_shapes(libmepb, name, n_hw):
    """Every launch shape of the tensor-core sweeps (mepb_sweep_mma.cu) against the oracle: one frequency tile per warp with
    several k-groups per CTA (n_hw = 16, 17), partly filled tiles (17, 40, 130), full CTAs (64), the shared-tile shapes of the
    photo sweep (100, 200: 3 tiles per warp + 1 shared; 130: 2 + 1), several grid.y passes (300);
    Hermitian-closed items (FeMn: 9 weights) and general items (Rashba: 18 weights), the photo sweep included."""
    m, p = load_case(name, mp_grid=(6, 5, 4), n_hw=n_hw, do_ahc=True, do_opt=True, hw_min=ev(0.04), hw_max=ev(5.0), eta=ev(0.1))
    ctx = libmepb.Context(m, p)
    out = ctx.run_mesh()
    ctx.close()
    ref, scale, sens = oracle_reference(m, p)
    assert_parity(out, ref, scale, sens, tol=TENS_TOL, label=f"{name} n_hw={n_hw}")


@pytest.mark.parametrize("name", ["femn_d0.900", "femn_rashba"])
def test_dmma_sweeps_at_finite_temperature(libmepb, name):
    """T = 300 K: every band pair has f_n != f_m, so each k-point emits the maximum number of items (28 AHC pairs, 28 + 28 photo
    items in paired order) — the full-slot case of the tensor-core sweeps, against the oracle."""
    m, p = load_case(name, mp_grid=(5, 4, 6), n_hw=40, do_ahc=True, do_opt=True, hw_min=ev(0.04), hw_max=ev(5.0), eta=ev(0.1),
                     T_kelvin=300.0)
    ctx = libmepb.Context(m, p)
    out = ctx.run_mesh()
    ctx.close()
    ref, scale, sens = oracle_reference(m, p)
    assert_parity(out, ref, scale, sens, tol=TENS_TOL, label=f"{name} T=300K")


# ------------------------------------------------------------------------------------------------ band sums on identical eigenpairs
BAND_SUM_TOL = 1e-12


def _band_sum_cases():
    jobs = dict(do_mep=True, do_kubo=True, do_ahc=True, do_opt=True, n_hw=21, hw_min=ev(0.03), hw_max=ev(6.0), eta=ev(0.1))
    yield "femn_T0", load_case("femn_d0.900", mp_grid=(12, 12, 12), valence_bands=4, **jobs), None
    yield "femn_T300_neF2", load_case("femn_d0.900", mp_grid=(10, 10, 10), valence_bands=4, T_kelvin=300.0, n_eF=2, eF_min=ev(-0.3),
                                      eF_max=ev(0.2), **jobs), None
    yield "rashba_T0", load_case("femn_rashba", mp_grid=(10, 10, 10), valence_bands=4, **jobs), None
    yield "rashba_T300", load_case("femn_rashba", mp_grid=(9, 8, 7), valence_bands=4, T_kelvin=300.0, **jobs), None
    yield "souza_T0", load_case("souza_phi0", mp_grid=(12, 12, 12), **dict(jobs, do_opt=True)), None
    yield "syn_nw16_T200", (synthetic_model(16, 3), mo.Params(mp_grid=(6, 6, 6), valence_bands=7, T_kelvin=200.0, **dict(jobs, n_hw=17))), None
    yield "syn_nw32_T300", (synthetic_model(32, 4), mo.Params(mp_grid=(3, 3, 3), valence_bands=14, T_kelvin=300.0, **dict(jobs, n_hw=16))), None
    yield "syn_nw64_T300_no_photo", (synthetic_model(64, 4), mo.Params(mp_grid=(3, 3, 2), valence_bands=30, T_kelvin=300.0,
                                                                      **dict(jobs, n_hw=16, do_opt=False))), None
    yield "femn_general_path", load_case("femn_d0.900", mp_grid=(8, 8, 8), valence_bands=4, **jobs), "large"
    # mep_sums_kernel (mepb_mep.cu): lanes whose band has no weight help their partner band (g ^ NW/2).  valence_bands above, at and
    # below NW/2, T = 0 (occupied set varies over the mesh: kubo-only owners above valence_bands) and T > 0 (nobody helps), one job only
    mj = dict(do_mep=True, do_kubo=True, do_ahc=False, do_opt=False, n_hw=1, hw_max=0.0)
    s8, s4 = synthetic_model(8, 3, tb_centers=True), synthetic_model(4, 3)
    yield "mep_nw8_nval6_T0", (s8, mo.Params(mp_grid=(7, 6, 5), valence_bands=6, **mj)), None
    yield "mep_nw8_nval2_T0_neF2", (s8, mo.Params(mp_grid=(6, 5, 7), valence_bands=2, n_eF=2, eF_min=ev(-0.8), eF_max=ev(0.6), **mj)), None
    yield "mep_nw8_nval5_T150", (s8, mo.Params(mp_grid=(5, 5, 5), valence_bands=5, T_kelvin=150.0, **mj)), None
    yield "mep_nw8_niu_only_nval3", (s8, mo.Params(mp_grid=(6, 6, 5), valence_bands=3, **dict(mj, do_kubo=False))), None
    yield "mep_nw8_kubo_only_T0", (s8, mo.Params(mp_grid=(6, 6, 5), valence_bands=4, **dict(mj, do_mep=False))), None
    yield "mep_nw4_nval1_T0", (s4, mo.Params(mp_grid=(8, 7, 6), valence_bands=1, **mj)), None
    yield "mep_nw4_nval3_T0", (s4, mo.Params(mp_grid=(8, 7, 6), valence_bands=3, **mj)), None


BS_CASES = {n: (c, f) for n, c, f in _band_sum_cases()}


@pytest.mark.parametrize("name", list(BS_CASES))
def test_band_sums_on_oracle_eigenpairs(libmepb, name, capsys):
    """kubo_ahc_tens, velo_ahc_tens, kubo_ohc_tens, photo_2nd_cond, mep_niu_IC/LC, kubo_mep_IC/LC (kubo.f90:37-198, photo.f90:44-81,
    mep_niu.f90:47-159, kubo_mep.f90:51-149) through mepb_band_sums: the production band-sum kernels and frequency sweeps fed the
    ORACLE'S OWN eigenpairs (zheevd energies, zgemm-rotated velocities).  The eigensolver is out of the comparison, so the
    near-degenerate FeMn band pairs no longer need the 20-sigma conditioning allowance of test_run_mesh_parity: every tensor must
    agree to 1e-12 of its un-cancelled magnitude sum |term|."""
    (model, params), force = BS_CASES[name]
    kp = mo.mesh_kpts(params.mp_grid)
    e, V = oracle_eigenpairs(model, params, kp)
    ref, scale, sl = band_sum_reference(model, params, e, V)
    with force_path(force):
        ctx = libmepb.Context(model, params)
    ctx.reset()
    half = e.shape[1] // 2                                   # two calls: the sums accumulate like mepb_accumulate
    ctx.band_sums(e[:, :half], V[..., :half])
    ctx.band_sums(e[:, half:], V[..., half:])
    out = ctx.get_partials().copy()
    ctx.close()
    assert out[0] == ref[0] == e.shape[1] and out[1] == ref[1]
    worst = {}
    for k, v in sl.items():
        if k in ("n_k", "n_degenerate") or out[v].size == 0:
            continue
        s = max(scale[k], 1e-300)
        worst[k] = float(np.abs(out[v] - ref[v]).max()) / s
    with capsys.disabled():
        print(f"\n[band_sums {name}] " + " ".join(f"{k}={w:.1e}" for k, w in worst.items()))
    for k, w in worst.items():
        assert w <= BAND_SUM_TOL, f"{name} {k}: {w:.3e} of sum|term| (scale {scale[k]:.3e})"


def test_update_params_sweep_equals_fresh_contexts(libmepb):
    """mepb_update_params (scripts/probe_hw.py as a library call): frequency window, smearing, Fermi level and temperature changed
    in ONE live context give bit-identical tensors to a fresh context per setting; fixed quantities are refused."""
    import copy
    from mepinterp_b200 import host
    m, p = load_case("femn_rashba", mp_grid=(8, 8, 8), n_hw=24, do_ahc=True, do_opt=True, do_kubo=True)
    sweeps = [dict(hw_min=ev(0.05), hw_max=ev(3.0)), dict(eta=ev(0.05)), dict(eF_min=ev(0.2), eF_max=ev(0.2), T_kelvin=250.0),
              dict(hw_min=ev(0.03), hw_max=ev(6.0), eta=ev(0.1), T_kelvin=0.0, kubo_tol=1e-6)]
    got = host.probe_hw(m, p, sweeps)
    q = copy.copy(p)
    for kw, res in zip(sweeps, got):
        for k, v in kw.items():
            setattr(q, k, v)
        ctx = libmepb.Context(m, copy.copy(q))
        fresh = ctx.run_mesh()
        ctx.close()
        for k, v in fresh.items():
            assert np.array_equal(np.asarray(v), np.asarray(res[k])), (kw, k)
    ctx = libmepb.Context(m, p)
    for bad in (dict(n_hw=25), dict(mp_grid=(8, 8, 9)), dict(do_mep=True), dict(n_eF=2)):
        with pytest.raises(libmepb.MepbError) as ei:
            ctx.update_params(**bad)
        assert ei.value.code == libmepb.api.ERR_ARG
    ctx.close()


def test_update_model_keeps_the_closure_flag(libmepb):
    """a model whose Hermitian closure differs from the one the context was sized for must be refused (the pair-item
    format, launch shapes and Fourier plan depend on it), on both paths; mepb_model_info reports the flag and the reason."""
    m, p = load_case("souza_phi0", mp_grid=(6, 6, 6), do_kubo=True)
    for force in (None, "large"):
        with force_path(force):
            ctx = libmepb.Context(m, p)
        info = ctx.model_info()
        assert info["herm_closed"] and info["path"] == ("general" if force else "sub")
        ref = ctx.run_mesh()
        H2 = m.H.copy()
        H2[0, 1, 0] += 1e-9                                  # one asymmetric last digit, as a 6-decimal _hr.dat can carry
        with pytest.raises(libmepb.MepbError) as ei:
            ctx.update_model(H2)
        assert ei.value.code == libmepb.api.ERR_UNSUPPORTED
        ctx.update_model(m.H)                                # the context is still usable and unchanged
        again = ctx.run_mesh()
        for k in ("ahc_tens", "mep_bands", "kubo_mep_ic"):
            assert np.array_equal(ref[k], again[k]), k
        ctx.close()
    mr, pr = load_case("femn_rashba", mp_grid=(4, 4, 4))
    ctx = libmepb.Context(mr, pr)
    info = ctx.model_info()
    ctx.close()
    assert not info["herm_closed"] and info["why"]


def test_general_path_update_model_on_device(libmepb):
    """mepb_update_model on the general path repacks the element tiles of H(R) with a device kernel: a sweep through one live
    context equals fresh contexts bit for bit (nw = 24, mesh-separable Fourier transform)."""
    p = mo.Params(mp_grid=(6, 5, 4), do_ahc=True, do_mep=True, valence_bands=9, n_hw=3, hw_max=ev(2.0), eta=ev(0.1), T_kelvin=100.0)
    models = [synthetic_model(24, 3, seed=s) for s in (1, 2, 3)]
    ctx = libmepb.Context(models[0], p)
    for i, m in enumerate(models):
        if i:
            ctx.update_model(m.H)
        live = ctx.run_mesh()
        c2 = libmepb.Context(m, p)
        fresh = c2.run_mesh()
        c2.close()
        for k, v in fresh.items():
            assert np.array_equal(np.asarray(v), np.asarray(live[k])), (i, k)
    ctx.close()


def test_device_reciprocals_accuracy(libmepb):
    """fast_rcp / fast_rcp3 / fast_rsqrt3 (mepb_internal.cuh: MUFU seed + Newton / cubic refinement, used instead of IEEE division in
    the sweeps, the MEP loops and the eigensolvers) on 10^6 values over 600 decades, both signs: relative error <= 2^-51."""
    m, p = load_case("souza_phi0", mp_grid=(2, 2, 2))
    ctx = libmepb.Context(m, p)
    rng = np.random.default_rng(12)
    x = rng.uniform(1.0, 2.0, size=1_000_000) * 10.0 ** rng.uniform(-300, 300, size=1_000_000) * rng.choice([-1.0, 1.0], size=1_000_000)
    r2, r3, rs = ctx.selftest_math(x)
    ctx.close()
    ulp = 2.0 ** -52
    assert np.abs(r2 * x - 1.0).max() <= 2 * ulp
    assert np.abs(r3 * x - 1.0).max() <= 2 * ulp
    pos = x > 0
    assert np.abs(rs[pos] * rs[pos] * x[pos] - 1.0).max() <= 4 * ulp


def test_probe_spiral_live_context_equals_fixture_run(libmepb):
    """scripts/FeMn_WX_rashba.py:loop_spiral as a library call: the in-memory generator at theta = 54.7 deg in a
    live context gives the tensors of the reference-generated fixture run directory bit for bit (same H(R); the lattice agrees to the
    8 decimals of the fixture's .latt copy), and other canting angles go through mepb_update_model."""
    from mepinterp_b200 import host
    m, p = load_case("femn_rashba", mp_grid=(8, 8, 8), n_hw=12)
    res = host.probe_spiral([20.0, 54.7], p)
    ctx = libmepb.Context(m, p)
    fix = ctx.run_mesh()
    ctx.close()
    for k in ("ahc_tens", "ahcVELO", "photoC_2nd"):
        assert np.abs(res[1][k] - fix[k]).max() <= 1e-7 * np.abs(fix[k]).max(), k
    assert np.abs(res[0]["ahcVELO"] - res[1]["ahcVELO"]).max() > 1e-3 * np.abs(fix["ahcVELO"]).max()
