"""CPU tests of the host-side mirror of the reference's file contract (mepinterp_b200/host.py): the product's own
input.cfg / _hr.dat readers against the oracle's line-by-line restatement on every golden run directory, and the
out/*.npy writer against the format the reference's save_npy emits (NPY, little-endian, fortran_order True)."""
import os

import numpy as np
import pytest

from mepinterp_b200 import host
from oracle import mep_oracle as mo
from tests.cases import GOLDEN

RUNS = ["femn_d0.900", "femn_d1.000", "femn_d1.100", "femn_rashba", "souza_phi0", "souza_phi03_r", "souza_phi03_tb"]


@pytest.mark.parametrize("name", RUNS)
def test_readers_match_oracle(name):
    d = os.path.join(GOLDEN, name)
    model, params, meta = host.load_run_dir(d)
    om, op = mo.load_model(d)
    assert (model.nw, model.nR) == (om.nw, om.nR)
    np.testing.assert_array_equal(model.R, om.R)
    np.testing.assert_array_equal(model.H, om.H)                 # bit-identical: same eV -> Hartree division
    np.testing.assert_array_equal(model.a_latt, om.a_latt)
    if om.wf_centers is None:
        assert model.wf_centers is None
    else:
        np.testing.assert_array_equal(model.wf_centers, om.wf_centers)
    for k in ("mp_grid", "do_gauge_trafo", "use_cart_velo", "do_mep", "do_kubo", "do_ahc", "do_opt", "do_write_mep_bands",
              "valence_bands", "n_hw", "n_eF", "hw_min", "hw_max", "eF_min", "eF_max", "T_kelvin", "eta", "kubo_tol"):
        assert getattr(params, k) == getattr(op, k), k
    assert meta["seed_name"] == op.seed_name
    if om.r is None:
        assert model.r is None
    else:                                                        # position operator: `_r.dat` / `_tb.dat` readers, Bohr
        np.testing.assert_array_equal(model.r, om.r)
        assert bool(params.do_gyro) == bool(op.do_gyro)


def test_cfg_dialect(tmp_path):
    """m_config syntax: comments, comma/quote separators, Fortran logicals, N_hw clamp, missing optional keys."""
    (tmp_path / "input.cfg").write_text(
        "# comment\n[jobs]\n  do_ahc = .true. ; trailing\n  do_mep= F\n[unitCell]\n  a1= 1,0,0\n  a2= 0 2 0\n  a3= '0' \"0\" 3\n  a0 = 2.0\n"
        "[wannBase]\n  seed_name= abc\n[wannInterp]\n  mp_grid= 4 5 6\n[Fermi]\n  eta_smearing= 0.1\n[Laser]\n  N_hw= 0\n  hw_max= 27.211385\n")
    p, a, wfc, meta = host.read_input_cfg(str(tmp_path / "input.cfg"))
    assert p.do_ahc and not p.do_mep and p.mp_grid == (4, 5, 6) and p.n_hw == 1 and wfc is None
    np.testing.assert_array_equal(a, np.diag([2.0, 4.0, 6.0]))
    assert p.hw_max == 1.0 and abs(p.eta - 0.1 / 27.211385) < 1e-18 and meta["seed_name"] == "abc"
    (tmp_path / "bad.cfg").write_text("[wannInterp]\n  mp_grid= 4 5\n")
    with pytest.raises((ValueError, KeyError)):
        host.read_input_cfg(str(tmp_path / "bad.cfg"))


def test_hr_reader_rejects_truncated_file(tmp_path):
    src = open(os.path.join(GOLDEN, "souza_phi0", "w90files", "wf1_hr.dat")).read().splitlines()
    (tmp_path / "t_hr.dat").write_text("\n".join(src[:-3]) + "\n")
    with pytest.raises(ValueError):
        host.read_hr_dat(str(tmp_path / "t_hr.dat"))


def test_output_writer_format(tmp_path):
    """file_io.f90:186-398 + npy.F90:1108-1136: names, shapes, fortran_order True, <f8/<c16; hw_lst in eV for n_hw > 1."""
    p = host.api.make_params(n_hw=3, hw_min=0.0, hw_max=2.0 / host.AU_TO_EV, n_eF=2, eF_min=-0.1, eF_max=0.1, do_ahc=True, do_opt=True)
    rng = np.random.default_rng(0)
    res = {"Ne": np.array([3.5, 4.25]), "ahc_tens": np.asfortranarray(rng.normal(size=(3, 3))),
           "ahcVELO": np.asfortranarray(rng.normal(size=(3, 3, 3)) + 1j * rng.normal(size=(3, 3, 3))),
           "photoC_2nd": np.asfortranarray(rng.normal(size=(3, 3, 3, 3, 2)) + 0j)}
    out = host.write_outputs(str(tmp_path), p, res)
    assert sorted(os.listdir(out)) == ["ahcVELO.npy", "ahc_tens.npy", "hw_lst.npy", "occ.dat", "photoC_2nd.npy"]
    with open(os.path.join(out, "ahcVELO.npy"), "rb") as fh:
        head = fh.read(128)
    assert head[:6] == b"\x93NUMPY" and b"'fortran_order': True" in head and b"<c16" in head
    np.testing.assert_array_equal(np.load(os.path.join(out, "ahc_tens.npy")), res["ahc_tens"])
    np.testing.assert_allclose(np.load(os.path.join(out, "hw_lst.npy")), [0.0, 1.0, 2.0], atol=1e-14)
    rows = open(os.path.join(out, "occ.dat")).read().splitlines()
    assert len(rows) == 2 and rows[0].split() == ["-2.72113850", "3.50000000"]


def test_souza_generator_equals_the_reference_generated_fixture():
    """mepinterp_b200.models.souza_tb against tests/golden/souza_phi0, which make_fixtures.py produced by
    running the reference's scripts/souza_tb_model.py + tb_input_writer.py unchanged: bit-identical H(R), R, lattice."""
    import numpy as np
    from mepinterp_b200 import host, models
    from tests.cases import GOLDEN
    ref, _, _ = host.load_run_dir(os.path.join(GOLDEN, "souza_phi0"))
    gen = models.souza_tb(0.0)
    assert gen.nw == ref.nw == 8 and gen.nR == ref.nR == 7 and gen.wf_centers is None and ref.wf_centers is None
    assert np.array_equal(np.asarray(gen.R), np.asarray(ref.R))
    assert np.array_equal(np.asarray(gen.a_latt), np.asarray(ref.a_latt))
    assert np.array_equal(np.asarray(gen.H), np.asarray(ref.H))
    other = models.souza_tb(0.35)                                   # phi enters exactly one hopping phase (x bond of site 1)
    changed = np.argwhere(np.asarray(other.H) != np.asarray(gen.H))
    assert {(int(a), int(b)) for a, b, _ in changed} == {(0, 1), (1, 0)} and {int(r) for _, _, r in changed} == {3}
    Hh = np.asarray(other.H)
    for r, rm in ((0, 6), (1, 5), (2, 4), (3, 3)):                  # Hermitian closure H(-R) = H(R)^dagger
        assert np.array_equal(Hh[:, :, r], Hh[:, :, rm].conj().T)


def test_femn_rashba_generator_equals_the_reference_generated_fixture():
    """mepinterp_b200.models.femn_rashba_tb (BASELINE config 4) against tests/golden/femn_rashba, which
    make_fixtures.py produced by running the reference's scripts/FeMn_WX_rashba.py unchanged: bit-identical H(R), R, centres; the
    strained lattice bit-identical to the full-precision values the reference generator wrote into its own input.cfg
    (input.cfg.as_generated; the fixture's input.cfg carries the 8-decimal copy of the generator's .latt file)."""
    from mepinterp_b200 import models
    ref, _, _ = host.load_run_dir(os.path.join(GOLDEN, "femn_rashba"))
    gen = models.femn_rashba_tb()
    assert gen.nw == ref.nw == 8 and gen.nR == ref.nR == 27
    assert np.array_equal(np.asarray(gen.R), np.asarray(ref.R))
    assert np.array_equal(np.asarray(gen.H), np.asarray(ref.H))
    assert np.array_equal(np.asarray(gen.wf_centers), np.asarray(ref.wf_centers))
    assert np.abs(np.asarray(gen.a_latt) - np.asarray(ref.a_latt)).max() < 1e-8
    rows = [ln for ln in open(os.path.join(GOLDEN, "femn_rashba", "input.cfg.as_generated")) if ln.strip().startswith(("a1=", "a2=", "a3="))]
    full = np.array([[float(x) for x in ln.split("=")[1].split()] for ln in rows])
    assert np.array_equal(np.asarray(gen.a_latt), full)
    other = models.femn_rashba_tb(theta_deg=20.0)                  # the canting angle only enters the on-site exchange block
    changed = {int(r) for _, _, r in np.argwhere(np.asarray(other.H) != np.asarray(gen.H))}
    assert changed == {13}                                          # R = (0, 0, 0) is the 14th of the 27 cells
