"""The C++ `mepInterp` driver (mepinterp_b200/csrc/driver): input parsing checked on CPU against the oracle's readers;
the full run (GPU) against the oracle's tensors, through the files the reference would write."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import cpu_ref
from oracle import mep_oracle as mo
from tests.cases import GOLDEN, assert_parity, oracle_reference

RUNS = ["femn_d0.900", "femn_d1.100", "femn_rashba", "souza_phi0"]


@pytest.fixture(scope="module")
def driver():
    from mepinterp_b200 import build as mb_build
    return mb_build.build_driver()


def _parse_echo(txt):
    d = {}
    for ln in txt.splitlines():
        k, *v = ln.split()
        d[k] = v
    return d


@pytest.mark.parametrize("name", RUNS)
def test_driver_parses_inputs_like_the_reference_readers(driver, name):
    d = os.path.join(GOLDEN, name)
    r = subprocess.run([driver, "--dir", d, "--check-input"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    e = _parse_echo(r.stdout)
    m, p = mo.load_model(d)
    assert [int(e["nw"][0]), int(e["nw"][2])] == [m.nw, m.nR] and e["nw"][4] == p.seed_name
    g = [int(x) for x in e["mp_grid"][:3]]
    assert tuple(g) == tuple(p.mp_grid)
    flags = [int(x) for x in (e["mp_grid"][4:8] + [e["mp_grid"][9], e["mp_grid"][11], e["mp_grid"][13], e["mp_grid"][15], e["mp_grid"][17], e["mp_grid"][19]])]
    assert flags == [int(p.do_mep), int(p.do_kubo), int(p.do_ahc), int(p.do_opt), int(p.do_gauge_trafo), int(p.use_cart_velo),
                     p.valence_bands, p.n_hw, p.n_eF, int(p.do_write_mep_bands)]
    sc = [float(x) for x in e["scalars"]]
    assert sc == [p.hw_min, p.hw_max, p.eF_min, p.eF_max, p.T_kelvin, p.eta, p.kubo_tol]          # bit-identical unit conversion
    al = np.array([float(x) for x in e["a_latt"]]).reshape(3, 3, order="F")
    np.testing.assert_array_equal(al, m.a_latt)
    wf = [float(x) for x in e["wf_centers"][1:]]
    if m.wf_centers is None:
        assert wf == []
    else:
        np.testing.assert_array_equal(np.array(wf).reshape(-1, 3).T, m.wf_centers)
    Rf = np.asfortranarray(m.R).ravel(order="F")
    Hf = np.asfortranarray(m.H).ravel(order="F").view(np.float64)
    cs = [float(x) for x in e["checksums"]]
    assert abs(cs[0] - float((Rf * (np.arange(Rf.size) % 7 + 1)).sum())) <= 1e-12 * max(1.0, np.abs(Rf).sum())
    assert abs(cs[1] - float((Hf * (np.arange(Hf.size) % 13 + 1)).sum())) <= 1e-12 * np.abs(Hf).sum() * 13
    assert abs(cs[2] - float(np.abs(Hf).sum())) <= 1e-12 * np.abs(Hf).sum()


def test_driver_errors_are_loud(driver, tmp_path):
    r = subprocess.run([driver, "--dir", str(tmp_path), "--check-input"], capture_output=True, text=True)
    assert r.returncode == 1 and "input.cfg" in r.stderr
    shutil.copytree(os.path.join(GOLDEN, "souza_phi0"), tmp_path / "run")
    hr = tmp_path / "run" / "w90files" / "wf1_hr.dat"
    lines = hr.read_text().splitlines()
    hr.write_text("\n".join(lines[:-5]) + "\n")
    r = subprocess.run([driver, "--dir", str(tmp_path / "run"), "--check-input"], capture_output=True, text=True)
    assert r.returncode == 1 and "read_hr_file" in r.stderr


@pytest.mark.parametrize("name", ["souza_phi03_r", "souza_phi03_tb"])
def test_driver_reads_position_operator_inputs(driver, name):
    """read_tb_basis (file_io.f90:800-860): `<seed>_tb.dat` wins, else `_hr.dat` + `_r.dat`; the driver's C++ readers give the arrays
    the Python host reader and the oracle's reader give (checksums of r in Bohr), and jobs%do_gyro is honoured, not ignored."""
    from mepinterp_b200 import host
    run = os.path.join(GOLDEN, name)
    r = subprocess.run([driver, "--dir", run, "--check-input"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("posop")][0].split()
    assert line[1] == "1" and line[3] == "1"
    model, params, _ = host.load_run_dir(run)
    z = np.asarray(model.r).ravel(order="F")                                 # Fortran order (3,nw,nw,nR), complex interleaved
    flat = np.stack([z.real, z.imag], axis=1).ravel()
    w = (np.arange(flat.size) % 11 + 1).astype(np.float64)
    assert abs(float(line[5]) - float((flat * w).sum())) < 1e-9 and abs(float(line[6]) - float(np.abs(flat).sum())) < 1e-9
    assert params.do_gyro


@pytest.mark.gpu
@pytest.mark.parametrize("name,mesh", [("femn_d0.900", 12), ("souza_phi0", 10), ("femn_rashba", 6), ("souza_phi03_r", 8), ("souza_phi03_tb", 6)])
def test_driver_end_to_end_files(driver, tmp_path, name, mesh):
    run = tmp_path / "run"
    shutil.copytree(os.path.join(GOLDEN, name), run)
    cfg = (run / "input.cfg").read_text().splitlines()
    cfg = [("    mp_grid= %d %d %d" % (mesh, mesh, mesh)) if ln.strip().startswith("mp_grid") else ln for ln in cfg]
    cfg = [("    N_hw= 9") if ln.strip().startswith("N_hw") else ln for ln in cfg]
    (run / "input.cfg").write_text("\n".join(cfg) + "\n")
    r = subprocess.run([driver, "--dir", str(run)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    m, p = mo.load_model(str(run))
    ref, scale, sens = oracle_reference(m, p)
    out = {"n_k": ref["n_k"], "eF_idx": ref["eF_idx"], "n_degenerate": ref["n_degenerate"], "Ne": ref["Ne"]}
    for k, v in ref.items():
        if isinstance(v, np.ndarray) and k != "Ne":
            f = run / "out" / (k + ".npy")
            assert f.exists(), k
            arr = np.load(f)
            assert arr.shape == v.shape and arr.flags["F_CONTIGUOUS"]
            out[k] = arr
    assert_parity(out, ref, scale, sens, tol=1e-8, label=name)
    hw = np.load(run / "out" / "hw_lst.npy")
    assert hw.shape == (9,) and abs(hw[-1] - p.hw_max * mo.AU_TO_EV) < 1e-12
    occ = (run / "out" / "occ.dat").read_text().split()
    assert abs(float(occ[1]) - ref["Ne"][0]) < 1e-7


@pytest.mark.gpu
def test_driver_band_mode_files(driver, tmp_path):
    """plot_bands = T (band_calc.f90:30-92): `kpts` file -> raw/enK.<ki> (raw f8 stream), out/eBands.dat, out/velo/velo_Vka.<ki>."""
    run = tmp_path / "run"
    shutil.copytree(os.path.join(GOLDEN, "femn_d0.900"), run)
    cfg = (run / "input.cfg").read_text().splitlines()
    cfg = [("    plot_bands= T") if ln.strip().startswith("plot_bands") else ln for ln in cfg]
    cfg = [("    do_write_velo= T") if ln.strip().startswith("do_write_velo") else ln for ln in cfg]
    (run / "input.cfg").write_text("\n".join(cfg) + "\n")
    rng = np.random.default_rng(7)
    kpts = np.vstack([np.zeros((1, 3)), rng.uniform(-0.5, 0.5, size=(36, 3))])          # Gamma + a random path
    with open(run / "kpts", "w") as f:                                                    # kptsgen.pl format, file_io.f90:451-478
        f.write("%d 1.0\n" % len(kpts))
        for k in kpts:
            f.write("%.15f %.15f %.15f 1.0\n" % tuple(k))
    r = subprocess.run([driver, "--dir", str(run)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    m, p = mo.load_model(str(run))
    rows = np.loadtxt(run / "out" / "eBands.dat", comments="#")
    assert rows.shape == (len(kpts) * m.nw, 5)
    e_all = mo.get_wann_interp(m, kpts, do_gauge_trafo=p.do_gauge_trafo, use_cart_velo=p.use_cart_velo)[0]
    for ki, k in enumerate(kpts):
        e_ref = e_all[ki]
        e_raw = np.fromfile(run / "raw" / ("enK.%07d" % (ki + 1)), dtype="<f8")
        assert e_raw.shape == (m.nw,)
        np.testing.assert_allclose(e_raw, e_ref, rtol=0, atol=1e-10 * max(1.0, np.abs(e_ref).max()))
        blk = rows[ki * m.nw:(ki + 1) * m.nw]
        assert np.all(blk[:, 0] == ki + 1) and np.allclose(blk[:, 1:4], k, atol=5e-9) and np.allclose(blk[:, 4], e_ref, atol=5e-9)
    velo = (run / "out" / "velo" / "velo_Vka.0000001").read_text().splitlines()
    assert len([ln for ln in velo if not ln.lstrip().startswith("#")]) == 3 * m.nw * m.nw


@pytest.mark.gpu
def test_driver_devices_all_gpus(driver, tmp_path):
    """`mepInterp --devices 0,1,..` (mepb_create_multi: round-robin blocks per device, NCCL all-reduce inside the library) on every
    GPU of the box against the oracle; on a one-GPU box it degenerates to --devices 0."""
    import ctypes
    n = ctypes.c_int(0)
    ctypes.CDLL("libcudart.so.12").cudaGetDeviceCount(ctypes.byref(n))
    devs = ",".join(str(i) for i in range(max(1, min(n.value, 8))))
    run = tmp_path / "run"
    shutil.copytree(os.path.join(GOLDEN, "femn_d0.900"), run)
    cfg = (run / "input.cfg").read_text().splitlines()
    cfg = [("    mp_grid= 14 13 12") if ln.strip().startswith("mp_grid") else ln for ln in cfg]
    cfg = [("    N_hw= 9") if ln.strip().startswith("N_hw") else ln for ln in cfg]
    (run / "input.cfg").write_text("\n".join(cfg) + "\n")
    r = subprocess.run([driver, "--dir", str(run), "--devices", devs], capture_output=True, text=True, env=dict(os.environ, MEPB_CHUNK_LOG2="10"))
    assert r.returncode == 0, r.stdout + r.stderr
    m, p = mo.load_model(str(run))
    ref, scale, sens = oracle_reference(m, p)
    out = {"n_k": ref["n_k"], "eF_idx": ref["eF_idx"], "n_degenerate": ref["n_degenerate"], "Ne": ref["Ne"]}
    for k, v in ref.items():
        if isinstance(v, np.ndarray) and k != "Ne":
            out[k] = np.load(run / "out" / (k + ".npy"))
    assert_parity(out, ref, scale, sens, tol=1e-8, label="driver --devices " + devs)
