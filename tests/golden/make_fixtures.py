#!/usr/bin/env python
"""Regenerates tests/golden/* from the read-only reference checkout (run in the build container only).

    python tests/golden/make_fixtures.py [/root/reference]

* femn_d{0.900,1.000,1.100}/ : the reference's own FeMn 3q tight-binding inputs (config/FeMn_TB/delta*/w90files/
  FeMn_hr.dat + *.latt) and its golden AHC outputs (ahc_tens.npy, ahcVELO_sol*.npy), with an input.cfg carrying the
  parameters of tests/test_FeMn_AHE.f90:127-217 (wf_centers :144-150, 64^3, n_hw=121 over 0-6 eV, eta=0.1 eV, T=0).
* souza_phi0/   : scripts/souza_tb_model.py:get_souza_tb(0.0) written through scripts/tb_input_writer.py:write_hr_file
  (BASELINE config 1; Wannier FT convention, no position operator).
* souza_phi03_r/: the same generator at phi = 0.3 WITH its position operator: get_souza_Pos through write_r_file
  (scripts/tb_input_writer.py:75-97) -> wf1_r.dat (on-site positions of the eight orbitals); all jobs on, use_cart_velo = T.
* souza_phi03_tb/: the same model as ONE wannier90-format wf1_tb.dat (lattice + H(R) + r(R); written here from the two files above —
  the reference has a reader for the format, file_io.f90:555-658, but no writer).
* femn_rashba/  : scripts/FeMn_WX_rashba.py run unchanged (theta=54.7 deg; BASELINE config 4), its wf1_hr.dat and the
  input.cfg it emits, with the jobs switched to do_opt/do_ahc as SURVEY.md 8(d) states.
The generators are executed from the reference tree; nothing of their source is copied here — only their outputs.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

CFG = """[jobs]
    plot_bands= False
    debug_mode= False
    R_vect_float= {rfloat}
    do_write_velo= False
    do_mep= {do_mep}
    do_kubo= {do_kubo}
    do_ahc= {do_ahc}
    do_opt= {do_opt}
    do_gyro= False
[unitCell]
    a1= {a1}
    a2= {a2}
    a3= {a3}
    a0= 1.0
[wannBase]
    seed_name= {seed}
    N_wf= {n_wf}
{centers}[wannInterp]
    use_cart_velo= False
    doGaugeTrafo= True
    mp_grid= {mp} {mp} {mp}
[Fermi]
    N_eF= 1
    eF_min= 0.0
    eF_max= 0.0
    Tkelvin= 0.0
    eta_smearing= 0.1
    kuboTol= 1e-05
[MEP]
    valence_bands= {nval}
    do_write_mep_bands= True
[Laser]
    N_hw= {n_hw}
    hw_min= {hw_min}
    hw_max= {hw_max}
"""


def fmt(v):
    return " ".join(repr(float(x)) for x in v)


def centers_block(c):
    if c is None:
        return ""
    return "".join("    wf_centers_%s= %s\n" % (ax, fmt(c[i])) for i, ax in enumerate("xyz"))


def write_case(name, hr_src, seed, a, centers, **kw):
    d = os.path.join(HERE, name)
    os.makedirs(os.path.join(d, "w90files"), exist_ok=True)
    shutil.copyfile(hr_src, os.path.join(d, "w90files", seed + "_hr.dat"))
    args = dict(rfloat="False", do_mep="False", do_kubo="False", do_ahc="False", do_opt="False", seed=seed,
                n_wf=0 if centers is None else len(centers[0]), centers=centers_block(centers), mp=16, nval=4, n_hw=1,
                hw_min=0.0, hw_max=0.0, a1=fmt(a[0]), a2=fmt(a[1]), a3=fmt(a[2]))
    args.update(kw)
    with open(os.path.join(d, "input.cfg"), "w") as fh:
        fh.write(CFG.format(**args))
    return d


def main():
    femn_centers = [[0.0, 0.5, 0.5, 0.0] * 2, [0.0, 0.5, 0.0, 0.0] * 2, [0.0, 0.0, 0.5, 0.5] * 2]   # test_FeMn_AHE.f90:144-150
    for delta, tag in (("0.900", "09"), ("1.000", "10"), ("1.100", "11")):
        src = os.path.join(REF, "config", "FeMn_TB", "delta" + delta)
        latt = np.loadtxt(glob.glob(os.path.join(src, "*.latt"))[0])
        d = write_case("femn_d" + delta, os.path.join(src, "w90files", "FeMn_hr.dat"), "FeMn", latt, femn_centers,
                       do_ahc="True", mp=64, n_hw=121, hw_min=0.0, hw_max=6.0, rfloat="True")
        shutil.copyfile(os.path.join(src, "ahc_tens.npy"), os.path.join(d, "ahc_tens.npy"))
        shutil.copyfile(os.path.join(src, "ahcVELO_sol%s.npy" % tag), os.path.join(d, "ahcVELO.npy"))
    # ---- Souza model through the reference's own generator + writer
    tmp = tempfile.mkdtemp()
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from souza_tb_model import get_souza_tb\nfrom tb_input_writer import write_hr_file\n"
            "nWfs, nrpts, tHopp, rHopp, ax, ay, az, a0 = get_souza_tb(0.0)\nwrite_hr_file('wf1', nWfs, nrpts, tHopp)\n"
            % os.path.join(REF, "scripts"))
    subprocess.check_call([sys.executable, "-W", "ignore", "-c", code], cwd=tmp, stdout=subprocess.DEVNULL)
    write_case("souza_phi0", os.path.join(tmp, "wf1_hr.dat"), "wf1", np.diag([2.0, 2.0, 2.0]), None,
               do_ahc="True", do_mep="True", mp=32, nval=2, n_hw=1, hw_min=0.0, hw_max=0.0)
    # ---- Souza model with the position operator the generator provides (write_r_file), phi = 0.3
    tmp3 = tempfile.mkdtemp()
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from souza_tb_model import get_souza_tb\nfrom tb_input_writer import write_hr_file, write_r_file\n"
            "nWfs, nrpts, tHopp, rHopp, ax, ay, az, a0 = get_souza_tb(0.3)\nwrite_hr_file('wf1', nWfs, nrpts, tHopp)\n"
            "write_r_file('wf1', nWfs, rHopp)\n" % os.path.join(REF, "scripts"))
    subprocess.check_call([sys.executable, "-W", "ignore", "-c", code], cwd=tmp3, stdout=subprocess.DEVNULL)
    pos_jobs = dict(do_ahc="True", do_mep="True", do_kubo="True", do_opt="True", mp=12, nval=2, n_hw=6, hw_min=0.05, hw_max=4.0)
    d = write_case("souza_phi03_r", os.path.join(tmp3, "wf1_hr.dat"), "wf1", np.diag([2.0, 2.0, 2.0]), None, **pos_jobs)
    shutil.copyfile(os.path.join(tmp3, "wf1_r.dat"), os.path.join(d, "w90files", "wf1_r.dat"))
    for name in ("souza_phi03_r",):
        cfg = os.path.join(HERE, name, "input.cfg")
        txt = open(cfg).read().replace("use_cart_velo= False", "use_cart_velo= True").replace("do_gyro= False", "do_gyro= True")
        txt = txt.replace("Tkelvin= 0.0", "Tkelvin= 300.0")
        open(cfg, "w").write(txt)
    # the same model in the wannier90 _tb.dat format (Angstrom lattice + hoppings in eV + positions in Angstrom)
    d2 = os.path.join(HERE, "souza_phi03_tb")
    os.makedirs(os.path.join(d2, "w90files"), exist_ok=True)
    shutil.copyfile(os.path.join(d, "input.cfg"), os.path.join(d2, "input.cfg"))
    hr = open(os.path.join(tmp3, "wf1_hr.dat")).read().splitlines()
    rr = open(os.path.join(tmp3, "wf1_r.dat")).read().splitlines()
    nw, nR = int(hr[1].split()[0]), int(hr[2].split()[0])
    ndl = -(-nR // 15)
    hrows = [ln.split() for ln in hr[3 + ndl:]]
    rrows = [ln.split() for ln in rr[2:]]
    with open(os.path.join(d2, "w90files", "wf1_tb.dat"), "w") as fh:
        fh.write("# written by tests/golden/make_fixtures.py from the reference generator's hr/r output\n")
        for row in np.diag([2.0, 2.0, 2.0]) * 0.52917721092:
            fh.write("  %.12f  %.12f  %.12f\n" % tuple(row))
        fh.write("%d\n%d\n" % (nw, nR))
        for i in range(ndl):
            fh.write(" ".join(["1"] * min(15, nR - 15 * i)) + "\n")
        for sc in range(nR):
            blk = hrows[sc * nw * nw:(sc + 1) * nw * nw]
            fh.write("\n%s %s %s\n" % tuple(blk[0][0:3]))
            for t in blk:
                fh.write("%s %s %s %s\n" % (t[3], t[4], t[5], t[6]))
        for sc in range(nR):
            blk = rrows[sc * nw * nw:(sc + 1) * nw * nw]
            fh.write("\n%s %s %s\n" % tuple(blk[0][0:3]))
            for t in blk:
                fh.write(" ".join([t[3], t[4]] + t[5:11]) + "\n")
    # ---- FeMn Rashba model: run the generator script unchanged
    tmp2 = tempfile.mkdtemp()
    env = dict(os.environ, PYTHONPATH=os.path.join(REF, "scripts"))
    subprocess.check_call([sys.executable, "-W", "ignore", os.path.join(REF, "scripts", "FeMn_WX_rashba.py")], cwd=tmp2, env=env,
                          stdout=subprocess.DEVNULL)
    run = glob.glob(os.path.join(tmp2, "new_spiral_run_*", "theta54.700"))[0]
    latt = np.loadtxt(glob.glob(os.path.join(run, "w90files", "*.latt"))[0])
    cfg_lines = open(os.path.join(run, "input.cfg")).read().splitlines()
    cen = [[float(x) for x in ln.split("=")[1].split()] for ln in cfg_lines if "wf_centers_" in ln]
    write_case("femn_rashba", os.path.join(run, "w90files", "wf1_hr.dat"), "wf1", latt, cen,
               do_ahc="True", do_opt="True", mp=16, nval=4, n_hw=200, hw_min=0.03, hw_max=6.0)
    shutil.copyfile(os.path.join(run, "input.cfg"), os.path.join(HERE, "femn_rashba", "input.cfg.as_generated"))
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
