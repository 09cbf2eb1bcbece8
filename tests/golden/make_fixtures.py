#!/usr/bin/env python
"""Regenerates tests/golden/* from the read-only reference checkout (run in the build container only).

    python tests/golden/make_fixtures.py [/root/reference]

* femn_d{0.900,1.000,1.100}/ : the reference's own FeMn 3q tight-binding inputs (config/FeMn_TB/delta*/w90files/
  FeMn_hr.dat + *.latt) and its golden AHC outputs (ahc_tens.npy, ahcVELO_sol*.npy), with an input.cfg carrying the
  parameters of tests/test_FeMn_AHE.f90:127-217 (wf_centers :144-150, 64^3, n_hw=121 over 0-6 eV, eta=0.1 eV, T=0).
* souza_phi0/   : scripts/souza_tb_model.py:get_souza_tb(0.0) written through scripts/tb_input_writer.py:write_hr_file
  (BASELINE config 1; Wannier FT convention, no position operator).
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
