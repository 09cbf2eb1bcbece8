#!/usr/bin/env python
"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                    # noqa: E402
import mepinterp_b200 as mb                           # noqa: E402
from mepinterp_b200 import host, models, api          # noqa: E402

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run(model, params, tag):
    ctx = mb.Context(model, params, device=0)
    out = ctx.run_mesh()
    ctx.close()
    print(tag, "n_k", out["n_k"], "ok", flush=True)


m, p, _ = host.load_run_dir(os.path.join(G, "femn_d0.900"))
p.mp_grid, p.n_hw, p.do_mep, p.do_kubo, p.do_opt = (5, 4, 3), 40, True, True, True
p.hw_min = 0.002
run(m, p, "femn all jobs (eig_hq, bands_sub, mep_sums, DMMA sweeps)")
m, p, _ = host.load_run_dir(os.path.join(G, "femn_rashba"))
p.mp_grid, p.n_hw = (4, 3, 3), 24
run(m, p, "rashba general items + photo")
m, p, _ = host.load_run_dir(os.path.join(G, "souza_phi0"))
p.mp_grid = (4, 4, 4)
run(m, p, "souza (n_hw = 1: FMA sweeps)")
for nw, rmax2, grid in ((64, 5, (3, 2, 2)), (32, 3, (3, 3, 2)), (13, 3, (3, 2, 2))):
    mm = models.synthetic_tb(nw, rmax2)
    pp = api.make_params(mp_grid=grid, do_ahc=True, do_mep=True, do_kubo=True, valence_bands=nw // 3, T_kelvin=300.0, n_hw=3,
                         hw_min=0.002, hw_max=0.1, eta=0.004)
    run(mm, pp, f"general path nw={nw}")
