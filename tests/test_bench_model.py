"""CPU tests of bench.py's flop model (no GPU, no library): the per-unit figures behind `roofline.achieved` follow the
formulation the kernels use and never credit more work than is executed."""
import importlib.util
import os

import numpy as np

from tests.cases import GOLDEN, synthetic_model
from oracle import mep_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_sparse_hermitian_model_counts_the_upper_triangle_only():
    bench = _bench()
    model, params = mo.load_model(os.path.join(GOLDEN, "femn_d0.900"))
    fm = bench.flops_model(model, params)
    H = np.asarray(model.H)
    nnz = int((H != 0).sum())
    nnz_up = int(sum((H[p, q, :] != 0).sum() for q in range(model.nw) for p in range(q + 1)))
    assert fm["sparse"] and fm["herm"] and nnz_up < nnz
    base = 48.0 * model.nw ** 3 + 60.0 * model.nw ** 2              # rotation + AHC pair sums (do_ahc, gauge transformation on)
    assert abs(fm["interp_per_k"] - (20.0 * nnz_up + base)) < 1e-6
    assert fm["sweep_ahc_per_item_hw"] == 56.0 and fm["mep_per_band"] == 0.0


def test_rashba_model_is_not_hermitian_closed_and_counts_every_entry():
    bench = _bench()
    model, params = mo.load_model(os.path.join(GOLDEN, "femn_rashba"))
    fm = bench.flops_model(model, params)
    assert not fm["herm"] and fm["sweep_ahc_per_item_hw"] == 92.0
    nnz = int((np.asarray(model.H) != 0).sum())
    assert fm["interp_per_k"] >= 20.0 * nnz


def test_mep_flops_are_per_band_on_the_small_path_and_per_kpoint_elsewhere():
    bench = _bench()
    model, params = mo.load_model(os.path.join(GOLDEN, "femn_d0.900"))
    params.do_mep = params.do_kubo = True
    fm = bench.flops_model(model, params)
    assert fm["mep_per_band"] == 99.0 * (model.nw - 1) ** 2         # mep_sums_kernel: bands with work only (measure() multiplies)
    big = synthetic_model(16, 2)
    p16 = mo.Params(mp_grid=(4, 4, 4), do_mep=True, do_kubo=True, valence_bands=7)
    fm16 = bench.flops_model(big, p16)
    assert fm16["mep_per_band"] == 0.0 and fm16["interp_per_k"] > 130.0 * 16 * 15 ** 2
