"""Subprocess worker of the full-size cross-check tests: run one golden case through the C ABI with the environment the
parent set (MEPB_SMALL_EIG=jacobi, MEPB_SWEEP_FMA=1 select the independent kernel variants; they are read once per process)
and dump the tensors.  usage: run_case_worker.py <golden case> <json overrides> <out.npz>"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepinterp_b200 as mb                         # noqa: E402
from tests.cases import load_case                   # noqa: E402


def main():
    name, over, out = sys.argv[1], json.loads(sys.argv[2]), sys.argv[3]
    if "mp_grid" in over:
        over["mp_grid"] = tuple(over["mp_grid"])
    m, p = load_case(name, **over)
    ctx = mb.Context(m, p, device=0)
    res = ctx.run_mesh()
    launches = ctx.kernel_launches
    ctx.close()
    assert launches > 0
    np.savez(out, **{k: np.asarray(v) for k, v in res.items()})


if __name__ == "__main__":
    main()
