"""CPU tests of the N>1 host logic (mepinterp_b200/dist.py) with world_size-2 gloo process groups: slab partition +
ONE all-reduce of the packed partial sums + redundant finalize on every rank == single-rank run.
The per-slab sums come from the oracle here (no GPU in this container); on the B200 box the same host logic drives
libmepb (tests/test_gpu_parity.py::test_sharded_*)."""
import os
import socket

import numpy as np
import pytest

from mepinterp_b200 import dist as mdist


def test_slab_partition_covers_mesh_exactly():
    for nk in (1, 7, 64, 1000, 128 ** 3):
        for world in (1, 2, 3, 4, 8):
            cuts = [mdist.slab(r, world, nk) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nk
            for (a0, a1), (b0, b1) in zip(cuts[:-1], cuts[1:]):
                assert a1 == b0 and a0 <= a1
            assert sum(b - a for a, b in cuts) == nk
    with pytest.raises(ValueError):
        mdist.slab(2, 2, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def blocks(rank: int, world: int, nk: int, block: int):
    """[begin, end) ranges rank owns under round-robin blocks (what mepb_accumulate_strided schedules)."""
    out, b0 = [], rank * block
    while b0 < nk:
        out.append((b0, min(nk, b0 + block)))
        b0 += world * block
    return out


def test_round_robin_blocks_cover_mesh_exactly():
    for nk in (1, 7, 210, 4096, 128 ** 3):
        for world in (1, 2, 3, 8):
            for block in (1, 5, 1024, max(1, -(-nk // world))):
                seen = sorted(r for rank in range(world) for r in blocks(rank, world, nk, block))
                assert seen[0][0] == 0 and seen[-1][1] == nk
                assert all(a[1] == b[0] for a, b in zip(seen[:-1], seen[1:]))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import cpu_ref
    from tests.cases import ev, load_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, p = load_case("femn_d0.900", mp_grid=(6, 5, 7), n_hw=4, hw_max=ev(3.0), do_mep=True, do_kubo=True, n_eF=3,
                         eF_min=ev(-0.4), eF_max=ev(0.4), T_kelvin=50.0)
        h = cpu_ref.Handle(m, p)
        nk = 6 * 5 * 7
        part = h.accumulate(*mdist.slab(rank, world, nk))
        part_blocks = np.zeros_like(part)                      # the same share dealt as round-robin blocks of 16 k-points
        for b0, b1 in blocks(rank, world, nk, 16):
            h.accumulate(b0, b1, partials=part_blocks)
        tb = torch.from_numpy(part_blocks)
        mdist.allreduce_partials(tb)
        t = torch.from_numpy(part)
        mdist.allreduce_partials(t)
        out = h.finalize(t.numpy())
        out["__blocks_vs_slabs"] = float(np.abs(tb.numpy() - t.numpy()).max() / np.abs(t.numpy()).max())
        q.put((rank, {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in out.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gloo_sharded_equals_single(world):
    import torch.multiprocessing as mp
    from oracle import cpu_ref
    from tests.cases import assert_parity, ev, load_case, oracle_reference
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    m, p = load_case("femn_d0.900", mp_grid=(6, 5, 7), n_hw=4, hw_max=ev(3.0), do_mep=True, do_kubo=True, n_eF=3,
                     eF_min=ev(-0.4), eF_max=ev(0.4), T_kelvin=50.0)
    # yardsticks: un-cancelled k-sum magnitude + the oracle's own movement under a 2-ulp input change (kubo_mep on FeMn
    # is ill-conditioned: OpenBLAS' alignment-dependent rounding alone moves it by ~1e-10 between processes)
    ref, scale, sens = oracle_reference(m, p, nthreads=1)
    for r in range(world):
        out = res[r]
        assert out.pop("__blocks_vs_slabs") < 1e-9          # any partition of the mesh gives the same sums (up to reassociation)
        assert out["n_k"] == 210
        assert_parity(out, ref, scale, sens, tol=1e-12)
    # every rank holds the identical reduced result (no root + broadcast needed)
    for k, v in res[0].items():
        if isinstance(v, np.ndarray):
            np.testing.assert_array_equal(v, res[1][k])
