"""torchrun worker for test_nccl_sharded_if_multi_gpu: mesh sharded over the GPUs of the box, ONE all-reduce of the packed partials —
through the library's own NCCL communicator (mepb_comm_init_rank + mepb_allreduce_partials) with round-robin blocks and with
slabs, and through torch.distributed on a zero-copy view of the device buffer; every variant against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepinterp_b200 as mb                        # noqa: E402
from mepinterp_b200 import dist as mdist           # noqa: E402
from tests.cases import assert_parity, load_case, oracle_reference   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["MEPB_CHUNK_LOG2"] = "10"            # waves of 1024 k-points: 8 round-robin blocks on the 20^3 mesh
    m, p = load_case("femn_d0.900", mp_grid=(20, 20, 20), n_hw=9)
    ref, scale, sens = oracle_reference(m, p, nthreads=4)
    ctx = mb.Context(m, p, device=local)
    out_torch = mdist.run_mesh_sharded(ctx, rank, world, partition="slabs")        # no communicator yet: torch all-reduce
    assert_parity(out_torch, ref, scale, sens, tol=1e-8, label="torch all-reduce, slabs")
    mdist.init_comm(ctx, rank, world)
    assert ctx.comm_size == world
    for part in ("blocks", "slabs"):
        out = mdist.run_mesh_sharded(ctx, rank, world, partition=part)
        assert_parity(out, ref, scale, sens, tol=1e-8, label=f"library NCCL all-reduce, {part}")
        t = ctx.timings()
        assert t["accumulate_ms"] > 0.0 and t["allreduce_ms"] > 0.0
    # every rank holds the same reduced tensors (no root + broadcast)
    a = torch.from_numpy(np.ascontiguousarray(out["ahcVELO"]).view(np.float64)).cuda()
    lo, hi = a.clone(), a.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("NCCL_SHARDED_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
