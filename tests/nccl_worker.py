"""torchrun worker for test_nccl_sharded_if_multi_gpu: mesh sharded over 2 GPUs, one NCCL all-reduce of the packed partials."""
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
    m, p = load_case("femn_d0.900", mp_grid=(20, 20, 20), n_hw=9)
    ctx = mb.Context(m, p, device=local)
    out = mdist.run_mesh_sharded(ctx, rank, world)
    ctx.close()
    ref, scale, sens = oracle_reference(m, p, nthreads=4)
    assert_parity(out, ref, scale, sens, tol=1e-8)
    dist.barrier()
    if rank == 0:
        print("NCCL_SHARDED_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
