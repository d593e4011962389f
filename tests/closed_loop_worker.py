"""Worker of test_nccl_exchange_equals_gloo_exchange_on_two_gpus: a 256-agent closed loop sharded over WORLD_SIZE GPUs;
rank 0 prints the table checksum of every step."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_agent_pkgs_b200 import scenarios as sc  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from multi_agent_pkgs_b200.swarm import ClosedLoop
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    sw = sc.config5_random(seed=9, n_rob=256, side=50.0)
    loop = ClosedLoop(sw, world, rank, dev, 64, None)
    loop.preroll(5)
    valid = (loop.valid if world > 1 else loop.t["have_plan"]).cpu().numpy()
    if rank == 0:
        print("SUMS " + " ".join(loop.sums) + f" valid={int(valid.sum())}", flush=True)
    loop.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
