"""Launched under torchrun (one rank per GPU): slab-decomposed run vs the same run on one GPU
(julia_b200/slabcheck.py), every case, bit for bit."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from julia_b200 import slabcheck  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="", help="comma-separated subset of slabcheck.CASES")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    res = slabcheck.run_all(rank, world, local, dist, [c for c in a.cases.split(",") if c] or None)
    if rank == 0:
        print("[multi_gpu_check] " + json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
