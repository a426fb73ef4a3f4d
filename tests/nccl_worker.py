"""Worker of tests/test_gpu_nccl.py: one process per GPU (torchrun), the library's NCCL transport.  Every rank builds
its range of one index; rank 0 also builds the whole index on its GPU alone and compares the concatenation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vstree_b200 import Esa, synth  # noqa: E402
from vstree_b200.capi import comm_unique_id, merge_shards  # noqa: E402

TABLES = ("suf", "lcp", "bwt", "bck", "sti1", "sti")  # sti1 / sti: slices of the text (reduce-scatter between the ranks)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    cases = [("human_like", synth.human_like_dna(6_000_000, seed=3, records=5), 4, 9),
             ("repeat_deep", synth.repeat_rich_dna(3_000_000, seed=8, copies=12, divergence=0.002, tandem_frac=0.05,
                                                   tandem_len=6000), 4, 8),  # lcp values beyond the deep threshold
             ("protein", synth.random_protein(2_000_000, seed=4, records=500, xfrac=0.001), 20, 4),
             ("tiny", np.array([0, 1, 2, 3, 254, 0, 1, 2, 255, 0, 1, 2, 3, 255, 254, 254, 0, 1], np.uint8), 4, 1)]
    esa = Esa(local)
    esa.attach_nccl(box[0], world, rank)
    parts_all = []
    for name, tis, sigma, k in cases:
        tabs, st = esa.build(tis, sigma, k, TABLES)
        gathered = [None] * world
        dist.all_gather_object(gathered, tabs)
        parts_all.append((name, tis, sigma, k, gathered, st))
    esa.close()
    if rank == 0:
        with Esa(local) as one:
            for name, tis, sigma, k, gathered, st in parts_all:
                want, _ = one.build(tis, sigma, k, TABLES)
                got = merge_shards(gathered)
                for t in ("suf", "lcp", "llv", "bwt", "bck", "sti1", "sti"):
                    assert np.array_equal(got[t], want[t]), f"{name}: {t} differs between {world} ranks and one GPU"
                for sc in ("longest", "maxbranchdepth", "largelcpvalues"):
                    assert got[sc] == want[sc], f"{name}: {sc}"
                assert st["kernel"]["comm"]["launches"] > 0
        print(f"NCCL-SHARDED-OK world={world}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
