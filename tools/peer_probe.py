"""Sharded build through the in-process transport with one rank per DEVICE (peer access, no IPC): per-rank stage times.
    python tools/peer_probe.py G [n_per_gpu]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vstree_b200 import synth  # noqa: E402
from vstree_b200.capi import EsaGroup  # noqa: E402

G = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
tis = synth.random_dna(n * G, seed=42)
with EsaGroup(list(range(G))) as grp:
    grp.upload(tis, 4)
    for rep in range(3):
        t0 = time.perf_counter()
        stats = grp.construct(12 if n * G > 268_000_000 else 11)
        wall = (time.perf_counter() - t0) * 1e3
    for r, st in enumerate(stats):
        print(f"G={G} rank {r} direct={st['comm_direct']} ms_total={st['ms_total']:.2f} wall={wall:.1f} shard={st['shard_count']} " +
              " ".join(f"{k}={v['ms']:.2f}" for k, v in st["kernel"].items() if v["launches"]))
