#!/usr/bin/env python
"""bench.py -- mkvtree index build throughput (suf+lcp+bwt) on B200, next to the reference mkvtree on the host.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5|small]

One "step" = one enhanced-suffix-array build (esa_construct: suffix array, inverse suffix array, lcp/llv, bwt,
64-bit suftab) of the workload's transformed sequence.  `value` times the device-resident build (text already in
HBM, CUDA events on the library's launching stream); `e2e` times esa_build through the C ABI with HOST buffers
(pinned input, H2D + build + D2H of every demanded table inside the timed region).  The dominant kernel
(radix_pass) is timed per launch with CUDA events inside the library and reported against the measured HBM peak.
N > 1 (torchrun, one process per GPU): ONE index of N x the single-GPU sequence length is built by all ranks together
(weak scaling): the suffix array is cut into N ranges along first-symbols bins, suffixes travel to their range's rank
in an NCCL all-to-all, every rank sorts its range and derives lcp/bwt for it.  value = symbols of the
whole index / max-over-ranks device time.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, n, alphabet size, flag, generator name, generator kwargs, prefix length (auto value))
    "c2": ("synthetic 100 Mbp random DNA (BASELINE.json configs[1]), -dna -pl 11 -suf -lcp -bwt", 100_000_000, 4,
           "-dna", "random_dna", {"seed": 42}),
    "c3": ("synthetic 250 Mbp repeat-rich DNA (configs[2])", 250_000_000, 4, "-dna", "repeat_rich_dna", {"seed": 7}),
    "c4": ("synthetic 500 M-residue protein (configs[3])", 500_000_000, 20, "-protein", "random_protein", {"seed": 11}),
    "c5": ("synthetic 3 Gbp human-genome-scale DNA (configs[4]), strong scaling", 3_000_000_000, 4, "-dna",
           "human_like_dna", {"seed": 2026, "records": 24}),
    "small": ("synthetic 4 Mbp random DNA (smoke size)", 4_000_000, 4, "-dna", "random_dna", {"seed": 42}),
}
STRONG = {"c5"}  # workloads whose total size is fixed when GPUs are added
TABLES = ("suf", "lcp", "bwt")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recommended_prefixlength(sigma, n):
    """vm_recommendedprefixlength (src/kurtz/detpfxlen.c:31-61) in Python floats = the same libm doubles; used by the
    reference arm, which must not load the product library."""
    import math
    v = n / 16.0
    k = 1 if v <= sigma else int(math.floor(math.log(v) / math.log(float(sigma))))
    return max(1, k)


def survey_model_bytes(n, sigma, k, tables, rounds_active_sum, nllv, with_pack=True):
    """SURVEY.md 8d's stage table: the algorithmic bytes of one device-resident build (4-byte indices), with the
    implementation's own data-dependent figures (sum of active-set sizes over the refinement rounds, number of large lcp
    values) plugged in.  Never contains replicated work or sector over-fetch."""
    N = n + 1
    b = 0.375 if sigma <= 4 else 0.625
    st = {"S0_pack": n * (1 + b) if with_pack else 0.0,
          "S1_bck": (n * b + 16.0 * sigma ** k) if "bck" in tables else 0.0,
          "S2_initial_sort": 76.0 * N,
          "S3_refinement_rounds": 92.0 * rounds_active_sum,
          "S4_inverse_sa": 8.0 * N,
          "S5_lcp": (25.0 * N + n * b + 16.0 * nllv) if "lcp" in tables else 0.0,
          "S6_bwt": 6.0 * N if "bwt" in tables else 0.0,
          "S7_sti1": 6.0 * N if "sti1" in tables else 0.0,
          "S8_widen_suf": 12.0 * N if "suf" in tables else 0.0}
    return st


def table_checksum(first, suf, lcp, bwt, llv):
    """Order-sensitive 64-bit checksums of a range of the tables: sums (mod 2^64) of a mix of (global index, value), so
    that the ranges of a sharded build add up to the checksum of the whole table."""
    import numpy as np
    K1, K2, K3 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F), np.uint64(0x165667B19E3779F9)

    def mix(idx, val):
        x = idx * K1 + val * K2
        x ^= x >> np.uint64(31)
        x *= K3
        x ^= x >> np.uint64(29)
        return int(x.sum(dtype=np.uint64))

    out = {"suf": 0, "lcp": 0, "bwt": 0, "llv": 0}
    step = 1 << 23

    def chunk(o):
        with np.errstate(over="ignore"):
            idx = np.arange(first + o, first + o + len(suf[o:o + step]), dtype=np.uint64)
            return (mix(idx, suf[o:o + step].astype(np.uint64)), mix(idx, lcp[o:o + step].astype(np.uint64)),
                    mix(idx, bwt[o:o + step].astype(np.uint64)))

    from concurrent.futures import ThreadPoolExecutor
    workers = max(1, min(16, (os.cpu_count() or 2) // max(1, int(os.environ.get("WORLD_SIZE", "1")))))
    with ThreadPoolExecutor(workers) as ex:  # numpy releases the GIL inside the array operations
        for a, b, c in ex.map(chunk, range(0, len(suf), step)):
            out["suf"] = (out["suf"] + a) & (2 ** 64 - 1)
            out["lcp"] = (out["lcp"] + b) & (2 ** 64 - 1)
            out["bwt"] = (out["bwt"] + c) & (2 ** 64 - 1)
    with np.errstate(over="ignore"):
        pairs = np.asarray(llv, dtype=np.uint64).reshape(-1, 2)
        if len(pairs):
            out["llv"] = mix(pairs[:, 0].copy(), pairs[:, 1].copy()) & (2 ** 64 - 1)
    return out


def make_tis(workload, world=1):
    """The sequence of the whole index: every rank generates the same one (it is replicated on the GPUs)."""
    from vstree_b200 import synth
    desc, n, sigma, flag, gen, kw = WORKLOADS[workload]
    if workload not in STRONG:
        n *= world
    return getattr(synth, gen)(n, **dict(kw)), sigma, flag


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md 'clocks line'): NVML polled every few
    milliseconds from a thread (the timed region of a 100 Mbp build is shorter than nvidia-smi's own loop period);
    nvidia-smi -lms is the fallback when the NVML binding is unusable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.nvml = None
        self.stop_flag = False
        self.sm_max = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.003)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm = sorted(x[0] for x in self.samples)
            reasons = sorted(k for k, bit in names.items() if any(x[1] & bit for x in self.samples))
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smmax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def write_fasta(tis, flag, path):
    from vstree_b200 import synth
    with open(path, "wb") as f:
        f.write(synth.to_fasta(tis, "dna" if flag == "-dna" else "protein", width=1 << 20))


def tmpfs_dir():
    return "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None


REF_MKVTREE = os.path.join(ROOT, "oracle", "_ref", "mkvtree.x")
OUR_MKVTREE = os.path.join(ROOT, "vstree_b200", "bin", "mkvtree")


def time_cli(binary, workdir, flag, k, repeats=1):
    """Wall time, process start -> all files closed, of `mkvtree -db in.fna -indexname idx <flag> -pl k -suf -lcp -bwt`
    in workdir (tmpfs), best of `repeats`."""
    best = None
    for _ in range(repeats):
        for fn in os.listdir(workdir):
            if fn.startswith("idx."):
                os.unlink(os.path.join(workdir, fn))
        t0 = time.perf_counter()
        pr = subprocess.run([binary, "-db", "in.fna", "-indexname", "idx", flag, "-pl", str(k), "-suf", "-lcp", "-bwt"],
                            cwd=workdir, capture_output=True)
        dt = time.perf_counter() - t0
        if pr.returncode != 0:
            raise RuntimeError(f"{binary} failed: " + pr.stderr.decode(errors="replace"))
        best = dt if best is None else min(best, dt)
    return best


def time_reference_cli(tis, sigma, flag, k, repeats=1, also_ours=False):
    """Wall time of the UNMODIFIED reference mkvtree (oracle/_ref/mkvtree.x) on a FASTA rendering of tis; falls back
    to the C restatement (oracle port) when the binary did not travel.  Returns (seconds, kind, cores[, ours_seconds])."""
    if os.access(REF_MKVTREE, os.X_OK):
        with tempfile.TemporaryDirectory(dir=tmpfs_dir()) as d:
            write_fasta(tis, flag, os.path.join(d, "in.fna"))
            best = time_cli(REF_MKVTREE, d, flag, k, repeats)
            ours = None
            if also_ours and os.access(OUR_MKVTREE, os.X_OK):
                time_cli(OUR_MKVTREE, d, flag, k, 1)  # first run of the binary on this box: page cache, module load
                ours = time_cli(OUR_MKVTREE, d, flag, k, 2)
        return (best, "reference", 1, ours) if also_ours else (best, "reference", 1)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_util as ou
    t0 = time.perf_counter()
    ou.oracle_build(tis, sigma, k, bck=False, sti1=False)
    dt = time.perf_counter() - t0
    return (dt, "port", 1, None) if also_ours else (dt, "port", 1)


def run_reference_arm(args, rank, world):
    """bench.py --impl reference: the reference's own CPU implementation of the path on the host cores, on the SAME
    workload as the product arm (same generator, seed, length and prefix length; the whole sequence every step).  The
    product library is not loaded here.  The reference is single-threaded C: cores = 1."""
    if rank != 0:
        return
    desc, n, sigma, flag, gen, kw = WORKLOADS[args.workload]
    if args.workload not in STRONG:
        n *= world
    tis, _, _ = make_tis(args.workload, world)
    k = recommended_prefixlength(sigma, n)
    unit = "Mbp/s" if flag == "-dna" else "Mres/s"
    # every step is the whole workload; when one step is so long that W + K of them would not end within the budget
    # (the reference is superlinear: N x 100 Mbp at N = 8), warm-ups go first, then timed steps (never below one)
    budget_s = float(os.environ.get("BENCH_REFERENCE_BUDGET_S", "540"))  # the driver allows 870 s per N
    times, kind, cores = [], "reference", 1
    t_begin = time.perf_counter()
    if os.access(REF_MKVTREE, os.X_OK):
        with tempfile.TemporaryDirectory(dir=tmpfs_dir()) as d:
            write_fasta(tis, flag, os.path.join(d, "in.fna"))
            first = time_cli(REF_MKVTREE, d, flag, k)  # also the first warm-up
            warm_left = max(0, args.warmup - 1)
            if first * (args.warmup + args.steps) > budget_s:
                warm_left = 0
            for _ in range(warm_left):
                time_cli(REF_MKVTREE, d, flag, k)
            if args.warmup == 0:
                times.append(first)
            while len(times) < args.steps:
                if times and (time.perf_counter() - t_begin) + times[-1] > budget_s:
                    break
                times.append(time_cli(REF_MKVTREE, d, flag, k))
    else:
        kind = "port"
        for i in range(args.warmup + args.steps):
            dt, kind, cores = time_reference_cli(tis, sigma, flag, k)
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = n * len(times) / total / 1e6
    line = {"impl": "reference", "metric": "mkvtree index build throughput (suf+lcp+bwt)", "value": value,
            "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True,
            "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": desc, "n": n, "prefixlength": int(k), "tables": list(TABLES),
                       "steps_timed": len(times),
                       "seam": "whole CLI process (FASTA parse, sort, table files on tmpfs): the reference has no "
                               "in-memory entry that skips the parse"},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind,
                             "sample": f"the whole workload ({n} symbols) every step, {len(times)} timed steps, CLI wall "
                                       "time incl. FASTA parse and file output to tmpfs"},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~20 s host baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default: --steps)")
    ap.add_argument("--verify", default=None, choices=["full", "sums", "off"],
                    help="N > 1, outside the timed regions: 'full' compares order-sensitive checksums of the sharded tables "
                         "with a single-GPU build of the same text on rank 0 (default for the weak-scaling workloads); "
                         "'sums' only reports the checksums (default for c5: 3 Gbp on one GPU next to a live sharded "
                         "context does not fit; tools/run_config.py prints the single-GPU ones); 'off' skips both")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.e2e_steps <= 0:
        args.e2e_steps = args.steps
    if args.verify is None:
        args.verify = "sums" if args.workload in STRONG else "full"
    # libraries (NCCL's version banner) write to fd 1: keep stdout for the ONE JSON line
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from vstree_b200 import Esa, load_library
    from vstree_b200.capi import comm_unique_id

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the host arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    desc, n, sigma, flag, gen, kw = WORKLOADS[args.workload]
    if args.workload not in STRONG:
        n *= world
    lib = load_library()
    k = int(lib.esa_recommended_prefixlength(sigma, n))
    if world == 1:
        tis = np.ascontiguousarray(make_tis(args.workload, world)[0])
        pinned = torch.from_numpy(tis).pin_memory()  # the e2e leg copies from pinned host memory
    else:
        # the sequence is replicated: rank 0 generates it, the other ranks receive it (host memory of the box is
        # finite: eight generators of a 3 Gbp sequence would not fit next to the pinned tables)
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        if rank == 0:
            dev.copy_(torch.from_numpy(np.ascontiguousarray(make_tis(args.workload, world)[0])))
        dist.broadcast(dev, src=0)
        pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
        pinned.copy_(dev)
        del dev
        torch.cuda.empty_cache()
        tis = pinned.numpy()
    esa = Esa(local_rank)
    if world > 1:
        # the library's own NCCL communicator; its id travels over torch.distributed
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        esa.attach_nccl(box[0], world, rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2 (126 MB)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg -------------------------------------------------------------------------------
    esa.upload(pinned, sigma)
    for _ in range(args.warmup):
        esa.construct(k, TABLES)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    kern = {}
    last = None
    for _ in range(args.steps):
        flush.fill_(1)  # evict the previous step's tables from L2 (outside the event-timed region)
        torch.cuda.synchronize()
        st = esa.construct(k, TABLES)
        dev_ms += st["ms_total"]
        launches += st["launches"]
        for name, ks in st["kernel"].items():
            a = kern.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0.0})
            for f in a:
                a[f] += ks[f]
        last = st
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    clocks = sampler.stop()
    if world == 1:
        esa.check()  # the tables of the last timed step are consistent (device-side property checks)

    # ---- end-to-end leg: esa_build through the C ABI with host buffers --------------------------------------
    e2e_ms, h2d, d2h, e2e_detail, e2e_steps_ms = 0.0, 0, 0, None, []
    verify = None
    for i in range(1 + args.e2e_steps):
        barrier()
        t1 = time.perf_counter()
        tabs, st = esa.build(pinned, sigma, k, TABLES, want_stats=True, copy=False)
        dt = (time.perf_counter() - t1) * 1000.0
        checksum = int(tabs["suf"][:: max(1, n // 1000)].sum())  # touch the result on the host
        if i == 0 and world > 1:
            # sharded build: the ranges tile the table and together hold a permutation of 0..n
            part = torch.tensor([int(tabs["suf"].sum(dtype=np.uint64)) % (1 << 56), tabs["shard_count"]],
                                dtype=torch.int64, device="cuda")
            dist.all_reduce(part)
            N = n + 1
            assert int(part[1]) == N and int(part[0]) % (1 << 56) == (N * (N - 1) // 2) % (1 << 56), "sharded suftab is not a permutation"
            if args.verify != "off":
                # outside every timed region: the tables of the sharded build against ONE GPU building the same text.
                # Order-sensitive checksums over (global index, value) add up over the ranges.
                cs = table_checksum(tabs["shard_first"], tabs["suf"], tabs["lcp"], tabs["bwt"], tabs["llv"])
                mine = [cs["suf"], cs["lcp"], cs["bwt"], cs["llv"], tabs["largelcpvalues"]]
                gathered = [None] * world
                dist.all_gather_object(gathered, mine)
                if rank == 0:
                    tot = [sum(g[j] for g in gathered) & (2 ** 64 - 1) for j in range(5)]
                    verify = {"sharded": tot, "maxbranchdepth": tabs["maxbranchdepth"], "longest": tabs["longest"]}
        esa.free(tabs)
        if i > 0:
            e2e_ms += dt
            e2e_steps_ms.append(round(dt, 2))
            h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
            e2e_detail = {"h2d_ms": round(st["ms_h2d"], 3), "device_ms": round(st["ms_total"], 3),
                          "d2h_tail_ms": round(st["ms_d2h"], 3), "wall_ms": round(dt, 3),
                          "host_cpus": st["host_cpus"], "host_widen_threads": st["host_widen_threads"]}
    # headline: the MEAN over the timed steps; the median and every step are reported next to it (the box shows sporadic
    # host-side stalls of tens of ms in upload or driver calls)
    e2e_mean_step = e2e_ms / args.e2e_steps
    e2e_median_step = float(np.median(e2e_steps_ms)) if e2e_steps_ms else e2e_mean_step

    # model inputs that live on the ranks: large lcp values and active-set sums of the whole index
    agg = torch.tensor([float(last["nllv"]), float(last["active_sum"]), float(last["active_first"]),
                        float(sum(v["bytes"] for name_, v in kern.items() if name_ not in ("comm",)) / args.steps)],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg)
    nllv_all, active_sum_all, active_first_all, own_bytes_all = (float(x) for x in agg)

    t = torch.tensor([dev_ms, e2e_mean_step, e2e_median_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, e2e_median_max = float(t[0]), float(t[1]), float(t[2])

    if verify is not None and rank == 0 and args.verify == "sums":
        verify = {"against": None, "checksums_suf_lcp_bwt_llv_nllv": ["%016x" % x for x in verify["sharded"][:4]]
                  + [verify["sharded"][4]], "maxbranchdepth": verify["maxbranchdepth"], "longest": verify["longest"],
                  "identical": None}
    elif verify is not None and rank == 0:
        # the same text on this GPU alone (a second context without communicator); the sharded context stays alive
        with Esa(local_rank) as one:
            t_one, _ = one.build(pinned, sigma, k, TABLES, want_stats=False, copy=False)
            cs = table_checksum(0, t_one["suf"], t_one["lcp"], t_one["bwt"], t_one["llv"])
            want = [cs["suf"], cs["lcp"], cs["bwt"], cs["llv"], t_one["largelcpvalues"]]
            same = (want == verify["sharded"] and t_one["maxbranchdepth"] == verify["maxbranchdepth"]
                    and t_one["longest"] == verify["longest"])
            one.free(t_one)
        verify = {"against": "single-GPU build of the same text on rank 0 (outside the timed regions)",
                  "checksums_suf_lcp_bwt_llv_nllv": ["%016x" % x for x in want[:4]] + [want[4]], "identical": bool(same)}
        assert same, f"sharded tables differ from the single-GPU build: {verify}"

    cpu, e2e_cli = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # (at N > 1 the reference arm of the driver is the CPU figure)
        try:
            sec, kind, cores, ours_sec = time_reference_cli(tis, sigma, flag, k, also_ours=True)
            cpu = {"value": n / sec / 1e6, "unit": "Mbp/s" if flag == "-dna" else "Mres/s", "cores": cores,
                   "kind": kind, "sample": f"the whole workload once ({n} symbols), CLI wall time incl. FASTA parse and "
                                           f"file output to tmpfs: {sec:.2f} s"}
            if ours_sec is not None:
                # the like-for-like process seam: both binaries, same FASTA file and command line, process start -> files
                # closed, tmpfs (the product's time is dominated by CUDA start-up, parse and file output, not the build)
                e2e_cli = {"ours_s": round(ours_sec, 3), "reference_s": round(sec, 3), "ratio": round(sec / ours_sec, 2),
                           "value": n / ours_sec / 1e6, "unit": "Mbp/s" if flag == "-dna" else "Mres/s",
                           "command": f"mkvtree -db in.fna -indexname idx {flag} -pl {k} -suf -lcp -bwt"}
        except Exception as e:  # the baseline is a report, not a gate
            cpu = {"value": None, "unit": "Mbp/s", "cores": 1, "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        unit = "Mbp/s" if flag == "-dna" else "Mres/s"
        peak, peak_src = measured_peak_gbs()
        rp = kern.get("radix_pass", {"launches": 0, "ms": 0.0, "bytes": 0.0})
        achieved = rp["bytes"] / (rp["ms"] * 1e-3) / 1e9 if rp["ms"] > 0 else 0.0
        traffic, traffic_text = None, None
        tp = os.path.join(ROOT, "profiles", "radix_pass_traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    tj = json.load(f)
                traffic = tj.get("dram_bytes_per_launch")
                traffic_text = (tj.get("text_pass") or {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # the first pass of the initial DNA sort is a kernel of its own (keys computed from the packed text inside the
        # pass): reported beside the dominant kernel, never mixed into its average
        rt = kern.get("radix_pass_text", {"launches": 0, "ms": 0.0, "bytes": 0.0})
        first_pass = None
        if rt["launches"]:
            a = rt["bytes"] / (rt["ms"] * 1e-3) / 1e9
            first_pass = {"kernel": "radix_pass_dna_kernel", "achieved": a, "peak": peak, "unit": "GB/s",
                          "frac": a / peak if peak else None, "traffic": traffic_text,
                          "launches_per_step": rt["launches"] // args.steps,
                          "avg_launch_ms": rt["ms"] / rt["launches"],
                          "algorithmic_bytes_per_launch": rt["bytes"] / rt["launches"],
                          "share_of_step": rt["ms"] / dev_ms if dev_ms else None,
                          "note": "reads 0.25 B per symbol of packed text, writes 8 B per sorted suffix; replaces the "
                                  "emitted key/position list (8 B written + 8 B read per suffix) of round 1"}
        stage = {name: {"launches": v["launches"] // args.steps, "ms": v["ms"] / args.steps,
                        "GB/s": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None)}
                 for name, v in kern.items() if v["launches"]}
        # the whole build against the HBM roofline (north star): SURVEY.md 8d's stage bytes of the WHOLE index (with this
        # implementation's measured active-set sums and large-lcp count) over the step time, against the measured copy
        # peak of all GPUs together; next to it the sum of the kernel families' own algorithmic bytes
        try:
            step_s = dev_ms_max / args.steps * 1e-3
            stages_model = survey_model_bytes(n, sigma, k, TABLES, active_sum_all, nllv_all)
            model_bytes = sum(stages_model.values())
            build_roofline = {"model": "SURVEY.md 8d stage table", "stage_bytes": stages_model,
                              "model_bytes_per_step": model_bytes, "model_bytes_per_symbol": model_bytes / n,
                              "achieved": model_bytes / step_s / 1e9, "peak": peak * world, "unit": "GB/s (all GPUs)",
                              "frac": model_bytes / step_s / 1e9 / (peak * world),
                              "frac_of_nominal_8TBs": model_bytes / step_s / 1e9 / (8000.0 * world),
                              "own_algorithmic_bytes_per_step": own_bytes_all,
                              "own_bytes_per_symbol": own_bytes_all / n,
                              "own_frac": own_bytes_all / step_s / 1e9 / (peak * world),
                              "refinement_rounds": last["rounds"], "active_sum": active_sum_all,
                              "active_first": active_first_all, "largelcpvalues": nllv_all}
        except Exception as e:  # a report next to the headline, never a reason to lose the line
            build_roofline = {"error": str(e)}
        line = {
            "metric": "mkvtree index build throughput (suf+lcp+bwt)",
            "value": n * args.steps / (dev_ms_max * 1e-3) / 1e6,
            "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": desc, "n": n, "prefixlength": k, "tables": list(TABLES),
                       "parallelism": (f"one index sharded over {world} GPUs by suffix-array range (NCCL all-to-all of "
                                       "suffixes, owner-routed rank requests per doubling round); per-rank figures "
                                       "below are rank 0's") if world > 1 else "single GPU",
                       "shard_count_rank0": last["shard_count"], "comm_bytes_sent_rank0": last["comm_bytes_sent"],
                       "exchange": ("partition pass fused with the all-to-all: stores into the owners' buffers through "
                                    "CUDA IPC peer mappings; doubling-round rank requests/answers by NCCL send/recv")
                                   if last["comm_direct"] else ("NCCL send/recv" if world > 1 else None),
                       "l2": "256 MiB flush buffer written between timed steps; working set >> 126 MB L2",
                       "e2e_path": "esa_build: pinned tis H2D, build, D2H of suf/lcp/llv/bwt into pinned tables; the "
                                   "suftab crosses PCIe as 32-bit entries from the moment the sort is done and is "
                                   "widened to the reference's 8-byte Uint by host threads when >= 4 are free per rank",
                       "timing": "sum of per-step CUDA-event times on the library stream; wall incl. flush "
                                 f"{wall_ms / args.steps:.2f} ms/step",
                       "h0": last["h0"], "doubling_rounds": last["rounds"], "active_first": last["active_first"],
                       "active_sum": last["active_sum"], "device_bytes_peak": last["device_bytes_peak"]},
            "roofline": {"bound": "hbm", "kernel": "radix_pass_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": peak_src, "launches_per_step": rp["launches"] // args.steps,
                         "avg_launch_ms": rp["ms"] / rp["launches"] if rp["launches"] else None,
                         "algorithmic_bytes_per_launch": rp["bytes"] / rp["launches"] if rp["launches"] else None,
                         "share_of_step": rp["ms"] / dev_ms if dev_ms else None, "first_pass": first_pass},
            "stages": stage,
            "build_roofline": build_roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": n / (e2e_ms_max * 1e-3) / 1e6, "unit": unit, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max,
                    "statistic": "mean over the timed steps (max over ranks)",
                    "median_ms_per_step": round(e2e_median_max, 3), "steps": args.e2e_steps,
                    "steps_ms_rank0": e2e_steps_ms, "last_step_rank0": e2e_detail},
            "e2e_cli": e2e_cli,
            "verify": verify,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    esa.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
