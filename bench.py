#!/usr/bin/env python
"""bench.py -- mkvtree index build throughput (suf+lcp+bwt) on B200, next to the reference mkvtree on the host.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|small]

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


def time_reference_cli(tis, sigma, flag, k, repeats=1):
    """Wall time of the UNMODIFIED reference mkvtree (oracle/_ref/mkvtree.x) on a FASTA rendering of tis; falls back
    to the C restatement (oracle port) when the binary did not travel.  Returns (seconds, kind, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_util as ou
    from vstree_b200 import synth
    if ou.have_ref():
        shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        with tempfile.TemporaryDirectory(dir=shm) as d:
            path = os.path.join(d, "in.fna")
            with open(path, "wb") as f:
                f.write(synth.to_fasta(tis, "dna" if flag == "-dna" else "protein", width=1 << 20))
            best = None
            for _ in range(repeats):
                for fn in os.listdir(d):
                    if fn.startswith("idx."):
                        os.unlink(os.path.join(d, fn))
                t0 = time.perf_counter()
                pr = subprocess.run([ou.REF_MKVTREE, "-db", "in.fna", "-indexname", "idx", flag, "-pl", str(k), "-suf",
                                     "-lcp", "-bwt"], cwd=d, capture_output=True)
                dt = time.perf_counter() - t0
                if pr.returncode != 0:
                    raise RuntimeError("reference mkvtree failed: " + pr.stderr.decode(errors="replace"))
                best = dt if best is None else min(best, dt)
        return best, "reference", 1
    t0 = time.perf_counter()
    ou.oracle_build(tis, sigma, k, bck=False, sti1=False)
    return time.perf_counter() - t0, "port", 1


def run_reference_arm(args, rank, world):
    """bench.py --impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    desc, n, sigma, flag, gen, kw = WORKLOADS[args.workload]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from vstree_b200 import load_library
    lib = load_library()  # host-side scalar math only (no GPU call)
    # bounded sample: the whole run should end within a few minutes at ~5 Msymbols/s on one core
    budget = 150.0
    sample = int(min(n, max(2_000_000, budget * 5e6 / (args.steps + args.warmup))))
    from vstree_b200 import synth
    tis = getattr(synth, gen)(sample, **kw)
    k = lib.esa_recommended_prefixlength(sigma, sample)
    times = []
    kind, cores = "reference", 1
    for i in range(args.warmup + args.steps):
        dt, kind, cores = time_reference_cli(tis, sigma, flag, k)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total / 1e6
    line = {"impl": "reference", "metric": "mkvtree index build throughput (suf+lcp+bwt)", "value": value,
            "unit": "Mbp/s" if flag == "-dna" else "Mres/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": desc, "sample": f"first-principles regeneration at n={sample} (same generator/seed)",
                       "prefixlength": int(k), "tables": list(TABLES)},
            "cpu_baseline": {"value": value, "unit": "Mbp/s" if flag == "-dna" else "Mres/s", "cores": cores,
                             "kind": kind, "sample": f"{sample} symbols per step, CLI wall time incl. parse and file "
                                                     "output to tmpfs"},
            "e2e": {"value": value, "unit": "Mbp/s" if flag == "-dna" else "Mres/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
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
    ap.add_argument("--e2e-steps", type=int, default=8)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
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
        esa.free(tabs)
        if i > 0:
            e2e_ms += dt
            e2e_steps_ms.append(round(dt, 2))
            h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]
            e2e_detail = {"h2d_ms": round(st["ms_h2d"], 3), "device_ms": round(st["ms_total"], 3),
                          "d2h_tail_ms": round(st["ms_d2h"], 3), "wall_ms": round(dt, 3),
                          "host_cpus": st["host_cpus"], "host_widen_threads": st["host_widen_threads"]}
    # the box shows sporadic host-side stalls (upload or driver calls held up for tens of ms, also in runs that do
    # nothing else): the headline is the MEDIAN step, the mean and every step are reported next to it
    e2e_mean_step = e2e_ms / args.e2e_steps
    e2e_ms_step = float(np.median(e2e_steps_ms)) if e2e_steps_ms else e2e_mean_step

    t = torch.tensor([dev_ms, e2e_ms_step], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            sec, kind, cores = time_reference_cli(tis, sigma, flag, k)
            cpu = {"value": n / sec / 1e6, "unit": "Mbp/s" if flag == "-dna" else "Mres/s", "cores": cores,
                   "kind": kind, "sample": f"the whole workload once ({n} symbols), CLI wall time incl. FASTA parse and "
                                           f"file output to tmpfs: {sec:.2f} s"}
        except Exception as e:  # the baseline is a report, not a gate
            cpu = {"value": None, "unit": "Mbp/s", "cores": 1, "kind": "reference", "sample": f"failed: {e}"}

    if rank == 0:
        unit = "Mbp/s" if flag == "-dna" else "Mres/s"
        peak, peak_src = measured_peak_gbs()
        rp = kern.get("radix_pass", {"launches": 0, "ms": 0.0, "bytes": 0.0})
        achieved = rp["bytes"] / (rp["ms"] * 1e-3) / 1e9 if rp["ms"] > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "radix_pass_traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        stage = {name: {"launches": v["launches"] // args.steps, "ms": v["ms"] / args.steps,
                        "GB/s": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None)}
                 for name, v in kern.items() if v["launches"]}
        # the whole build against the HBM roofline (north star): the algorithmic bytes of every kernel family of rank 0
        # (DESIGN.md section 5 lists the per-unit figures) over the step time, against the measured copy peak per GPU
        try:
            build_bytes = sum(v["bytes"] for name_, v in kern.items() if name_ not in ("comm",)) / args.steps
            step_s = dev_ms_max / args.steps * 1e-3
            build_roofline = {"algorithmic_bytes_per_step_rank0": build_bytes,
                              "bytes_per_symbol": build_bytes / max(1, last["shard_count"]),
                              "achieved": build_bytes / step_s / 1e9 if step_s > 0 else None, "peak": peak,
                              "unit": "GB/s per GPU",
                              "frac": (build_bytes / step_s / 1e9 / peak) if (peak and step_s > 0) else None}
        except Exception:  # a report next to the headline, never a reason to lose the line
            build_roofline = None
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
                         "share_of_step": rp["ms"] / dev_ms if dev_ms else None},
            "stages": stage,
            "build_roofline": build_roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": n / (e2e_ms_max * 1e-3) / 1e6, "unit": unit, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max, "statistic": "median over steps",
                    "mean_ms_per_step_rank0": round(e2e_mean_step, 3), "steps": args.e2e_steps,
                    "steps_ms_rank0": e2e_steps_ms, "last_step_rank0": e2e_detail},
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
