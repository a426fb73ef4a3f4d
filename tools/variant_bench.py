"""Builds one workload with several library variants (build/variants/*/libvstree_b200.so, made here with
tools/build_variant.sh) in subprocesses and prints one line per variant: build time and the radix pass family.
    python tools/variant_bench.py c2 4 name1 name2 ...      (name "default" = the in-tree library)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload, reps = sys.argv[1], sys.argv[2]
out = {}
for name in sys.argv[3:]:
    env = dict(os.environ)
    if name != "default":
        env["VSTREE_B200_LIB"] = os.path.join(ROOT, "build", "variants", name, "libvstree_b200.so")
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_run.py"), workload, reps], env=env,
                        capture_output=True, text=True, timeout=600)
    lines = [l for l in pr.stdout.splitlines() if l.startswith("rep")]
    if pr.returncode != 0 or not lines:
        print(name, "FAILED", pr.stderr[-400:].replace("\n", " | "), flush=True)
        continue
    best = None
    for l in lines[1:] or lines:
        f = dict(kv.split("=") for kv in l.split() if "=" in kv)
        ms = float(l.split("ms_total")[1].split()[0])
        if best is None or ms < best[0]:
            best = (ms, f)
    out[name] = {"ms_total": best[0], **{k: float(v) for k, v in best[1].items()}}
    print("%-28s total %.3f  radix_pass %.3f  hist %.3f  regroup %.3f  finalize %.3f  classify %.3f" % (
        name, best[0], out[name].get("radix_pass", 0), out[name].get("radix_hist", 0), out[name].get("regroup", 0),
        out[name].get("finalize", 0), out[name].get("classify", 0)), flush=True)
print(json.dumps(out))
