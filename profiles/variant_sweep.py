#!/usr/bin/env python
"""Build-macro sweeps of libmbe.so in two halves (the GPU box has the tree but GPU time is scarce):

  here (no GPU):   python profiles/variant_sweep.py build MBE_UPT_BLOCKS_SMALL=6,7,8 [MORE=..]
                   -> mobile_env_gan_b200/csrc/libmbe_<tag>.so per combination (git-ignored, travels with gpurun)
  on the GPU box:  python profiles/variant_sweep.py run "mobile-medium-central-v0:65536,mobile-medium-ma-v0:131072"
                   -> one bench.py run per (variant, workload) through MBE_LIB_PATH, a table at the end
  afterwards:      python profiles/variant_sweep.py clean

The launch-bound tables in profiles/README.md came from sweeps of this form."""
import glob
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mobile_env_gan_b200", "csrc")
sys.path.insert(0, ROOT)


def build(specs):
    from mobile_env_gan_b200.csrc.build import build as build_lib

    axes = []
    for spec in specs:
        name, values = spec.split("=", 1)
        axes.append([(name, v) for v in values.split(",")])
    for combo in itertools.product(*axes):
        tag = "_".join(f"{n.replace('MBE_', '')}{v}" for n, v in combo)
        out = os.path.join(CSRC, f"libmbe_{tag}.so")
        build_lib(force=True, out=out, defines=[f"{n}={v}" for n, v in combo])
        print("built", os.path.relpath(out, ROOT))


def run(workloads, steps="1024"):
    libs = [os.path.join(CSRC, "libmbe.so")] + sorted(glob.glob(os.path.join(CSRC, "libmbe_*.so")))
    rows = []
    for lib in libs:
        for item in workloads.split(","):
            wl, envs = item.split(":")
            env = dict(os.environ, MBE_LIB_PATH=lib)
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--envs", envs, "--steps", steps,
                   "--no-cpu-baseline"]
            res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
            try:
                d = json.loads(res.stdout.strip().splitlines()[-1])
                two = (d.get("two_env_groups_in_flight") or {}).get("ms_per_step")
                rows.append((os.path.basename(lib), wl, envs, d["ms_per_step"] * 1e3, d["roofline"]["frac"],
                             two * 1e3 if two else float("nan")))
            except Exception as exc:  # noqa: BLE001
                rows.append((os.path.basename(lib), wl, envs, float("nan"), float("nan"), float("nan")))
                print("failed:", lib, wl, exc, res.stderr[-300:], file=sys.stderr)
            print("%-34s %-28s %8s  %8.2f us  frac %.3f  two-groups %.2f us" % rows[-1], flush=True)
    return rows


def clean():
    for lib in glob.glob(os.path.join(CSRC, "libmbe_*.so")):
        os.remove(lib)
        print("removed", os.path.relpath(lib, ROOT))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) >= 3 and sys.argv[1] == "run":
        run(*sys.argv[2:4])
    elif len(sys.argv) == 2 and sys.argv[1] == "clean":
        clean()
    else:
        print(__doc__)
