"""TEST/BENCH INFRASTRUCTURE ONLY -- recipe that installs the UNMODIFIED reference into ``oracle/_ref/``.

The reference (a fork of mobile-env 2.0.1) is pure Python, so "building" it is a pip install of its
own ``setup.py`` package from where it lies under ``/root/reference`` (read-only: pip builds the wheel
from a copy under the system temp directory) with ``--no-deps`` (its rendering dependencies --
matplotlib, pygame, shapely, svgpath2mpl -- are not in the offline wheelhouse; ``oracle/ref_harness.py``
puts inert stand-ins into ``sys.modules`` and restates ``shapely.geometry.Point.distance``, the only
one with arithmetic on the path).  Output goes only into ``oracle/_ref/`` -- git-ignored, but shipped
to the GPU box with the snapshot -- so that ``bench.py --impl reference`` and ``cpu_baseline`` can time
the reference's own ``MComCore.step`` (mobile_env/core/base.py:230-296) on the box's host cores
(``cpu_baseline.kind == "reference"``).  No reference source is copied into the repository's history.

    python -m oracle.build_ref          # done by __graft_entry__.build() where /root/reference exists
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("MBE_REFERENCE_SRC", "/root/reference")
REF_OUT = os.path.join(ROOT, "oracle", "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(REF_OUT, "mobile_env", "core", "base.py"))


def _newest(path: str) -> float:
    return max((os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(path) for f in fs if f.endswith(".py")),
               default=0.0)


def build_ref(force: bool = False) -> str | None:
    """Installs the reference into oracle/_ref (returns the path), or returns the existing install /
    None when ``/root/reference`` is absent (the GPU box only uses what travelled with the snapshot)."""
    if not os.path.isdir(os.path.join(REF_SRC, "mobile_env", "core")):
        return REF_OUT if installed() else None
    if installed() and not force and _newest(os.path.join(REF_OUT, "mobile_env")) >= _newest(
            os.path.join(REF_SRC, "mobile_env")):
        return REF_OUT
    with tempfile.TemporaryDirectory(prefix="mbe_ref_") as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns("*.ipynb", ".git"))
        shutil.rmtree(REF_OUT, ignore_errors=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", REF_OUT, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n" + res.stderr[-2000:])
    if not installed():
        raise RuntimeError(f"reference install left no mobile_env package under {REF_OUT}")
    return REF_OUT


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
