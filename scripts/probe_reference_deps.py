"""Probe: can the UNMODIFIED reference NAT / mel path run on this machine?

SURVEY.md §8c(iii) / VERDICT r1 item 1: the only place the real `AcousticModel.inference`,
`DurationModel`, `AcousticModel.__call__` and `MelFilter` could produce golden vectors is a box with
jax + dm-haiku + librosa. This script records, for the container it runs in (build container or GPU box):
which of the reference's third-party imports resolve, what any offline wheel source holds, and whether a
reference install exists under baseline/_ref.  Output: one JSON document on stdout.
"""
from __future__ import annotations

import glob
import importlib
import json
import os
import platform
import subprocess
import sys

MODULES = ["jax", "jaxlib", "haiku", "optax", "librosa", "soundfile", "textgrid", "fire", "einops",
           "numba", "scipy", "torch", "torchaudio", "numpy", "tensorflow", "flax", "chex", "jmp"]


def main() -> None:
    out = {"python": sys.version.split()[0], "platform": platform.platform(), "host": platform.node(),
           "cpu_count": os.cpu_count(), "modules": {}, "wheel_sources": {}, "pip_download": {}}
    for m in MODULES:
        try:
            mod = importlib.import_module(m)
            out["modules"][m] = getattr(mod, "__version__", "present")
        except Exception as e:  # noqa: BLE001
            out["modules"][m] = f"ABSENT ({type(e).__name__})"
    for d in ["/opt/wheelhouse", "/wheelhouse", "/opt/wheels", os.path.expanduser("~/.cache/pip")]:
        if os.path.isdir(d):
            names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(d, "**", "*.whl"), recursive=True))
            hits = [n for n in names if any(k in n.lower() for k in ("jax", "haiku", "librosa", "optax", "soundfile"))]
            out["wheel_sources"][d] = {"n_wheels": len(names), "matching": hits}
        else:
            out["wheel_sources"][d] = None
    for pkg in ["jax", "dm-haiku", "librosa"]:
        r = subprocess.run([sys.executable, "-m", "pip", "download", "--no-deps", "-d", "/tmp/_probe_dl",
                            "--find-links", "/opt/wheelhouse", "--no-index", pkg],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        tail = [l for l in r.stdout.strip().splitlines() if l.strip()][-1:] or [""]
        out["pip_download"][pkg] = {"rc": r.returncode, "tail": tail[0][:200]}
    # is there any route to an index at all? (expected: no network)
    r = subprocess.run([sys.executable, "-m", "pip", "download", "--no-deps", "-d", "/tmp/_probe_dl", "--timeout", "5",
                        "--retries", "0", "dm-haiku"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    out["pip_index"] = {"rc": r.returncode, "tail": ([l for l in r.stdout.strip().splitlines() if l.strip()][-1:] or [""])[0][:200]}
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    out["baseline_ref"] = sorted(os.listdir(ref))[:20] if os.path.isdir(ref) else None
    out["reference_tree_present"] = os.path.isdir("/root/reference")
    out["verdict"] = ("reference NAT/mel path importable"
                      if all("ABSENT" not in out["modules"][m] for m in ("jax", "haiku", "librosa"))
                      else "reference NAT/mel path NOT importable: jax / dm-haiku / librosa absent and not installable offline")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
