"""Recipe for oracle/_ref/: the reference's OWN scorer files, compiled where they lie (TEST INFRASTRUCTURE).

    python -m oracle.build_ref          # in the build container, where /root/reference is mounted

The reference is Python, so "compiling its own few source files" is `py_compile`: the files below are byte-compiled
straight from /root/reference into oracle/_ref/*.refbc. No reference SOURCE enters the repository: oracle/_ref/ is
git-ignored (built artefact, like our own .so files) but not gpurun-ignored, so the bytecode travels to the GPU box and
`bench.py` can time the reference's own implementation on that box's host cores (`cpu_baseline.kind = "reference"`,
`--impl reference`), which /root/reference itself cannot do because it does not exist there. Only bench.py's CPU legs and
tests/ load these modules (oracle/ref_loader.py); nothing under videogpa_b200/ does.

Files:
    metrics/base.py               Metric ABC that mvcs.py imports
    metrics/mvcs.py               MVCSMetric.compute                      (metrics/mvcs.py:12-114)
    utils/projection_utils.py     project_points / batch_reproject        (utils/projection_utils.py:12-101)
    train/loss.py                 DPOLoss / create_loss_strategy          (train/loss.py:25-155)
"""
from __future__ import annotations

import json
import py_compile
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
FILES = {
    "metrics_base": "metrics/base.py",
    "metrics_mvcs": "metrics/mvcs.py",
    "utils_projection_utils": "utils/projection_utils.py",
    "train_loss": "train/loss.py",
}


def build(verbose: bool = False) -> bool:
    """Byte-compile the reference files into oracle/_ref/. Returns False (and leaves _ref untouched) without /root/reference."""
    if not REF.is_dir():
        return False
    OUT.mkdir(parents=True, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "files": {}}
    for name, rel in FILES.items():
        src = REF / rel
        dst = OUT / f"{name}.refbc"
        py_compile.compile(str(src), cfile=str(dst), dfile=f"<reference>/{rel}", doraise=True)
        manifest["files"][name] = rel
        if verbose:
            print(f"{src} -> {dst}")
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref built" if ok else "no /root/reference here: nothing built")
