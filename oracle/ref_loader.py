"""Loader for the byte-compiled reference files under oracle/_ref/ (see oracle/build_ref.py) — TEST INFRASTRUCTURE.

`load()` returns a namespace with the reference's own `MVCSMetric`, `project_points`, `batch_reproject`, `DPOLoss`,
`create_loss_strategy`, or None when oracle/_ref/ is absent or was compiled by another Python (the callers then fall back to
the numpy restatement and label the number "port")."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import json
import sys
import types
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"
_CACHE = {}


def _load_pyc(modname: str, path: Path):
    loader = importlib.machinery.SourcelessFileLoader(modname, str(path))
    spec = importlib.util.spec_from_loader(modname, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    loader.exec_module(mod)
    return mod


def available() -> bool:
    man = REF_DIR / "MANIFEST.json"
    if not man.is_file():
        return False
    try:
        return json.loads(man.read_text()).get("python") == sys.version.split()[0]
    except Exception:
        return False


def load():
    if "ns" in _CACHE:
        return _CACHE["ns"]
    ns = None
    if available():
        try:
            if "metrics" not in sys.modules:                       # mvcs.py does `from metrics.base import Metric`
                pkg = types.ModuleType("metrics")
                pkg.__path__ = []
                sys.modules["metrics"] = pkg
            _load_pyc("metrics.base", REF_DIR / "metrics_base.refbc")
            mvcs = _load_pyc("metrics.mvcs", REF_DIR / "metrics_mvcs.refbc")
            proj = _load_pyc("_ref_projection_utils", REF_DIR / "utils_projection_utils.refbc")
            loss = _load_pyc("_ref_train_loss", REF_DIR / "train_loss.refbc")
            ns = types.SimpleNamespace(MVCSMetric=mvcs.MVCSMetric, project_points=proj.project_points,
                                       batch_reproject=getattr(proj, "batch_reproject", None), DPOLoss=loss.DPOLoss,
                                       create_loss_strategy=loss.create_loss_strategy)
        except Exception:
            ns = None
    _CACHE["ns"] = ns
    return ns
