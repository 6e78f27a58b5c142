"""CPU: the C-ABI library builds, loads without a GPU and exports exactly what include/*.h declares."""
import ctypes
import os
import re

from videogpa_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^\s*(?:const\s+)?(?:int|size_t|const char\*|char\*)\s+\*?\s*(vgpa_[a-z0-9_]+)\s*\(", re.M)


def header_symbols():
    text = open(os.path.join(ROOT, "include", "videogpa_b200.h")).read()
    return sorted(set(DECL.findall(text)))


def test_header_declares_entry_points():
    syms = header_symbols()
    for must in ("vgpa_linear_bf16", "vgpa_attention_bf16", "vgpa_layernorm_modulate_bf16", "vgpa_mvcs_batch",
                 "vgpa_reproject_batch", "vgpa_pointcloud_filter", "vgpa_epipolar_batch", "vgpa_dpo_loss_forward",
                 "vgpa_dpo_loss_backward", "vgpa_cfg_scheduler_step", "vgpa_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/videogpa_b200.h but not exported"


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_loads_without_gpu_and_reports_errors(lib):
    assert lib.vgpa_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any CUDA call, so it is checkable on a CPU-only host
    rc = lib.vgpa_linear_bf16(None, None)
    assert rc != 0 and b"null args" in lib.vgpa_last_error()
    rc = lib.vgpa_mvcs_batch(None, None, None, 1, 2, 4, 4, 3, 3, None, 0, None, None, None, None)
    assert rc != 0 and b"null pointer" in lib.vgpa_last_error()
    assert lib.vgpa_mvcs_workspace_bytes(1, 8, 256, 256) > 0
    assert lib.vgpa_reproject_workspace_bytes(10, 504, 504) >= 10 * 504 * 504 * 8


def test_struct_layouts_match_header_sizes():
    # spot-check ctypes struct sizes against the C layout rules (8-byte pointers, natural alignment)
    assert ctypes.sizeof(_lib.AttentionArgs) == 4 * 8 + 5 * 4 + 4 + 8 * 8 + 8 + 16     # ... + lse pointer + workspace
    assert ctypes.sizeof(_lib.AttentionBwdArgs) == 9 * 8 + 5 * 4 + 4 + 16 * 8 + 8 + 8
    assert ctypes.sizeof(_lib.SchedArgs) == 7 * 8 + 8 + 4 + 7 * 4
    assert ctypes.sizeof(_lib.DpoArgs) % 8 == 0


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "videogpa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"


def test_graft_entry_build_hook_runs():
    # the driver calls __graft_entry__.build() on a CPU-only host every round
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    ge = importlib.import_module("__graft_entry__")
    ge.build()


def test_header_abi_version_matches_bindings():
    text = open(os.path.join(ROOT, "include", "videogpa_b200.h")).read()
    m = re.search(r"#define\s+VGPA_ABI_VERSION\s+(\d+)", text)
    assert m and int(m.group(1)) == _lib.ABI_VERSION
