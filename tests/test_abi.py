"""The C-ABI library builds, loads and exports every symbol include/css_b200.h declares, with the argument counts the
ctypes binding assumes.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "css_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|unsigned\s+long\s+long|const\s+char\s*\*)\s+(css_\w+)\s*\(([^)]*)\)\s*;", src):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_header_declares_the_path():
    fns = header_functions()
    for name in ["css_sim_map", "css_upsample_label_fuse", "css_select", "css_rep_pass", "css_class_stats", "css_proto_ema", "css_sample",
                 "css_score_ce", "css_grad_scatter", "css_threshold_glue", "css_version", "css_last_error"]:
        assert name in fns


def test_library_builds_loads_and_exports_every_symbol():
    from css_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    fns = header_functions()
    assert set(fns) == set(_lib.SIGNATURES), "ctypes SIGNATURES and the header disagree"
    for name, nargs in fns.items():
        assert hasattr(lib, name), f"{name} not exported"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args"
    assert _lib.load().css_version() == 100
    assert _lib.load().css_select_tiles(257) == 2


def test_argument_errors_are_reported_without_a_gpu():
    from css_b200 import _lib
    lib = _lib.load()
    rc = lib.css_sim_map(None, 0, None, None, 1, 21, 256, 4, 4, 0, 0.5, None, None)
    assert rc == -1 and b"null" in lib.css_last_error()
    with pytest.raises(RuntimeError, match="css_sim_map"):
        _lib.check(rc, "css_sim_map")
    one = ctypes.c_void_p(16)
    assert lib.css_sim_map(one, 0, one, one, 1, 21, 128, 4, 4, 0, 0.5, one, None) == -2      # D != 256
    assert lib.css_sim_map(one, 0, one, one, 1, 33, 256, 4, 4, 0, 0.5, one, None) == -2      # C > 32
    assert lib.css_sim_map(one, 7, one, one, 1, 21, 256, 4, 4, 0, 0.5, one, None) == -3      # unknown rep dtype


def test_ops_refuse_cpu_tensors():
    import torch
    import css_b200
    with pytest.raises(RuntimeError, match="CUDA"):
        css_b200.ops.cos_sim_map(torch.zeros(1, 256, 4, 4), torch.zeros(21, 256))
    crit = css_b200.Contrast_Loss(num_queries=4, num_negatives=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        crit(torch.zeros(1, 256, 4, 4), torch.zeros(1, 3, 4, 4), torch.zeros(1, 1, 4, 4), torch.zeros(1, 3, 4, 4),
             torch.zeros(3, 256))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "css_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_score_kernel_issues_all_row_loads_before_the_fmas():
    """Performance guard that needs no GPU: in the gradient variant of the scorer the eight 128-bit loads of a candidate
    row must be issued back to back.  Under a tighter register budget ptxas sinks the last load below the first FFMA2s and
    the row latency is paid twice per step (measured: 264 -> 332 us); this pins the schedule the timings were taken with."""
    import shutil
    import subprocess
    from css_b200 import build
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    funcs = re.findall(r"Function : (\S*score_ce_kernelILb1E\S*)\n(.*?)(?=Function : |\Z)", sass, flags=re.S)
    assert funcs, "gradient variants of score_ce_kernel not found"
    for name, body in funcs:
        ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
        row_load = "LDG.E.64" if "bfloat16" in name else "LDG.E.128"       # bf16 rows: 8 bytes per lane per chunk
        best = run = 0
        for op in ops:
            if op.startswith(row_load):
                run += 1
                best = max(best, run)
            elif op.startswith(("FFMA2", "LDS")):
                run = 0
        assert best >= 8, f"{name}: row loads are split by compute (longest run of {row_load} = {best})"
        assert any(op.startswith("FFMA2") for op in ops), "packed fp32x2 FMAs (sm_100 FFMA2) expected in the scorer"
