"""The C-ABI boundary: generated layout header is current, the ctypes mirror of WsiModelDesc
matches the C struct, and the built library exports every symbol include/wsi_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layout_header_is_current():
    from wsi_b200.backend import layout

    with open(os.path.join(ROOT, "include", "wsi_layout.h")) as f:
        assert f.read() == layout.c_header(), "run tools/gen_layout.py"


def test_ctypes_struct_matches_c_struct(tmp_path):
    from wsi_b200.backend.capi import WsiModelDesc

    names = [n for n, _ in WsiModelDesc._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/wsi_b200.h"', "int main(){",
           'printf("%zu\\n", sizeof(WsiModelDesc));']
    src += [f'printf("%zu\\n", offsetof(WsiModelDesc, {n}));' for n in names]
    src += ["return 0;}"]
    c = tmp_path / "off.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "off"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([gcc, str(c), "-o", str(exe)])
    vals = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert vals[0] == C.sizeof(WsiModelDesc)
    for n, off in zip(names, vals[1:]):
        assert getattr(WsiModelDesc, n).offset == off, n


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "wsi_b200.h")).read()
    return sorted(set(re.findall(r"\b(wsi_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from wsi_b200 import build as wb
    from wsi_b200.backend import capi

    wb.build_extension()
    lib = C.CDLL(capi.library_path())
    syms = _declared_symbols()
    assert {"wsi_version", "wsi_last_error", "wsi_validate", "wsi_step", "wsi_run_steps", "wsi_sched_range",
            "wsi_launch_count", "wsi_num_blocks", "wsi_desc_size"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), s
    assert capi.load_library().wsi_version() == capi.ABI_VERSION


def test_validate_rejects_bad_descriptors():
    from parity_util import load_golden
    from wsi_b200.backend import capi

    lib = capi.load_library()
    fm, _ = load_golden("quickstart")
    buf = capi.ModelBuffers(fm, device="numpy")
    assert lib.wsi_validate(C.byref(buf.desc)) == 0
    buf.desc.V = 5
    assert lib.wsi_validate(C.byref(buf.desc)) == 1
    assert b"P/V" in lib.wsi_last_error()
    buf.desc.V = buf.desc.P + 1
    buf.desc.abi_version = 99
    assert lib.wsi_validate(C.byref(buf.desc)) == 1


def test_product_does_not_import_oracle_or_hostsim():
    """No CPU fallback: nothing under wsi_b200/ may reference oracle/ or tests/hostsim."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wsi_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+(oracle|hostsim)\b", text, re.M) or "libwsi_oracle" in text \
                        or "libwsi_hostsim" in text:
                    bad.append(f)
    assert not bad, bad


def test_missing_extension_fails_loudly(monkeypatch):
    from wsi_b200.backend import capi

    monkeypatch.setattr(capi, "_LIB", None)
    monkeypatch.setattr(capi, "_LIB_PATH", "/nonexistent/libwsi_b200.so")
    with pytest.raises(capi.ExtensionMissingError):
        capi.load_library()
