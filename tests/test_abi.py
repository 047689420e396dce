"""The C-ABI library: it loads, exports every symbol include/nrcu.h declares, and fails loudly
without a GPU.  No compute calls here (CPU box)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import REPO


def declared_symbols():
    hdr = open(os.path.join(REPO, "include", "nrcu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nrcu_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    from nrenderer_b200 import api
    assert declared_symbols() == sorted(api.ABI_SYMBOLS)


def test_library_builds_and_exports_every_symbol():
    from nrenderer_b200 import build, api
    so = build.build_cuda()
    assert os.path.exists(so)
    lib = C.CDLL(so)
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym
    assert api.load_library().nrcu_abi_version() == 2


def test_struct_layouts_match_the_header():
    from nrenderer_b200 import api, flatscene
    assert C.sizeof(flatscene.NrcuMaterial) == 24 * 4
    assert C.sizeof(api.NrcuRenderParams) == 32
    assert C.sizeof(api.NrcuStats) == 72
    # compile a tiny C program printing sizeof/offsetof and compare
    import subprocess, tempfile
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "nrcu.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(nrcu_scene), offsetof(nrcu_scene, materials), offsetof(nrcu_scene, texture_rgba),
        sizeof(nrcu_material), sizeof(nrcu_render_params), sizeof(nrcu_stats)); return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.run(["gcc", f"-I{REPO}/include", os.path.join(td, "t.c"), "-o", os.path.join(td, "t")], check=True)
        out = subprocess.run([os.path.join(td, "t")], capture_output=True, text=True, check=True).stdout.split()
    S = flatscene.NrcuScene
    assert [int(x) for x in out] == [C.sizeof(S), S.materials.offset, S.texture_rgba.offset, 96, 32, 72]


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    from nrenderer_b200 import api
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    from oracle import pyoracle as po
    import pyemu
    for c, k, want in kat:
        assert list(api.philox4x32(c, k)) == want
        assert list(po.philox4x32(c, k)) == want
        assert list(pyemu.philox4x32(c, k)) == want


def test_no_gpu_means_loud_failure():
    """No CPU fallback: on a machine without CUDA the context cannot be created."""
    from nrenderer_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.NrcuError, match="no CUDA device"):
        api.Context(0)


def test_product_does_not_touch_the_oracle():
    """Nothing under nrenderer_b200/ may include, import, link or load oracle/ or tests/host_emu."""
    pkg = os.path.join(REPO, "nrenderer_b200")
    pats = [r'#\s*include\s*[<"][^>"]*(oracle|emu)', r'^\s*(from|import)\s+\S*(oracle|pyemu)', r'(dlopen|CDLL)\s*\([^)]*(oracle|emu)',
            r'libnroracle|libnrcu_emu']
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                for pat in pats:
                    assert not re.search(pat, txt, flags=re.M), (f, pat)
    out = os.popen(f"ldd {pkg}/libnrcuda.so").read()
    assert "oracle" not in out and "emu" not in out


def test_missing_library_is_a_loud_failure(tmp_path):
    """The binding never falls back: a library path that does not exist (NRCU_LIBRARY names an experiment build) raises."""
    code = ("import os; os.environ['NRCU_LIBRARY'] = %r\n"
            "from nrenderer_b200 import api\n"
            "try:\n    api.load_library()\nexcept api.NrcuError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n") % str(tmp_path / "libnrcuda.nope.so")
    import subprocess, sys
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REPO).stdout
    assert "RAISED True" in out
