"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the
headers declare; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    declared = _declared("points_transfer.h")
    assert declared == set(pkg.ABI_SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), sym
    # the synthetic-workload generators are bench scaffolding: their own library, and the
    # drop-in does not export them
    S = pkg.synth_lib()
    assert _declared("pt_synth.h") == set(pkg.SYNTH_SYMBOLS)
    for sym in pkg.SYNTH_SYMBOLS:
        assert hasattr(S, sym), sym
        assert not hasattr(L, sym), sym


def test_headers_are_plain_c(tmp_path):
    # the boundary must be usable from C: no C++/torch types in the signatures
    src = tmp_path / "t.c"
    src.write_text('#include "points_transfer.h"\n#include "pt_synth.h"\n'
                   "int main(void){ pt_build_opts o = {0}; (void)o; return sizeof(pt_cand) == 32 "
                   "&& sizeof(pt_attr) == 16 ? 0 : 1; }\n")
    exe = tmp_path / "t"
    import subprocess
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_struct_mirrors(pkg):
    assert pkg.POINT_DTYPE.itemsize == 80 and pkg.ATTR_DTYPE.itemsize == 16
    assert pkg.CAND_DTYPE.itemsize == 32
    assert ctypes.sizeof(pkg.api.IndexInfo) == 8 + 8 + 16 + 48 + 8 + 16


def test_status_strings_and_options(pkg):
    assert pkg.status_string(0) == "ok"
    assert "no CPU fallback" in pkg.status_string(3)
    pkg.set_option("knn_variant", 0)
    assert pkg.get_option("knn_variant") == 0
    pkg.set_option("knn_variant", -1)                           # back to auto (the default)
    with pytest.raises(pkg.PointsTransferError):
        pkg.set_option("no_such_option", 1)


def test_distance_radius_transform(pkg):
    assert pkg.Distance.transformed_distance(3.0) == 9.0       # src/Distance.h:97
    assert pkg.Distance.inverse_of_transformed_distance(9.0) == 3.0


def test_no_cpu_fallback(pkg):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    pts = pkg.synth.cloud_host(100, 1)
    with pytest.raises(pkg.PointsTransferError) as e:
        pkg.Tree(pts)
    assert e.value.status == 3
    with pytest.raises(TypeError):
        pkg.Tree(np.zeros((10, 3)))   # not Point records


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "3d-reconstruction-from-point-cloud_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ('#include "pt_oracle', "libpt_oracle", "from oracle", "import oracle",
                               "oracle/pto", "pto_"):
                    assert needle not in text, (needle, os.path.join(dirpath, f))
