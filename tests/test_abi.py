"""The C-ABI library loads and exports every symbol include/ekfvio_c.h declares, the Python
binding covers them all, and without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ekfvio_c.h")
LIB = os.path.join(ROOT, "ekf_vio_b200", "libekfvio_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ekfvio_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    names = declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    from ekf_vio_b200 import capi
    assert sorted(capi.SIGNATURES) == declared_symbols()


def test_every_entry_point_cites_the_reference():
    src = open(HEADER).read()
    for ref in ("TightlyCoupledEKF.cpp:23-56", "TightlyCoupledEKF.cpp:58-94", "TightlyCoupledEKF.cpp:96-121", "TightlyCoupledEKF.cpp:475-628",
                "TightlyCoupledEKF.cpp:176-325", "TightlyCoupledEKF.cpp:699-714", "KLTTracker.cpp:61-64", "KLTTracker.cpp:72-92", "Feature.h:60-62"):
        assert ref in src, ref


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ekf_vio_b200 import capi
    with pytest.raises(capi.EkfvioError, match="no CUDA device"):
        capi.EkfBatch(1, 4)
    with pytest.raises(capi.EkfvioError, match="no CUDA device"):
        capi.KltTracker(64, 64, 1, 8)


def test_product_sources_do_not_reach_into_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ekf_vio_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle/" not in text and "oracle_lib" not in text and "libekf_oracle" not in text, os.path.join(dirpath, fn)
    for hdr in os.listdir(os.path.join(ROOT, "include", "ekf_vio")) + ["../ekfvio_c.h"]:
        text = open(os.path.join(ROOT, "include", "ekf_vio", hdr)).read()
        assert "oracle" not in text.lower()


def test_facade_headers_compile_standalone(tmp_path):
    """include/ekf_vio/*.h with the reference's class names compile without Eigen/OpenCV/ROS."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "ekf_vio/TightlyCoupledEKF.h"\n#include "ekf_vio/KLTTracker.h"\n'
                   "int f(TightlyCoupledEKF& e, KLTTracker& k, Frame& a, Frame& b) {\n"
                   "  std::vector<Eigen::Vector2f> m; std::vector<Eigen::Matrix2f> c; std::vector<bool> p;\n"
                   "  e.process(0.05f); k.findNewFeaturePositions(a, b, e.previousFeaturePositionVector(), e.features, m, c, p);\n"
                   "  e.updateWithFeaturePositions(m, c, p); return (int)e.features.size(); }\n")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), str(src)])
