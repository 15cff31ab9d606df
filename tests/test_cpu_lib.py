"""CPU: the C-ABI library loads and exports every symbol include/pcb200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "pcb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pcb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported():
    from person_capture_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert set(_lib.EXPORTS) == set(names)


def test_layout_pad_is_reported_by_the_library():
    """pcb_layout_pad needs no GPU: the host side takes the activation-layout padding from the library it loaded."""
    import ctypes as C
    from person_capture_b200 import _lib
    lib = _lib.load()
    lo, pad = C.c_int(-1), C.c_int(-1)
    lib.pcb_layout_pad(C.byref(lo), C.byref(pad))
    assert (lo.value, pad.value) == (_lib.P_PAD_LO, _lib.P_PAD)
    assert (lo.value, pad.value) in ((1, 2), (0, 1))          # ring (product) or trailing pad (build-time experiment)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from person_capture_b200 import _lib
    from person_capture_b200.face_embedder import FaceEmbedder
    with pytest.raises(_lib.PcbError):
        FaceEmbedder("cuda", "scrfd_2.5g_bnkps")
    with pytest.raises(RuntimeError):
        FaceEmbedder("cpu", "scrfd_2.5g_bnkps")
    with pytest.raises(RuntimeError):
        FaceEmbedder("cuda", "yolov8n-face.pt")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "person_capture_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_graph_macs_match_survey_tables():
    from person_capture_b200 import graphs
    g10, g25 = graphs.build_graph("scrfd_10g_bnkps"), graphs.build_graph("scrfd_2.5g_bnkps")
    assert abs(graphs.graph_macs(g10, 256, 256) / 1e9 - 8.54) < 0.01          # S=512
    assert abs(graphs.graph_macs(g10, 640, 640) / 1e9 - 53.36) < 0.01         # S=1280
    assert abs(graphs.graph_macs(g25, 208, 208) / 1e9 - 1.45) < 0.01          # S=416
    assert abs(graphs.graph_macs(graphs.build_graph("arcface_r100"), 112, 112) / 1e9 - 12.090) < 0.001
    assert abs(graphs.graph_macs(graphs.build_graph("arcface_r50"), 112, 112) / 1e9 - 6.309) < 0.001


def test_weights_are_deterministic_and_loadable():
    import numpy as np
    from person_capture_b200 import graphs, weights
    a = weights.arcface_random_weights("arcface_r50")
    b = weights.arcface_random_weights("arcface_r50")
    assert all(np.array_equal(a[k], b[k]) for k in a)
    for name in ("scrfd_2.5g_bnkps", "scrfd_10g_bnkps", "arcface_r50"):
        g = graphs.build_graph(name)
        ops, blob, outs, reg = graphs.pack(g, weights.load_params(name))
        assert len(blob) > 0 and len(ops) == len(g.ops)
