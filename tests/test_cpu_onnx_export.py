"""CPU: the oracle's networks written as ONNX files (oracle/onnx_export.py, protobuf by hand: no `onnx` package offline) and
executed by an independent engine, OpenCV's cv2.dnn, against the torch-CPU executors parity is measured with.  This is the
offline stand-in for north_star's "ONNX Runtime CPU execution of the same ONNX graphs": same graphs, same weights, another
inference engine."""
import os

import cv2
import numpy as np
import pytest
import torch

import pcb_test_helpers as H
from oracle import onnx_export as X
from oracle.scrfd_detect import SCRFDOracle
from person_capture_b200 import synth, weights


class _DnnScrfd:
    """`net.run(blob)` of SCRFDOracle on top of cv2.dnn running the exported ONNX file (nine InsightFace-layout outputs)."""

    def __init__(self, path, names):
        self.net, self.names = cv2.dnn.readNetFromONNX(path), names

    def run(self, blob):
        self.net.setInput(np.ascontiguousarray(blob, np.float32))
        return [np.asarray(o) for o in self.net.forward(self.names)]


@pytest.mark.parametrize("name,S", [("scrfd_2.5g_bnkps", 320), ("scrfd_10g_bnkps", 512)])
def test_scrfd_onnx_through_cv2_dnn_equals_the_torch_oracle(tmp_path, name, S):
    path = str(tmp_path / f"{name}.onnx")
    names = X.export_scrfd(name, weights.load_params(name), S, path, layout="insightface")
    assert len(names) == 9 and os.path.getsize(path) > 1_000_000
    img = synth.ClipSpec(S, S, 10, seed=3, target_segments=[(0, 9)]).frame(2)
    blob = cv2.dnn.blobFromImage(img, 1.0 / 128, (S, S), (127.5, 127.5, 127.5), swapRB=True)
    dnn = _DnnScrfd(path, names)
    got, ref = dnn.run(blob), H.oracle_scrfd(name).run(blob)
    for n, g, r in zip(names, got, ref):
        assert g.shape == r.shape, n
        np.testing.assert_allclose(g, r, rtol=0, atol=2e-4, err_msg=n)
    # and the whole detector on top of it: same faces from both executors
    a, b = SCRFDOracle(dnn), SCRFDOracle(H.oracle_scrfd(name))
    a.det_thresh = b.det_thresh = 0.5
    da, ka = a.detect(img, input_size=(S, S))
    db, kb = b.detect(img, input_size=(S, S))
    assert da.shape == db.shape and len(da) >= 1
    np.testing.assert_allclose(da, db, rtol=0, atol=2e-3)
    np.testing.assert_allclose(ka, kb, rtol=0, atol=2e-3)


@pytest.mark.parametrize("name", ["arcface_r50", "arcface_r100"])
def test_iresnet_onnx_through_cv2_dnn_equals_the_torch_oracle(tmp_path, name):
    path = str(tmp_path / f"{name}.onnx")
    out = X.export_iresnet(name, weights.load_params(name), path)
    net = cv2.dnn.readNetFromONNX(path)
    rng = np.random.default_rng(11)
    chips = rng.integers(0, 256, (2, 112, 112, 3), dtype=np.uint8)
    from oracle.face_embedder import arcface_preprocess
    for c in chips:
        x = arcface_preprocess(cv2.GaussianBlur(c, (0, 0), 1.5))[None]
        net.setInput(x)
        y = net.forward(out)
        ref = H.oracle_arcface(name).run(x)
        assert y.shape == ref.shape == (1, 512)
        assert H.cos(y, ref) >= 0.999999
        np.testing.assert_allclose(y, ref, rtol=0, atol=1e-3 * float(np.abs(ref).max()))
