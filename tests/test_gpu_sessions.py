"""Level-2 drop-ins: ScrfdSession.detect / ArcSession.run against the oracle's statement of the two
third-party calls the reference makes (InsightFace SCRFD.detect, face_embedder.py:2176-2187; arc_sess.run, :1369)."""
import cv2
import numpy as np
import pytest

import pcb_test_helpers as H
from person_capture_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,Hh,S,thr", [(640, 360, 640, 0.5), (416, 234, 416, 0.3), (300, 400, 512, 0.5)])
def test_scrfd_session_detect_matches_scrfd_oracle(engine_10g_r50, W, Hh, S, thr):
    from oracle.scrfd_detect import SCRFDOracle
    from person_capture_b200.sessions import ScrfdSession
    sess = ScrfdSession(engine_10g_r50)
    ora = SCRFDOracle(H.oracle_scrfd("scrfd_10g_bnkps"))
    clip = synth.ClipSpec(W, Hh, 40, seed=31, distractor_prob=1.0, target_segments=[(0, 39)])
    seen = 0
    for i in (2, 11, 23):
        img = clip.frame(i)
        sess.det_thresh = ora.det_thresh = thr
        det, kps = sess.detect(img, input_size=(S, S))
        odet, okps = ora.detect(img, (S, S))
        assert det.dtype == np.float32 and kps.dtype == np.float32 and det.shape == (len(odet), 5) and kps.shape == (len(odet), 5, 2)
        assert len(det) >= 1
        # fp16 convolutions vs the fp32 oracle: same detections in the same (score) order, sub-pixel agreement
        assert np.abs(det[:, :4] - odet[:, :4]).max() <= 0.75, np.abs(det[:, :4] - odet[:, :4]).max()
        assert np.abs(det[:, 4] - odet[:, 4]).max() <= 2e-2
        assert np.abs(kps - okps).max() <= 0.75
        seen += len(det)
    assert seen >= 4
    with pytest.raises(ValueError):
        sess.detect(clip.frame(0), input_size=(640, 480))


def test_arc_session_run_matches_oracle(engine_10g_r50):
    from oracle.face_embedder import arcface_preprocess
    from person_capture_b200.sessions import ArcSession
    sess = ArcSession(engine_10g_r50)
    rng = np.random.default_rng(5)
    chips = []
    for i in range(6):
        canvas = synth.background(rng, 150, 150, clutter=2)
        synth.paste_face(canvas, 60 + i, 75, 75, 112, float(rng.uniform(-5, 5)))
        chips.append(np.ascontiguousarray(canvas[19:131, 19:131]))
    chips = np.stack(chips)
    X = np.stack([arcface_preprocess(c) for c in chips])
    name = sess.get_inputs()[0].name
    (out,) = sess.run(None, {name: X})
    ref = H.oracle_arcface("arcface_r50").run(X)
    assert out.shape == (6, 512) and out.dtype == np.float32
    assert min(H.cos(a, b) for a, b in zip(out, ref)) >= 0.999
    e, ef = sess.run_chips(chips, flip=True)
    assert np.array_equal(e, out)                       # the blob path recovers the uint8 chip exactly
    ref_f = H.oracle_arcface("arcface_r50").run(np.stack([arcface_preprocess(cv2.flip(c, 1)) for c in chips]))
    assert min(H.cos(a, b) for a, b in zip(ef, ref_f)) >= 0.999
    with pytest.raises(ValueError):
        sess.run(None, {name: X + 0.003})               # not a blob of uint8 chips
