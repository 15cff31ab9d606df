"""ORACLE TOOLING: train the SCRFD graphs on synthetic faces and export folded fp16 weights.

The reference downloads trained `scrfd_*_bnkps.onnx` files (person_capture/face_embedder.py:55-65);
none are available offline, and random-init detectors would exercise only the fallback
branches of the path.  This script trains the architecture of oracle/models.py
on the planted faces of person_capture_b200/synth.py for a few thousand CPU steps and writes
weights/<name>.npz, which both the oracle and the CUDA path load.

usage: python -m oracle.train_scrfd scrfd_2.5g_bnkps --steps 3000 --threads 4
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.models import SCRFD  # noqa: E402
from person_capture_b200 import synth  # noqa: E402

LEVEL_RANGES = ((0, 56), (56, 200), (200, 10_000))  # face side (px) -> stride 8 / 16 / 32


def make_sample(rng: np.random.Generator, size: int, tex_cache: dict):
    # backgrounds at 1x..4x zoom-out: a 1080p frame letterboxed to 512 shows the background's structure ~4x
    # smaller than a native-resolution frame does; without these the stride-32 head fires on large smooth blobs
    zoom = int(rng.choice([1, 1, 2, 3, 4]))
    img = synth.background(rng, size * zoom, size * zoom, clutter=int(rng.integers(2, 9)))
    if zoom > 1:
        import cv2
        img = cv2.resize(img, (size, size), interpolation=cv2.INTER_AREA)
    # top-left anchored letterbox (InsightFace SCRFD.detect): a 16:9 frame fills only the top 56% of the square and
    # the rest is zero pad; half of the samples get such a pad so the heads learn that the content/pad edge is not a face
    hc, wc = size, size
    if rng.random() < 0.5:
        if rng.random() < 0.8:
            hc = int(size * rng.uniform(0.4, 0.95))
        else:
            wc = int(size * rng.uniform(0.4, 0.95))
    n = int(rng.choice([0, 1, 1, 1, 2, 2, 3, 5]))
    boxes, kpss = [], []
    placed = []
    for _ in range(n):
        side = float(np.exp(rng.uniform(np.log(12.0), np.log(min(hc, wc) * 0.95))))
        for _try in range(6):
            cx = float(rng.uniform(side * 0.35, wc - side * 0.35))
            cy = float(rng.uniform(side * 0.35, hc - side * 0.35))
            if all(abs(cx - x) > (side + s) * 0.55 or abs(cy - y) > (side + s) * 0.55 for x, y, s in placed):
                break
        else:
            continue
        placed.append((cx, cy, side))
        ident = int(rng.integers(0, 4000))
        bb, kp = synth.paste_face(img, ident, cx, cy, side, float(rng.uniform(-12, 12)),
                                  float(rng.uniform(0.75, 1.2)), None)
        boxes.append(bb)
        kpss.append(kp)
    if rng.random() < 0.3:
        img = cv2_blur(img, float(rng.uniform(0.4, 1.4)))
    if rng.random() < 0.5:
        noise = rng.normal(0, rng.uniform(1, 6), img.shape)
        img = np.clip(img.astype(np.float32) + noise, 0, 255).astype(np.uint8)
    img[hc:] = 0
    img[:, wc:] = 0
    return img, np.array(boxes, np.float32).reshape(-1, 4), np.array(kpss, np.float32).reshape(-1, 5, 2)


def cv2_blur(img, sigma):
    import cv2
    return cv2.GaussianBlur(img, (0, 0), sigma)


def build_targets(boxes, kpss, size):
    """Per level dense targets: cls [h,w], reg [h,w,4], kps [h,w,10], pos mask, ignore mask."""
    out = []
    for li, s in enumerate((8, 16, 32)):
        h = w = size // s
        cls = np.zeros((h, w), np.float32)
        ign = np.zeros((h, w), bool)
        reg = np.zeros((h, w, 4), np.float32)
        kps = np.zeros((h, w, 10), np.float32)
        ys, xs = np.mgrid[0:h, 0:w]
        cx = (xs * s).astype(np.float32)
        cy = (ys * s).astype(np.float32)
        # big faces first so small ones overwrite
        order = np.argsort([-(b[2] - b[0]) * (b[3] - b[1]) for b in boxes]) if len(boxes) else []
        for gi in order:
            b = boxes[gi]
            side = max(b[2] - b[0], b[3] - b[1])
            lo, hi = LEVEL_RANGES[li]
            bx, by = (b[0] + b[2]) / 2, (b[1] + b[3]) / 2
            inside = (cx > b[0]) & (cx < b[2]) & (cy > b[1]) & (cy < b[3])
            if lo <= side < hi:
                r = max(s * 0.75, 0.22 * side)
                pos = (np.abs(cx - bx) <= r) & (np.abs(cy - by) <= r) & inside
                if not pos.any():
                    iy = int(np.clip(round(by / s), 0, h - 1))
                    ix = int(np.clip(round(bx / s), 0, w - 1))
                    pos = np.zeros((h, w), bool)
                    pos[iy, ix] = True
                ign |= inside & ~pos
                cls[pos] = 1.0
                ign[pos] = False
                reg[pos] = np.stack([cx - b[0], cy - b[1], b[2] - cx, b[3] - cy], -1)[pos] / s
                k = kpss[gi]
                kt = np.stack([(k[j // 2, j % 2] - (cx if j % 2 == 0 else cy)) / s for j in range(10)], -1)
                kps[pos] = kt[pos]
            elif lo * 0.7 <= side < hi * 1.4:
                ign |= inside & (cls == 0)
        out.append((cls, ign, reg, kps))
    return out


def loss_fn(raws, targets_batch):
    """raws: list of [N,30,h,w]; targets_batch: list over images of per-level tuples."""
    total_cls = 0.0
    total_reg = 0.0
    total_kps = 0.0
    npos = 0.0
    for li, o in enumerate(raws):
        cls_t = torch.from_numpy(np.stack([t[li][0] for t in targets_batch]))
        ign = torch.from_numpy(np.stack([t[li][1] for t in targets_batch]))
        reg_t = torch.from_numpy(np.stack([t[li][2] for t in targets_batch]))
        kps_t = torch.from_numpy(np.stack([t[li][3] for t in targets_batch]))
        o = o.permute(0, 2, 3, 1)  # N,h,w,30
        pos = cls_t > 0.5
        for a in range(2):
            logit = o[..., a]
            p = torch.sigmoid(logit)
            ce = F.binary_cross_entropy_with_logits(logit, cls_t, reduction="none")
            pt = p * cls_t + (1 - p) * (1 - cls_t)
            alpha = 0.25 * cls_t + 0.75 * (1 - cls_t)
            fl = alpha * (1 - pt) ** 2 * ce
            total_cls = total_cls + (fl * (~ign)).sum()
            r = o[..., 2 + 4 * a: 6 + 4 * a]
            k = o[..., 10 + 10 * a: 20 + 10 * a]
            if pos.any():
                total_reg = total_reg + F.smooth_l1_loss(r[pos], reg_t[pos], beta=0.25, reduction="sum")
                total_kps = total_kps + F.smooth_l1_loss(k[pos], kps_t[pos], beta=0.25, reduction="sum")
        npos += float(pos.sum()) * 2
    npos = max(npos, 1.0)
    return total_cls / npos, total_reg / npos / 4, total_kps / npos / 10


def recalibrate_bn(net, rng, size: int, batches: int = 80, batch: int = 12):
    """Re-estimate BatchNorm running statistics as a cumulative average over fresh batches.

    With ~1.5k steps of a one-cycle schedule the exponential running stats lag the final weights
    (eval-mode recall collapses while batch-stat recall is ~95%); the exported folded scale/bias
    come from these statistics, so they are recomputed before export.
    """
    for mod in net.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.reset_running_stats()
            mod.momentum = None
    net.train()
    with torch.no_grad():
        for _ in range(batches):
            imgs = [make_sample(rng, size, None)[0] for _ in range(batch)]
            x = np.stack(imgs)[..., ::-1].astype(np.float32)
            x = (x - 127.5) / 128.0
            net.head_raw(torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2))))
    net.eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--seed", type=int, default=20240)
    ap.add_argument("--lr", type=float, default=2e-3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--recalibrate-only", action="store_true")
    ap.add_argument("--freeze-bn", action="store_true",
                    help="fine-tune with BatchNorm in eval mode (fixed population statistics) so that the "
                         "exported folded graph and the training-time graph are the same function")
    args = ap.parse_args()
    torch.set_num_threads(args.threads)
    torch.manual_seed(args.seed)
    rng = np.random.default_rng(args.seed)
    net = SCRFD(args.name)
    if args.resume:
        net.load_state_dict(torch.load(args.resume))
    out = args.out or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights", args.name + ".npz")
    if args.recalibrate_only:
        recalibrate_bn(net, rng, args.size)
        np.savez(out, **net.export_folded())
        print("recalibrated", out)
        return
    def set_train_mode():
        net.train()
        if args.freeze_bn:
            for mod in net.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.eval()

    if args.freeze_bn:
        recalibrate_bn(net, rng, args.size, batches=40)
    set_train_mode()
    opt = torch.optim.AdamW(net.parameters(), lr=args.lr, weight_decay=1e-4)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=args.lr, total_steps=args.steps, pct_start=0.1)
    out = args.out or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights", args.name + ".npz")
    t0 = time.time()
    for step in range(args.steps):
        imgs, tg = [], []
        for _ in range(args.batch):
            img, boxes, kpss = make_sample(rng, args.size, None)
            imgs.append(img)
            tg.append(build_targets(boxes, kpss, args.size))
        x = np.stack(imgs)[..., ::-1].astype(np.float32)  # BGR->RGB
        x = (x - 127.5) / 128.0
        x = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))
        raws = net.head_raw(x)
        lc, lr_, lk = loss_fn(raws, tg)
        loss = lc + 0.5 * lr_ + 0.5 * lk
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
        opt.step()
        sched.step()
        if step % 25 == 0 or step == args.steps - 1:
            print(f"{args.name} step {step} loss {float(loss):.4f} cls {float(lc):.4f} reg {float(lr_):.4f} kps {float(lk):.4f} "
                  f"t={time.time() - t0:.0f}s", flush=True)
        if (step % 500 == 499) or step == args.steps - 1:
            net.eval()
            P = net.export_folded()
            np.savez(out, **P)
            torch.save(net.state_dict(), out.replace(".npz", ".pt"))
            set_train_mode()
    if not args.freeze_bn:
        recalibrate_bn(net, rng, args.size)
    net.eval()
    np.savez(out, **net.export_folded())
    print("saved", out)


if __name__ == "__main__":
    main()
