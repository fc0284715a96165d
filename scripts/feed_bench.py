"""Time the data feed on the GPU beside the host cost of the same work through the libraries upstream calls.
One JSON line: frames/s of `feed.eval_batch` (BASELINE configs[1]'s batch of 32 evaluation frames) and of `feed.train_batch`
(configs[3]'s batch of 64 training frames with blur + jitter), device-timed with CUDA events after warm-up, inputs resident in
HBM; `cpu_reference` = the same samples through Pillow / torchvision / numpy on one host core (oracle/feed_oracle.py: test
infrastructure, used here as the measured baseline only).  Not run in the round it was written in (no GPU budget was left):
    /usr/local/graft/bin/gpurun -- python scripts/feed_bench.py
`run(device, ...)` takes any device so that tests/test_feed_wrappers.py can drive the same code on the CPU emulator."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _timed(fn, device, steps, warmup):
    for _ in range(warmup):
        fn()
    if device.type == "cuda":
        torch.cuda.synchronize()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        torch.cuda.synchronize()
        return start.elapsed_time(stop) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return 1e3 * (time.perf_counter() - t0) / steps


def run(device, eval_frames=32, train_frames=64, steps=20, warmup=3, n_hand=600, n_obj=200, cpu_samples=4):
    from PIL import Image, ImageFilter
    import torchvision.transforms.functional as TF
    from hoisdf_b200 import feed, ops
    from oracle import feed_oracle as FO
    from test_feed import product_sample

    raw = [FO.synthetic_eval_annotation(s) for s in range(eval_frames)]
    eval_samples = [feed.eval_geometry(ann, corners, (640, 480), 0.7) for _, ann, corners in raw]
    eval_frames_d = torch.from_numpy(np.stack([r[0] for r in raw])).to(device)
    host = [product_sample(s, n_hand, n_obj) for s in range(train_frames)]
    aug = [FO.synthetic_aug(s) for s in range(train_frames)]
    frames = torch.from_numpy(np.stack([a[0] for a in aug])).to(device)
    hand_masks = torch.from_numpy(np.stack([a[1] for a in aug])).to(device)
    obj_masks = torch.from_numpy(np.stack([a[2] for a in aug])).to(device)
    rows = torch.from_numpy(np.concatenate([h[1] for h in host])).to(device)
    offsets = torch.from_numpy(np.cumsum([0] + [len(h[1]) for h in host]).astype(np.int64)).to(device)
    samples = [h[0] for h in host]
    before = ops.STATS["launches"]
    feed.train_batch(frames, hand_masks, obj_masks, rows, offsets, samples, n_hand, n_obj, 6.2, 5.8)
    launches = ops.STATS["launches"] - before
    eval_ms = _timed(lambda: feed.eval_batch(eval_frames_d, eval_samples), device, steps, warmup)
    train_ms = _timed(lambda: feed.train_batch(frames, hand_masks, obj_masks, rows, offsets, samples, n_hand, n_obj, 6.2, 5.8),
                      device, steps, warmup)

    def cpu_train_sample(i):           # the library calls of ho3d.py:351-381,484-486,524-552 for one sample
        s, sdf = host[i]
        img, hm, om = aug[i][:3]
        coef = tuple(float(c) for c in s["coef"])
        pil = Image.fromarray(img).transform((256, 256), Image.AFFINE, coef).filter(ImageFilter.GaussianBlur(s["blur_radius"]))
        for name, f in s["jitter"]:
            pil = getattr(TF, "adjust_" + name)(pil, f)
        np.ascontiguousarray(np.asarray(pil).astype(np.float32).transpose(2, 0, 1)) / np.float32(255.0)
        for m in (hm, om):
            np.asarray(Image.fromarray(m).transform((256, 256), Image.AFFINE, coef).resize((128, 128), Image.NEAREST)).astype(np.float32)
        FO.sdf_point_sets(sdf, s["index"], n_hand, n_obj, s["mano_root"], s["obj_center_cam"], 6.2, 5.8, rot_mat=s["rot_mat"])

    t0 = time.perf_counter()
    for i in range(min(cpu_samples, train_frames)):
        cpu_train_sample(i)
    cpu_ms = 1e3 * (time.perf_counter() - t0) / min(cpu_samples, train_frames)
    return {"metric": "frames/s through hoisdf_b200.feed (inputs resident in HBM; host geometry and draws excluded)",
            "device": str(device), "eval_batch": {"frames": eval_frames, "ms": eval_ms, "frames_per_s": 1e3 * eval_frames / eval_ms},
            "train_batch": {"frames": train_frames, "ms": train_ms, "frames_per_s": 1e3 * train_frames / train_ms,
                            "launches": launches, "points": [n_hand, n_obj]},
            "cpu_reference": {"kind": "port", "cores": 1, "ms_per_train_sample": cpu_ms, "frames_per_s": 1e3 / cpu_ms,
                              "sample": "%d training samples through Pillow / torchvision / numpy (pixels, masks, point sets; "
                                        "draws and geometry excluded on both sides)" % min(cpu_samples, train_frames)},
            "steps": steps, "warmup": warmup, "data": "synthetic"}


if __name__ == "__main__":
    assert torch.cuda.is_available(), "feed_bench times the GPU path; there is no CPU fallback"
    print(json.dumps(run(torch.device("cuda:0"))))
