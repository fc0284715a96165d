"""Host cost of the pieces of upstream's training sample that hoisdf_b200/feed.py moves to the GPU, timed on THIS
machine's CPU through the same library calls upstream makes (Pillow, torchvision, numpy) on a 640 x 480 frame with DexYCB-sized
SDF files (tool/pre_process_sdf.py samples ~ 10^5 rows per frame): one DataLoader worker's time per sample.  Reference-side
evidence only: the GPU side of the feed has not been timed (no GPU budget was left when it was written).
Usage: python scripts/feed_reference_cost.py [repeats]"""
import os
import random
import sys
import time

import numpy as np
from PIL import Image, ImageFilter

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import feed_oracle as FO  # noqa: E402  (test infrastructure: this script is a measurement of the reference side)


def best(fn, n):
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return 1e3 * min(t), 1e3 * float(np.median(t))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    import torchvision.transforms.functional as TF
    import torchvision.transforms as T
    img, hs, os_, center, scale, rot = FO.synthetic_aug(0)
    affine, _ = FO.get_affine_transform(np.asarray(center), scale, [256, 256], rot=rot)
    pil, pil_h, pil_o = Image.fromarray(img), Image.fromarray(hs), Image.fromarray(os_)
    warped = FO.transform_img(pil, affine, [256, 256]).crop((0, 0, 256, 256))
    rows = np.random.default_rng(0).standard_normal((120000, 6)).astype(np.float32)
    nh = 70000

    def draws_upstream():
        np.random.choice(list(range(nh)), size=600, replace=False)
        np.random.choice(list(range(nh, len(rows))), size=200, replace=False)
        np.random.choice(np.where(np.abs(rows[:nh, 3]) < 0.5)[0], size=600, replace=False)
        np.random.choice(np.where(np.abs(rows[nh:, 4]) < 0.5)[0] + nh, size=200, replace=False)

    def draws_feed():
        from hoisdf_b200 import feed
        feed.draw_sdf_indices(rows, nh, 600, 200, 0.5)

    def jitter():
        p = warped
        for f, a in ((TF.adjust_brightness, 1.2), (TF.adjust_saturation, 0.8), (TF.adjust_hue, 0.1), (TF.adjust_contrast, 1.3)):
            p = f(p, a)

    idx = np.random.default_rng(1).integers(0, len(rows), 1600)
    rot_mat = np.eye(3, dtype=np.float32)
    root = np.zeros(3, np.float32)
    table = [
        ("frame warp (transform_img + crop)", lambda: FO.transform_img(pil, affine, [256, 256]).crop((0, 0, 256, 256))),
        ("2 mask warps + NEAREST shrinks", lambda: [FO.transform_img(m, affine, [256, 256]).crop((0, 0, 256, 256)).resize((64, 64), Image.NEAREST) for m in (pil_h, pil_o)]),
        ("GaussianBlur(0.3)", lambda: warped.filter(ImageFilter.GaussianBlur(0.3))),
        ("color_jitter (4 adjustments)", jitter),
        ("ToTensor(float32) / 255", lambda: T.ToTensor()(np.asarray(warped).astype(np.float32)) / 255.0),
        ("4 np.random.choice draws, upstream form (list(range(n)))", draws_upstream),
        ("the same draws, feed.draw_sdf_indices", draws_feed),
        ("SDF point sets (gather, rotate, normalise)", lambda: FO.sdf_point_sets(rows, idx, 600, 200, root, root, 6.2, 5.8, rot_mat=rot_mat)),
    ]
    print("| piece of one training sample (host, one core) | best ms | median ms |\n|---|---|---|")
    for name, fn in table:
        b, m = best(fn, n)
        print("| %s | %.2f | %.2f |" % (name, b, m))
    print("\ncores: %d; Pillow %s" % (os.cpu_count(), Image.__version__))


if __name__ == "__main__":
    random.seed(0)
    np.random.seed(0)
    main()
