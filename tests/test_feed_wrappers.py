"""hoisdf_b200/feed.py's tensor-level wrappers (argument marshalling, strides, temporaries, the order of the calls in
`train_batch`) driven on the CPU: the C entry points are taken from the EMULATED library (the same .cu files compiled for the
host, tests/emu) and the three GPU touch points of the module (`_require_gpu`, `_on`, `_stream`) are patched for the duration of
a test, so host tensors flow through exactly the Python code the GPU path runs.  The bodies are the ones of
tests/test_gpu_zzz_feed.py: each GPU test of the feed is executed here against the emulator with `cuda` = the CPU device.
Test infrastructure: the product never loads the emulated library."""
import contextlib

import pytest
import torch

import test_gpu_zzz_feed as G
from hoisdf_b200 import _capi, feed
from test_kernel_emulation import build_emulated

FEED_ENTRY_POINTS = ("hoisdf_image_crop_fwd", "hoisdf_sdf_rows_fwd", "hoisdf_gaussian_blur_params", "hoisdf_gaussian_blur_u8",
                     "hoisdf_color_jitter_u8", "hoisdf_train_image_smem_bytes", "hoisdf_train_image_fwd", "hoisdf_mask_crop_fwd")


@pytest.fixture()
def host(monkeypatch):
    emu = build_emulated("feed")
    for name in FEED_ENTRY_POINTS:
        restype, argtypes = _capi.SIGNATURES[name]
        getattr(emu, name).restype, getattr(emu, name).argtypes = restype, argtypes
    monkeypatch.setattr(feed, "lib", emu)
    monkeypatch.setattr(feed, "_require_gpu", lambda t: None)
    monkeypatch.setattr(feed, "_on", lambda dev: contextlib.nullcontext())
    monkeypatch.setattr(feed, "_stream", lambda: None)
    return torch.device("cpu")


def test_evaluation_crop_batch(host):
    G.test_evaluation_crop_batch_vs_oracle_and_fixture(host)


def test_rotated_warp_and_masks(host):
    G.test_rotated_warp_and_masks_vs_oracle_and_fixture(host)


@pytest.mark.parametrize("h,w,size", [(37, 53, 19), (5, 7, 33)])
def test_ragged_sizes(host, h, w, size):
    G.test_ragged_sizes_and_out_of_frame_windows(host, h, w, size)


def test_sdf_point_sets(host):
    G.test_sdf_point_sets_vs_oracle(host, True, True, False)
    G.test_sdf_point_sets_vs_oracle(host, False, False, True)


def test_training_item(host):
    G.test_training_item_reproduces_the_upstream_fixture(host)


def test_mirrored_warp(host):
    G.test_mirrored_warp_equals_warping_the_mirrored_frame(host)


def test_gaussian_blur(host):
    G.test_gaussian_blur_vs_pillow(host)


def test_color_jitter(host):
    G.test_color_jitter_vs_torchvision(host)


def test_training_image(host):
    G.test_training_image_reproduces_the_upstream_fixture(host)


def test_train_batch(host):
    G.test_train_batch_reproduces_the_upstream_item(host)


def test_eval_batch(host):
    G.test_eval_batch_reproduces_the_upstream_item(host)


def test_dexycb_eval_batch(host):
    G.test_dexycb_eval_batch_reproduces_the_upstream_item(host)


def test_train_batch_dexycb(host):
    G.test_train_batch_reproduces_the_upstream_dexycb_item(host)


def test_fused_training_image(host):
    G.test_fused_training_image_equals_the_step_by_step_calls(host)


def test_fused_masks(host):
    G.test_fused_mask_crop_equals_the_two_step_route(host)


def test_smoke_feed(host):
    """`__graft_entry__.smoke_feed` (the feed's line of the driver's smoke run) on the emulator."""
    import time
    import __graft_entry__ as entry
    from oracle import feed_oracle as FO
    entry.smoke_feed(host)

    def once(fn):
        t0 = time.perf_counter()
        fn()
        return 1e3 * (time.perf_counter() - t0)

    img, _, _, centre, scale, rot = FO.synthetic_aug(2)
    frame, K, bbox_hand, p2d = FO.synthetic_frame(2)
    coef = feed.pil_coefficients(feed.crop_affine(centre, scale, 256, rot))[None]
    entry.smoke_feed_timing(host, img, coef, 0.37, [("hue", 0.1)], frame, K, bbox_hand, p2d, n_train=2, n_eval=2, timed=once)


def test_feed_bench_script(host):
    """scripts/feed_bench.py (the GPU timing of the feed, not yet run on a GPU) driven on the emulator at a tiny size."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("feed_bench", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "scripts", "feed_bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    line = mod.run(host, eval_frames=2, train_frames=2, steps=1, warmup=0, n_hand=24, n_obj=8, cpu_samples=1)
    assert line["train_batch"]["launches"] == 3 and line["eval_batch"]["frames_per_s"] > 0
    assert line["cpu_reference"]["ms_per_train_sample"] > 0


def test_the_patches_are_gone_afterwards():
    assert feed.lib is _capi.lib and feed._stream.__module__ == "hoisdf_b200.feed"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        feed.to_tensor(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))
