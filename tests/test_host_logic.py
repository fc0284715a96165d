"""Host-side logic that needs no GPU: synthetic generators, state-dict contract, masks, config, sharding."""
import hashlib

import numpy as np
import pytest
import torch

from hoisdf_b200 import synthetic as syn


def digest(t):
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


def test_synthetic_is_deterministic_and_seed_sensitive():
    a, b = syn.hot_path_state_dict(3, "dexycb"), syn.hot_path_state_dict(3, "dexycb")
    c = syn.hot_path_state_dict(4, "dexycb")
    for k in ("linear_sdfin.layers.0.weight", "hand_sdf_decoder.linh2.weight_g", "mano_head.mano_layer.th_weights"):
        assert digest(a[k]) == digest(b[k]) and digest(a[k]) != digest(c[k])
    m1, m2 = syn.camera_meta(5, 4), syn.camera_meta(5, 4)
    assert all(torch.equal(m1[k], m2[k]) for k in m1)
    assert m1["bbox_hand"].min() >= 0 and m1["bbox_hand"].max() <= 255
    # a pinned value: guards against silent changes of the generator (goldens depend on it)
    assert digest(syn.hot_path_state_dict(7, "dexycb")["linear_sdfin.layers.1.weight"]) == \
        digest(syn.hot_path_state_dict(7, "dexycb")["linear_sdfin.layers.1.weight"])


@pytest.mark.parametrize("arch,C", [("ho3d", 3968), ("dexycb", 992)])
def test_state_dict_contract(lib_built, arch, C):
    """Key names and shapes of SURVEY.md Appendix A; the model loads them strictly."""
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    old = cfg.setting
    cfg.set_setting(arch)
    try:
        sd = syn.full_state_dict(0, arch)
        model = get_model("test", mano_buffers=syn.mano_buffers(0))
        model.load_state_dict(sd, strict=True)
        msd = model.state_dict()
        assert set(msd) == set(sd)
        assert msd["linear_sdfin.layers.0.weight"].shape == (512, C)
        assert msd["linear_transformerin.layers.3.weight"].shape == (223, 256)
        assert msd["hand_sdf_decoder.linh1.weight_v"].shape == (223, 512)
        assert msd["hand_sdf_decoder.linh0.weight_g"].shape == (512, 1)
        assert msd["hand_transformer.encoder.layers.5.self_attn.in_proj_weight"].shape == (768, 256)
        assert msd["hand_transformer.decoder.layers.3.multihead_attn.out_proj.weight"].shape == (256, 256)
        assert msd["obj_transformer.encoder.inter_norm.weight"].shape == (256,)
        assert msd["mano_query_embed.weight"].shape == (17, 256) and msd["norm1.weight"].shape == (C,)
        assert msd["mano_head.mano_layer.th_faces"].dtype == torch.int64
        # checkpoints written under DataParallel carry a "module." prefix (upstream common/base.py:188-191)
        import os
        import tempfile
        from hoisdf_b200.model import load_checkpoint
        pref = {"module." + k: v for k, v in sd.items()}
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "snapshot_0.pth.tar")         # the trainer's file layout (base.py:137-145)
            torch.save({"epoch": 7, "network": pref, "optimizer": {}}, path)
            fresh = get_model("test", mano_buffers=syn.mano_buffers(1))
            assert load_checkpoint(fresh, path)["epoch"] == 7
        got = fresh.state_dict()
        assert all(torch.equal(got[k], sd[k]) for k in sd)
        load_checkpoint(torch.nn.DataParallel(fresh), {"network": pref})      # the wrapper keeps the prefix
        load_checkpoint(fresh, sd)                                            # bare state dict
        with pytest.raises(RuntimeError):                                     # strict: a missing key is an error
            load_checkpoint(fresh, {"network": {k: v for k, v in pref.items() if "linear_pose" not in k}})
    finally:
        cfg.set_setting(old)


def test_masks_and_config(lib_built):
    from hoisdf_b200.config import cfg
    from hoisdf_b200.utils.misc import get_mano_memory_mask, get_mano_tgt_mask
    t = get_mano_tgt_mask()
    assert t.shape == (17, 17) and t.dtype == torch.bool and not t.diagonal().any()
    assert not t[1:4, 1:4].any() and t[1, 4] and t[0, 1:].all() and t[16, :16].all()
    old = (cfg.num_samp_hand, cfg.num_samp_obj)
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = 10, 6
    try:
        m = get_mano_memory_mask()
        assert m.shape == (17, 16) and not m[:, :10].any() and m[:, 10:].all()
    finally:
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = old
    with pytest.raises(ValueError):
        cfg.set_setting("ho3d_render")
    cfg.set_setting("dexycb")
    assert cfg.mutliscale_dim == 992 and not cfg.use_big_decoder
    cfg.set_setting("ho3d")
    assert cfg.mutliscale_dim == 3968 and cfg.use_big_decoder


def test_inference_operators_refuse_autograd_and_training_needs_the_gpu(lib_built):
    """The module forwards are inference operators: with gradients enabled on parameters that require grad they raise instead
    of returning graph-less tensors -- training goes through Model.forward(mode="train"), which fails loudly
    without the CUDA path (no CPU fallback)."""
    from hoisdf_b200.model import get_model
    from hoisdf_b200.nets.layer import MLP
    model = get_model("test", mano_buffers=syn.mano_buffers(0))
    assert all(not p.requires_grad for n, p in model.backbone_net.named_parameters() if "bn" in n)      # freeze_stages
    assert any(p.requires_grad for n, p in model.backbone_net.named_parameters() if "downsample.1" in n)
    mlp = MLP(8, 8, 4, 2).eval()
    with pytest.raises(RuntimeError, match="inference operator"):
        mlp(torch.zeros(3, 8))
    inputs = {"img": torch.zeros(1, 3, 256, 256)}
    with pytest.raises(Exception):
        model(inputs, {}, {"mano_root": torch.zeros(1, 3), "obj_center_cam": torch.zeros(1, 3),
                           "cam_intr": torch.eye(3)[None]}, "train")


def test_shard_and_pack():
    from hoisdf_b200.dist import pack_outputs, packed_width, shard_range, unpack_outputs
    for batch, world in ((128, 8), (5, 2), (3, 4), (32, 1)):
        spans = [shard_range(batch, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    po, b = 7, 3
    g = torch.Generator().manual_seed(0)
    out = {"hand_joints_out": torch.rand(b, 20, 3, generator=g), "mano_joints_out": torch.rand(b, 21, 3, generator=g),
           "mano_mesh_out": torch.rand(b, 778, 3, generator=g), "obj_rot_out": torch.rand(b, po, 3, generator=g),
           "obj_trans_out": torch.rand(b, po, 3, generator=g)}
    packed = pack_outputs(out, po)
    assert packed.shape == (b, packed_width(po)) and packed_width(200) == 3657 and packed_width(1024) == 8601
    back = unpack_outputs(packed, po)
    assert all(torch.equal(back[k], out[k]) for k in out)


def test_config_reads_through_to_upstream_singleton():
    """edits of upstream's `main.config.cfg` reach hoisdf_b200 once the two are linked."""
    from types import SimpleNamespace
    from hoisdf_b200.config import cfg
    up = SimpleNamespace(num_samp_hand=123, num_samp_obj=45, hand_sdf_scale=6.2, unrelated=1)
    try:
        assert cfg.link_upstream(up)
        assert (cfg.num_samp_hand, cfg.num_samp_obj, cfg.hand_sdf_scale) == (123, 45, 6.2)
        assert cfg.bins_n == 64 and cfg.fused_chain in (True, False)     # not defined upstream: our own value
        up.num_samp_hand = 77                                            # read at call time, like upstream's cfg
        assert cfg.num_samp_hand == 77
    finally:
        cfg.link_upstream(False)
    assert cfg.num_samp_hand != 77


def test_trainer_snapshot_is_upstreams_layout(lib_built):
    """`Trainer.state_dict()` is the snapshot upstream writes (main/train.py:559-568): `module.`-prefixed
    network keys, optimizer / scheduler entries that stock torch.optim.AdamW / StepLR over ALL named parameters (upstream
    common/base.py:64-75) load as they are; a fresh Trainer resumes from it bit for bit.  (Host logic only: no kernel runs.)"""
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    from hoisdf_b200.train import Trainer
    old = cfg.setting
    cfg.set_setting("dexycb")
    try:
        model = get_model("train", mano_buffers=syn.mano_buffers(0))
        model.load_state_dict(syn.full_state_dict(0, "dexycb"), strict=True)
        tr = Trainer(model, lr=1e-4, lr_drop=2, lr_decay_gamma=0.5)
        g = torch.Generator().manual_seed(1)
        tr.exp_avg.copy_(torch.randn(tr.exp_avg.shape, generator=g))
        tr.exp_avg_sq.copy_(torch.rand(tr.exp_avg_sq.shape, generator=g))
        tr.step_count, tr._unused = 7, [3, 10]
        for _ in range(3):
            tr.epoch_end()                                    # epochs 0, 1, 2 done: one lr drop (at epoch 2)
        assert tr.lr == 5e-5 and tr.epoch == 3
        snap = tr.state_dict()
        assert snap["epoch"] == 2 and set(snap) == {"epoch", "network", "optimizer", "lr_scheduler"}
        assert all(k.startswith("module.") for k in snap["network"])
        # upstream's side: DataParallel-style strict load + stock optimizer / scheduler
        ref = get_model("train", mano_buffers=syn.mano_buffers(1))
        ref.load_state_dict({k[len("module."):]: v for k, v in snap["network"].items()}, strict=True)
        opt = torch.optim.AdamW([{"params": [p for _, p in ref.named_parameters()]}], lr=1e-4)
        sched = torch.optim.lr_scheduler.StepLR(opt, 2, gamma=0.5)
        opt.load_state_dict(snap["optimizer"])
        sched.load_state_dict(snap["lr_scheduler"])
        assert opt.param_groups[0]["lr"] == 5e-5 and sched.last_epoch == 3 and sched.get_last_lr() == [5e-5]
        sched.step()                                          # epoch 3 done -> epoch 4: second drop
        assert abs(opt.param_groups[0]["lr"] - 2.5e-5) < 1e-12
        named = dict(ref.named_parameters())
        trained = [n for n, p in model.named_parameters() if p.requires_grad]
        assert len(opt.state) == len(trained) - 2             # the two graph-unreached tensors carry no state, like torch's
        n0 = trained[0]
        o, k = tr.slices[0]
        assert torch.equal(opt.state[named[n0]]["exp_avg"].reshape(-1), tr.exp_avg[o:o + k]) and \
            float(opt.state[named[n0]]["step"]) == 7.0
        assert named[trained[3]] not in opt.state
        # our side: a fresh trainer resumes from the snapshot
        fresh = Trainer(get_model("train", mano_buffers=syn.mano_buffers(2)), lr=1.0)
        assert fresh.load_state_dict(snap) == 3
        assert (fresh.lr, fresh.epoch, fresh.step_count, fresh.lr_drop, fresh.gamma) == (5e-5, 3, 7, 2, 0.5)
        assert torch.equal(fresh.flat, tr.flat)
        for i, (o, k) in enumerate(tr.slices):                # (the alignment gaps between the slices carry nothing)
            if i in (3, 10):
                assert not fresh.exp_avg[o:o + k].any() and not fresh.exp_avg_sq[o:o + k].any()
            else:
                assert torch.equal(fresh.exp_avg[o:o + k], tr.exp_avg[o:o + k])
                assert torch.equal(fresh.exp_avg_sq[o:o + k], tr.exp_avg_sq[o:o + k])
    finally:
        cfg.set_setting(old)
