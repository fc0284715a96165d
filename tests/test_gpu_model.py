"""End-to-end parity of the drop-in `Model` operators against the CPU oracle on identical seeded inputs:
sdf_infer (bit-exact candidate masks, identical selected index sets), sdf_forward, get_input_transformer,
the whole hot path and the full image-to-pose forward.  Tolerance for the outputs is the north star's
1e-3 relative (to the tensor's max-abs); measured margins are ~100x smaller."""
import pytest
import torch

from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O
from util import align_selection, aligned

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(b.abs().max(), 1e-30))


def to_dev(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


@pytest.fixture(scope="module", params=["dexycb", "ho3d"])
def setup(request, cuda):
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    arch = request.param
    old = (cfg.setting, cfg.num_samp_hand, cfg.num_samp_obj)
    cfg.set_setting(arch)
    type(cfg).dataset = "ho3d"          # eval extras of the dexycb DATASET are not part of this path
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = 96, 40
    seed, B = 5, 2
    sd = syn.full_state_dict(seed, arch)
    model = get_model("test", mano_buffers=syn.mano_buffers(seed))
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).eval()
    meta = syn.camera_meta(seed, B)
    pyr = syn.feature_pyramid(seed, B, arch)
    ocfg = O.default_cfg(num_samp_hand=96, num_samp_obj=40)
    yield dict(arch=arch, model=model, sd=sd, meta=meta, pyr=pyr, ocfg=ocfg, B=B, seed=seed, dev=cuda)
    cfg.set_setting(old[0])
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = old[1], old[2]


def check_selection(taps_dev, taps_ref, P):
    """identical candidate masks (bit-exact) and identical selected index lists; any mismatch must be
    explained by an |sdf| difference below the k-th gap (SURVEY.md section 7 'Top-k parity').
    Returns (max error of the first-stage SDF of ALL candidates, max error of the values the final selection ranked)."""
    offs = taps_dev["offsets"]
    cand = taps_dev["cand_index"].cpu().long()
    sdf = taps_dev["cand_sdf"].cpu()
    B = taps_ref["index"].shape[0]
    full = taps_dev["exact_sdf"].numel() == sdf.numel()          # no cascade: every candidate was ranked exactly
    if not full:
        ex_sdf, ex_idx = taps_dev["exact_sdf"].cpu().view(B, -1), taps_dev["exact_index"].cpu().long().view(B, -1)
    worst_all = worst = 0.0
    for b in range(B):
        got_c = cand[offs[b]:offs[b + 1]]
        assert torch.equal(got_c, taps_ref["cand_index"][b]), "candidate (bbox) mask differs"
        worst_all = max(worst_all, (sdf[offs[b]:offs[b + 1]] - taps_ref["cand_sdf"][b]).abs().max().item())
        if full:
            d = (taps_dev["exact_sdf"].cpu()[offs[b]:offs[b + 1]] - taps_ref["cand_sdf"][b]).abs().max().item()
        else:                                       # re-ranked subset: look the oracle values up by lattice index
            pos = torch.searchsorted(taps_ref["cand_index"][b].contiguous(), ex_idx[b].contiguous())
            assert torch.equal(taps_ref["cand_index"][b][pos], ex_idx[b])
            d = (ex_sdf[b] - taps_ref["cand_sdf"][b][pos]).abs().max().item()
        worst = max(worst, d)
        got, want = taps_dev["index"][b].cpu().long(), taps_ref["index"][b]
        if not torch.equal(got, want):
            # allow swaps only between candidates whose oracle |sdf| differ by less than the numeric noise
            ref_abs = dict(zip(taps_ref["cand_index"][b].tolist(), taps_ref["cand_sdf"][b].abs().tolist()))
            for g, w in zip(got.tolist(), want.tolist()):
                if g != w:
                    assert abs(ref_abs[g] - ref_abs[w]) < 4 * max(d, 1e-8), (b, g, w, ref_abs[g], ref_abs[w], d)
    return worst_all, worst


def test_sdf_infer_c_entry_equals_python_orchestration(setup, monkeypatch):
    """hoisdf_sdf_infer_fwd (ONE C call: compaction, stage A, screening, final stage, device-side verdict, top-P) against
    the Python orchestration of the same kernels: identical to the bit, including the verdict's gap / error values."""
    from hoisdf_b200.config import cfg
    from hoisdf_b200 import _capi
    s = setup
    model, dev = s["model"], s["dev"]
    meta = to_dev(s["meta"], dev)
    pyr = to_dev(s["pyr"], dev)
    for kind, ck, bk, P in (("hand", "mano_root", "bbox_hand", 96), ("obj", "obj_center_cam", "bbox_obj", 40)):
        res = {}
        for native in (True, False):
            monkeypatch.setattr(type(cfg), "native_sdf_infer", native)
            taps = {}
            with torch.no_grad():
                ctx = model._ctx(pyr)
                out = model.sdf_infer(ctx, meta[ck], meta["cam_intr"], meta[bk], 3.1, P, kind, taps=taps)
            assert bool(taps.get("native", False)) == native and taps["single_pass"]
            res[native] = (out, taps)
        (a, ta), (b, tb) = res[True], res[False]
        for x, y in zip(a[:3], b[:3]):
            assert torch.equal(x, y)
        for k in ("index", "cand_index", "cand_sdf", "exact_sdf", "exact_index", "screen_gap", "screen_rows"):
            assert torch.equal(ta[k].reshape(-1).to(tb[k].dtype), tb[k].reshape(-1)), k
        assert float(ta["screen_err"]) == float(tb["screen_err"]) and bool(ta["screen_verified"]) == bool(tb["screen_verified"])
    # error contract: fewer lattice points inside the bbox than num_points -> the upstream-like exception
    tiny = meta["bbox_hand"].clone()
    tiny[:, 2:] = tiny[:, :2] + 2.0
    monkeypatch.setattr(type(cfg), "native_sdf_infer", True)
    with pytest.raises(RuntimeError, match="fewer than num_points"), torch.no_grad():
        model.sdf_infer(model._ctx(pyr), meta["mano_root"], meta["cam_intr"], tiny, 3.1, 96, "hand")
    # workspace contract of the C entry: a too-small row budget is refused before anything is launched
    import ctypes as C
    from hoisdf_b200 import ops
    plan = model.plan_candidates(meta["mano_root"], meta["cam_intr"], meta["bbox_hand"], 3.1)
    host = plan.host_offsets()
    a = _capi.SdfInferArgs()
    ctx = model._ctx(pyr)
    sdfin, packed = model.linear_sdfin.packed(), model.hand_sdf_decoder.packed()
    gm, gm16 = ops.make_pyramid(ctx.gmaps, cfg.input_img_shape), ops._pyramid_h(ctx.gmaps16, cfg.input_img_shape)
    small = 1024
    nbytes = int(_capi.lib.hoisdf_sdf_infer_workspace_bytes(2, small, 96, 1024, 64))
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    outs = [torch.empty(2 * 96 * 30, device=dev) for _ in range(8)]
    a.center, a.cam_intr, a.bbox = plan.center.data_ptr(), plan.cam_intr.data_ptr(), plan.bbox.data_ptr()
    a.sdf_scale, a.bins, a.batch, a.num_points, a.margin, a.clamp = 3.1, 64, 2, 96, 1024, 0.15
    a.gmaps, a.gmaps16, a.bias0 = C.addressof(gm), C.addressof(gm16), sdfin[0].b.data_ptr()
    s1 = sdfin[1].h3
    a.s1_a, a.s1_b, a.s1_c, a.ld_s1, a.b_s1, a.s1_scale = s1.plane_ptr(0), s1.plane_ptr(1), s1.plane_ptr(2), s1.ld, \
        sdfin[1].b.data_ptr(), 1.0
    a.dec, a.workspace, a.workspace_bytes, a.max_rows = C.addressof(packed.struct_h3), ws.data_ptr(), nbytes, small
    a.planned, a.chunk_counts, a.offsets, a.host_offsets = 1, plan.counts.data_ptr(), plan.offsets.data_ptr(), host.data_ptr()
    a.points, a.sdf, a.posenc, a.sel_index, a.status_flag, a.screen_err, a.screen_gap, a.verified = \
        (t.data_ptr() for t in outs)
    assert int(host[-1]) > small
    assert _capi.lib.hoisdf_sdf_infer_fwd(C.byref(a), ops._stream()) == _capi.E_WORKSPACE
    # planned = 0: the entry plans, waits for the B + 1 offsets itself (its own counts / offsets / pinned buffer) and must
    # give exactly what the planned call gives
    rows = ops.round_up(int(host[-1]), 1 << 16)
    nb = int(_capi.lib.hoisdf_sdf_infer_workspace_bytes(2, rows, 96, 1024, 64))
    ws2 = torch.empty(nb, device=dev, dtype=torch.uint8)
    res = {}
    for planned in (1, 0):
        o = dict(points=torch.empty(2, 96, 3, device=dev), sdf=torch.empty(2, 96, 1, device=dev),
                 posenc=torch.empty(2, 96, 30, device=dev), sel=torch.empty(2, 96, device=dev, dtype=torch.int32),
                 flag=torch.zeros(1, device=dev, dtype=torch.int32), err=torch.empty(1, device=dev),
                 gap=torch.empty(2, device=dev), ok=torch.empty(1, device=dev, dtype=torch.int32))
        pinned = torch.empty(3, dtype=torch.int64).pin_memory()
        nf = torch.zeros(2, dtype=torch.int64)
        a.workspace, a.workspace_bytes, a.max_rows, a.planned = ws2.data_ptr(), nb, rows, planned
        if planned:
            a.chunk_counts, a.offsets, a.host_offsets = plan.counts.data_ptr(), plan.offsets.data_ptr(), host.data_ptr()
        else:
            a.chunk_counts, a.offsets, a.host_offsets = None, None, pinned.data_ptr()
        a.n_f = nf.data_ptr()
        a.points, a.sdf, a.posenc, a.sel_index, a.status_flag = (o[k].data_ptr() for k in ("points", "sdf", "posenc", "sel", "flag"))
        a.screen_err, a.screen_gap, a.verified = o["err"].data_ptr(), o["gap"].data_ptr(), o["ok"].data_ptr()
        assert _capi.lib.hoisdf_sdf_infer_fwd(C.byref(a), ops._stream()) == 0
        torch.cuda.synchronize()
        assert torch.equal(nf, host[1:] - host[:-1]) and int(o["ok"]) == 1
        if not planned:
            assert torch.equal(pinned, host)
        res[planned] = o
    for k in res[1]:
        assert torch.equal(res[1][k], res[0][k]), k
    a.workspace, a.workspace_bytes, a.max_rows, a.planned = ws.data_ptr(), nbytes, small, 1
    a.chunk_counts, a.offsets, a.host_offsets, a.n_f = plan.counts.data_ptr(), plan.offsets.data_ptr(), host.data_ptr(), None
    a.workspace_bytes = 1024
    assert _capi.lib.hoisdf_sdf_infer_fwd(C.byref(a), ops._stream()) == _capi.E_WORKSPACE
    a.workspace = None
    assert _capi.lib.hoisdf_sdf_infer_fwd(C.byref(a), ops._stream()) == -1
    assert _capi.lib.hoisdf_sdf_infer_workspace_bytes(2, 1 << 16, 96, 1024, 64) > 0
    assert _capi.lib.hoisdf_sdf_infer_keep(96, 1024) == 1120 and _capi.lib.hoisdf_sdf_infer_keep(8000, 1024) == 8192


@pytest.mark.parametrize("final_stage", ["h3", "fma"])
def test_sdf_infer(setup, final_stage, monkeypatch):
    from hoisdf_b200.config import cfg
    monkeypatch.setattr(type(cfg), "final_stage", final_stage)
    m, s = setup["model"], setup
    dev = s["dev"]
    pyr_d, meta_d = to_dev(s["pyr"], dev), to_dev(s["meta"], dev)
    for kind, ck, bk, P in (("hand", "mano_root", "bbox_hand", 96), ("obj", "obj_center_cam", "bbox_obj", 40)):
        taps, otaps = {}, {}
        with torch.no_grad():
            pts, sdf, pe, cls = m.sdf_infer(pyr_d, meta_d[ck], meta_d["cam_intr"], meta_d[bk], 3.1, P, kind, taps=taps)
        opts, osdf, ope, _ = O.sdf_infer(dict(s["sd"]), s["pyr"], s["meta"][ck], s["meta"]["cam_intr"], s["meta"][bk],
                                         3.1, P, kind, s["ocfg"], otaps)
        worst_all, worst = check_selection(taps, otaps, P)
        assert worst < 5e-6, worst                      # the values the final selection ranked, absolute (|sdf| < 1)
        if "screen_gap" in taps:
            # coarse-to-fine cascade: every candidate ranked on the tensor cores (single fp16 product: `worst_all`
            # ~1e-4; FP16x3: ~1e-6), the best rows re-ranked by the next, more accurate stage, the last one being the
            # fp32 FMA kernels.  The result is provably an all-fp32 selection when each stage's error is below the
            # |sdf| gap it leaves -- checked on the device.
            assert bool(taps["screen_verified"])
            assert float(taps["screen_gap"].min()) > 3 * float(taps["screen_err"]) > 0
            if taps["single_pass"]:
                first = "pre" if "pre_gap" in taps else "screen"      # the step that follows the single-product stage
                assert bool(taps[first + "_verified"])
                assert float(taps[first + "_gap"].min()) > 3 * float(taps[first + "_err"]) > 0
                assert worst_all < 3e-3 and float(taps[first + "_gap"].min()) > worst_all, (worst_all, taps[first + "_gap"])
            else:
                assert worst_all < 5e-6 and float(taps["screen_gap"].min()) > 3 * worst_all
        else:
            assert worst_all < 5e-6, worst_all
        assert cls is None and pts.shape == (s["B"], P, 3) and sdf.shape == (s["B"], P, 1) and pe.shape == (s["B"], P, 30)
        if torch.equal(taps["index"].cpu().long(), otaps["index"]):
            assert torch.equal(pts.cpu(), opts)         # lattice coordinates are bit-exact
            assert (sdf.cpu() - osdf).abs().max() < 5e-6 and (pe.cpu() - ope).abs().max() < 2e-6


@pytest.mark.parametrize("level", [1, 2])
def test_sdf_infer_cascade_levels(setup, level):
    """The fallback levels of the selection cascade (1: FP16x3 on every candidate -> final stage; 2: fp32 FMA kernels on
    every candidate) select what the oracle selects."""
    m, s = setup["model"], setup
    dev = s["dev"]
    pyr_d, meta_d = to_dev(s["pyr"], dev), to_dev(s["meta"], dev)
    taps, otaps = {}, {}
    with torch.no_grad():
        m.sdf_infer(pyr_d, meta_d["mano_root"], meta_d["cam_intr"], meta_d["bbox_hand"], 3.1, 96, "hand", taps=taps,
                    level=level)
    O.sdf_infer(dict(s["sd"]), s["pyr"], s["meta"]["mano_root"], s["meta"]["cam_intr"], s["meta"]["bbox_hand"], 3.1, 96,
                "hand", s["ocfg"], otaps)
    worst_all, worst = check_selection(taps, otaps, 96)
    assert not taps["single_pass"] and worst_all < 5e-6 and worst < 5e-6
    assert ("screen_gap" in taps) == (level == 1)


def test_unverified_cascade_escalates(setup, monkeypatch):
    """Margins too small for the device-side check to pass: `hot_path` reads the verdict after queueing the forward and
    re-runs it at the next cascade level until the selection is verified (level 1 if its own check happens to pass with
    a margin of one row, else level 2: no screening at all); a direct
    `sdf_infer` call escalates synchronously.  Either way the result is the oracle's selection."""
    from hoisdf_b200.config import cfg
    m, s = setup["model"], setup
    dev = s["dev"]
    monkeypatch.setattr(type(cfg), "screen_margin_single", 2)
    monkeypatch.setattr(type(cfg), "screen_margin_safe", 1)
    out = m.hot_path(to_dev(s["pyr"], dev), to_dev(s["meta"], dev))
    taps = m.last_taps
    for kind in ("hand", "obj"):
        assert not taps[kind]["single_pass"]                                             # stage A was abandoned
        assert "screen_gap" not in taps[kind] or bool(taps[kind]["screen_verified"])     # level 1 verified, or level 2
    otaps = {}
    with torch.no_grad():
        oout = O.hot_path_eval(dict(s["sd"]), s["pyr"], s["meta"], s["ocfg"], otaps)
    check_selection(taps["hand"], otaps["hand"], 96)
    check_selection(taps["obj"], otaps["obj"], 40)
    for k in ("mano_mesh_out", "mano_joints_out", "hand_joints_out"):
        assert rel(out[k], oout[k]) < 1e-4, (k, rel(out[k], oout[k]))
    pyr_d, meta_d = to_dev(s["pyr"], dev), to_dev(s["meta"], dev)
    with torch.no_grad():
        pts, _, _, _ = m.sdf_infer(pyr_d, meta_d["mano_root"], meta_d["cam_intr"], meta_d["bbox_hand"], 3.1, 96, "hand")
    assert torch.equal(pts, taps["hand_points"])


def test_sdf_infer_too_few_candidates(setup):
    m, s = setup["model"], setup
    dev = s["dev"]
    meta_d = to_dev(s["meta"], dev)
    bbox = meta_d["bbox_hand"].clone()
    bbox[1] = torch.tensor([10.0, 10.0, 10.5, 10.5], device=dev)     # (almost) empty box for sample 1
    with pytest.raises(RuntimeError):
        m.sdf_infer(to_dev(s["pyr"], dev), meta_d["mano_root"], meta_d["cam_intr"], bbox, 3.1, 96, "hand")


def test_sdf_forward_and_point_features(setup):
    m, s = setup["model"], setup
    dev = s["dev"]
    B, P = s["B"], 77
    g = torch.Generator().manual_seed(3)
    pts = torch.rand(B, P, 3, generator=g) * 2 - 1
    pyr_d, meta_d = to_dev(s["pyr"], dev), to_dev(s["meta"], dev)
    sd = dict(s["sd"])
    with torch.no_grad():
        sdf, cls, pe = m.sdf_forward(pyr_d, pts.to(dev), meta_d["obj_center_cam"], meta_d["cam_intr"], 3.1, "obj")
        lat, cam = m.get_input_transformer(pyr_d, pts.to(dev), meta_d["mano_root"], meta_d["cam_intr"], 3.1)
    osdf, _, ope = O.sdf_forward(sd, s["pyr"], pts, s["meta"]["obj_center_cam"], s["meta"]["cam_intr"], 3.1, "obj", s["ocfg"])
    olat, ocam = O.get_input_transformer(sd, s["pyr"], pts, s["meta"]["mano_root"], s["meta"]["cam_intr"], 3.1, s["ocfg"])
    assert sdf.shape == (B, P, 1) and pe.shape == (B, P, 30) and lat.shape == (B, P, 223)
    assert (sdf.cpu() - osdf).abs().max() < 5e-6 and (pe.cpu() - ope).abs().max() < 2e-6
    # K = 3968 contraction on the tensor cores: accumulate-truncation error grows with K (1.5e-5 measured at 3968)
    assert rel(lat, olat) < 2.5e-5 and torch.equal(cam.cpu(), ocam)


def test_hot_path_outputs(setup):
    m, s = setup["model"], setup
    dev = s["dev"]
    out = m.hot_path(to_dev(s["pyr"], dev), to_dev(s["meta"], dev))
    otaps = {}
    with torch.no_grad():
        oout = O.hot_path_eval(dict(s["sd"]), s["pyr"], s["meta"], s["ocfg"], otaps)
    taps = m.last_taps
    check_selection(taps["hand"], otaps["hand"], 96)
    check_selection(taps["obj"], otaps["obj"], 40)
    # identical selected SETS; order identical up to swaps of |sdf| near-ties (tests/util.py)
    hp = align_selection(taps["hand"]["index"], otaps["hand"]["index"], otaps["hand_sdf"])
    op = align_selection(taps["obj"]["index"], otaps["obj"]["index"], otaps["obj_sdf"])
    S, Ph = 96 + 40, 96
    # token order inside each of the four groups follows the selection order: align before comparing
    def tok_perm(first, second, n_first):
        return [torch.cat([a, b + n_first]) for a, b in zip(first, second)]
    hand_tok, obj_tok = tok_perm(hp, op, 96), tok_perm(op, hp, 40)
    assert rel(aligned(taps["hand_transformer_in"], hand_tok), otaps["hand_transformer_in"].transpose(0, 1)) < 1e-5
    assert rel(aligned(taps["obj_transformer_in"], obj_tok), otaps["obj_transformer_in"].transpose(0, 1)) < 1e-5
    enc = taps["hand_encoder_out"].cpu()
    enc = torch.stack([aligned(enc[l], hand_tok) for l in range(enc.shape[0])])
    assert rel(enc, otaps["hand_encoder_out"].transpose(1, 2)) < 1e-4
    assert rel(taps["hs"], otaps["hs"].transpose(1, 2)) < 1e-4
    for k in oout:
        got = aligned(out[k], op) if k in ("obj_rot_out", "obj_trans_out") else out[k]
        assert got.shape == oout[k].shape, k
        assert rel(got, oout[k]) < 1e-3, (k, rel(got, oout[k]))     # north-star tolerance
        assert rel(got, oout[k]) < 1e-4, (k, rel(got, oout[k]))     # what the fp32-grade kernels actually deliver


def test_full_forward_from_image(setup):
    """Image -> pose through the cuDNN backbone + our hot path vs the all-CPU oracle.  cuDNN and MKL-DNN
    convolutions differ at ~1e-6 relative, which can legitimately flip near-tied selections, so this gate is
    looser: >= 97 % of the selected points must coincide, per-point outputs are compared on the common points,
    global outputs (joints, mesh) to 2e-2 of their range."""
    m, s = setup["model"], setup
    dev = s["dev"]
    img = syn.image_batch(s["seed"], s["B"])
    otaps = {}
    with torch.no_grad():
        oout = O.model_eval(dict(s["sd"]), img, s["meta"], s["ocfg"], s["arch"], otaps)

    def compare(out):
        for k in ("loss_joint_3d", "loss_joint_cls", "loss_all_joint_3d", "obj_rot", "obj_trans"):
            assert k in out and out[k].dim() == 0
        taps = m.last_taps
        for kind in ("hand", "obj"):
            got, want = taps[kind]["index"].cpu().long(), otaps[kind]["index"]
            common = sum(len(set(g.tolist()) & set(w.tolist())) for g, w in zip(got, want))
            assert common >= 0.97 * want.numel(), (kind, common, want.numel())
        for k in ("mano_mesh_out", "mano_joints_out", "hand_joints_out"):
            assert out[k].shape == oout[k].shape and rel(out[k], oout[k]) < 2e-2, (k, rel(out[k], oout[k]))
        got, want = taps["obj"]["index"].cpu().long(), otaps["obj"]["index"]
        for k in ("obj_rot_out", "obj_trans_out"):
            assert out[k].shape == oout[k].shape
            for b in range(s["B"]):
                pos = {int(i): j for j, i in enumerate(want[b].tolist())}
                pairs = [(j, pos[int(i)]) for j, i in enumerate(got[b].tolist()) if int(i) in pos]
                gi, wi = zip(*pairs)
                assert rel(out[k][b, list(gi)], oout[k][b, list(wi)]) < 2e-2, k

    compare(m({"img": img.to(dev)}, to_dev(syn.eval_targets(s["B"]), dev), to_dev(s["meta"], dev), "eval"))
    # channels_last backbone: pyramid consumed zero-copy
    m.channels_last_()
    compare(m({"img": img.to(dev)}, to_dev(syn.eval_targets(s["B"]), dev), to_dev(s["meta"], dev), "eval"))


def test_image_encoder_matches_oracle(setup):
    """ResNet-50 + U-Net on the FP16x3 tensor-core convolution kernels against the ORACLE's encoder
    (oracle backbone / unet_decoder, pinned to upstream by tests/golden/image_*.npz), every pyramid level and the
    heat-map / segmentation head, <= 1e-4 of the level's range, both architectures."""
    m, s = setup["model"], setup
    img = syn.image_batch(s["seed"] + 20, s["B"])
    with torch.no_grad():
        feat, skips = O.backbone(dict(s["sd"]), img)
        opyr, odec = O.unet_decoder(dict(s["sd"]), feat, skips, s["arch"])
        pyr, dec = m.run_image_encoder(img.to(s["dev"]))
    for name, want in opyr.items():
        got = pyr[name]
        assert tuple(got.shape) == tuple(want.shape), name
        assert rel(got, want) < 1e-4, (name, rel(got, want))
    assert tuple(dec.shape) == tuple(odec.shape) and rel(dec, odec) < 1e-4, rel(dec, odec)


def test_image_to_pose_tight(setup):
    """the exact path bench.py times (image -> pose), gated tightly.  The oracle is fed the pyramid the
    GPU encoder produced (its own encoder differs from ours at ~1e-6, which may flip near-tied selections), so the
    selected index SETS must be identical and every `*_out` within the north star's 1e-3 (measured ~1e-5)."""
    m, s = setup["model"], setup
    dev = s["dev"]
    img = syn.image_batch(s["seed"] + 21, s["B"])
    out = m({"img": img.to(dev)}, to_dev(syn.eval_targets(s["B"]), dev), to_dev(s["meta"], dev), "eval")
    taps = m.last_taps
    with torch.no_grad():
        pyr, _ = m.run_image_encoder(img.to(dev))
        cpu_pyr = {k: v.float().cpu().contiguous() for k, v in pyr.items()}
        otaps = {}
        oout = O.hot_path_eval(dict(s["sd"]), cpu_pyr, s["meta"], s["ocfg"], otaps)
    perms = {}
    for kind in ("hand", "obj"):
        want_sdf = torch.stack([torch.sort(c.abs())[0][:otaps[kind]["index"].shape[1]] for c in otaps[kind]["cand_sdf"]])
        perms[kind] = align_selection(taps[kind]["index"], otaps[kind]["index"], want_sdf)
    for k, want in oout.items():
        got = aligned(out[k], perms["obj"]) if k in ("obj_rot_out", "obj_trans_out") else out[k]
        assert got.shape == want.shape, k
        assert rel(got, want) < 1e-3, (k, rel(got, want))


def test_single_pass_screening_is_verified_or_falls_back(setup):
    """Opt-in single-pass TF32 screening: ~6e-5 error, so the device-side check (gap > 3 x observed error) decides
    between accepting it and redoing the screening with 3xTF32; either way the selection equals the oracle's."""
    from hoisdf_b200.config import cfg
    m, s = setup["model"], setup
    dev = s["dev"]
    old = cfg.screen_passes
    type(cfg).screen_passes = 1
    try:
        taps, otaps = {}, {}
        with torch.no_grad():
            m.sdf_infer(to_dev(s["pyr"], dev), to_dev(s["meta"], dev)["mano_root"], to_dev(s["meta"], dev)["cam_intr"],
                        to_dev(s["meta"], dev)["bbox_hand"], 3.1, 96, "hand", taps=taps)
        O.sdf_infer(dict(s["sd"]), s["pyr"], s["meta"]["mano_root"], s["meta"]["cam_intr"], s["meta"]["bbox_hand"],
                    3.1, 96, "hand", s["ocfg"], otaps)
        assert bool(taps["screen_verified"])
        align_selection(taps["index"], otaps["index"], torch.stack([torch.sort(c.abs())[0][:96] for c in otaps["cand_sdf"]]))
    finally:
        type(cfg).screen_passes = old


def test_dexycb_eval_branch(setup):
    """cfg.dataset == 'dexycb' (upstream model.py:370-422,606-654): GT-point SDF queries, heat-map / segmentation
    outputs, ground-truth MANO forward and every loss entry, from the image, against the oracle AND against the
    committed upstream fixture (tests/golden/dexycb_eval_seed14.npz)."""
    import os
    import numpy as np
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    if setup["arch"] != "dexycb":
        pytest.skip("the dexycb dataset branch runs on the dexycb architecture")
    dev = setup["dev"]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dexycb_eval_seed14.npz"))
    seed, B, ph, po = int(g["seed"]), int(g["batch"]), int(g["num_samp_hand"]), int(g["num_samp_obj"])
    old = (cfg.dataset, cfg.num_samp_hand, cfg.num_samp_obj)
    type(cfg).dataset, type(cfg).num_samp_hand, type(cfg).num_samp_obj = "dexycb", ph, po
    try:
        sd = syn.full_state_dict(seed, "dexycb")
        model = get_model("test", mano_buffers=syn.mano_buffers(seed))
        model.load_state_dict(sd, strict=True)
        model = model.to(dev).eval()
        img, meta = syn.image_batch(seed, B), syn.camera_meta(seed, B)
        inputs, targets = syn.dexycb_extras(seed, B, ph, po)
        tdev = to_dev(targets, dev)
        before = tdev["mano_param"].clone()
        out = model({"img": img.to(dev), **to_dev(inputs, dev)}, tdev, to_dev(meta, dev), "eval")
        assert torch.equal(before, tdev["mano_param"])
        with torch.no_grad():
            oout = O.model_eval_dexycb(sd, img, inputs, targets, meta,
                                       O.default_cfg(dataset="dexycb", num_samp_hand=ph, num_samp_obj=po))
        assert set(out) == set(oout), set(out) ^ set(oout)
        # entries that do not pass through the top-k selection: tight
        for k in ("sdfhand_loss", "sdfobj_loss", "joint_heatmap", "obj_seg", "hand_seg", "joint_heatmap_out",
                  "hand_seg_pred_out", "obj_seg_pred_out", "mano_joints_gt_out", "mano_mesh_gt_out"):
            assert out[k].shape == oout[k].shape, k
            assert rel(out[k], oout[k]) < 1e-3, (k, rel(out[k], oout[k]))
            assert rel(out[k], torch.from_numpy(g[k])) < 1e-3, (k, "vs upstream fixture")
        for k in ("hand_seg_gt_out", "obj_seg_gt_out"):
            assert torch.equal(out[k].cpu(), targets[k.replace("_gt_out", "")])
        # the same branch with its static stages replayed from CUDA graphs (bench.py --config 3): bit-identical, twice
        model.enable_cuda_graphs()
        for _ in range(2):
            gout = model({"img": img.to(dev), **to_dev(inputs, dev)}, tdev, to_dev(meta, dev), "eval")
            assert set(gout) == set(out)
            for k in out:
                assert torch.equal(gout[k], out[k]), k
        model.enable_cuda_graphs(False)
        # entries behind the selection (cuDNN vs MKL-DNN pyramids can flip near-ties): same gate as the image test
        for k in ("mano_mesh_out", "mano_joints_out", "hand_joints_out", "loss_all_joint_3d", "mano_mesh_loss",
                  "mano_joint_loss", "pose_param_loss", "shape_param_loss", "loss_joint_cls", "obj_rot", "obj_trans"):
            assert out[k].shape == oout[k].shape, k
            assert rel(out[k], oout[k]) < 2e-2, (k, rel(out[k], oout[k]))
    finally:
        type(cfg).dataset, type(cfg).num_samp_hand, type(cfg).num_samp_obj = old


def test_cuda_graph_forward_matches_eager(setup):
    """`Model.enable_cuda_graphs()`: the static stages replayed from CUDA graphs give the eager forward's results bit for
    bit, also when the graphs are replayed on new inputs, and are recaptured after a parameter changes."""
    from hoisdf_b200.config import cfg
    if setup["arch"] != "ho3d":
        pytest.skip("one architecture is enough")
    m, s = setup["model"], setup
    dev = s["dev"]
    batches = [({"img": syn.image_batch(sd_, s["B"]).to(dev)}, to_dev(syn.eval_targets(s["B"]), dev),
                to_dev(syn.camera_meta(sd_, s["B"]), dev)) for sd_ in (s["seed"], s["seed"] + 1)]
    eager = [m(*bt, "eval") for bt in batches]
    m.enable_cuda_graphs()
    try:
        for rep in range(2):                     # second round: pure replays
            for bt, ref in zip(batches, eager):
                out = m(*bt, "eval")
                assert set(out) == set(ref)
                for k in ref:
                    assert torch.equal(out[k], ref[k]), (rep, k)
        assert len(m._graphs) == 2               # "weights" + one shape key
        with torch.no_grad():
            m.linear_pose.layers[0].bias.add_(0.01)          # parameter version changes -> graphs are rebuilt
        out = m(*batches[0], "eval")
        m.enable_cuda_graphs(False)
        ref = m(*batches[0], "eval")
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
        assert not torch.equal(ref["mano_mesh_out"], eager[0]["mano_mesh_out"])
    finally:
        m.enable_cuda_graphs(False)
        with torch.no_grad():
            m.linear_pose.layers[0].bias.sub_(0.01)


def test_drop_in_under_the_upstream_harness(setup, tmp_path):
    """drive the drop-in exactly like upstream's Tester (common/base.py:179-193) and test loop
    (main/test.py:119-129): get_model("test") -> .cuda() -> DataParallel -> torch.load(snapshot) ->
    load_state_dict(ckpt["network"], strict=True) with `module.`-prefixed keys -> .eval() -> model(inputs, targets,
    meta_info, "eval") on CPU batch tensors (DataParallel scatters them) -> split into `*_out` / losses, `.mean()`."""
    from torch.nn import DataParallel
    from hoisdf_b200.model import get_model
    s = setup
    snap = tmp_path / "snapshot_0_0.pth.tar"
    torch.save({"epoch": 0, "network": {"module." + k: v for k, v in s["sd"].items()}}, snap)   # common/base.py:113-118
    model = get_model("test", mano_buffers=syn.mano_buffers(s["seed"]))
    model = model.cuda()
    model = DataParallel(model, device_ids=[0])
    ckpt = torch.load(snap)
    model.load_state_dict(ckpt["network"], strict=True)
    model.eval()
    img = syn.image_batch(s["seed"] + 30, s["B"])
    inputs, targets, meta = {"img": img}, syn.eval_targets(s["B"]), s["meta"]
    with torch.no_grad():
        model_out = model(inputs, targets, meta, "eval")
    out = {k[:-4]: model_out[k] for k in model_out.keys() if "_out" in k}
    loss = {k: model_out[k] for k in model_out.keys() if "_out" not in k}
    loss = {k: loss[k].mean() for k in loss}
    assert set(out) == {"hand_joints", "mano_joints", "mano_mesh", "obj_rot", "obj_trans"}
    assert all(v.is_cuda and v.shape[0] == s["B"] for v in out.values()) and all(v.dim() == 0 for v in loss.values())
    # the wrapper adds nothing but the scatter: the wrapped module called directly on device tensors gives the same bits
    direct = model.module({"img": img.to(s["dev"])}, to_dev(targets, s["dev"]), to_dev(meta, s["dev"]), "eval")
    for k in out:
        assert torch.equal(out[k], direct[k + "_out"]), (k, float((out[k] - direct[k + "_out"]).abs().max()))
