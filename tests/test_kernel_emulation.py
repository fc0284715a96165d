"""The simple (non-tensor-core) CUDA kernels of csrc/metrics.cu, lattice.cu, topk.cu and heads.cu executed UNCHANGED on the host by a CPU thread emulator
(tests/emu/cuda_emu.h: one OS thread per CUDA thread, std::barrier for __syncthreads, an exchange buffer for warp
shuffles) and compared with the oracle -- so that the kernel source, its launch geometry and its C-ABI argument
handling are checked in the GPU-less suite too.  Test infrastructure: the emulated library is built from the same
.cu file with `g++ -DHOISDF_EMULATE` into tests/emu/_build/ and is never loaded by the product."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O
from util import rel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


def build_emulated(name):
    """g++ build of hoisdf_b200/csrc/<name>.cu against the emulator header -> tests/emu/_build/lib<name>_emu.so"""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    lib = os.path.join(out, "lib%s_emu.so" % name)
    src = os.path.join(ROOT, "hoisdf_b200", "csrc", name + ".cu")
    deps = [src, os.path.join(EMU, "cuda_emu.h"), os.path.join(ROOT, "include", "hoisdf_b200.h"),
            os.path.join(ROOT, "hoisdf_b200", "csrc", "common.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run([gxx, "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-DHOISDF_EMULATE", "-I" + EMU,
                        "-x", "c++", src, "-o", lib], check=True)
    return C.CDLL(lib)


@pytest.fixture(scope="module")
def emu():
    lib = build_emulated("metrics")
    lib.hoisdf_obj_metrics_workspace_bytes.restype = C.c_int64
    lib.hoisdf_obj_metrics_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_obj_metrics_fwd.argtypes = [vp, vp, i64, i64, vp, vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp]
    lib.hoisdf_mesh_metrics_fwd.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp, i64, vp]
    lib.hoisdf_hand_joint_metrics_fwd.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp]
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data


def f32(t):
    return np.ascontiguousarray(torch.as_tensor(t).numpy(), dtype=np.float32)


@pytest.mark.parametrize("B,N,votes,by_id", [(2, 300, 17, True), (1, 1, 1, True), (2, 1100, 300, False)])
def test_obj_metrics_kernels_on_the_emulator(emu, B, N, votes, by_id):
    m = syn.metric_inputs(40 + N, B, votes=votes, n_templates=3, n_verts=N)
    templates = torch.stack([t["verts"] for t in m["templates"]])
    ids = m["obj_cls_ids"] - 1
    args = (m["out"]["obj_rot"], m["out"]["obj_trans"], m["targets"]["obj_rot"], m["targets"]["rel_obj_trans"])
    want = O.obj_pose_metrics(templates, ids, *args)
    tm = f32(templates if by_id else templates[ids])
    idn = np.ascontiguousarray(ids.numpy(), dtype=np.int64) if by_id else None
    rp, tp, rg, tg = (f32(a) for a in args)
    out = np.full((4, B), np.nan, np.float32)
    nbytes = emu.hoisdf_obj_metrics_workspace_bytes(B, N)
    ws = np.zeros(nbytes // 4, np.float32)
    rc = emu.hoisdf_obj_metrics_fwd(ptr(tm), ptr(idn), tm.shape[0], N, ptr(rp), ptr(tp), votes, ptr(rg), ptr(tg), B,
                                    out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data, out[3].ctypes.data,
                                    ptr(ws), nbytes, None)
    assert rc == 0
    for name, got, ref in zip(("adds", "mme", "mce", "oce"), out, want):
        assert rel(got, ref) < 1e-5 or float(np.abs(got - ref.numpy()).max()) < 1e-9, (name, got, ref)
    # the same meshes given directly (compute_obj_metrics_* entry)
    tsel = templates[ids]
    pred = torch.bmm(tsel, O.batch_rodrigues(args[0].mean(1)).permute(0, 2, 1)) + args[1].mean(1)[:, None]
    tgt = torch.bmm(tsel, O.batch_rodrigues(args[2]).permute(0, 2, 1)) + args[3][:, None]
    out2 = np.full((3, B), np.nan, np.float32)
    pm, tmesh = f32(pred), f32(tgt)
    rc = emu.hoisdf_mesh_metrics_fwd(ptr(pm), ptr(tmesh), B, N, out2[0].ctypes.data, out2[1].ctypes.data,
                                     out2[2].ctypes.data, ptr(ws), nbytes, None)
    assert rc == 0
    for name, got, ref in zip(("adds", "mme", "mce"), out2, O.mesh_metrics(pred, tgt)):
        assert rel(got, ref) < 1e-5 or float(np.abs(got - ref.numpy()).max()) < 1e-9, (name, got, ref)
    assert emu.hoisdf_mesh_metrics_fwd(ptr(pm), ptr(tmesh), B, N, None, None, None, ptr(ws), 4, None) == -2


@pytest.mark.parametrize("J", [21, 300, 3])
def test_hand_joint_kernel_on_the_emulator(emu, J):
    gen = torch.Generator().manual_seed(J)
    B = 5
    gt = torch.randn(B, J, 3, generator=gen) * 0.08
    pred = gt * 1.2 + torch.randn(B, J, 3, generator=gen) * 0.01 + 0.02
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
    pred[1] = gt[1] @ q.T * 0.7 + 0.3
    pred[2] = gt[2] * torch.tensor([1.0, 1.0, -1.0])        # mirror image: the det < 0 branch (metrics.py:197-202)
    if J > 3:
        pred[3, :, 2] = 0.0
        gt[3, :, 2] = 0.0                                   # planar: rank-deficient cross-covariance
    p, g = f32(pred), f32(gt)
    mje, pamje, aligned = np.zeros(B, np.float32), np.zeros(B, np.float32), np.zeros((B, J, 3), np.float32)
    assert emu.hoisdf_hand_joint_metrics_fwd(ptr(p), ptr(g), B, J, ptr(mje), ptr(pamje), ptr(aligned), None) == 0
    omje, opamje = O.hand_joint_metrics(pred, gt)
    scale = float(gt.abs().max())
    assert rel(mje, omje) < 1e-5
    assert float(np.abs(pamje - opamje.numpy()).max()) < 1e-5 * scale
    for b in range(B):
        assert float(np.abs(aligned[b] - O.rigid_align(p[b], g[b])).max()) < 2e-5 * scale, b
    assert emu.hoisdf_hand_joint_metrics_fwd(None, None, B, J, None, None, None, None) == -1


def test_lattice_kernels_on_the_emulator():
    """Candidate generation (upstream main/model.py:257-302): the sheared lattice, projection, strict bbox test and the
    stable compaction of csrc/lattice.cu give the oracle's boolean mask and pixel coordinates BIT FOR BIT."""
    lib = build_emulated("lattice")
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_lattice_count.argtypes = [vp, vp, vp, C.c_float, i64, i32, vp, vp, vp]
    lib.hoisdf_lattice_compact.argtypes = [vp, vp, vp, C.c_float, i64, i32, vp, vp, vp, vp, vp]
    lib.hoisdf_project_points.argtypes = [vp, vp, vp, C.c_float, i64, i64, vp, vp, vp]
    B, bins = 2, 64
    meta = syn.camera_meta(9, B)
    center, K, bbox = f32(meta["obj_center_cam"]), f32(meta["cam_intr"]), f32(meta["bbox_obj"])
    chunks = lib.hoisdf_lattice_chunks(bins)
    counts, offsets = np.zeros(B * chunks, np.int32), np.zeros(B + 1, np.int64)
    assert lib.hoisdf_lattice_count(ptr(center), ptr(K), ptr(bbox), 3.1, B, bins, ptr(counts), ptr(offsets), None) == 0
    total = int(offsets[-1])
    cand, uv = np.full(total, -1, np.int32), np.full((total, 2), np.nan, np.float32)
    assert lib.hoisdf_lattice_compact(ptr(center), ptr(K), ptr(bbox), 3.1, B, bins, ptr(counts), ptr(offsets), ptr(cand),
                                      ptr(uv), None) == 0
    lat = O.lattice(bins)
    for b in range(B):
        mask, ouv = O.candidate_mask(lat, meta["obj_center_cam"][b], meta["cam_intr"][b], meta["bbox_obj"][b], 3.1)
        want = mask.nonzero().flatten().numpy()
        got = cand[offsets[b]:offsets[b + 1]]
        assert 3000 < len(want) < bins ** 3 and np.array_equal(got, want)                 # the index mask, exactly
        assert np.array_equal(uv[offsets[b]:offsets[b + 1]], ouv[mask].numpy())          # projected pixels, bit for bit
    # explicit points (model.py:148-150,190-192)
    P = 77
    pts = f32(torch.rand(B, P, 3, generator=torch.Generator().manual_seed(1)) * 2 - 1)
    cam, puv = np.zeros((B, P, 3), np.float32), np.zeros((B, P, 2), np.float32)
    assert lib.hoisdf_project_points(ptr(pts), ptr(center), ptr(K), 3.1, B, P, ptr(cam), ptr(puv), None) == 0
    ocam = torch.from_numpy(pts) / 3.1 + meta["obj_center_cam"][:, None]
    assert np.array_equal(cam, ocam.numpy())
    assert np.abs(puv - O.project(ocam, meta["cam_intr"]).numpy()).max() < 1e-4


def rnd(seed, *shape, lo=-1.0, hi=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return (g.random(size=shape, dtype=np.float32) * (hi - lo) + lo).astype(np.float32)


@pytest.mark.parametrize("P", [1, 37, 600])
def test_select_points_kernel_on_the_emulator(P):
    """Near-surface selection (upstream main/model.py:345-354): radix select + bitonic sort on the composite key give the
    stable |sdf| order BIT FOR BIT (ties -> lower row), lattice coordinates exactly, the clamp after the selection."""
    lib = build_emulated("topk")
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_select_points.argtypes = [vp, vp, vp, i64, i64, i32, C.c_float, i32, vp, vp, vp, vp, vp, vp, vp]
    B, n_f = 2, [max(P, 700), 3000]
    g = np.random.Generator(np.random.PCG64(41 + P))
    sdf = np.tanh(g.standard_normal(sum(n_f)).astype(np.float32) * 0.4).astype(np.float32)
    sdf[10] = sdf[20] = np.float32(1e-4)               # a tie among the selected: the lower row must win
    sdf[n_f[0] + 5] = -sdf[n_f[0] + 9]                 # |sdf| tie across signs
    offsets = np.concatenate([[0], np.cumsum(n_f)]).astype(np.int64)
    cand = np.concatenate([np.sort(g.choice(64 ** 3, n, replace=False)) for n in n_f]).astype(np.int32)
    lat = O.lattice(64)

    def run(p, by_row):
        sel, row = np.full((B, p), -7, np.int32), np.full((B, p), -7, np.int32)
        pts, osdf, pe = np.zeros((B, p, 3), np.float32), np.zeros((B, p), np.float32), np.zeros((B, p, 30), np.float32)
        flag = np.zeros(1, np.int32)
        assert lib.hoisdf_select_points(ptr(sdf), ptr(offsets), ptr(cand), B, p, 64, 0.15, int(by_row), ptr(sel), ptr(row),
                                        ptr(pts), ptr(osdf), ptr(pe), ptr(flag), None) == 0
        return sel, row, pts, osdf, pe, int(flag[0])

    sel, row, pts, osdf, pe, flag = run(P, False)
    assert flag == 0
    sel_r, row_r, *_ = run(P, True)
    assert np.array_equal(np.sort(row, axis=1), row_r) and np.array_equal(cand[row_r], sel_r)
    for b in range(B):
        s = torch.from_numpy(sdf[offsets[b]:offsets[b + 1]])
        order = torch.sort(s.abs(), stable=True)[1][:P]
        want = torch.from_numpy(cand[offsets[b]:offsets[b + 1]])[order].long()
        assert np.array_equal(sel[b], want.numpy())
        assert np.array_equal(pts[b], lat[want].numpy())
        assert np.array_equal(osdf[b], s[order].clamp(-0.15, 0.15).numpy())
        assert np.abs(pe[b] - O.nerf_embed(lat[want]).numpy()).max() < 2e-6
    if P > 1:
        assert run(n_f[0] + 1, False)[5] == 1          # too few candidates -> flag (upstream model.py:348 fails there)
    assert lib.hoisdf_select_points(ptr(sdf), ptr(offsets), ptr(cand), B, 8193, 64, 0.15, 0, ptr(sel), ptr(row), ptr(pts),
                                    ptr(osdf), ptr(pe), None, None) == -2


def test_vote_and_mano_kernels_on_the_emulator():
    """Joint voting (upstream common/nets/loss.py:31-36,54-57) and ManoHead + ManoLayer (mano_head.py:185-256,
    manolayer.py:111-276) of csrc/heads.cu against the oracle."""
    lib = build_emulated("heads")
    vp, i64 = C.c_void_p, C.c_int64

    class ManoModel(C.Structure):
        _fields_ = [(n, vp) for n in ("shapedirs", "posedirs", "v_template", "j_regressor", "weights", "hands_mean")]

    lib.hoisdf_vote_joints_fwd.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp]
    lib.hoisdf_mano_fwd.argtypes = [C.POINTER(ManoModel), vp, vp, i64, vp, vp, vp]
    lib.hoisdf_mano_aa_fwd.argtypes = [C.POINTER(ManoModel), vp, vp, i64, vp, vp, vp]
    L, B, P = 2, 2, 133
    pts, off, cls = rnd(71, B, P, 3, lo=-0.1, hi=0.1), rnd(72, L, B, P, 60, lo=-0.05, hi=0.05), rnd(73, L, B, P, 20, lo=-3, hi=3)
    joints = np.zeros((L, B, 20, 3), np.float32)
    assert lib.hoisdf_vote_joints_fwd(ptr(pts), ptr(off), ptr(cls), L, B, P, ptr(joints), None) == 0
    ref = O.vote_joints(torch.from_numpy(pts), torch.from_numpy(off).permute(0, 2, 1, 3), torch.from_numpy(cls).permute(0, 2, 1, 3))
    assert rel(joints, ref) < 2e-6

    sd = syn.hot_path_state_dict(74, "dexycb")
    bufs = {k: f32(v.reshape(-1)) for k, v in syn.mano_buffers(74).items() if v.dtype == torch.float32}
    model = ManoModel(*[ptr(bufs[k]) for k in ("th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor",
                                               "th_weights", "th_hands_mean")])
    pose6d, shape = torch.from_numpy(rnd(75, L, 16, B, 6)), torch.from_numpy(rnd(76, L, B, 10, lo=-2, hi=2))
    pose6d[0, 3, 0] = torch.tensor([1.0, 0, 0, 0, 1.0, 0])       # identity rotation -> the NaN->0 branch
    p6 = f32(pose6d.permute(0, 2, 1, 3).reshape(L * B, 16, 6))
    sh = f32(shape.reshape(L * B, 10))
    verts, jts = np.zeros((L * B, 778, 3), np.float32), np.zeros((L * B, 21, 3), np.float32)
    assert lib.hoisdf_mano_fwd(C.byref(model), ptr(p6), ptr(sh), L * B, ptr(verts), ptr(jts), None) == 0
    overts, ojoints = O.mano_head(sd, pose6d, shape)
    assert np.abs(verts - overts.reshape(L * B, 778, 3).numpy()).max() < 2e-6     # metres; hand extent ~0.2
    assert np.abs(jts - ojoints.reshape(L * B, 21, 3).numpy()).max() < 2e-6
    params = torch.cat([torch.from_numpy(rnd(77, B, 48, lo=-0.6, hi=0.6)), torch.from_numpy(rnd(78, B, 10, lo=-2, hi=2))], 1)
    ogt = O.mano_head_gt(sd, params.clone())
    pose = params[:, :48].clone()
    pose[:, 3:] -= syn.mano_buffers(74)["th_hands_mean"].reshape(-1)
    pa, be = f32(pose), f32(params[:, 48:])
    verts, jts = np.zeros((B, 778, 3), np.float32), np.zeros((B, 21, 3), np.float32)
    assert lib.hoisdf_mano_aa_fwd(C.byref(model), ptr(pa), ptr(be), B, ptr(verts), ptr(jts), None) == 0
    assert np.abs(verts - ogt["verts3d"].numpy()).max() < 2e-6 and np.abs(jts - ogt["joints3d"].numpy()).max() < 2e-6
