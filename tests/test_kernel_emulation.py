"""The simple (non-tensor-core) CUDA kernels of csrc/metrics.cu, lattice.cu, topk.cu, heads.cu, gather.cu, layernorm.cu,
sdf.cu, linear.cu (fp32 FMA GEMM), attention.cu (SIMT attention), narrow.cu and backward.cu executed UNCHANGED on the host
by a CPU emulator (tests/emu/cuda_emu.h: every CUDA thread of a block is a fiber, __syncthreads / shuffles / ballots are
cooperative barriers, `_Float16` stands in for `__half`) and compared with the oracle or with PyTorch autograd -- so that
the kernel sources, their launch geometry and their C-ABI argument handling are checked in the GPU-less suite too.  Test
infrastructure: the emulated library is built from the same .cu files with `g++ -DHOISDF_EMULATE` into tests/emu/_build/
and is never loaded by the product."""
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


EMULATED = ("metrics", "lattice", "topk", "heads", "gather", "layernorm", "sdf", "linear", "attention", "narrow", "backward",
            "feed", "augment")
STUBS = ("stubs_linear.cpp", "stubs_attention.cpp", "stubs_h3.cpp")
_LIB = {}


def build_emulated(name=None, extra=(), more=()):
    """ONE emulated library for the whole module: every file of EMULATED (hoisdf_b200/csrc/<name>.cu, compiled unchanged
    with g++ -DHOISDF_EMULATE against tests/emu/cuda_emu.h, in parallel) + the stub files that stand in for the
    tensor-core entry points -> tests/emu/_build/libhoisdf_emu.so.  `name` only documents which file a test exercises."""
    if "lib" in _LIB:
        return _LIB["lib"]
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    from concurrent.futures import ThreadPoolExecutor
    out = os.path.join(EMU, "_build")
    os.makedirs(out, exist_ok=True)
    csrc = os.path.join(ROOT, "hoisdf_b200", "csrc")
    common = [os.path.join(EMU, "cuda_emu.h"), os.path.join(ROOT, "include", "hoisdf_b200.h"),
              os.path.join(csrc, "common.cuh"), os.path.join(csrc, "tc_common.cuh")]
    units = [(os.path.join(csrc, n + ".cu"), os.path.join(out, n + ".o")) for n in EMULATED] + \
            [(os.path.join(EMU, n), os.path.join(out, n.replace(".cpp", ".o"))) for n in STUBS]

    def compile_one(unit):
        src, obj = unit
        if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(d) for d in common + [src]):
            return False
        subprocess.run([gxx, "-std=c++20", "-O1", "-pthread", "-fPIC", "-DHOISDF_EMULATE", "-I" + EMU, "-x", "c++", "-c",
                        src, "-o", obj], check=True)
        return True

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        rebuilt = any(list(ex.map(compile_one, units)))
    lib = os.path.join(out, "libhoisdf_emu.so")
    if rebuilt or not os.path.exists(lib):
        subprocess.run([gxx, "-shared", "-pthread", "-o", lib] + [o for _, o in units], check=True)
    _LIB["lib"] = C.CDLL(lib)
    return _LIB["lib"]


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


class Pyramid(C.Structure):
    _fields_ = [("map", C.c_void_p * 5), ("c", C.c_int32 * 5), ("h", C.c_int32 * 5), ("w", C.c_int32 * 5),
                ("levels", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32)]


def make_pyramid(maps_nhwc):
    p = Pyramid()
    for i, m in enumerate(maps_nhwc):
        p.map[i], p.c[i], p.h[i], p.w[i] = m.ctypes.data, m.shape[3], m.shape[1], m.shape[2]
    p.levels, p.img_h, p.img_w = len(maps_nhwc), 256, 256
    return p


def join_split(hi, lo):
    return hi.view(np.float16).astype(np.float32) + lo.view(np.float16).astype(np.float32) / 2048.0


def test_gather_kernels_on_the_emulator():
    """Fused bilinear gather (upstream F.grid_sample x5, bilinear / border / align_corners=True, main/model.py:164-175):
    CONCAT mode against ATen's grid_sample incl. out-of-image projections and the image corners; SUM mode (+bias, ReLU)
    with ragged row offsets; split-half output."""
    import torch.nn.functional as F
    lib = build_emulated("gather")
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_gather_fwd.argtypes = [C.POINTER(Pyramid), vp, i64, vp, i64, i64, i32, vp, i32, vp, i64, vp]
    lib.hoisdf_gather_split_fwd.argtypes = [C.POINTER(Pyramid), vp, i64, vp, i64, i64, i32, vp, i32, vp, vp, i64, vp]
    lib.hoisdf_nchw_to_nhwc.argtypes = [vp, vp, i64, i64, i64, i64, vp]
    B, P = 2, 21
    pyr = syn.feature_pyramid(21, B, "dexycb")
    uv = torch.from_numpy(rnd(22, B, P, 2, lo=-20.0, hi=275.0))
    uv[0, 0] = torch.tensor([0.0, 255.0]); uv[0, 1] = torch.tensor([255.0, 0.0]); uv[0, 2] = torch.tensor([127.5, 127.5])
    cfg = O.default_cfg()
    ref = O.gather_pyramid(pyr, O.grid_from_uv(uv, cfg))
    maps = []
    for n in O.LEVELS:
        src = f32(pyr[n])
        dst = np.ascontiguousarray(src.transpose(0, 2, 3, 1))
        if src.shape[2] <= 16:                           # NCHW -> NHWC through the kernel itself (the small levels)
            got = np.zeros_like(dst)
            assert lib.hoisdf_nchw_to_nhwc(ptr(src), ptr(got), B, src.shape[1], src.shape[2], src.shape[3], None) == 0
            assert np.array_equal(got, dst)
        maps.append(dst)
    Ctot = sum(m.shape[3] for m in maps)
    pstruct = make_pyramid(maps)
    uvf = f32(uv.reshape(-1, 2))
    out = np.zeros((B * P, Ctot), np.float32)
    assert lib.hoisdf_gather_fwd(C.byref(pstruct), ptr(uvf), B * P, None, B, P, 0, None, 0, ptr(out), Ctot, None) == 0
    assert np.abs(out.reshape(B, P, Ctot) - ref.numpy()).max() < 5e-6
    # SUM mode over equal-width maps, ragged rows
    gen = torch.Generator().manual_seed(3)
    gm = [torch.randn(B, 128, h, h, generator=gen) for h in (16, 8, 4)]
    bias = torch.randn(128, generator=gen)
    n0 = 9
    flat = uv.reshape(1, -1, 2)
    parts = []
    for s, sl in ((0, slice(0, n0)), (1, slice(n0, B * P))):
        g = O.grid_from_uv(flat[:, sl], cfg).unsqueeze(1)
        parts.append(sum(F.grid_sample(m[s:s + 1], g, padding_mode="border", align_corners=True) for m in gm)[0, :, 0].T)
    ref2 = torch.relu(torch.cat(parts) + bias)
    gmn = [np.ascontiguousarray(f32(m).transpose(0, 2, 3, 1)) for m in gm]
    ps = make_pyramid(gmn)
    offsets = np.array([0, n0, B * P], np.int64)
    bz = f32(bias)
    out2 = np.zeros((B * P, 128), np.float32)
    assert lib.hoisdf_gather_fwd(C.byref(ps), ptr(uvf), B * P, ptr(offsets), B, 0, 1, ptr(bz), 1, ptr(out2), 128, None) == 0
    assert np.abs(out2 - ref2.numpy()).max() < 5e-6 * max(1.0, float(ref2.abs().max()))
    hi, lo = np.zeros((B * P, 128), np.uint16), np.zeros((B * P, 128), np.uint16)
    assert lib.hoisdf_gather_split_fwd(C.byref(ps), ptr(uvf), B * P, ptr(offsets), B, 0, 1, ptr(bz), 1, ptr(hi), ptr(lo), 128,
                                       None) == 0
    assert np.abs(join_split(hi, lo) - out2).max() < 3e-7 * float(np.abs(out2).max())
    assert lib.hoisdf_gather_fwd(C.byref(ps), ptr(uvf), B * P, None, B, 0, 1, None, 0, ptr(out2), 128, None) == -2


def test_add_layernorm_kernel_on_the_emulator():
    """Residual + LayerNorm (+ the shared inter_norm) of upstream common/nets/transformer.py:296-301,196-197."""
    import torch.nn.functional as F
    lib = build_emulated("layernorm")
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_add_layernorm_split_fwd.argtypes = [vp] * 8 + [i64, i64, vp, vp, i64, vp, vp, i64, vp]
    rows, d = 19, 256
    x, res = rnd(1, rows, d, lo=-2, hi=2), rnd(2, rows, d)
    g1, b1, g2, b2 = rnd(3, d, lo=0.5, hi=1.5), rnd(4, d), rnd(5, d, lo=0.5, hi=1.5), rnd(6, d)
    y, y2 = np.zeros((rows, d), np.float32), np.zeros((rows, d), np.float32)
    hi, lo = np.zeros((rows, d), np.uint16), np.zeros((rows, d), np.uint16)
    hi2, lo2 = np.zeros((rows, d), np.uint16), np.zeros((rows, d), np.uint16)
    assert lib.hoisdf_add_layernorm_split_fwd(ptr(x), ptr(res), ptr(g1), ptr(b1), ptr(y), ptr(g2), ptr(b2), ptr(y2), rows, d,
                                              ptr(hi), ptr(lo), d, ptr(hi2), ptr(lo2), d, None) == 0
    t = torch.from_numpy
    r1 = F.layer_norm(t(x) + t(res), (d,), t(g1), t(b1), 1e-5)
    r2 = F.layer_norm(r1, (d,), t(g2), t(b2), 1e-5)
    assert np.abs(y - r1.numpy()).max() < 3e-6 and np.abs(y2 - r2.numpy()).max() < 3e-6
    assert np.abs(join_split(hi, lo) - y).max() < 3e-7 * float(np.abs(y).max())
    assert np.abs(join_split(hi2, lo2) - y2).max() < 3e-7 * float(np.abs(y2).max())
    assert lib.hoisdf_add_layernorm_split_fwd(ptr(x), None, ptr(g1), ptr(b1), ptr(y), None, None, None, rows, 128,
                                              None, None, 0, None, None, 0, None) == -4


def test_posenc_and_token_kernels_on_the_emulator():
    """NeRF embedding + row-buffer tail (upstream common/utils/sdf_utils.py:96-141, main/model.py:218-219,332-333), the
    SDFDecoder input padding and the token assembly with the SDF activation (model.py:123-126,520-531) of csrc/sdf.cu."""
    lib = build_emulated("sdf", extra=("stubs_sdf.cpp",))
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_posenc_fwd.argtypes = [vp, vp, i64, i32, vp, i64, i64, vp]
    lib.hoisdf_posenc_split_fwd.argtypes = [vp, vp, i64, i32, vp, vp, i64, vp]
    lib.hoisdf_sdf_pad_input.argtypes = [vp, i64, vp, i64, vp]
    lib.hoisdf_tokens_fwd.argtypes = [vp, vp, vp, i64, vp, vp, i64, i64, vp, i64, i64, vp]
    rows = 29
    lat = O.lattice(64)
    idx = np.random.Generator(np.random.PCG64(5)).choice(64 ** 3, rows, replace=False).astype(np.int32)
    want = torch.cat([O.nerf_embed(lat[idx.astype(np.int64)]), lat[idx.astype(np.int64)]], 1).numpy()      # posenc | xyz
    buf = np.full((rows, 516), 7.0, np.float32)
    assert lib.hoisdf_posenc_fwd(ptr(idx), None, rows, 64, ptr(buf), 516, 256, None) == 0
    assert np.abs(buf[:, 256:286] - want[:, :30]).max() < 2e-6 and np.array_equal(buf[:, 286:289], want[:, 30:])
    assert not buf[:, 289:292].any() and not buf[:, 515].any() and (buf[:, :256] == 7.0).all() and (buf[:, 292:515] == 7.0).all()
    pts = rnd(6, rows, 3)
    assert lib.hoisdf_posenc_fwd(None, ptr(pts), rows, 64, ptr(buf), 516, 256, None) == 0
    assert np.abs(buf[:, 256:286] - O.nerf_embed(torch.from_numpy(pts)).numpy()).max() < 2e-6
    hi, lo = np.zeros((rows, 520), np.uint16), np.zeros((rows, 520), np.uint16)
    assert lib.hoisdf_posenc_split_fwd(ptr(idx), None, rows, 64, ptr(hi), ptr(lo), 520, None) == 0
    assert np.abs(join_split(hi, lo)[:, 256:289] - want).max() < 3e-7 * 1.1 + 2e-6
    assert not hi[:, 289:296].any() and not hi[:, 519].any()
    x = rnd(7, rows, 289)
    padded = np.full((rows, 516), 7.0, np.float32)
    assert lib.hoisdf_sdf_pad_input(ptr(x), rows, ptr(padded), 516, None) == 0
    assert np.array_equal(padded[:, :289], x) and not padded[:, 289:292].any() and not padded[:, 515].any()
    # tokens = cat[xyz (3), posenc (30), fea (223) * sigmoid(sdf / beta) / beta] at token offset t0 of an S-token sequence
    B, P, S, t0 = 2, 5, 9, 3
    xyz, pe, fea, sdf = rnd(8, B, P, 3), rnd(9, B, P, 30), rnd(10, B, P, 223), rnd(11, B, P, lo=-0.15, hi=0.15)
    beta = np.array([0.1], np.float32)
    tokens = np.zeros((B, S, 256), np.float32)
    assert lib.hoisdf_tokens_fwd(ptr(xyz), ptr(pe), ptr(fea), 223, ptr(sdf), ptr(beta), B, P, ptr(tokens), S, t0, None) == 0
    sig = O.sdf_activation({"hand_sigmoid_beta": torch.tensor([0.1])}, "hand_sigmoid_beta", torch.from_numpy(sdf)[..., None])
    ref = torch.cat([torch.from_numpy(xyz), torch.from_numpy(pe), torch.from_numpy(fea) * sig], 2).numpy()
    assert np.abs(tokens[:, t0:t0 + P] - ref).max() < 2e-6 * float(np.abs(ref).max())
    assert not tokens[:, :t0].any() and not tokens[:, t0 + P:].any()
    assert lib.hoisdf_tokens_fwd(ptr(xyz), ptr(pe), ptr(fea), 223, ptr(sdf), ptr(beta), B, P, ptr(tokens), S, S - 2, None) == -2


class LinearArgs(C.Structure):
    _fields_ = [("x", C.c_void_p), ("ldx", C.c_int64), ("x_rows_per_batch", C.c_int64), ("x_batch_stride", C.c_int64),
                ("w", C.c_void_p), ("ldw", C.c_int64), ("bias", C.c_void_p), ("residual", C.c_void_p),
                ("y", C.c_void_p), ("ldy", C.c_int64), ("y_rows_per_batch", C.c_int64), ("y_batch_stride", C.c_int64),
                ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64), ("act", C.c_int32), ("w_lo", C.c_void_p),
                ("tf32_passes", C.c_int32)]


def aligned(shape, dtype=np.float32):
    """16-byte aligned array (the kernels use 128-bit loads)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + 16, np.uint8)
    off = (-raw.ctypes.data) % 16
    return raw[off:off + n].view(dtype).reshape(shape)


def al(a):
    out = aligned(a.shape, a.dtype)
    out[...] = a
    return out


@pytest.mark.parametrize("m,n,k,act,res", [(33, 60, 256, 1, False), (130, 223, 512, 0, True), (1, 1, 4, 1, False)])
def test_fp32_linear_kernel_on_the_emulator(m, n, k, act, res):
    """nn.Linear (+ReLU, + residual) of upstream common/nets/layer.py:192-201 on the fp32 FMA kernel of csrc/linear.cu."""
    lib = build_emulated("linear", extra=("stubs_linear.cpp",))
    lib.hoisdf_linear_fwd.argtypes = [C.POINTER(LinearArgs), C.c_void_p]
    x, w, b = al(rnd(1, m, k)), al(rnd(2, n, k, lo=-0.1, hi=0.1)), al(rnd(3, n))
    r = al(rnd(4, m, n)) if res else None
    y = aligned((m, n))
    a = LinearArgs(ptr(x), k, 0, 0, ptr(w), k, ptr(b), ptr(r), ptr(y), n, 0, 0, m, n, k, act, None, 0)
    assert lib.hoisdf_linear_fwd(C.byref(a), None) == 0
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    if act:
        ref = np.maximum(ref, 0)
    if res:
        ref = ref + r
    assert np.abs(y - ref).max() < 2e-6 * max(1.0, float(np.abs(ref).max()))
    a.k = 6
    assert lib.hoisdf_linear_fwd(C.byref(a), None) == -3            # K must be a multiple of 4


def test_sdf_decoder_chain_on_the_emulator():
    """The whole fp32 SDFDecoder (upstream common/nets/sdf_net.py:87-122: weight-norm folding, 289 -> 512 -> 223 (+ skip
    concat) -> 512 -> 512 -> 1, tanh) through hoisdf_fold_weight_norm + hoisdf_sdf_decoder_fwd on the emulator: the
    padded row layout and the permuted skip-connection weight reproduce the oracle."""
    lib = build_emulated("sdf", more=("linear",), extra=("stubs_linear.cpp", "stubs_h3.cpp"))
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_fold_weight_norm.argtypes = [vp, vp, i64, i64, vp, i64, vp, i64, vp]
    lib.hoisdf_sdf_pad_input.argtypes = [vp, i64, vp, i64, vp]

    class SdfWeights(C.Structure):
        _fields_ = [(n, vp) for n in ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4", "w0_lo", "w1_lo",
                                       "w2_lo", "w3_lo")] + [("tf32_passes", C.c_int32)]

    lib.hoisdf_sdf_decoder_fwd.argtypes = [C.POINTER(SdfWeights), vp, i64, i64, vp, vp, vp, C.c_float, vp]
    sd = syn.hot_path_state_dict(7, "dexycb")
    pre = "hand_sdf_decoder."

    def fold(layer, rows, cols, ld, src_col=None):
        g, v = al(f32(sd[pre + "linh%d.weight_g" % layer]).reshape(-1)), al(f32(sd[pre + "linh%d.weight_v" % layer]))
        out = aligned((rows, ld))
        sc = None if src_col is None else np.ascontiguousarray(src_col, np.int32)
        assert lib.hoisdf_fold_weight_norm(ptr(g), ptr(v), rows, cols, ptr(out), ld, ptr(sc), ld, None) == 0
        return out

    # packed layouts of include/hoisdf_b200.h: w0 (512, 292), w1 (223, 512), w2 (512, 516) skip-permuted, w3 (512, 512)
    w0 = fold(0, 512, 289, 292, list(range(289)) + [-1] * 3)
    w1 = fold(1, 223, 512, 512)
    # upstream linh2 reads cat([relu(linh1) (223), input (289)]); the row buffer holds [input 289 | 0 x3 | relu(linh1) 223 | 0]
    w2 = fold(2, 512, 512, 516, [223 + c for c in range(289)] + [-1] * 3 + list(range(223)) + [-1])
    w3 = fold(3, 512, 512, 512)
    assert rel(w1, O.fold_weight_norm(sd[pre + "linh1.weight_g"], sd[pre + "linh1.weight_v"])) < 1e-6
    w4, b4 = al(f32(sd[pre + "linh4.weight"]).reshape(-1)), al(f32(sd[pre + "linh4.bias"]))
    bs = [al(f32(sd[pre + "linh%d.bias" % i])) for i in range(4)]
    rows = 21
    x = rnd(9, rows, 289)
    buf, ha, hb, out = aligned((rows, 516)), aligned((rows, 512)), aligned((rows, 512)), aligned((rows,))
    assert lib.hoisdf_sdf_pad_input(ptr(x), rows, ptr(buf), 516, None) == 0
    wts = SdfWeights(ptr(w0), ptr(bs[0]), ptr(w1), ptr(bs[1]), ptr(w2), ptr(bs[2]), ptr(w3), ptr(bs[3]), ptr(w4), ptr(b4),
                     None, None, None, None, 0)
    assert lib.hoisdf_sdf_decoder_fwd(C.byref(wts), ptr(buf), 516, rows, ptr(ha), ptr(hb), ptr(out), 0.0, None) == 0
    with torch.no_grad():
        ref = O.sdf_decoder(sd, "hand_sdf_decoder", torch.from_numpy(x))[:, 0].numpy()
    assert np.abs(out - ref).max() < 2e-6, np.abs(out - ref).max()
    clamped = aligned((rows,))
    assert lib.hoisdf_sdf_decoder_fwd(C.byref(wts), ptr(buf), 516, rows, ptr(ha), ptr(hb), ptr(clamped), 0.01, None) == 0
    assert np.abs(clamped - np.clip(ref, -0.01, 0.01)).max() < 2e-6


def test_simt_attention_kernels_on_the_emulator():
    """nn.MultiheadAttention core (upstream common/nets/transformer.py:294,378,383) on the SIMT kernels of
    csrc/attention.cu: the small masked kernel (decoder: 17 queries, bool mask True = blocked, key-validity limit) and
    the fp32 flash kernel (streaming softmax) against an fp64 softmax."""
    lib = build_emulated("attention", extra=("stubs_attention.cpp",))
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_attention_fwd.argtypes = [vp, i64, vp, vp, i64, vp, i64, i64, i64, i64, i64, i64, vp, vp, i64, vp]
    B, H, d = 1, 2, 128

    def reference(q, k, v, mask, kv_valid):
        Lq, S = q.shape[1], k.shape[1]
        qq = torch.from_numpy(q).double().view(B, Lq, H, 64).transpose(1, 2)
        kk = torch.from_numpy(k).double().view(B, S, H, 64).transpose(1, 2)
        vv = torch.from_numpy(v).double().view(B, S, H, 64).transpose(1, 2)
        sc = qq @ kk.transpose(-1, -2) / 8.0
        if mask is not None:
            sc = sc.masked_fill(torch.from_numpy(mask.astype(bool)), float("-inf"))
        sc[..., kv_valid:] = float("-inf")
        return (torch.softmax(sc, -1) @ vv).transpose(1, 2).reshape(B, Lq, d).numpy()

    Lq, S, valid = 17, 70, 50
    q, k, v = al(rnd(52, B, Lq, d)), al(rnd(53, B, S, d)), al(rnd(54, B, S, d))
    mask = np.zeros((Lq, S), np.uint8); mask[:, 40:] = 1; mask[3, 5] = 1
    out = aligned((B, Lq, d))
    assert lib.hoisdf_attention_fwd(ptr(q), d, ptr(k), ptr(v), d, ptr(out), d, B, H, Lq, S, valid, ptr(mask), None, 0, None) == 0
    assert np.abs(out - reference(q, k, v, mask, valid)).max() < 2e-6
    Lq, S, valid = 130, 150, 97                      # flash kernel: two query tiles, three key tiles, ragged both ways
    q, k, v = al(rnd(55, B, Lq, d)), al(rnd(56, B, S, d)), al(rnd(57, B, S, d))
    out = aligned((B, Lq, d))
    assert lib.hoisdf_attention_fwd(ptr(q), d, ptr(k), ptr(v), d, ptr(out), d, B, H, Lq, S, valid, None, None, 0, None) == 0
    assert np.abs(out - reference(q, k, v, None, valid)).max() < 2e-6
    assert lib.hoisdf_attention_fwd(ptr(q), d, ptr(k), ptr(v), d, ptr(out), 6, B, H, Lq, S, valid, None, None, 0, None) == -2


@pytest.mark.parametrize("n,act", [(3, 0), (12, 2), (20, 1)])
def test_narrow_linear_kernel_on_the_emulator(n, act):
    """Last layer of the small heads (upstream main/model.py:81-90) on csrc/narrow.cu: split-half input, fp32 weights."""
    lib = build_emulated("narrow")
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_linear_narrow_split_fwd.argtypes = [vp, vp, i64, i64, vp, i64, vp, i64, i64, C.c_int32, vp, i64, vp]
    m, k = 37, 256
    x = rnd(1, m, k, lo=-2, hi=2)
    hi = x.astype(np.float16)
    lo = ((x - hi.astype(np.float32)) * 2048.0).astype(np.float16)
    xh, xl = al(hi.view(np.uint16)), al(lo.view(np.uint16))
    w, b = al(rnd(2, n, k, lo=-0.1, hi=0.1)), al(rnd(3, n))
    y = aligned((m, n))
    assert lib.hoisdf_linear_narrow_split_fwd(ptr(xh), ptr(xl), k, m, ptr(w), k, ptr(b), n, k, act, ptr(y), n, None) == 0
    ref = join_split(xh, xl).astype(np.float64) @ w.astype(np.float64).T + b
    ref = np.maximum(ref, 0) if act == 1 else (1 / (1 + np.exp(-ref)) if act == 2 else ref)
    assert np.abs(y - ref).max() < 2e-6 * max(1.0, float(np.abs(ref).max()))
    assert np.abs(join_split(xh, xl) - x).max() < 3e-7 * 2


# ---------------------------------------------------------------------------------------------------------------------
# backward kernels of the SDF branch (csrc/backward.cu) against PyTorch autograd
# ---------------------------------------------------------------------------------------------------------------------
def backward_lib():
    lib = build_emulated("backward")
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_gemm_f32.argtypes = [vp, i64, i32, vp, i64, i32, vp, i64, i64, i64, i64, i32, vp]
    lib.hoisdf_act_bias_bwd.argtypes = [vp, i64, vp, i64, i64, i64, i32, vp, i32, vp]
    lib.hoisdf_weight_norm_bwd.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, i32, vp]
    lib.hoisdf_gather_bwd.argtypes = [C.POINTER(Pyramid), vp, i64, vp, i64, i64, vp, i64, vp]
    lib.hoisdf_sdf_loss_bwd.argtypes = [vp, vp, i64, C.c_float, C.c_float, vp, vp]
    return lib


def gemm(lib, a, ta, b, tb, accumulate_into=None):
    """C = op(A) . op(B) through hoisdf_gemm_f32 (a, b: 2-D float32 arrays as STORED)."""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    m, k = (a.shape[1], a.shape[0]) if ta else a.shape
    n = b.shape[0] if tb else b.shape[1]
    c = np.zeros((m, n), np.float32) if accumulate_into is None else accumulate_into
    assert lib.hoisdf_gemm_f32(ptr(a), a.shape[1], int(ta), ptr(b), b.shape[1], int(tb), ptr(c), n, m, n, k,
                               int(accumulate_into is not None), None) == 0
    return c


@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_f32_kernel_on_the_emulator(ta, tb):
    lib = backward_lib()
    m, n, k = 70, 45, 37                                         # ragged in every dimension
    a = rnd(1, *((k, m) if ta else (m, k)))
    b = rnd(2, *((n, k) if tb else (k, n)))
    ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    c = gemm(lib, a, ta, b, tb)
    assert np.abs(c - ref).max() < 2e-6 * float(np.abs(ref).max())
    c2 = gemm(lib, a, ta, b, tb, accumulate_into=c.copy())
    assert np.abs(c2 - 2 * ref).max() < 4e-6 * float(np.abs(ref).max())
    assert lib.hoisdf_gemm_f32(ptr(a), 1, int(ta), ptr(b), b.shape[1], int(tb), ptr(c), n, m, n, k, 0, None) == -2


def test_elementwise_backward_kernels_on_the_emulator():
    """ReLU mask + bias gradient, weight-norm gradient, clamp / L1 / tanh gradient against autograd."""
    lib = backward_lib()
    t = torch.from_numpy
    m, n = 53, 70
    y = np.maximum(rnd(1, m, n), 0).astype(np.float32)           # a ReLU output: zeros where the unit was off
    dy = rnd(2, m, n)
    want_dz = dy * (y > 0)
    dz, db = dy.copy(), np.zeros(n, np.float32)
    assert lib.hoisdf_act_bias_bwd(ptr(dz), n, ptr(y), n, m, n, 1, ptr(db), 0, None) == 0
    assert np.array_equal(dz, want_dz) and np.abs(db - want_dz.sum(0)).max() < 1e-5
    dz2, db2 = dy.copy(), db.copy()
    assert lib.hoisdf_act_bias_bwd(ptr(dz2), n, None, 0, m, n, 0, ptr(db2), 1, None) == 0        # no activation, accumulate
    assert np.array_equal(dz2, dy) and np.abs(db2 - (want_dz.sum(0) + dy.sum(0))).max() < 1e-5
    # weight norm (nn.utils.weight_norm, dim 0): W = g * v / |v|
    rows, cols = 23, 289
    g, v, dw = t(rnd(3, rows, 1, lo=0.5, hi=1.5)).requires_grad_(), t(rnd(4, rows, cols)).requires_grad_(), rnd(5, rows, cols)
    (O.fold_weight_norm(g, v) * t(dw)).sum().backward()
    dg, dv = np.zeros(rows, np.float32), np.zeros((rows, cols), np.float32)
    assert lib.hoisdf_weight_norm_bwd(ptr(f32(g.detach()).reshape(-1)), ptr(f32(v.detach())), ptr(dw), cols, rows, cols,
                                      ptr(dg), ptr(dv), 0, None) == 0
    assert np.abs(dg - g.grad.numpy().reshape(-1)).max() < 1e-5 and np.abs(dv - v.grad.numpy()).max() < 1e-6
    # SepSDFLoss on the pre-tanh value, with the clamps of sdf_forward / Model.forward
    nrow, clamp = 200, 0.05
    z = t(rnd(6, nrow, lo=-0.2, hi=0.2)).requires_grad_()
    gt = rnd(7, nrow, lo=-0.1, hi=0.1)
    loss = torch.nn.functional.l1_loss(torch.clamp(torch.tanh(z), -clamp, clamp), torch.clamp(t(gt), -clamp, clamp))
    (3.0 * loss).backward()
    dzl = np.zeros(nrow, np.float32)
    assert lib.hoisdf_sdf_loss_bwd(ptr(f32(z.detach())), ptr(gt), nrow, clamp, 3.0, ptr(dzl), None) == 0
    assert (z.grad == 0).any() and (z.grad != 0).any()           # both sides of the clamp are exercised
    assert np.abs(dzl - z.grad.numpy()).max() < 1e-7


def test_gather_backward_kernel_on_the_emulator():
    """Scatter-add of the bilinear gather (F.grid_sample backward w.r.t. the feature maps; the grid is detached upstream)."""
    lib = backward_lib()
    B, P = 2, 19
    gen = torch.Generator().manual_seed(4)
    maps = [torch.randn(B, c, h, h, generator=gen).requires_grad_() for c, h in ((12, 16), (20, 8), (8, 4))]
    uv = torch.from_numpy(rnd(22, B, P, 2, lo=-20.0, hi=275.0))
    uv[0, 0] = torch.tensor([0.0, 255.0]); uv[0, 1] = torch.tensor([255.0, 0.0]); uv[1, 2] = torch.tensor([127.5, 127.5])
    cfg = O.default_cfg()
    g = O.grid_from_uv(uv, cfg).unsqueeze(1)
    feats = torch.cat([torch.nn.functional.grid_sample(m, g, padding_mode="border", align_corners=True) for m in maps], 1)
    feats = feats.squeeze(2).permute(0, 2, 1)                    # (B, P, C) as O.gather_pyramid
    dout = rnd(23, B, P, feats.shape[2])
    (feats * torch.from_numpy(dout)).sum().backward()
    grads = [np.zeros((B, m.shape[2], m.shape[3], m.shape[1]), np.float32) for m in maps]
    ps = make_pyramid(grads)
    uvf = f32(uv.reshape(-1, 2))
    do = np.ascontiguousarray(dout.reshape(B * P, -1))
    assert lib.hoisdf_gather_bwd(C.byref(ps), ptr(uvf), B * P, None, B, P, ptr(do), do.shape[1], None) == 0
    for got, m in zip(grads, maps):
        want = m.grad.permute(0, 2, 3, 1).numpy()
        assert np.abs(got - want).max() < 2e-6 * max(1.0, float(np.abs(want).max()))
    # ragged rows: sample 0 owns the first 7 rows only
    grads2 = [np.zeros_like(x) for x in grads]
    ps2 = make_pyramid(grads2)
    offsets = np.array([0, 7, B * P], np.int64)
    assert lib.hoisdf_gather_bwd(C.byref(ps2), ptr(uvf), B * P, ptr(offsets), B, 0, ptr(do), do.shape[1], None) == 0
    assert not np.array_equal(grads2[0], grads[0]) and abs(float(grads2[0].sum() - grads[0].sum())) < 1e-3


def test_sdf_decoder_backward_chain_on_the_emulator():
    """The whole SDFDecoder backward (upstream common/nets/sdf_net.py:87-122 under the SDF L1 loss) composed from the
    entry points of csrc/backward.cu -- GEMMs, ReLU masks / bias sums, the skip concat, weight-norm gradients, the loss
    head -- against autograd of the oracle: gradients of every parameter and of the decoder input."""
    lib = backward_lib()
    sd = {k: v.clone() for k, v in syn.hot_path_state_dict(7, "dexycb").items() if k.startswith("hand_sdf_decoder.")}
    pre = "hand_sdf_decoder."
    rows, clamp = 37, 0.15
    x_t = torch.from_numpy(rnd(9, rows, 289)).requires_grad_()
    gt = rnd(10, rows, lo=-0.2, hi=0.2)
    params = {k: v.requires_grad_() for k, v in sd.items()}
    pred = O.sdf_decoder(params, "hand_sdf_decoder", x_t)
    loss = torch.nn.functional.l1_loss(torch.clamp(pred, -clamp, clamp), torch.clamp(torch.from_numpy(gt), -clamp, clamp)[:, None])
    loss.backward()

    # ---- forward on the kernels (activations kept, as a training forward must)
    W = [O.fold_weight_norm(sd[pre + "linh%d.weight_g" % i].detach(), sd[pre + "linh%d.weight_v" % i].detach()).numpy()
         for i in range(4)] + [f32(sd[pre + "linh4.weight"].detach())]
    bias = [f32(sd[pre + "linh%d.bias" % i].detach()) for i in range(5)]
    x = f32(x_t.detach())
    h0 = np.maximum(gemm(lib, x, False, W[0], True) + bias[0], 0)
    h1 = np.maximum(gemm(lib, h0, False, W[1], True) + bias[1], 0)
    cat = np.ascontiguousarray(np.concatenate([h1, x], 1))       # torch.cat([xh, input], 1), sdf_net.py:97-98
    h2 = np.maximum(gemm(lib, cat, False, W[2], True) + bias[2], 0)
    h3 = np.maximum(gemm(lib, h2, False, W[3], True) + bias[3], 0)
    z4 = (gemm(lib, h3, False, W[4], True) + bias[4]).reshape(-1)
    assert np.abs(np.tanh(z4) - pred.detach().numpy()[:, 0]).max() < 2e-6

    # ---- backward on the kernels
    dz4 = np.zeros(rows, np.float32)
    assert lib.hoisdf_sdf_loss_bwd(ptr(np.ascontiguousarray(z4)), ptr(gt), rows, clamp, 1.0, ptr(dz4), None) == 0
    dW, dB = [None] * 5, [None] * 5

    def layer_bwd(i, dout, out, inp, relu):
        d = np.ascontiguousarray(dout, np.float32)
        dB[i] = np.zeros(d.shape[1], np.float32)
        assert lib.hoisdf_act_bias_bwd(ptr(d), d.shape[1], ptr(out) if relu else None, out.shape[1] if relu else 0,
                                       d.shape[0], d.shape[1], 1 if relu else 0, ptr(dB[i]), 0, None) == 0
        dW[i] = gemm(lib, d, True, inp, False)                   # dW = dZ^T . X
        return gemm(lib, d, False, W[i], False)                  # dX = dZ . W

    dh3 = layer_bwd(4, dz4[:, None], None, h3, False)
    dh2 = layer_bwd(3, dh3, h3, h2, True)
    dcat = layer_bwd(2, dh2, h2, cat, True)
    dh0 = layer_bwd(1, dcat[:, :223], h1, h0, True)
    dx = layer_bwd(0, dh0, h0, x, True) + dcat[:, 223:]

    def close(got, want, what):
        want = want.numpy()
        assert np.abs(got.reshape(want.shape) - want).max() < 1e-5 * max(float(np.abs(want).max()), 1e-3), what

    close(dx, x_t.grad, "input")
    close(dW[4], params[pre + "linh4.weight"].grad, "linh4.weight")
    for i in range(5):
        close(dB[i], params[pre + "linh%d.bias" % i].grad, "linh%d.bias" % i)
    for i in range(4):
        g, v = f32(sd[pre + "linh%d.weight_g" % i].detach()).reshape(-1), f32(sd[pre + "linh%d.weight_v" % i].detach())
        dg, dv = np.zeros(g.shape[0], np.float32), np.zeros_like(v)
        assert lib.hoisdf_weight_norm_bwd(ptr(g), ptr(v), ptr(dW[i]), v.shape[1], v.shape[0], v.shape[1], ptr(dg), ptr(dv), 0,
                                          None) == 0
        close(dg, params[pre + "linh%d.weight_g" % i].grad, "linh%d.weight_g" % i)
        close(dv, params[pre + "linh%d.weight_v" % i].grad, "linh%d.weight_v" % i)


def transformer_backward_lib():
    lib = backward_lib()
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_layernorm_bwd.argtypes = [vp, vp, vp, i64, i64, vp, vp, vp, vp, i32, vp]
    lib.hoisdf_softmax_rows_fwd.argtypes = [vp, i64, i64, i64, i64, vp, i64, vp, i64, vp]
    lib.hoisdf_softmax_rows_bwd.argtypes = [vp, i64, vp, i64, i64, i64, vp, i64, vp]
    lib.hoisdf_softmax_dropout_rows_fwd.argtypes = [vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, i64, C.c_float, C.c_uint64, vp]
    lib.hoisdf_softmax_dropout_rows_bwd.argtypes = [vp, i64, vp, i64, i64, i64, vp, i64, C.c_float, C.c_uint64, vp]
    return lib


def test_layernorm_and_softmax_backward_kernels_on_the_emulator():
    lib = transformer_backward_lib()
    t = torch.from_numpy
    rows, d = 45, 256
    h, gamma, beta = t(rnd(1, rows, d, lo=-2, hi=2)).requires_grad_(), t(rnd(2, d, lo=0.5, hi=1.5)).requires_grad_(), \
        t(rnd(3, d)).requires_grad_()
    dy = rnd(4, rows, d)
    (torch.nn.functional.layer_norm(h, (d,), gamma, beta, 1e-5) * t(dy)).sum().backward()
    dh, dg, dbt, stats = np.zeros((rows, d), np.float32), np.zeros(d, np.float32), np.zeros(d, np.float32), \
        np.zeros(rows * 2, np.float32)
    assert lib.hoisdf_layernorm_bwd(ptr(f32(h.detach())), ptr(f32(gamma.detach())), ptr(dy), rows, d, ptr(dh), ptr(dg), ptr(dbt),
                                    ptr(stats), 0, None) == 0
    assert np.abs(dh - h.grad.numpy()).max() < 5e-6 and np.abs(dg - gamma.grad.numpy()).max() < 2e-5
    assert np.abs(dbt - beta.grad.numpy()).max() < 2e-5
    assert lib.hoisdf_layernorm_bwd(ptr(f32(h.detach())), ptr(f32(gamma.detach())), ptr(dy), rows, 128, ptr(dh), None, None, None,
                                    0, None) == -4
    # row softmax with a key-validity limit, and its backward
    r, c, valid = 19, 70, 50
    s = t(rnd(5, r, c, lo=-3, hi=3)).requires_grad_()
    dp = rnd(6, r, c)
    p_ref = torch.softmax(s[:, :valid], -1)
    (p_ref * t(dp[:, :valid])).sum().backward()
    p = np.full((r, c), 7.0, np.float32)
    assert lib.hoisdf_softmax_rows_fwd(ptr(f32(s.detach())), c, r, c, valid, None, 0, ptr(p), c, None) == 0
    assert np.abs(p[:, :valid] - p_ref.detach().numpy()).max() < 1e-6 and not p[:, valid:].any()
    ds = dp.copy()
    assert lib.hoisdf_softmax_rows_bwd(ptr(p), c, ptr(ds), c, r, c, ptr(ds), c, None) == 0          # in place
    assert np.abs(ds[:, :valid] - s.grad.numpy()[:, :valid]).max() < 1e-6 and not ds[:, valid:].any()


def test_softmax_dropout_kernels_on_the_emulator():
    """hoisdf_softmax_dropout_rows_fwd / _bwd: nn.MultiheadAttention's dropout on the probabilities with hashed keep decisions.
    The forward's pd is p * keep / (1 - q) with a keep rate of 1 - q, reproducible from the seed and different for another
    seed; the backward equals autograd of softmax -> mask -> scale for exactly that mask."""
    lib = transformer_backward_lib()
    t = torch.from_numpy
    r, c, valid, q = 64, 200, 180, 0.1
    s = t(rnd(7, r, c, lo=-3, hi=3)).requires_grad_()
    p, pd, pd2, pd3 = (np.zeros((r, c), np.float32) for _ in range(4))
    args = (ptr(f32(s.detach())), c, r, c, valid, None, 0)
    assert lib.hoisdf_softmax_dropout_rows_fwd(*args, ptr(p), c, ptr(pd), c, q, 1234, None) == 0
    assert lib.hoisdf_softmax_dropout_rows_fwd(*args, None, 0, ptr(pd2), c, q, 1234, None) == 0
    assert lib.hoisdf_softmax_dropout_rows_fwd(*args, None, 0, ptr(pd3), c, q, 1235, None) == 0
    p_ref = torch.softmax(s[:, :valid], -1).detach().numpy()
    assert np.abs(p[:, :valid] - p_ref).max() < 1e-6 and not p[:, valid:].any() and not pd[:, valid:].any()
    keep = pd[:, :valid] != 0
    assert abs(keep.mean() - (1 - q)) < 0.02 and np.array_equal(pd, pd2) and (pd3 != pd).any()
    assert np.abs(pd[:, :valid] - p_ref * keep / (1 - q)).max() < 1e-6
    dpd = rnd(8, r, c)
    (torch.softmax(s[:, :valid], -1) * t(keep.astype(np.float32)) / (1 - q) * t(dpd[:, :valid])).sum().backward()
    ds = dpd.copy()
    assert lib.hoisdf_softmax_dropout_rows_bwd(ptr(p), c, ptr(ds), c, r, c, ptr(ds), c, q, 1234, None) == 0      # in place
    assert np.abs(ds[:, :valid] - s.grad.numpy()[:, :valid]).max() < 1e-6 and not ds[:, valid:].any()
    assert lib.hoisdf_softmax_dropout_rows_fwd(*args, None, 0, ptr(pd), c, 1.0, 1, None) == -2                    # p_drop < 1


def test_encoder_layer_backward_chain_on_the_emulator():
    """One post-norm transformer encoder layer (upstream common/nets/transformer.py:286-302: MHA with 4 heads of 64, residual,
    LayerNorm, FFN with ReLU, residual, LayerNorm) differentiated with the entry points of csrc/backward.cu -- per-head
    GEMMs + row softmax for the attention core, GEMMs + ReLU masks for the projections and the FFN, LayerNorm backward --
    against autograd of the oracle's encoder layer: gradients of the layer input and of all 12 parameter tensors."""
    lib = transformer_backward_lib()
    t = torch.from_numpy
    S, d, H, ffn = 23, 256, 4, 1024
    full = syn.hot_path_state_dict(7, "dexycb")
    pre = "hand_transformer.encoder.layers.0."
    P = {k: v.clone().requires_grad_() for k, v in full.items() if k.startswith(pre)}
    src = t(rnd(1, S, 1, d)).requires_grad_()                      # (S, B = 1, d) as upstream
    out = O.encoder_layer(P, pre[:-1], src, torch.zeros_like(src), H)
    dout = rnd(2, S, d)
    (out[:, 0] * t(dout)).sum().backward()

    W = {k[len(pre):]: f32(v.detach()) for k, v in P.items()}
    x = f32(src.detach())[:, 0]

    def lin(a, w, b):
        return gemm(lib, a, False, w, True) + b

    def lin_bwd(dz, a, w):
        """-> dX, dW, db of Z = A . W^T + b given dZ (no activation)."""
        dz = np.ascontiguousarray(dz, np.float32)
        db = np.zeros(dz.shape[1], np.float32)
        assert lib.hoisdf_act_bias_bwd(ptr(dz), dz.shape[1], None, 0, dz.shape[0], dz.shape[1], 0, ptr(db), 0, None) == 0
        return gemm(lib, dz, False, w, False), gemm(lib, dz, True, a, False), db

    def ln(hh, g, b):
        return torch.nn.functional.layer_norm(t(hh), (d,), t(g), t(b), 1e-5).numpy()

    def ln_bwd(hh, g, dy):
        dh, dg, dbt, stats = np.zeros_like(hh), np.zeros(d, np.float32), np.zeros(d, np.float32), np.zeros(2 * len(hh), np.float32)
        assert lib.hoisdf_layernorm_bwd(ptr(hh), ptr(g), ptr(np.ascontiguousarray(dy)), len(hh), d, ptr(dh), ptr(dg), ptr(dbt),
                                        ptr(stats), 0, None) == 0
        return dh, dg, dbt

    # ---- forward (activations kept); pos = 0, so q = k = v input = src
    qkv = lin(x, W["self_attn.in_proj_weight"], W["self_attn.in_proj_bias"])
    heads, probs = [], []
    for hd in range(H):
        q, k, v = (np.ascontiguousarray(qkv[:, i * d + hd * 64:i * d + hd * 64 + 64]) for i in range(3))
        sc = gemm(lib, q * np.float32(0.125), False, k, True)
        p = np.zeros_like(sc)
        assert lib.hoisdf_softmax_rows_fwd(ptr(sc), S, S, S, S, None, 0, ptr(p), S, None) == 0
        probs.append(p)
        heads.append(gemm(lib, p, False, v, False))
    attn = np.ascontiguousarray(np.concatenate(heads, 1))
    proj = lin(attn, W["self_attn.out_proj.weight"], W["self_attn.out_proj.bias"])
    h1 = np.ascontiguousarray(x + proj)
    y1 = ln(h1, W["norm1.weight"], W["norm1.bias"])
    f1 = np.maximum(lin(y1, W["linear1.weight"], W["linear1.bias"]), 0)
    f2 = lin(f1, W["linear2.weight"], W["linear2.bias"])
    h2 = np.ascontiguousarray(y1 + f2)
    y2 = ln(h2, W["norm2.weight"], W["norm2.bias"])
    assert np.abs(y2 - out.detach().numpy()[:, 0]).max() < 5e-6

    # ---- backward
    G = {}
    dh2, G["norm2.weight"], G["norm2.bias"] = ln_bwd(h2, W["norm2.weight"], dout)
    df1, G["linear2.weight"], G["linear2.bias"] = lin_bwd(dh2, f1, W["linear2.weight"])
    dz1 = np.ascontiguousarray(df1)
    G["linear1.bias"] = np.zeros(ffn, np.float32)
    assert lib.hoisdf_act_bias_bwd(ptr(dz1), ffn, ptr(f1), ffn, S, ffn, 1, ptr(G["linear1.bias"]), 0, None) == 0
    G["linear1.weight"] = gemm(lib, dz1, True, y1, False)
    dy1 = gemm(lib, dz1, False, W["linear1.weight"], False) + dh2
    dh1, G["norm1.weight"], G["norm1.bias"] = ln_bwd(h1, W["norm1.weight"], dy1)
    dattn, G["self_attn.out_proj.weight"], G["self_attn.out_proj.bias"] = lin_bwd(dh1, attn, W["self_attn.out_proj.weight"])
    dqkv = np.zeros_like(qkv)
    for hd in range(H):
        q, k, v = (np.ascontiguousarray(qkv[:, i * d + hd * 64:i * d + hd * 64 + 64]) for i in range(3))
        do = np.ascontiguousarray(dattn[:, hd * 64:hd * 64 + 64])
        dv = gemm(lib, probs[hd], True, do, False)               # dV = P^T dO
        dp = gemm(lib, do, False, v, True)                       # dP = dO V^T
        assert lib.hoisdf_softmax_rows_bwd(ptr(probs[hd]), S, ptr(dp), S, S, S, ptr(dp), S, None) == 0
        dq = gemm(lib, dp, False, k, False) * np.float32(0.125)  # dQ = dS K / 8
        dk = gemm(lib, dp, True, q, False) * np.float32(0.125)   # dK = dS^T Q / 8
        for i, blk in enumerate((dq, dk, dv)):
            dqkv[:, i * d + hd * 64:i * d + hd * 64 + 64] = blk
    dx_attn, G["self_attn.in_proj_weight"], G["self_attn.in_proj_bias"] = lin_bwd(dqkv, x, W["self_attn.in_proj_weight"])
    dx = dx_attn + dh1

    def close(got, want, what):
        want = want.numpy().reshape(got.shape)
        assert np.abs(got - want).max() < 2e-5 * max(float(np.abs(want).max()), 1e-3), (what, np.abs(got - want).max())

    close(dx, src.grad[:, 0], "src")
    assert set(G) == set(W)
    for k in G:
        close(G[k], P[pre + k].grad, k)


def test_adamw_kernel_on_the_emulator():
    """torch.optim.AdamW (upstream common/base.py:68: lr 1e-4, PyTorch defaults otherwise) over three updates."""
    lib = backward_lib()
    vp, f = C.c_void_p, C.c_float
    lib.hoisdf_adamw_step.argtypes = [vp, vp, vp, vp, C.c_int64, f, f, f, f, f, C.c_int64, vp]
    n = 1000
    w = torch.from_numpy(rnd(1, n)).requires_grad_()
    opt = torch.optim.AdamW([w], lr=1e-4)
    p, m, v = f32(w.detach()).copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for step in range(1, 4):
        g = rnd(10 + step, n, lo=-3, hi=3)
        w.grad = torch.from_numpy(g.copy())
        opt.step()
        assert lib.hoisdf_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), n, 1e-4, 0.9, 0.999, 1e-8, 0.01, step, None) == 0
        assert np.abs(p - w.detach().numpy()).max() < 2e-7
    st = opt.state[w]
    assert np.abs(m - st["exp_avg"].numpy()).max() < 1e-6 and np.abs(v - st["exp_avg_sq"].numpy()).max() < 1e-6
    assert lib.hoisdf_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), n, 1e-4, 0.9, 0.999, 1e-8, 0.01, 0, None) == -2


def test_decoder_layer_backward_chain_on_the_emulator():
    """One post-norm transformer decoder layer (upstream common/nets/transformer.py:366-395): masked self-attention over
    the 17 MANO queries (block-diagonal tgt_mask, common/utils/misc.py:11-36), cross-attention to the encoder memory under
    the memory mask (object tokens blocked, misc.py:39-47), FFN, three LayerNorms -- differentiated with csrc/backward.cu
    against autograd of the oracle: gradients of tgt, memory, query_pos and of all 18 parameter tensors."""
    lib = transformer_backward_lib()
    t = torch.from_numpy
    Lq, S, d, H, ffn, Ph = 17, 23, 256, 4, 1024, 15
    full = syn.hot_path_state_dict(7, "dexycb")
    pre = "hand_transformer.decoder.layers.1."
    P = {k: v.clone().requires_grad_() for k, v in full.items() if k.startswith(pre)}
    cfg = O.default_cfg(num_samp_hand=Ph, num_samp_obj=S - Ph)
    tgt_mask, mem_mask = O.mano_tgt_mask(cfg), O.mano_memory_mask(cfg)
    tgt, memory, qpos = (t(rnd(i, n, 1, d)).requires_grad_() for i, n in ((1, Lq), (2, S), (3, Lq)))
    out = O.decoder_layer(P, pre[:-1], tgt, memory, torch.zeros_like(memory), qpos, tgt_mask, mem_mask, H)
    dout = rnd(4, Lq, d)
    (out[:, 0] * t(dout)).sum().backward()
    W = {k[len(pre):]: f32(v.detach()) for k, v in P.items()}
    x, mem, qp = f32(tgt.detach())[:, 0], f32(memory.detach())[:, 0], f32(qpos.detach())[:, 0]
    masks = {"self_attn": np.ascontiguousarray(tgt_mask.numpy().astype(np.uint8)),
             "multihead_attn": np.ascontiguousarray(mem_mask.numpy().astype(np.uint8))}

    def lin(a, w, b):
        return gemm(lib, a, False, w, True) + b

    def lin_bwd(dz, a, w):
        dz = np.ascontiguousarray(dz, np.float32)
        db = np.zeros(dz.shape[1], np.float32)
        assert lib.hoisdf_act_bias_bwd(ptr(dz), dz.shape[1], None, 0, dz.shape[0], dz.shape[1], 0, ptr(db), 0, None) == 0
        return gemm(lib, dz, False, w, False), gemm(lib, dz, True, a, False), db

    def ln(hh, name):
        return torch.nn.functional.layer_norm(t(hh), (d,), t(W[name + ".weight"]), t(W[name + ".bias"]), 1e-5).numpy()

    def ln_bwd(hh, name, dy, G):
        dh, dg, dbt, st = np.zeros_like(hh), np.zeros(d, np.float32), np.zeros(d, np.float32), np.zeros(2 * len(hh), np.float32)
        assert lib.hoisdf_layernorm_bwd(ptr(hh), ptr(W[name + ".weight"]), ptr(np.ascontiguousarray(dy)), len(hh), d, ptr(dh),
                                        ptr(dg), ptr(dbt), ptr(st), 0, None) == 0
        G[name + ".weight"], G[name + ".bias"] = dg, dbt
        return dh

    def mha_fwd(name, qin, kin, vin):
        """-> (output, saved activations) of nn.MultiheadAttention with the bool mask of this attention."""
        Wi, bi = W[name + ".in_proj_weight"], W[name + ".in_proj_bias"]
        q, k, v = lin(qin, Wi[:d], bi[:d]), lin(kin, Wi[d:2 * d], bi[d:2 * d]), lin(vin, Wi[2 * d:], bi[2 * d:])
        m = masks[name]
        probs, heads = [], []
        for hd in range(H):
            sl = slice(hd * 64, hd * 64 + 64)
            sc = gemm(lib, np.ascontiguousarray(q[:, sl]) * np.float32(0.125), False, np.ascontiguousarray(k[:, sl]), True)
            p = np.zeros_like(sc)
            assert lib.hoisdf_softmax_rows_fwd(ptr(sc), sc.shape[1], sc.shape[0], sc.shape[1], sc.shape[1], ptr(m), m.shape[0],
                                               ptr(p), sc.shape[1], None) == 0
            probs.append(p)
            heads.append(gemm(lib, p, False, np.ascontiguousarray(v[:, sl]), False))
        cat = np.ascontiguousarray(np.concatenate(heads, 1))
        return lin(cat, W[name + ".out_proj.weight"], W[name + ".out_proj.bias"]), (qin, kin, vin, q, k, v, probs, cat)

    def mha_bwd(name, dy, saved, G):
        """-> gradients of the three inputs (query side, key side, value side)."""
        qin, kin, vin, q, k, v, probs, cat = saved
        dcat, G[name + ".out_proj.weight"], G[name + ".out_proj.bias"] = lin_bwd(dy, cat, W[name + ".out_proj.weight"])
        dq, dk, dv = np.zeros_like(q), np.zeros_like(k), np.zeros_like(v)
        for hd in range(H):
            sl = slice(hd * 64, hd * 64 + 64)
            do = np.ascontiguousarray(dcat[:, sl])
            dv[:, sl] = gemm(lib, probs[hd], True, do, False)
            dp = gemm(lib, do, False, np.ascontiguousarray(v[:, sl]), True)
            n = dp.shape[1]
            assert lib.hoisdf_softmax_rows_bwd(ptr(probs[hd]), n, ptr(dp), n, dp.shape[0], n, ptr(dp), n, None) == 0
            dq[:, sl] = gemm(lib, dp, False, np.ascontiguousarray(k[:, sl]), False) * np.float32(0.125)
            dk[:, sl] = gemm(lib, dp, True, np.ascontiguousarray(q[:, sl]), False) * np.float32(0.125)
        Wi = W[name + ".in_proj_weight"]
        dqin, gq, bq = lin_bwd(dq, qin, Wi[:d])
        dkin, gk, bk = lin_bwd(dk, kin, Wi[d:2 * d])
        dvin, gv, bv = lin_bwd(dv, vin, Wi[2 * d:])
        G[name + ".in_proj_weight"], G[name + ".in_proj_bias"] = np.concatenate([gq, gk, gv]), np.concatenate([bq, bk, bv])
        return dqin, dkin, dvin

    # ---- forward
    qk = np.ascontiguousarray(x + qp)
    a1, s1 = mha_fwd("self_attn", qk, qk, x)
    h1 = np.ascontiguousarray(x + a1)
    y1 = ln(h1, "norm1")
    q2 = np.ascontiguousarray(y1 + qp)
    a2, s2 = mha_fwd("multihead_attn", q2, mem, mem)
    h2 = np.ascontiguousarray(y1 + a2)
    y2 = ln(h2, "norm2")
    f1 = np.maximum(lin(y2, W["linear1.weight"], W["linear1.bias"]), 0)
    h3 = np.ascontiguousarray(y2 + lin(f1, W["linear2.weight"], W["linear2.bias"]))
    y3 = ln(h3, "norm3")
    assert np.abs(y3 - out.detach().numpy()[:, 0]).max() < 5e-6
    # the memory mask gives the blocked (object-side) keys exactly zero weight
    assert all(not p[:, Ph:].any() for p in s2[6])

    # ---- backward
    G = {}
    dh3 = ln_bwd(h3, "norm3", dout, G)
    df1, G["linear2.weight"], G["linear2.bias"] = lin_bwd(dh3, f1, W["linear2.weight"])
    dz1 = np.ascontiguousarray(df1)
    G["linear1.bias"] = np.zeros(ffn, np.float32)
    assert lib.hoisdf_act_bias_bwd(ptr(dz1), ffn, ptr(f1), ffn, Lq, ffn, 1, ptr(G["linear1.bias"]), 0, None) == 0
    G["linear1.weight"] = gemm(lib, dz1, True, y2, False)
    dy2 = gemm(lib, dz1, False, W["linear1.weight"], False) + dh3
    dh2 = ln_bwd(h2, "norm2", dy2, G)
    dq2, dmem_k, dmem_v = mha_bwd("multihead_attn", dh2, s2, G)
    dy1 = dh2 + dq2
    dh1 = ln_bwd(h1, "norm1", dy1, G)
    dqk_q, dqk_k, dx_v = mha_bwd("self_attn", dh1, s1, G)
    dx = dh1 + dqk_q + dqk_k + dx_v
    dqp = dq2 + dqk_q + dqk_k
    dmem = dmem_k + dmem_v

    def close(got, want, what):
        want = want.numpy().reshape(got.shape)
        assert np.abs(got - want).max() < 2e-5 * max(float(np.abs(want).max()), 1e-3), (what, np.abs(got - want).max())

    close(dx, tgt.grad[:, 0], "tgt")
    close(dmem, memory.grad[:, 0], "memory")
    close(dqp, qpos.grad[:, 0], "query_pos")
    assert set(G) == set(W)
    for k in G:
        close(G[k], P[pre + k].grad, k)


def test_vote_loss_backward_kernel_on_the_emulator():
    """JointvoteLoss (upstream common/nets/loss.py:22-61): gradients of loss_joint_3d + loss_joint_cls + loss_all_joint_3d
    w.r.t. the vote offsets and class logits against autograd of the oracle's restatement."""
    lib = backward_lib()
    vp, i64, f = C.c_void_p, C.c_int64, C.c_float
    lib.hoisdf_vote_loss_bwd.argtypes = [vp, vp, vp, vp, i64, i64, i64, f, f, f, f, vp, vp, vp, vp]
    t = torch.from_numpy
    L, B, Pn = 2, 2, 70
    pts = rnd(1, B, Pn, 3, lo=-0.1, hi=0.1)
    gt = rnd(2, B, 20, 3, lo=-100, hi=100)                        # millimetres
    off, cls = rnd(3, L, B, Pn, 60, lo=-0.05, hi=0.05), rnd(4, L, B, Pn, 20, lo=-3, hi=3)
    cfg = O.default_cfg(hand_cls_dist=0.06)
    off_t, cls_t = t(off).requires_grad_(), t(cls).requires_grad_()
    off_u, cls_u = off_t.permute(0, 2, 1, 3), cls_t.permute(0, 2, 1, 3)          # upstream layout (L, P, B, .)
    joints = O.vote_joints(t(pts), off_u, cls_u)
    l1, l2, l3 = O.joint_vote_losses(t(pts), off_u, cls_u, joints, t(gt), cfg)
    gw = (1.0, 0.5, 2.0)
    (gw[0] * l1 + gw[1] * l2 + gw[2] * l3).backward()
    d_off, d_cls, npos = np.zeros_like(off), np.zeros_like(cls), np.zeros(1, np.float32)
    assert lib.hoisdf_vote_loss_bwd(ptr(pts), ptr(off), ptr(cls), ptr(gt), L, B, Pn, cfg.hand_cls_dist, gw[0], gw[1], gw[2],
                                    ptr(d_off), ptr(d_cls), ptr(npos), None) == 0
    mask = (torch.norm(t(pts).unsqueeze(2) - t(gt).unsqueeze(1) / 1000, dim=-1) < cfg.hand_cls_dist)
    assert 0 < int(mask.sum()) < mask.numel() and float(npos[0]) == float(mask.sum())
    assert np.abs(d_off - off_t.grad.numpy()).max() < 2e-5 * float(off_t.grad.abs().max())
    assert np.abs(d_cls - cls_t.grad.numpy()).max() < 2e-5 * float(cls_t.grad.abs().max())


def test_tokens_backward_kernel_on_the_emulator():
    """Token assembly with the SDF activation (upstream main/model.py:123-126,520-531): gradients of the point features,
    the SDF values and the learnable beta against autograd of the oracle's sdf_activation."""
    lib = backward_lib()
    vp, i64 = C.c_void_p, C.c_int64
    lib.hoisdf_tokens_bwd_workspace_bytes.restype = i64
    lib.hoisdf_tokens_bwd_workspace_bytes.argtypes = [i64, i64]
    lib.hoisdf_tokens_bwd.argtypes = [vp, i64, i64, vp, i64, vp, vp, i64, i64, vp, i64, vp, vp, C.c_int32, vp, i64, vp]
    t = torch.from_numpy
    B, Pn, S, t0 = 2, 13, 20, 4
    fea, sdf = t(rnd(1, B, Pn, 223)).requires_grad_(), t(rnd(2, B, Pn, lo=-0.15, hi=0.15)).requires_grad_()
    beta = torch.tensor([0.1], requires_grad=True)
    d_tok = rnd(3, B, S, 256)
    sig = O.sdf_activation({"b": beta.detach().clone()}, "b", sdf.detach()[..., None])       # value check of the oracle path
    tok_fea = fea * (torch.sigmoid(sdf[..., None] / beta) / beta)
    assert torch.allclose(tok_fea.detach(), fea.detach() * sig, atol=1e-6)
    (tok_fea * t(d_tok[:, t0:t0 + Pn, 33:])).sum().backward()
    d_fea, d_sdf, d_beta = np.zeros((B * Pn, 223), np.float32), np.zeros(B * Pn, np.float32), np.zeros(1, np.float32)
    nbytes = lib.hoisdf_tokens_bwd_workspace_bytes(B, Pn)
    ws = np.zeros(nbytes // 4, np.float32)
    assert lib.hoisdf_tokens_bwd(ptr(d_tok), S, t0, ptr(f32(fea.detach()).reshape(B * Pn, 223)), 223,
                                 ptr(f32(sdf.detach()).reshape(-1)), ptr(f32(beta.detach())), B, Pn, ptr(d_fea), 223, ptr(d_sdf),
                                 ptr(d_beta), 0, ptr(ws), nbytes, None) == 0
    assert np.abs(d_fea - fea.grad.numpy().reshape(B * Pn, 223)).max() < 1e-5 * float(fea.grad.abs().max())
    assert np.abs(d_sdf - sdf.grad.numpy().reshape(-1)).max() < 1e-5 * float(sdf.grad.abs().max())
    assert abs(float(d_beta[0]) - float(beta.grad)) < 1e-4 * abs(float(beta.grad))
    assert lib.hoisdf_tokens_bwd(ptr(d_tok), S, S - 2, ptr(d_fea), 223, ptr(d_sdf), ptr(d_beta), B, Pn, ptr(d_fea), 223, None,
                                 ptr(d_beta), 0, ptr(ws), nbytes, None) == -2


@pytest.mark.parametrize("m,k,n,relu", [(300, 512, 1, False), (77, 256, 3, True), (1000, 100, 10, False), (65, 33, 16, True)])
def test_thin_linear_kernels_on_the_emulator(m, k, n, relu):
    """hoisdf_thin_linear_fwd / _dw: the n <= 16 heads (upstream sdf_net.py:53-64 `linh4`, model.py:82-91) as streaming passes
    over x; pitched x / dz views, the accumulate flag and the error contract."""
    lib = backward_lib()
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.hoisdf_thin_linear_fwd.argtypes = [vp, i64, vp, i64, vp, i64, i64, i64, i32, vp, i64, vp]
    lib.hoisdf_thin_linear_dw.argtypes = [vp, i64, vp, i64, i64, i64, i64, vp, i64, i32, vp]
    rng = np.random.default_rng(m + n)
    ldx, lddz = k + 5, n + 3
    xb = rng.standard_normal((m, ldx)).astype(np.float32)
    w = rng.standard_normal((n, k)).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    dzb = rng.standard_normal((m, lddz)).astype(np.float32)
    x, dz = xb[:, :k], dzb[:, :n]
    y = np.empty((m, n), np.float32)
    assert lib.hoisdf_thin_linear_fwd(ptr(xb), ldx, ptr(w), k, ptr(b), m, k, n, 1 if relu else 0, ptr(y), n, None) == 0
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    ref = np.maximum(ref, 0) if relu else ref
    assert np.abs(y - ref).max() <= 1e-5 * max(np.abs(ref).max(), 1.0)
    assert lib.hoisdf_thin_linear_fwd(ptr(xb), ldx, ptr(w), k, None, m, k, n, 0, ptr(y), n, None) == 0          # no bias
    assert np.abs(y - x.astype(np.float64) @ w.astype(np.float64).T).max() <= 1e-5 * max(np.abs(ref).max(), 1.0)
    dw = np.full((n, k), 7.0, np.float32)
    assert lib.hoisdf_thin_linear_dw(ptr(xb), ldx, ptr(dzb), lddz, m, k, n, ptr(dw), k, 0, None) == 0
    dref = dz.astype(np.float64).T @ x.astype(np.float64)
    assert np.abs(dw - dref).max() <= 1e-5 * np.abs(dref).max()
    assert lib.hoisdf_thin_linear_dw(ptr(xb), ldx, ptr(dzb), lddz, m, k, n, ptr(dw), k, 1, None) == 0           # accumulate
    assert np.abs(dw - 2 * dref).max() <= 2e-5 * np.abs(dref).max()
    dwp = np.full((n, k + 4), 7.0, np.float32)                                                                    # pitched dw
    assert lib.hoisdf_thin_linear_dw(ptr(xb), ldx, ptr(dzb), lddz, m, k, n, ptr(dwp), k + 4, 0, None) == 0
    assert np.abs(dwp[:, :k] - dref).max() <= 1e-5 * np.abs(dref).max() and (dwp[:, k:] == 7.0).all()
    assert lib.hoisdf_thin_linear_fwd(ptr(xb), ldx, ptr(w), k, None, m, k, 17, 0, ptr(y), 17, None) == -4          # n > 16
    assert lib.hoisdf_thin_linear_dw(ptr(xb), k - 1, ptr(dzb), lddz, m, k, n, ptr(dw), k, 0, None) == -2           # ldx < k
    assert lib.hoisdf_thin_linear_dw(None, ldx, ptr(dzb), lddz, m, k, n, ptr(dw), k, 0, None) == -1
