"""Stage-isolated parity of every hoisdf_b200 kernel against the CPU oracle (oracle/hoisdf_oracle.py), called
through the C ABI.  Bit-exact for index/mask work, fp32 tolerances (stated per test) for the arithmetic."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / max(b.abs().max(), 1e-30))


def rnd(seed, *shape, lo=-1.0, hi=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.random(size=shape, dtype=np.float32) * (hi - lo) + lo).astype(np.float32))


# ---------------------------------------------------------------- linear
@pytest.mark.parametrize("m,n,k", [(1, 1, 4), (127, 60, 256), (128, 128, 16), (333, 223, 512), (1000, 512, 292),
                                   (257, 1024, 3968), (64, 3, 256), (513, 768, 256), (300, 256, 1024)])
@pytest.mark.parametrize("act", [0, 1])
@pytest.mark.parametrize("impl", ["fma", "tf32x3", "fp16x3"])
def test_linear_matches_fp64(cuda, m, n, k, act, impl, monkeypatch):
    """The three Linear kernels against fp64.  fp32 FMA: error ~ sqrt(K) * 2^-24.  tcgen05 3xTF32 / FP16x3: the tensor
    core truncates on every accumulate, so their error grows ~linearly in K but stays fp32-grade (22-bit operands)."""
    from hoisdf_b200 import ops
    monkeypatch.setattr(ops, "USE_TENSOR_CORES", impl != "fma")
    monkeypatch.setattr(ops, "TC_MODE", "h3" if impl == "fp16x3" else "tf32")
    x, w, b = rnd(1, m, k), rnd(2, n, k, lo=-0.1, hi=0.1), rnd(3, n)
    res = rnd(4, m, ops.round_up(n, 4))[:, :n]
    pw = ops.PackedLinear.pack(w.to(cuda), b.to(cuda))
    assert pw.w_lo is not None
    y = ops.linear(x.to(cuda), pw, act)
    ref = x.double() @ w.double().T + b.double()
    if act:
        ref = ref.relu()
    tol = {"fma": 2e-6 * max(1.0, math.sqrt(k / 256.0)), "tf32x3": 1.5e-6 + 3e-9 * k, "fp16x3": 1.5e-6 + 4.5e-9 * k}[impl]
    assert rel_err(y, ref) < tol
    # with fused residual (same pitch as the output)
    out = torch.zeros(m, ops.round_up(n, 4), device=cuda)[:, :n]
    resd_p = torch.zeros(m, ops.round_up(n, 4), device=cuda)
    resd_p[:, :n] = res.to(cuda)
    y2 = ops.linear(x.to(cuda), pw, 0, out=out, residual=resd_p[:, :n])
    ref2 = x.double() @ w.double().T + b.double() + res.double()
    assert rel_err(y2, ref2) < tol


@pytest.mark.parametrize("wmax", [15.0, 40.0, 700.0, 3.0e4])
def test_linear_h3_large_weights(cuda, wmax):
    """a BatchNorm fold with a tiny running variance can push |w| beyond the fp16 range of the packed planes
    (w_hi * 2^11 < 65504).  PackedLinearH3.pack divides such a layer by a power of two and the epilogue multiplies it
    back; the result stays fp32-grade."""
    from hoisdf_b200 import ops
    m, n, k = 300, 96, 256
    x, w, b = rnd(1, m, k), rnd(2, n, k, lo=-0.1, hi=0.1), rnd(3, n)
    w[5] *= wmax / 0.1                      # one output channel with a huge folded scale
    w[17, 3] = wmax
    pw = ops.PackedLinearH3.pack(w.to(cuda), b.to(cuda))
    assert (pw.scale > 1.0) == (wmax >= 16.0)
    assert torch.isfinite(pw.planes.float()).all()
    y = ops.linear_h3(ops.split_rows(x.to(cuda)), pw, 1)
    ref = (x.double() @ w.double().T + b.double()).relu()
    assert rel_err(y, ref) < 4e-6
    # and the implicit-GEMM convolution path (1x1 convolution = the same GEMM)
    xs = ops.split_rows(rnd(4, 2 * 8 * 8, k).to(cuda))
    yc = torch.empty(2 * 8 * 8, n, device=cuda)
    ops.conv_h3(xs, 2, 8, 8, k, pw, [(0, 0)], 8, 8, out=yc)
    refc = rnd(4, 2 * 8 * 8, k).double() @ w.double().T + b.double()
    assert rel_err(yc, refc) < 4e-6


def test_linear_batched_rows_and_errors(cuda):
    from hoisdf_b200 import ops, _capi
    L, T, P, d = 3, 10, 7, 256
    x = rnd(5, L, T, d).to(cuda)
    w, b = rnd(6, 20, d, lo=-0.1, hi=0.1), rnd(7, 20)
    pw = ops.PackedLinear.pack(w.to(cuda), b.to(cuda))
    y = torch.empty(L * P, 20, device=cuda)
    ops.linear_raw(x.data_ptr() + 2 * d * 4, d, L * P, pw, y.data_ptr(), 20, 0, None, x_batch=(P, T * d))
    ref = x[:, 2:2 + P].reshape(-1, d).cpu().double() @ w.double().T + b.double()
    assert rel_err(y, ref) < 2e-6
    # error convention: misaligned K -> HOISDF_E_ALIGN (-3), raised as HoisdfError by the binding
    bad = ops.PackedLinear(pw.w, pw.b, 20, 254, 256)
    with pytest.raises(_capi.HoisdfError):
        ops.linear_raw(x.data_ptr(), d, 4, bad, y.data_ptr(), 20)


def test_fold_weight_norm(cuda):
    from hoisdf_b200 import ops
    v, g = rnd(8, 223, 512), rnd(9, 223, 1, lo=0.5, hi=1.5)
    w = ops.fold_weight_norm(g.to(cuda), v.to(cuda))
    assert rel_err(w, O.fold_weight_norm(g, v)) < 1e-6


# ---------------------------------------------------------------- lattice (bit-exact)
def test_lattice_candidates_bit_exact(cuda):
    from hoisdf_b200 import ops
    B = 6
    meta = syn.camera_meta(11, B)
    samples = O.lattice(64)
    for key_c, key_b in (("mano_root", "bbox_hand"), ("obj_center_cam", "bbox_obj")):
        c, K, bb = meta[key_c], meta["cam_intr"], meta[key_b]
        counts, offsets = ops.lattice_count(c.to(cuda), K.to(cuda), bb.to(cuda), 3.1, 64)
        total = int(offsets[-1])
        idx, uv = ops.lattice_compact(c.to(cuda), K.to(cuda), bb.to(cuda), 3.1, 64, counts, offsets, total)
        offs = offsets.cpu()
        all_idx = torch.arange(64 ** 3)
        for b in range(B):
            m, ouv = O.candidate_mask(samples, c[b], K[b], bb[b], 3.1)
            got = idx[offs[b]:offs[b + 1]].cpu().long()
            assert torch.equal(got, all_idx[m]), "candidate index mask differs for sample %d" % b
            assert torch.equal(uv[offs[b]:offs[b + 1]].cpu(), ouv[m]), "projected uv not bit-identical"


def test_lattice_known_answers(cuda):
    """the lattice is sheared -- samples[1] = (2.4414e-4*2/63-1 ...); column maxima 1.0317/1.03125/1."""
    from hoisdf_b200 import ops
    s = O.lattice(64)
    assert abs(float(s[:, 0].max()) - 1.0317) < 1e-3 and abs(float(s[:, 1].max()) - 1.03125) < 1e-4
    # device lattice == oracle lattice, bit for bit, through the posenc kernel's xyz columns
    idx = torch.arange(0, 64 ** 3, 37, dtype=torch.int32, device=cuda)
    rows = torch.zeros(idx.numel(), ops.ROW_LD, device=cuda)
    ops.posenc(rows, lattice_index=idx, bins=64)
    assert torch.equal(rows[:, 286:289].cpu(), s[idx.cpu().long()])
    pe = O.nerf_embed(s[idx.cpu().long()])
    assert (rows[:, 256:286].cpu() - pe).abs().max() < 2e-6     # sinf/cosf vs ATen CPU: <= 2 ulp near 1
    assert float(rows[:, 289:292].abs().max()) == 0.0 and float(rows[:, 515].abs().max()) == 0.0


def test_project_points(cuda):
    from hoisdf_b200 import ops
    B, P = 3, 50
    meta = syn.camera_meta(12, B)
    pts = rnd(13, B, P, 3)
    cam, uv = ops.project_points(pts.to(cuda), meta["mano_root"].to(cuda), meta["cam_intr"].to(cuda), 3.1)
    ocam = pts / 3.1 + meta["mano_root"][:, None]
    ouv = O.project(ocam, meta["cam_intr"])
    assert torch.equal(cam.cpu(), ocam)
    assert (uv.cpu().view(B, P, 2) - ouv).abs().max() < 1e-3   # pixels; bmm accumulation order is unspecified


# ---------------------------------------------------------------- gather
@pytest.mark.parametrize("arch", ["dexycb", "ho3d"])
def test_gather_concat_matches_grid_sample(cuda, arch):
    from hoisdf_b200 import ops
    B, P = 2, 300
    pyr = syn.feature_pyramid(21, B, arch)
    uv = rnd(22, B, P, 2, lo=-20.0, hi=275.0)      # includes out-of-image projections (border padding)
    uv[0, 0] = torch.tensor([0.0, 255.0]); uv[0, 1] = torch.tensor([255.0, 0.0]); uv[0, 2] = torch.tensor([127.5, 127.5])
    cfg = O.default_cfg()
    ref = O.gather_pyramid(pyr, O.grid_from_uv(uv, cfg))
    maps = [ops.to_nhwc(pyr[n].to(cuda)) for n in O.LEVELS]
    C = sum(m.shape[3] for m in maps)
    out = torch.empty(B * P, C, device=cuda)
    ops.gather(maps, uv.view(-1, 2).to(cuda).contiguous(), B, mode=ops.GATHER_CONCAT, out=out, rows_per_sample=P)
    assert (out.cpu().view(B, P, C) - ref).abs().max() < 5e-6
    # channels_last input is consumed zero-copy
    cl = pyr["stride4"].to(cuda).contiguous(memory_format=torch.channels_last)
    assert ops.to_nhwc(cl).data_ptr() == cl.data_ptr()


def test_gather_sum_is_projected_linear(cuda):
    """SUM mode over per-level projected maps == linear_sdfin layer 0 applied after the upstream gather."""
    from hoisdf_b200 import ops
    B, P, arch = 2, 257, "dexycb"
    pyr = syn.feature_pyramid(23, B, arch)
    C = syn.multiscale_dim(arch)
    w, bias = rnd(24, 512, C, lo=-0.03, hi=0.03), rnd(25, 512)
    uv = rnd(26, B, P, 2, lo=0.0, hi=255.0)
    cfg = O.default_cfg()
    feats = O.gather_pyramid(pyr, O.grid_from_uv(uv, cfg))
    ref = F.relu(feats.double() @ w.double().T + bias.double())
    maps = [ops.to_nhwc(pyr[n].to(cuda)) for n in O.LEVELS]
    pw = ops.PackedLinear.pack(w.to(cuda), bias.to(cuda))
    gm, off = [], 0
    for m in maps:
        b, h, ww, c = m.shape
        g = torch.empty(b, h, ww, 512, device=cuda)
        ops.linear(m.view(-1, c), ops.fma_only(pw.cols(off, off + c)), 0, out=g.view(-1, 512))   # as Model does
        gm.append(g); off += c
    out = torch.empty(B * P, 512, device=cuda)
    uv2 = uv.view(-1, 2).to(cuda).contiguous()
    ops.gather(gm, uv2, B, mode=ops.GATHER_SUM, out=out, rows_per_sample=P, bias=pw.b, act=1)
    assert rel_err(out.view(B, P, 512), ref) < 3e-6
    # ragged rows through row_offsets: sample 0 owns the first 100 rows, sample 1 all the others
    n0 = 100
    offsets = torch.tensor([0, n0, B * P], dtype=torch.int64, device=cuda)
    ops.gather(gm, uv2, B, mode=ops.GATHER_SUM, out=out, row_offsets=offsets, bias=pw.b, act=1)
    flat = uv.view(1, -1, 2)
    f0 = O.gather_pyramid(pyr, O.grid_from_uv(flat[:, :n0], cfg), sample=0)[0]
    f1 = O.gather_pyramid(pyr, O.grid_from_uv(flat[:, n0:], cfg), sample=1)[0]
    ref2 = F.relu(torch.cat([f0, f1]).double() @ w.double().T + bias.double())
    assert rel_err(out, ref2) < 3e-6
    # split-half output (what the FP16x3 Linear reads): same values to 2^-22
    outs = ops.SplitRows.empty(B * P, 512, cuda)
    ops.gather(gm, uv2, B, mode=ops.GATHER_SUM, out=outs, row_offsets=offsets, bias=pw.b, act=1)
    assert rel_err(outs.float(), out) < 3e-7


# ---------------------------------------------------------------- SDF decoder
@pytest.mark.parametrize("impl", ["fma", "tf32x3", "fp16x3"])
def test_sdf_decoder_matches_oracle(cuda, impl, monkeypatch):
    from hoisdf_b200 import ops
    from hoisdf_b200.nets.sdf_net import SDFDecoder
    monkeypatch.setattr(ops, "USE_TENSOR_CORES", impl != "fma")
    monkeypatch.setattr(ops, "TC_MODE", "h3" if impl == "fp16x3" else "tf32")
    sd = syn.hot_path_state_dict(31, "dexycb")
    dec = SDFDecoder(256, 33).to(cuda).eval()
    dec.load_state_dict({k[len("hand_sdf_decoder."):]: v for k, v in sd.items() if k.startswith("hand_sdf_decoder.")})
    x = rnd(32, 1000, 289)
    with torch.no_grad():
        got, _ = dec(x.to(cuda))
    ref = O.sdf_decoder(sd, "hand_sdf_decoder", x)
    assert got.shape == (1000, 1)
    # |sdf| < 1; fp32 accumulation-order differences only (FP16x3: one accumulator, 22-bit operands)
    assert (got.cpu() - ref).abs().max() < (4e-6 if impl == "fp16x3" else 2e-6)


# ---------------------------------------------------------------- split-half format / FP16x3 Linear
def test_split_half_round_trip_and_posenc(cuda):
    from hoisdf_b200 import ops
    x = torch.cat([rnd(41, 300, 289, lo=-50, hi=50), rnd(42, 300, 289, lo=-1e-3, hi=1e-3)])
    xs = ops.split_rows(x.to(cuda))
    back = xs.float().cpu()
    assert ((back - x).abs() <= x.abs() * 2.0 ** -21 + 1e-12).all()          # 22 significand bits
    # zero-padded tail columns
    xs2 = ops.SplitRows.empty(600, 296, cuda)
    xs2.buf.fill_(float("nan"))
    ops.split_rows(x.to(cuda), out=xs2.window(0, 289), kpad=292)
    assert torch.equal(xs2.buf[:, :, 289:292].cpu(), torch.zeros(600, 2, 3, dtype=torch.float16))
    # posenc written in split-half format == the fp32 kernel's columns
    idx = torch.arange(0, 262144, 997, dtype=torch.int32, device=cuda)
    rows32 = torch.zeros(idx.numel(), ops.ROW_LD, device=cuda)
    ops.posenc(rows32, lattice_index=idx, bins=64)
    rows16 = ops.SplitRows.empty(idx.numel(), ops.ROWH_LD, cuda)
    rows16.buf.fill_(float("nan"))
    ops.posenc(rows16, lattice_index=idx, bins=64)
    got = rows16.window(256, 40).float()
    assert rel_err(got[:, :33], rows32[:, 256:289]) < 3e-7 and float(got[:, 33:].abs().max()) == 0.0
    assert float(rows16.window(512, 8).float()[:, 7].abs().max()) == 0.0     # column 519


def test_linear_h3_modes(cuda):
    """FP16x3 Linear: split-half output chaining, strided row groups, ragged M / N / K, residual path."""
    from hoisdf_b200 import ops
    m, k, n1, n2 = 777, 289, 512, 223
    x, w1, b1 = rnd(51, m, k), rnd(52, n1, k, lo=-0.1, hi=0.1), rnd(53, n1)
    w2, b2 = rnd(54, n2, n1, lo=-0.1, hi=0.1), rnd(55, n2)
    p1, p2 = ops.PackedLinearH3.pack(w1.to(cuda), b1.to(cuda)), ops.PackedLinearH3.pack(w2.to(cuda), b2.to(cuda))
    h = ops.linear_h3(ops.split_rows(x.to(cuda)), p1, ops.ACT_RELU, split_out=True)
    y = ops.linear_h3(h, p2, ops.ACT_NONE)
    href = (x.double() @ w1.double().T + b1.double()).relu()
    assert rel_err(h.float(), href) < 3e-6
    assert rel_err(y, href @ w2.double().T + b2.double()) < 5e-6
    # output into a column window of a wider split-half buffer (the decoder's skip slot)
    wide = ops.SplitRows.empty(m, ops.ROWH_LD, cuda)
    wide.buf.zero_()
    ops.linear_h3(h, p2, ops.ACT_RELU, out=wide.window(ops.SKIP_OFF_H, n2))
    assert rel_err(wide.window(ops.SKIP_OFF_H, n2).float(), (href @ w2.double().T + b2.double()).relu()) < 5e-6
    assert float(wide.window(0, ops.SKIP_OFF_H).float().abs().max()) == 0.0
    # strided row groups: rows [2, 2+P) of every group of T rows
    L, T, P, d = 5, 300, 128, 256
    xb, w3, b3 = rnd(56, L * T, d), rnd(57, 60, d, lo=-0.1, hi=0.1), rnd(58, 60)
    xs = ops.split_rows(xb.to(cuda))
    p3 = ops.PackedLinearH3.pack(w3.to(cuda), b3.to(cuda))
    for P in (128, 100):      # TMA-store epilogue / direct-store epilogue (groups not a multiple of the tile)
        yb = ops.linear_h3(ops.SplitRows(xs.buf[2:], d), p3, 0, x_batch=(P, T * xs.ld), m=L * P)
        ref = xb.view(L, T, d)[:, 2:2 + P].reshape(-1, d).double() @ w3.double().T + b3.double()
        assert rel_err(yb, ref) < 3e-6
    # |w| >= 16: packed with a power-of-two scale (test_linear_h3_large_weights); non-finite weights are refused
    assert ops.PackedLinearH3.pack(torch.full((4, 8), 40.0, device=cuda), None).scale == 8.0
    with pytest.raises(ValueError):
        ops.PackedLinearH3.pack(torch.full((4, 8), float("inf"), device=cuda), None)


# ---------------------------------------------------------------- selection (bit-exact)
@pytest.mark.parametrize("P", [1, 37, 600, 4096])
def test_select_points_bit_exact(cuda, P):
    from hoisdf_b200 import ops
    B = 3
    n_f = [5000, P, 70000]
    g = np.random.Generator(np.random.PCG64(41 + P))
    sdf = torch.from_numpy(np.tanh(g.standard_normal(sum(n_f)).astype(np.float32) * 0.05))
    sdf[10] = sdf[20]                                  # a tie: lower row must win
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(n_f)]), dtype=torch.int64)
    cand = torch.cat([torch.sort(torch.from_numpy(g.choice(64 ** 3, n, replace=False)))[0] for n in n_f]).int()
    sel, pts, osdf, pe, flag, row = ops.select_points(sdf.to(cuda), offsets.to(cuda), cand.to(cuda), B, P, 64, 0.15)
    assert int(flag) == 0
    # screening mode: same set, ascending row order
    sel_r, _, _, _, _, row_r = ops.select_points(sdf.to(cuda), offsets.to(cuda), cand.to(cuda), B, P, 64, 0.15,
                                                  order_by_row=True)
    assert torch.equal(torch.sort(row, dim=1)[0], row_r) and torch.equal(cand.to(cuda)[row_r.long()], sel_r)
    lat = O.lattice(64)
    for b in range(B):
        s = sdf[offsets[b]:offsets[b + 1]]
        order = torch.sort(s.abs(), stable=True)[1][:P]
        want = cand[offsets[b]:offsets[b + 1]][order].long()
        assert torch.equal(sel[b].cpu().long(), want)
        assert torch.equal(pts[b].cpu(), lat[want])
        assert torch.equal(osdf[b, :, 0].cpu(), s[order].clamp(-0.15, 0.15))
        assert (pe[b].cpu() - O.nerf_embed(lat[want])).abs().max() < 2e-6
    # too few candidates -> flag, like the upstream shape-mismatch failure (model.py:348)
    if P + 1 <= 4096:
        _, _, _, _, flag, _ = ops.select_points(sdf.to(cuda), offsets.to(cuda), cand.to(cuda), B, P + 1, 64, 0.15)
        assert int(flag) == 1


# ---------------------------------------------------------------- attention / layernorm / transformer layers
@pytest.mark.parametrize("S", [64, 200, 801])
@pytest.mark.parametrize("impl", ["fma", "bf16x3"])
def test_attention_flash_matches_softmax(cuda, S, impl, monkeypatch):
    """Streaming-softmax attention against fp64.  fp32 FMA kernel: 2e-6.  tcgen05 kernel (bf16 hi/lo split, three
    products, fp32 accumulate): 2^-16-grade operands -> 3e-5 for unit-scale scores."""
    from hoisdf_b200 import ops
    monkeypatch.setattr(ops, "USE_TENSOR_CORES", impl == "bf16x3")
    B, H, d = 2, 4, 256
    qkv = rnd(51, B, S, 3 * d).to(cuda)
    q2 = qkv.view(B * S, 3 * d)
    q, k, v = [t.cpu().double().view(B, S, H, 64).transpose(1, 2) for t in qkv.split(d, dim=2)]
    for kv_valid in (None, max(1, S // 3)):
        out = torch.empty(B * S, d, device=cuda)
        ops.attention(q2, 3 * d, q2[:, d:], q2[:, 2 * d:], 3 * d, out, d, B, H, S, S, kv_valid=kv_valid)
        sc = q @ k.transpose(-1, -2) / 8.0
        if kv_valid is not None:
            sc[..., kv_valid:] = float("-inf")
        ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(B, S, d)
        assert rel_err(out.view(B, S, d), ref) < (2e-6 if impl == "fma" else 3e-5)


def test_attention_small_masks(cuda):
    from hoisdf_b200 import ops
    B, H, d, Lq, S, valid = 2, 4, 256, 17, 150, 90
    q, kv = rnd(52, B, Lq, d).to(cuda), rnd(53, B, S, 2 * d).to(cuda)
    mask = torch.zeros(Lq, S, dtype=torch.bool); mask[:, valid:] = True; mask[3, 5] = True
    out = torch.empty(B * Lq, d, device=cuda)
    kv2 = kv.view(B * S, 2 * d)
    ops.attention(q.view(-1, d), d, kv2, kv2[:, d:], 2 * d, out, d, B, H, Lq, S, mask=mask.to(cuda).to(torch.uint8))
    qq = q.cpu().double().view(B, Lq, H, 64).transpose(1, 2)
    kk = kv[..., :d].cpu().double().view(B, S, H, 64).transpose(1, 2)
    vv = kv[..., d:].cpu().double().view(B, S, H, 64).transpose(1, 2)
    sc = (qq @ kk.transpose(-1, -2) / 8.0).masked_fill(mask, float("-inf"))
    ref = (torch.softmax(sc, -1) @ vv).transpose(1, 2).reshape(B, Lq, d)
    assert rel_err(out.view(B, Lq, d), ref) < 2e-6


def test_add_layernorm(cuda):
    from hoisdf_b200 import ops
    x, r = rnd(54, 1001, 256), rnd(55, 1001, 256)
    g, b, g2, b2 = rnd(56, 256), rnd(57, 256), rnd(58, 256), rnd(59, 256)
    y2 = torch.empty(1001, 256, device=cuda)
    y = ops.add_layernorm(x.to(cuda), r.to(cuda), g.to(cuda), b.to(cuda), gamma2=g2.to(cuda), beta2=b2.to(cuda), out2=y2)
    ref = F.layer_norm((x + r).double(), (256,), g.double(), b.double(), 1e-5)
    ref2 = F.layer_norm(ref, (256,), g2.double(), b2.double(), 1e-5)
    assert rel_err(y, ref) < 2e-6 and rel_err(y2, ref2) < 2e-6


def _load_prefix(module, sd, prefix):
    module.load_state_dict({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, strict=True)


def test_transformer_matches_oracle(cuda):
    from hoisdf_b200.config import cfg
    from hoisdf_b200.nets.transformer import Transformer, VoteTransformer
    from hoisdf_b200.utils.misc import get_mano_memory_mask, get_mano_tgt_mask
    sd = syn.hot_path_state_dict(61, "dexycb")
    ocfg = O.default_cfg(num_samp_hand=70, num_samp_obj=30)
    old = (cfg.num_samp_hand, cfg.num_samp_obj)
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = 70, 30
    try:
        S, B = 100, 2
        src = rnd(62, S, B, 256)
        tr = Transformer(256, 4, 6, 4, 1024, 0.1, "relu", False, True).to(cuda).eval()
        _load_prefix(tr, sd, "hand_transformer.")
        with torch.no_grad():
            hs, mem, inter, attn = tr(src.to(cuda), None, sd["mano_query_embed.weight"].to(cuda), None,
                                      tgt_mask=get_mano_tgt_mask().to(cuda), memory_mask=get_mano_memory_mask().to(cuda))
        ohs, omem, ointer = O.transformer(sd, "hand_transformer", src, sd["mano_query_embed.weight"],
                                          torch.zeros_like(src), O.mano_tgt_mask(ocfg), O.mano_memory_mask(ocfg), ocfg)
        assert attn is None and hs.shape == ohs.shape and inter.shape == ointer.shape
        assert rel_err(mem, omem) < 2e-5 and rel_err(inter, ointer) < 2e-5 and rel_err(hs, ohs) < 2e-5
        # non-zero positional embedding path
        pos = rnd(63, S, B, 256) * 0.1
        vt = VoteTransformer(256, 4, 3, 1024, 0.1, "relu", False, True).to(cuda).eval()
        _load_prefix(vt, sd, "obj_transformer.")
        with torch.no_grad():
            m2, i2 = vt(src.to(cuda), None, pos.to(cuda))
        om2, oi2 = O.vote_transformer(sd, "obj_transformer", src, pos, ocfg)
        assert rel_err(m2, om2) < 2e-5 and rel_err(i2, oi2) < 2e-5
    finally:
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = old


def test_encoder_c_entry_equals_python_layers(cuda, monkeypatch):
    """hoisdf_encoder_fwd (the whole encoder stack in ONE C call, caller-owned workspace) against the per-layer Python path
    over the same kernels: last output, every inter_norm output and their split-half copies, identical to the bit."""
    from hoisdf_b200.config import cfg
    from hoisdf_b200.nets.transformer import Transformer
    sd = syn.hot_path_state_dict(61, "dexycb")
    tr = Transformer(256, 4, 6, 4, 1024, 0.1, "relu", False, True).to(cuda).eval()
    _load_prefix(tr, sd, "hand_transformer.")
    x = rnd(64, 3, 200, 256).to(cuda)
    res = {}
    for native in (True, False):
        monkeypatch.setattr(type(cfg), "native_encoder", native)
        with torch.no_grad():
            out, inter = tr.encoder.forward_bm(x, None)
        res[native] = (out.clone(), inter.clone(), tr.encoder.last_out_split.float(), tr.encoder.last_inter_split.float())
    for a, b in zip(res[True], res[False]):
        assert a.shape == b.shape and torch.equal(a, b)
    from hoisdf_b200 import _capi
    assert _capi.lib.hoisdf_encoder_workspace_bytes(3, 200, 1024, 4) > 0 and _capi.lib.hoisdf_encoder_workspace_bytes(3, 20, 1024, 4) == 0
    # the decoder stack (hoisdf_decoder_fwd) the same way: hs of all 4 layers identical to the bit, through the whole
    # Transformer.forward_bm (encoder + decoder, masked self-attention, key-range-limited cross-attention)
    from hoisdf_b200.utils.misc import get_mano_memory_mask, get_mano_tgt_mask
    old = (cfg.num_samp_hand, cfg.num_samp_obj)
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = 150, 50
    try:
        outs = {}
        for native in (True, False):
            monkeypatch.setattr(type(cfg), "native_decoder", native)
            with torch.no_grad():
                hs, mem, inter = tr.forward_bm(x, sd["mano_query_embed.weight"].to(cuda), None, get_mano_tgt_mask().to(cuda),
                                               get_mano_memory_mask())
            outs[native] = hs.clone()
        assert outs[True].shape == (4, 3, 17, 256) and torch.equal(outs[True], outs[False])
        assert float(outs[True].abs().max()) > 0
        assert _capi.lib.hoisdf_decoder_workspace_bytes(3, 17, 200, 1024, 4) > 0
    finally:
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = old


# ---------------------------------------------------------------- heads
def test_vote_and_mano(cuda):
    from hoisdf_b200 import ops
    from hoisdf_b200.nets.mano_head import ManoHead, ManoLayer
    L, B, P = 3, 2, 333
    pts, off, cls = rnd(71, B, P, 3, lo=-0.1, hi=0.1), rnd(72, L, B, P, 60, lo=-0.05, hi=0.05), rnd(73, L, B, P, 20, lo=-3, hi=3)
    got = ops.vote_joints(pts.to(cuda), off.to(cuda), cls.to(cuda))
    ref = O.vote_joints(pts, off.permute(0, 2, 1, 3), cls.permute(0, 2, 1, 3))
    assert rel_err(got, ref) < 2e-6
    sd = syn.hot_path_state_dict(74, "dexycb")
    head = ManoHead(ManoLayer.from_buffers(syn.mano_buffers(74))).to(cuda)
    pose6d, shape = rnd(75, L, 16, B, 6), rnd(76, L, B, 10, lo=-2, hi=2)
    pose6d[0, 3, 0] = torch.tensor([1.0, 0, 0, 0, 1.0, 0])       # identity rotation -> the NaN->0 branch
    res, _ = head(pose6d.to(cuda), shape.to(cuda))
    overts, ojoints = O.mano_head(sd, pose6d, shape)
    assert (res["verts3d"].cpu() - overts).abs().max() < 2e-6    # metres; hand extent ~0.2
    assert (res["joints3d"].cpu() - ojoints).abs().max() < 2e-6
    # ground-truth branch (axis-angle parameters, upstream mano_head.py:258-276) through hoisdf_mano_aa_fwd
    params = torch.cat([rnd(77, B, 48, lo=-0.6, hi=0.6), rnd(78, B, 10, lo=-2, hi=2)], 1)
    keep = params.clone()
    _, gt = head(pose6d.to(cuda), shape.to(cuda), mano_params=params.to(cuda))
    ogt = O.mano_head_gt(sd, params)
    assert torch.equal(params, keep)
    for k in ("verts3d", "joints3d", "mano_pose", "mano_shape"):
        assert gt[k].shape == ogt[k].shape, k
        assert (gt[k].cpu() - ogt[k]).abs().max() < 2e-6, k
    assert rel_err(res["mano_pose"], O.rot6d_to_mat(pose6d.permute(0, 2, 1, 3).reshape(-1, 6)).view(L, B, 16, 3, 3)) < 2e-6


# ---------------------------------------------------------------- U-Net decoder on the FP16x3 convolution kernels
@pytest.mark.parametrize("arch", ["dexycb", "ho3d"])
def test_unet_h3_matches_cudnn(cuda, arch):
    """nets/unet_h3.py (implicit-GEMM 3x3 / 1x1 convolutions, 4-parity transposed convolutions, BN folded, NHWC
    split-half activations) against the same modules on cuDNN fp32.  Tolerance: the tensor core truncates on every
    accumulate, so the error grows with K (up to 9 x 2048 here): measured 1.4e-5 (dexycb) / 3.4e-5 (ho3d) of range."""
    from hoisdf_b200.nets.module import Decoder, Decoder_big
    from hoisdf_b200.nets.unet_h3 import UNetH3
    B = 3           # odd: the 8x8 level packs two images per 128-pixel tile, the last tile is half empty
    dec = (Decoder_big if arch == "ho3d" else Decoder)()
    pre = "decoder_net.resnet_decoder."
    dec.load_state_dict({k[len(pre):]: v for k, v in syn.full_state_dict(61, arch).items() if k.startswith(pre)})
    dec = dec.to(cuda).eval()
    feat = torch.relu(rnd(62, B, 2048, 8, 8)).to(cuda)
    skips = {n: torch.relu(rnd(63 + i, B, c, s, s)).to(cuda)
             for i, (n, c, s) in enumerate((("stride16", 1024, 16), ("stride8", 512, 32), ("stride4", 256, 64),
                                            ("stride2", 64, 128)))}
    with torch.no_grad():
        ref_pyr, ref_out = dec(feat, skips)
        pyr, out = UNetH3(dec)(feat, skips)
        # channels_last inputs take the row-wise split instead of the transposing one
        cl = lambda t: t.contiguous(memory_format=torch.channels_last)  # noqa: E731
        pyr2, out2 = UNetH3(dec)(cl(feat), {k: cl(v) for k, v in skips.items()})
    assert set(pyr) == set(ref_pyr)
    for k in ref_pyr:
        assert pyr[k].shape == ref_pyr[k].shape, k
        assert rel_err(pyr[k], ref_pyr[k]) < 6e-5, (k, rel_err(pyr[k], ref_pyr[k]))
        assert rel_err(pyr2[k], pyr[k]) < 1e-6, k
    assert out.shape == ref_out.shape and rel_err(out, ref_out) < 6e-5 and rel_err(out2, out) < 1e-6


# ---------------------------------------------------------------- ResNet-50 encoder on the FP16x3 kernels
def _nhwc_rows(t):
    """(B, C, H, W) -> (B*H*W, C) fp32 contiguous."""
    b, c, h, w = t.shape
    return t.permute(0, 2, 3, 1).reshape(b * h * w, c).contiguous()


def test_stem_im2col_and_maxpool_split(cuda):
    """hoisdf_stem_im2col_split + FP16x3 Linear == conv 7x7 s2 p3; hoisdf_maxpool3x3s2_split == MaxPool2d(3, 2, 1)."""
    from hoisdf_b200 import ops
    B, H, W = 3, 64, 32
    img, w, b = rnd(81, B, 3, H, W), rnd(82, 64, 3, 7, 7, lo=-0.1, hi=0.1), rnd(83, 64)
    cols = ops.stem_im2col(img.to(cuda))
    assert cols.rows == B * (H // 2) * (W // 2) and cols.cols == ops.STEM_K_PAD
    ref_cols = F.unfold(img.double(), 7, padding=3, stride=2)                       # (B, 3*49, L), index c*49 + tap
    ref_cols = ref_cols.view(B, 3, 49, -1).permute(0, 3, 2, 1).reshape(-1, 147)     # -> tap*3 + c
    got = cols.float().cpu()
    assert torch.equal(got[:, 147:], torch.zeros(got.shape[0], 13))
    assert ((got[:, :147].double() - ref_cols).abs() <= ref_cols.abs() * 2.0 ** -21 + 3e-11).all()
    pw = ops.PackedLinearH3.pack(w.permute(0, 2, 3, 1).reshape(64, 147).contiguous().to(cuda), b.to(cuda))
    y = ops.linear_h3(cols, pw, ops.ACT_RELU, split_out=True)
    ref = F.conv2d(img.double(), w.double(), b.double(), stride=2, padding=3).relu()
    assert rel_err(y.float(), _nhwc_rows(ref)) < 3e-6
    # max-pool of the split-half map, input read through a column window of a wider buffer
    wide = ops.SplitRows.empty(y.rows, 96, cuda)
    wide.buf.fill_(float("nan"))
    ops.split_rows(_nhwc_rows(ref).float().to(cuda), out=wide.window(0, 64))
    pooled = ops.maxpool3x3s2(wide.window(0, 64), B, H // 2, W // 2, 64)
    want = _nhwc_rows(F.max_pool2d(ref.float(), 3, 2, 1))
    assert ((pooled.float().cpu() - want).abs() <= want.abs() * 2.0 ** -21 + 3e-11).all()


@pytest.mark.parametrize("B", [1, 3])
def test_conv_h3_stride2_and_split_residual(cuda, B):
    """Implicit-GEMM convolution with stride 2 (3x3 pad 1 and 1x1) and the split-half residual epilogue of both the
    convolution and the Linear entry -- the pieces the ResNet bottleneck adds to the U-Net's kernel -- against fp64."""
    from hoisdf_b200 import ops
    from hoisdf_b200.nets.unet_h3 import TAPS_1X1, TAPS_3X3
    cin, cout, H = 64, 96, 16
    x = rnd(91, B, cin, H, H)
    xs = ops.split_rows(_nhwc_rows(x).to(cuda))
    w3, b3 = rnd(92, cout, cin, 3, 3, lo=-0.1, hi=0.1), rnd(93, cout)
    p3 = ops.PackedLinearH3.pack(w3.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(cuda), b3.to(cuda))
    res = rnd(94, B, cout, H // 2, H // 2)
    rs = ops.split_rows(_nhwc_rows(res).to(cuda))
    y = ops.conv_h3(xs, B, H, H, cin, p3, TAPS_3X3, H // 2, H // 2, stride=2, act=ops.ACT_RELU,
                    out=ops.SplitRows.empty(B * (H // 2) ** 2, cout, cuda), residual_split=rs)
    ref = (F.conv2d(x.double(), w3.double(), b3.double(), stride=2, padding=1) + res.double()).relu()
    assert rel_err(y.float(), _nhwc_rows(ref)) < 4e-6
    # stride-1 3x3 without residual, fp32 output
    y1 = torch.empty(B * H * H, cout, device=cuda)
    ops.conv_h3(xs, B, H, H, cin, p3, TAPS_3X3, H, H, act=ops.ACT_NONE, out=y1)
    assert rel_err(y1, _nhwc_rows(F.conv2d(x.double(), w3.double(), b3.double(), padding=1))) < 4e-6
    # 1x1 stride 2 (projection shortcut)
    w1, b1 = rnd(95, cout, cin, 1, 1, lo=-0.1, hi=0.1), rnd(96, cout)
    p1 = ops.PackedLinearH3.pack(w1.reshape(cout, cin).contiguous().to(cuda), b1.to(cuda))
    yd = ops.conv_h3(xs, B, H, H, cin, p1, TAPS_1X1, H // 2, H // 2, stride=2,
                     out=ops.SplitRows.empty(B * (H // 2) ** 2, cout, cuda))
    assert rel_err(yd.float(), _nhwc_rows(F.conv2d(x.double(), w1.double(), b1.double(), stride=2))) < 3e-6
    # Linear with a split-half residual read through a window of a wider buffer, ragged M
    m = B * H * H - 5
    wide = ops.SplitRows.empty(m, 160, cuda)
    r2 = rnd(97, m, cout)
    ops.split_rows(r2.to(cuda), out=wide.window(32, cout))
    yl = ops.linear_h3(ops.SplitRows(xs.buf[:m], cin), p1, ops.ACT_RELU, split_out=True, residual_split=wide.window(32, cout))
    refl = (_nhwc_rows(x)[:m].double() @ w1.reshape(cout, cin).double().T + b1.double() + r2.double()).relu()
    assert rel_err(yl.float(), refl) < 3e-6
    yf = ops.linear_h3(ops.SplitRows(xs.buf[:m], cin), p1, ops.ACT_NONE, residual_split=wide.window(32, cout))
    assert rel_err(yf, _nhwc_rows(x)[:m].double() @ w1.reshape(cout, cin).double().T + b1.double() + r2.double()) < 3e-6


def test_add_layernorm_split_outputs(cuda):
    from hoisdf_b200 import ops
    x, r = rnd(54, 333, 256), rnd(55, 333, 256)
    g, b, g2, b2 = rnd(56, 256), rnd(57, 256), rnd(58, 256), rnd(59, 256)
    y2 = torch.empty(333, 256, device=cuda)
    s1, s2 = ops.SplitRows.empty(333, 256, cuda), ops.SplitRows.empty(333, 256, cuda)
    y = ops.add_layernorm(x.to(cuda), r.to(cuda), g.to(cuda), b.to(cuda), gamma2=g2.to(cuda), beta2=b2.to(cuda), out2=y2,
                          out_split=s1, out2_split=s2)
    y_plain = ops.add_layernorm(x.to(cuda), r.to(cuda), g.to(cuda), b.to(cuda))
    assert torch.equal(y, y_plain)
    for s, f in ((s1, y), (s2, y2)):
        # 22 significand bits; below fp16's normal range (|x| < 6e-5) the absolute floor is 2^-25 / 2^11 = 1.5e-11
        assert ((s.float() - f).abs() <= f.abs() * 2.0 ** -21 + 3e-11).all()


def test_resnet_h3_matches_cudnn(cuda):
    """nets/resnet_h3.py (stem im2col + Linear, max-pool, 16 bottlenecks with the shortcut added in the GEMM epilogue,
    BN folded) against the same module on cuDNN fp32, including writing the stage outputs into concat-buffer slots."""
    from hoisdf_b200 import ops
    from hoisdf_b200.nets.module import BackboneNet
    from hoisdf_b200.nets.resnet_h3 import ResNetH3
    B = 3
    net = BackboneNet(50)
    pre = "backbone_net."
    net.load_state_dict({k[len(pre):]: v for k, v in syn.full_state_dict(61, "dexycb").items() if k.startswith(pre)})
    net = net.to(cuda).eval()
    img = syn.image_batch(5, B).to(cuda)
    with torch.no_grad():
        ref_feat, ref_skips = net(img)
        feat, skips = ResNetH3(net.resnet)(img)
        wide = ops.SplitRows.empty(B * 64 * 64, 512, cuda)
        feat2, skips2 = ResNetH3(net.resnet)(img, {"stride4": wide.window(0, 256)})
    assert (feat.b, feat.h, feat.w, feat.c) == tuple(ref_feat.shape[i] for i in (0, 2, 3, 1))
    assert rel_err(feat.nchw(), ref_feat) < 1e-4, rel_err(feat.nchw(), ref_feat)
    for k in ("stride2", "stride4", "stride8", "stride16"):
        assert rel_err(skips[k].nchw(), ref_skips[k]) < 1e-4, (k, rel_err(skips[k].nchw(), ref_skips[k]))
    assert torch.equal(skips2["stride4"].nchw(), skips["stride4"].nchw()) and torch.equal(feat2.nchw(), feat.nchw())


@pytest.mark.parametrize("m,n,k,act", [(1000, 1, 64, 2), (333, 3, 256, 0), (4097, 20, 256, 1), (77, 24, 34, 0)])
def test_linear_narrow_split(cuda, m, n, k, act):
    """hoisdf_linear_narrow_split_fwd (last layer of the small heads) against fp64; x read through a column window."""
    from hoisdf_b200 import ops
    x, w, b = rnd(101, m, k), rnd(102, n, k, lo=-0.2, hi=0.2), rnd(103, n)
    wide = ops.SplitRows.empty(m, ops.round_up(k, 8) + 16, cuda)
    wide.buf.fill_(float("nan"))
    xs = ops.split_rows(x.to(cuda), out=wide.window(8, k), kpad=k if k % 4 == 0 else None) if k % 4 == 0 else None
    if xs is None:          # k not a multiple of 4: split into an exact-width buffer
        xs = ops.split_rows(x.to(cuda))
    y = ops.linear_narrow(ops.SplitRows(xs.buf, k, xs.col0), w.to(cuda), b.to(cuda), act)
    ref = x.double() @ w.double().T + b.double()
    ref = ref.relu() if act == 1 else (torch.sigmoid(ref) if act == 2 else ref)
    assert y.shape == (m, n) and rel_err(y, ref) < 2e-6


@pytest.fixture
def h3_pair_forced(cuda):
    """Force tcgen05.mma.cta_group::2 for every 2-CTA launch of the FP16x3 GEMM (default: only where it pays)."""
    from hoisdf_b200 import _capi
    _capi.lib.hoisdf_debug_h3_pair(2)
    yield
    _capi.lib.hoisdf_debug_h3_pair(1)


def test_linear_h3_cta_pair(cuda, h3_pair_forced):
    """The cta_group::2 form of the FP16x3 GEMM (a PAIR of CTAs computes an M = 256 tile, each staging half of W): every
    output mode and shape class -- ragged M / N / K, odd tile counts, N tiles of 64 / 128 / 256 rows, chunk 1 and 4,
    split-half chaining, residual, strided row groups, implicit-GEMM convolution -- against fp64; and the automatic mode
    picks it for the fat shapes (same results)."""
    from hoisdf_b200 import ops, _capi
    for (m, k, n, chunk) in ((777, 289, 512, 4), (129, 147, 64, 1), (1000, 1024, 223, 4), (4096, 3968, 1024, 4),
                             (300, 256, 60, 1), (2048, 2048, 512, 1), (128 * 5, 512, 384, 4)):
        x, w, b = rnd(201, m, k), rnd(202, n, k, lo=-0.1, hi=0.1), rnd(203, n)
        pw = ops.PackedLinearH3.pack(w.to(cuda), b.to(cuda))
        ref = (x.double() @ w.double().T + b.double()).relu()
        y = ops.linear_h3(ops.split_rows(x.to(cuda)), pw, ops.ACT_RELU, chunk_kb=chunk)
        ys = ops.linear_h3(ops.split_rows(x.to(cuda)), pw, ops.ACT_RELU, split_out=True, chunk_kb=chunk)
        assert rel_err(y, ref) < 4e-6 and rel_err(ys.float(), ref) < 4e-6, (m, k, n, chunk)
    # split-half residual (ResNet shortcut) and the direct-store epilogue (strided row groups)
    m, k, n = 1000, 256, 256
    x, w, b, r = rnd(204, m, k), rnd(205, n, k, lo=-0.1, hi=0.1), rnd(206, n), rnd(207, m, n)
    pw = ops.PackedLinearH3.pack(w.to(cuda), b.to(cuda))
    y = ops.linear_h3(ops.split_rows(x.to(cuda)), pw, ops.ACT_RELU, split_out=True, residual_split=ops.split_rows(r.to(cuda)))
    assert rel_err(y.float(), (x.double() @ w.double().T + b.double() + r.double()).relu()) < 4e-6
    L, T, P, d = 5, 300, 100, 256
    xb, w3, b3 = rnd(208, L * T, d), rnd(209, 60, d, lo=-0.1, hi=0.1), rnd(210, 60)
    xs = ops.split_rows(xb.to(cuda))
    yb = ops.linear_h3(ops.SplitRows(xs.buf[2:], d), ops.PackedLinearH3.pack(w3.to(cuda), b3.to(cuda)), 0,
                       x_batch=(P, T * xs.ld), m=L * P)
    assert rel_err(yb, xb.view(L, T, d)[:, 2:2 + P].reshape(-1, d).double() @ w3.double().T + b3.double()) < 4e-6
    # 3x3 convolution, zero padding, into fp32 (vs ATen conv2d in fp64)
    B, H, W, cin, cout = 2, 16, 16, 64, 256
    xi, wc, bc = rnd(211, B, cin, H, W), rnd(212, cout, cin, 3, 3, lo=-0.05, hi=0.05), rnd(213, cout)
    xs = ops.split_rows(xi.permute(0, 2, 3, 1).reshape(B * H * W, cin).contiguous().to(cuda))
    wk = wc.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    taps = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    yc = torch.empty(B * H * W, cout, device=cuda)
    ops.conv_h3(xs, B, H, W, cin, ops.PackedLinearH3.pack(wk.to(cuda), bc.to(cuda)), taps, H, W, out=yc)
    refc = F.conv2d(xi.double(), wc.double(), bc.double(), padding=1).permute(0, 2, 3, 1).reshape(B * H * W, cout)
    assert rel_err(yc, refc) < 4e-6
    # automatic mode: the fat shape takes the pair form, bit-identical to the forced run
    m, k, n = 4096, 3968, 1024
    x, w = rnd(214, m, k).to(cuda), rnd(215, n, k, lo=-0.1, hi=0.1).to(cuda)
    pw, xs = ops.PackedLinearH3.pack(w, None), ops.split_rows(x)
    forced = ops.linear_h3(xs, pw, 0).clone()
    _capi.lib.hoisdf_debug_h3_pair(1)
    auto = ops.linear_h3(xs, pw, 0).clone()
    _capi.lib.hoisdf_debug_h3_pair(0)
    single = ops.linear_h3(xs, pw, 0).clone()
    assert torch.equal(forced, auto) and rel_err(single, forced.double()) < 2e-6


def test_linear_h3_single_product(cuda):
    """single_pass = 1: ONE fp16 tensor-core product per K step (candidate pre-screening): 11-bit operands -> ~5e-4."""
    from hoisdf_b200 import ops
    m, k, n = 1000, 519, 512
    x, w, b = rnd(111, m, k), rnd(112, n, k, lo=-0.1, hi=0.1), rnd(113, n)
    pw = ops.PackedLinearH3.pack(w.to(cuda), b.to(cuda))
    xs = ops.split_rows(x.to(cuda))
    ref = (x.double() @ w.double().T + b.double()).relu()
    y1 = ops.linear_h3(xs, pw, ops.ACT_RELU, split_out=True, single=True, chunk_kb=ops.SCREEN_CHUNK_KB)
    y3 = ops.linear_h3(xs, pw, ops.ACT_RELU, split_out=True)
    # the single-product mode writes the hi plane only (fp16 values)
    e1, e3 = rel_err(y1.buf[:, 0, :n].float(), ref), rel_err(y3.float(), ref)
    assert 1e-5 < e1 < 2e-3 and e3 < 3e-6, (e1, e3)
    yf = ops.linear_h3(xs, pw, ops.ACT_NONE, single=True)                      # fp32 output, default chunking
    assert rel_err(yf, x.double() @ w.double().T + b.double()) < 2e-3


def test_linear_h3_split_k(cuda):
    """Split-K form of the FP16x3 GEMM (hoisdf_linear_h3_args.split_k: the weight gradients of the training step -- a
    handful of output tiles over a contraction as long as the batch has rows): partial products added into the output by TMA
    reductions, vs fp64; bias added once, the device-scalar y_scale applied, ragged M / N / K; shapes that do not qualify
    (activation, many tiles, short K) ignore the flag and stay bit-identical to the plain launch."""
    from hoisdf_b200 import ops
    for i, (m, k, n) in enumerate(((256, 51200, 256), (289, 38400, 223), (1024, 12800, 256), (256, 230400, 60),
                                   (3968, 12800, 512), (100, 2048, 20), (512, 5000, 1024))):
        x, w, b = rnd(301 + i, m, k), rnd(311 + i, n, k, lo=-0.1, hi=0.1), rnd(321 + i, n)
        pw = ops.PackedLinearH3.pack(w.to(cuda), b.to(cuda))
        xs = ops.split_rows(x.to(cuda))
        sc = torch.tensor([0.25], device=cuda)
        ref = (x.double() @ w.double().T) * 0.25 + b.double()
        plain = ops.linear_h3(xs, pw, ops.ACT_NONE, y_scale=sc)
        split = ops.linear_h3(xs, pw, ops.ACT_NONE, y_scale=sc, split_k=True)
        assert rel_err(plain, ref) < 4e-6 and rel_err(split, ref) < 4e-6, (m, k, n, rel_err(split, ref))
    # not eligible: ReLU epilogue / enough tiles / short contraction -> the flag changes nothing
    for (m, k, n, act) in ((256, 51200, 256, ops.ACT_RELU), (20000, 4096, 512, ops.ACT_NONE), (256, 1024, 256, ops.ACT_NONE)):
        x, w = rnd(331, m, k), rnd(332, n, k, lo=-0.1, hi=0.1)
        pw, xs = ops.PackedLinearH3.pack(w.to(cuda), None), ops.split_rows(x.to(cuda))
        assert torch.equal(ops.linear_h3(xs, pw, act), ops.linear_h3(xs, pw, act, split_k=True)), (m, k, n)
