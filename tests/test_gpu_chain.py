"""The fused candidate chain kernel (csrc/sdf_chain.cu, `hoisdf_sdf_chain_fwd`): linear_sdfin.layers.1 -> NeRF embedding
-> SDFDecoder in ONE persistent tcgen05 kernel, against (a) the oracle's fp32 arithmetic (upstream main/model.py:330-346,
common/nets/sdf_net.py:87-122) within the single-product fp16 screening tolerance and (b) the unfused single-product
launches it replaces (same operands, same fp32 TMEM accumulation)."""
import numpy as np
import pytest
import torch

from hoisdf_b200 import synthetic as syn
from oracle import hoisdf_oracle as O

pytestmark = pytest.mark.gpu


def rnd(seed, *shape, lo=-1.0, hi=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.random(size=shape, dtype=np.float32) * (hi - lo) + lo).astype(np.float32))


def _decoder(cuda, sd, name="hand_sdf_decoder"):
    from hoisdf_b200.nets.sdf_net import SDFDecoder
    dec = SDFDecoder(256, 33).to(cuda).eval()
    dec.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")})
    return dec


# single-product fp16 operands (11-bit significands), fp32 accumulation: measured ~5e-5 absolute on |sdf| < 1
SCREEN_TOL = 4e-4


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 1000, 4096 + 77])
def test_decoder_mode_matches_oracle(cuda, rows):
    """SDFDecoder.forward in isolation through the fused kernel (decoder-only mode) against the oracle."""
    from hoisdf_b200 import ops
    sd = syn.hot_path_state_dict(31, "dexycb")
    dec = _decoder(cuda, sd)
    x = rnd(32 + rows, rows, 289)
    buf = ops.sdf_pad_input(x.to(cuda))
    assert isinstance(buf, ops.SplitRows)
    out = torch.full((rows,), float("nan"), device=cuda)
    ops.sdf_chain(dec.packed(), out, x=buf)
    ref = O.sdf_decoder(sd, "hand_sdf_decoder", x).view(-1)
    err = (out.cpu() - ref).abs().max()
    assert torch.isfinite(out).all()
    assert err < SCREEN_TOL, float(err)
    # the unfused single-product chain computes the same thing (same operands; accumulation order may differ)
    unf = ops.sdf_decoder(dec.packed(), buf, chunk_kb=ops.SCREEN_CHUNK_KB, single=True)
    assert (out - unf).abs().max() < 1e-4      # the unfused chain rounds relu(linh3) to fp16 before the head


@pytest.mark.parametrize("rows", [64, 128 * 3, 2500])
def test_rows_mode_matches_oracle(cuda, rows):
    """linear_sdfin.layers.1 + embedding + decoder from the relu(layers.0) rows and lattice indices."""
    from hoisdf_b200 import ops
    from hoisdf_b200.nets.layer import MLP
    sd = syn.hot_path_state_dict(33, "dexycb")
    dec = _decoder(cuda, sd, "obj_sdf_decoder")
    w1 = sd["linear_sdfin.layers.1.weight"].to(cuda)
    b1 = sd["linear_sdfin.layers.1.bias"].to(cuda)
    pw = ops.PackedLinear.pack(w1, b1)
    h = rnd(90 + rows, rows, 512, lo=0.0, hi=0.5)                # relu output of layer 0
    idx = torch.from_numpy(np.random.Generator(np.random.PCG64(rows)).integers(0, 64 ** 3, size=rows).astype(np.int32))
    hs = ops.split_rows(h.to(cuda))
    out = torch.full((rows,), float("nan"), device=cuda)
    ops.sdf_chain(dec.packed(), out, sdfin1=pw, a0=hs, lattice_index=idx.to(cuda), bins=64)
    # oracle: fea = relu(h W1^T + b1); input = [fea | posenc(lattice point) | xyz]
    pts = O.lattice(64)[idx.long()]
    fea = torch.relu(h @ w1.cpu().T + b1.cpu())
    inp = torch.cat([fea, O.nerf_embed(pts), pts], dim=1)
    ref = O.sdf_decoder(sd, "obj_sdf_decoder", inp).view(-1)
    assert torch.isfinite(out).all()
    err = (out.cpu() - ref).abs().max()
    assert err < SCREEN_TOL, float(err)
    # against the launches it replaces
    rs = ops.SplitRows.empty(rows, ops.ROWH_LD, cuda)
    ops.linear(hs, pw, ops.ACT_RELU, out=rs.window(0, 256), chunk_kb=ops.SCREEN_CHUNK_KB, single=True)
    ops.posenc(rs, lattice_index=idx.to(cuda), bins=64)
    unf = ops.sdf_decoder(dec.packed(), rs, chunk_kb=ops.SCREEN_CHUNK_KB, single=True)
    assert (out - unf).abs().max() < 1e-4      # the unfused chain rounds relu(linh3) to fp16 before the head


def test_rows_mode_with_points_and_clamp(cuda):
    from hoisdf_b200 import ops
    sd = syn.hot_path_state_dict(35, "dexycb")
    dec = _decoder(cuda, sd)
    pw = ops.PackedLinear.pack(sd["linear_sdfin.layers.1.weight"].to(cuda), sd["linear_sdfin.layers.1.bias"].to(cuda))
    rows = 300
    h = rnd(7, rows, 512, lo=0.0, hi=0.5)
    pts = rnd(8, rows, 3)
    out = torch.empty(rows, device=cuda)
    ops.sdf_chain(dec.packed(), out, sdfin1=pw, a0=ops.split_rows(h.to(cuda)), points=pts.to(cuda).contiguous(), clamp=0.01)
    fea = torch.relu(h @ sd["linear_sdfin.layers.1.weight"].T + sd["linear_sdfin.layers.1.bias"])
    ref = O.sdf_decoder(sd, "hand_sdf_decoder", torch.cat([fea, O.nerf_embed(pts), pts], 1)).view(-1).clamp(-0.01, 0.01)
    assert (out.cpu() - ref).abs().max() < SCREEN_TOL
    assert out.abs().max() <= 0.01


def test_chain_argument_errors(cuda):
    from hoisdf_b200 import _capi
    import ctypes as C
    a = _capi.SdfChainArgs()
    assert _capi.lib.hoisdf_sdf_chain_fwd(C.byref(a), None) == -1      # HOISDF_E_NULL


@pytest.mark.parametrize("rows,offsets", [(700, True), (128 * 5, False), (3000, True)])
def test_gather_mode_matches_unfused_gather(cuda, rows, offsets):
    """Gather mode: the kernel's own 8 gather warps interpolate the fp16 projected maps (upstream's 5 x F.grid_sample + cat +
    linear_sdfin.layers.0, main/model.py:316-330, in the commuted form of DESIGN.md 2.2).  Against (a) the stand-alone
    gather kernel on the SAME fp16-rounded maps followed by the rows-mode chain (only the fp32 summation order inside a
    level differs: fp16 round-off flips), (b) the gather on the fp32 maps (adds the fp16 rounding of the maps)."""
    from hoisdf_b200 import ops
    sd = syn.hot_path_state_dict(37, "dexycb")
    dec = _decoder(cuda, sd)
    pw = ops.PackedLinear.pack(sd["linear_sdfin.layers.1.weight"].to(cuda), sd["linear_sdfin.layers.1.bias"].to(cuda))
    B = 3
    sizes = [(24, 24), (12, 12), (6, 6), (3, 3), (2, 2)]
    maps = [rnd(50 + i, B, h, w, 512, lo=-0.2, hi=0.2).to(cuda) for i, (h, w) in enumerate(sizes)]
    maps16 = ops.maps_to_half(maps)
    bias0 = rnd(60, 512, lo=-0.05, hi=0.05).to(cuda)
    uv = rnd(61 + rows, rows, 2, lo=-20.0, hi=275.0).to(cuda).contiguous()          # incl. pixels outside the image (border)
    idx = torch.from_numpy(np.random.Generator(np.random.PCG64(rows)).integers(0, 64 ** 3, size=rows).astype(np.int32)).to(cuda)
    if offsets:
        cuts = sorted(np.random.Generator(np.random.PCG64(7)).integers(1, rows, size=B - 1).tolist())
        row_offsets = torch.tensor([0] + cuts + [rows], dtype=torch.int64, device=cuda)
        rps = 0
    else:
        row_offsets, rps = None, rows // B if rows % B == 0 else None
        if rps is None:
            row_offsets, rps = torch.tensor([0, rows // 3, 2 * rows // 3, rows], dtype=torch.int64, device=cuda), 0
    out = torch.full((rows,), float("nan"), device=cuda)
    ops.sdf_chain(dec.packed(), out, sdfin1=pw, gmaps16=maps16, uv=uv, row_offsets=row_offsets, batch=B,
                  rows_per_sample=rps, bias0=bias0, lattice_index=idx, bins=64)
    assert torch.isfinite(out).all()

    def unfused(ms):
        hs = ops.SplitRows.empty(rows, 512, cuda)
        ops.gather(ms, uv, B, mode=ops.GATHER_SUM, out=hs, row_offsets=row_offsets, rows_per_sample=rps, bias=bias0,
                   act=ops.ACT_RELU)
        ref = torch.empty(rows, device=cuda)
        ops.sdf_chain(dec.packed(), ref, sdfin1=pw, a0=hs, lattice_index=idx, bins=64)
        return ref

    same_maps = unfused([m.float() for m in maps16])
    assert (out - same_maps).abs().max() < 5e-5, float((out - same_maps).abs().max())
    assert (out - unfused(maps)).abs().max() < SCREEN_TOL
    # the stand-alone fp16 gather (hoisdf_gather_sum_h16_fwd: fp16 maps in, fp16 hi plane out) + rows-mode chain: same values
    hs = ops.SplitRows.empty(rows, 512, cuda)
    ops.gather_h16(maps16, uv, B, hs, row_offsets=row_offsets, rows_per_sample=rps, bias=bias0, act=ops.ACT_RELU)
    ref = torch.empty(rows, device=cuda)
    ops.sdf_chain(dec.packed(), ref, sdfin1=pw, a0=hs, lattice_index=idx, bins=64)
    assert (ref - same_maps).abs().max() < 5e-5, float((ref - same_maps).abs().max())
