"""`Model` -- the drop-in boundary: same constructor, method signatures, parameter names and output dict as
upstream main/model.py:28-665, with the hot path (sdf_infer -> sdf_forward -> get_input_transformer ->
transformers -> heads -> MANO / vote) running on the hoisdf_b200 sm_100a kernels.

What is different by design (B200-first):
  * no per-sample Python loop and no `.cpu()` round trips: the whole batch goes through a handful of launches;
    the only host read-back is the per-sample candidate COUNT (needed to size the row buffers), issued before
    the backbone so that it never stalls the GPU;
  * `linear_sdfin` layer 0 is applied to the pyramid once per image (a 1x1 projection of every level) and the
    bilinear gather then interpolates 512 projected channels instead of 3968 raw ones -- interpolation is
    linear, so  W.(sum_t w_t F_t) = sum_t w_t (W.F_t); 37x fewer FLOPs for that layer at N_f ~ 19k;
  * the image encoder (ResNet-50 + U-Net) runs on the same FP16x3 tensor-core GEMM as implicit-GEMM convolutions
    (nets/resnet_h3.py, nets/unet_h3.py), NHWC split-half activations end to end; the cuDNN modules stay as the
    reference path (cfg.tc_backbone / cfg.tc_unet = False);
  * the near-surface selection is a verified coarse-to-fine cascade (single-product FP16 -> FP16x3), `sdf_infer`;
  * the static-shape stages can be replayed from CUDA graphs (`enable_cuda_graphs`).
`mode="train"` (forward with the autograd tape of hoisdf_b200/autograd.py) lives in hoisdf_b200/train.py.
"""
from __future__ import annotations

import contextlib
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .config import cfg
from .nets.layer import MLP, _require_inference
from .nets.loss import dexycb_losses, eval_losses
from .nets.mano_head import ManoHead, ManoLayer
from .nets.module import BackboneNet, DecoderNet, DecoderNet_big
from .nets.sdf_net import SDFDecoder
from .nets.transformer import Transformer, VoteTransformer
from .nets.unet_h3 import UNetH3
from .utils.misc import get_mano_memory_mask, get_mano_tgt_mask


@contextlib.contextmanager
def _tensor_cores(enabled: bool):
    old = ops.USE_TENSOR_CORES
    ops.USE_TENSOR_CORES = enabled
    try:
        yield
    finally:
        ops.USE_TENSOR_CORES = old


class PyramidContext:
    """Channels-last views of one forward's feature pyramid + the per-level projection through
    `linear_sdfin.layers[0]` (computed lazily, once, shared by every SDF query of the forward)."""

    def __init__(self, feature_pyramid: Dict[str, torch.Tensor], model: "Model"):
        self.maps = [ops.to_nhwc(feature_pyramid[name]) for name in cfg.mutliscale_layers]
        self.split = getattr(feature_pyramid, "split", {})        # split-half copies the U-Net already produced
        self.batch = self.maps[0].shape[0]
        self.channels = sum(m.shape[3] for m in self.maps)
        self._model = model
        self._gmaps = None
        self._gmaps16 = None

    @property
    def gmaps16(self):
        """fp16 copy of the projected maps: what the fused candidate chain's gather warps interpolate."""
        if self._gmaps16 is None:
            self._gmaps16 = ops.maps_to_half(self.gmaps)
        return self._gmaps16

    @property
    def gmaps(self):
        if self._gmaps is None:
            w0 = self._model.linear_sdfin.packed()[0]
            if w0.k != self.channels:
                raise RuntimeError("pyramid has %d channels, linear_sdfin expects %d" % (self.channels, w0.k))
            # K up to 2048 and the result feeds the top-k-critical candidate SDF.  Default: the FP16x3 GEMM draining
            # its TMEM accumulator every K block (fp32-FMA-grade result, cfg.projection_chunk_kb); the projected maps
            # are shared by every stage of the selection cascade, so their rounding cannot reorder it.
            # cfg.tc_projection = False: the fp32 FMA kernel.
            tc = bool(cfg.tc_projection) and ops.use_h3() and w0.h3 is not None
            if not tc:
                w0 = ops.PackedLinear(w0.w, w0.b, w0.n, w0.k, w0.ldw, None, None)
            g, off = [], 0
            for name, m in zip(cfg.mutliscale_layers, self.maps):
                b, h, w, c = m.shape
                out = torch.empty(b, h, w, w0.n, device=m.device, dtype=torch.float32)
                wl = w0.cols(off, off + c)
                if tc and wl.h3 is not None:
                    xs = self.split.get(name)
                    if xs is None:
                        xs = ops.split_rows(m.view(b * h * w, c))
                    ops.linear_h3(xs, wl.h3, ops.ACT_NONE, out=out.view(b * h * w, w0.n),
                                  chunk_kb=int(cfg.projection_chunk_kb))
                else:
                    ops.linear(m.view(b * h * w, c), ops.fma_only(wl), ops.ACT_NONE, out=out.view(b * h * w, w0.n))
                g.append(out)
                off += c
            self._gmaps = g
        return self._gmaps


class CandidatePlan:
    """Result of pass 1 of the candidate generation: device chunk offsets + host copy of the sample offsets."""

    def __init__(self, center, cam_intr, bbox, sdf_scale):
        self.center = center.detach().to(torch.float32).contiguous()
        self.cam_intr = cam_intr.detach().to(torch.float32).contiguous()
        self.bbox = bbox.detach().to(torch.float32).contiguous()
        self.sdf_scale = float(sdf_scale)
        self.counts, self.offsets = ops.lattice_count(self.center, self.cam_intr, self.bbox, self.sdf_scale, cfg.bins_n)
        self._host = torch.empty(self.offsets.shape, dtype=torch.int64, pin_memory=True)
        self._host.copy_(self.offsets, non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record()

    def host_offsets(self):
        self._event.synchronize()
        return self._host


class GraphedForward:
    """CUDA graphs of the two static-shape stages of one eval forward (same kernels, replayed without the ~400
    Python / ctypes launches): G1 = image encoder + pyramid projection, G2 = everything after the point selection
    (for the dexycb dataset branch also the two SDF queries at the supervision points and the ground-truth MANO forward).
    The selection itself stays eager: its row counts are data dependent (they size the candidate buffers)."""

    def __init__(self, model: "Model", img, root, objc, K, extras=None):
        self.model = model
        self.img, self.root, self.objc, self.K = img.clone(), root.clone(), objc.clone(), K.clone()
        self.extras = None if extras is None else {k: v.clone() for k, v in extras.items()}
        self.pool = torch.cuda.graph_pool_handle()
        self.g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g1, pool=self.pool, capture_error_mode="relaxed"):
            pyr, self.decoder_out = model.run_image_encoder(self.img)
            self.ctx = PyramidContext(pyr, model)
            self.ctx.gmaps                       # projection of the pyramid through linear_sdfin layer 0
            if cfg.fused_chain and (cfg.fused_gather or cfg.gather_h16) and cfg.screen_single:
                self.ctx.gmaps16                 # + its fp16 copy for the candidate screening stage
        self.g2 = self.sel = self.out = self.taps = self.dex_sdf = None

    def capture_pose(self, sel):
        self.sel = {k: v.clone() for k, v in sel.items()}
        self.g2 = torch.cuda.CUDAGraph()
        m, ex = self.model, self.extras
        with torch.cuda.graph(self.g2, pool=self.pool, capture_error_mode="relaxed"):
            self.out, self.taps = m._pose_from_points(self.ctx, self.sel, self.root, self.objc, self.K,
                                                      None if ex is None else ex["mano_param"])
            if ex is not None:                   # upstream model.py:376-391: SDF at the supervision points
                hand_s, _, _ = m.sdf_forward(self.ctx, ex["hand_sdf_points"], self.root, self.K, cfg.hand_sdf_scale, "hand")
                obj_s, _, _ = m.sdf_forward(self.ctx, ex["obj_sdf_points"], self.objc, self.K, cfg.obj_sdf_scale, "obj")
                self.dex_sdf = (hand_s, obj_s)


class Model(nn.Module):
    def __init__(self, backbone_net, decoder_net, hand_sdf_decoder, obj_sdf_decoder, hand_transformer,
                 obj_transformer, mano_layer):
        super().__init__()
        self.backbone_net = backbone_net
        self.decoder_net = decoder_net
        self.hand_sdf_decoder = hand_sdf_decoder
        self.obj_sdf_decoder = obj_sdf_decoder
        self.hand_transformer = hand_transformer
        self.obj_transformer = obj_transformer

        self.hand_sigmoid_beta = nn.Parameter(0.1 * torch.ones(1))
        self.obj_sigmoid_beta = nn.Parameter(0.1 * torch.ones(1))

        d = cfg.hidden_dim
        self.norm1 = nn.LayerNorm(cfg.mutliscale_dim)  # unused upstream too (model.py:55); kept for strict loading
        self.linear_transformerin = MLP(cfg.mutliscale_dim, [1024, 512, 256], d - cfg.PointFeatSize, 4, True)
        self.linear_sdfin = MLP(cfg.mutliscale_dim, [512], int(d), 2, True)
        if cfg.use_inverse_kinematics:
            raise NotImplementedError("the ho3d_render / inverse-kinematics setting is out of scope")
        self.mano_query_embed = nn.Embedding(cfg.mano_num_queries, d)
        self.mano_head = ManoHead(mano_layer, coord_change_mat=torch.tensor(
            [[1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]], dtype=torch.float32))
        self.linear_pose = MLP(d, d, 6, 3)
        self.linear_shape = MLP(d, d, 10, 3)
        self.linear_handvote = MLP(d, d, 20 * 3, 4)
        self.linear_handcls = MLP(d, d, 20, 3)
        self.linear_objvote = MLP(d, d, 8 * 3, 4)   # dead upstream as well (model.py:86-87)
        self.linear_objcls = MLP(d, d, 8, 3)
        self.linear_obj_rel_trans = MLP(d, d, 3, 3)
        self.linear_obj_rot = MLP(d, d, 3, 3)
        self.freeze_stages()
        self.last_taps = None      # diagnostics of the most recent forward (selected lattice indices, N_f, ...)
        self._graphs = None        # CUDA-graph mode (enable_cuda_graphs): {shape key: GraphedForward}

    def freeze_stages(self):
        """upstream model.py:114-121: the backbone's BatchNorm affine parameters are not trained."""
        for name, param in self.backbone_net.named_parameters():
            if "bn" in name:
                param.requires_grad = False

    # ------------------------------------------------------------------------------------------------
    # layout helper
    # ------------------------------------------------------------------------------------------------
    def channels_last_(self):
        """Run ResNet + U-Net in channels_last so the pyramid comes out NHWC (no transpose before the gather)."""
        self.backbone_net.to(memory_format=torch.channels_last)
        self.decoder_net.to(memory_format=torch.channels_last)
        self._channels_last = True
        return self

    def enable_cuda_graphs(self, enabled: bool = True):
        """Replay the static-shape stages of the eval forward (image encoder + projection; point features ->
        transformers -> heads) from CUDA graphs.  Same kernels and results; removes ~400 host-side launches per
        forward, so the step no longer depends on how fast the host can enqueue.  Graphs are captured per input shape on
        first use and dropped whenever a parameter or buffer changes."""
        self._graphs = {} if enabled else None
        return self

    def _weights_key(self):
        return hash(tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers())))

    # ------------------------------------------------------------------------------------------------
    # upstream-signature operators
    # ------------------------------------------------------------------------------------------------
    def sdf_activation(self, input, beta):
        """upstream model.py:123-126 (in-place floor of beta, then sigmoid(sdf/beta)/beta)."""
        beta.data.clamp_(min=2e-3)
        return torch.sigmoid(input / beta) / beta

    def _ctx(self, feature_pyramid):
        return feature_pyramid if isinstance(feature_pyramid, PyramidContext) else PyramidContext(feature_pyramid, self)

    def get_input_transformer(self, feature_pyramid, sdf_points, center_joint, cam_intr, sdf_scale):
        """upstream model.py:145-179 -> (transformer_latent (B,P,223), cam_sdf_points (B,P,3))."""
        _require_inference(self, sdf_points)
        ctx = self._ctx(feature_pyramid)
        pts = sdf_points.detach().to(torch.float32).contiguous()
        b, p, _ = pts.shape
        cam, uv = ops.project_points(pts, center_joint.contiguous(), cam_intr.contiguous(), sdf_scale)
        if ops.use_h3():    # the gather writes the (rows, C) matrix directly in the split-half format the MLP reads
            feats = ops.SplitRows.empty(b * p, ctx.channels, pts.device)
        else:
            feats = torch.empty(b * p, ctx.channels, device=pts.device, dtype=torch.float32)
        ops.gather(ctx.maps, uv, b, mode=ops.GATHER_CONCAT, out=feats, rows_per_sample=p, img_hw=cfg.input_img_shape)
        latent = self.linear_transformerin.forward_rows(feats)
        return latent.view(b, p, -1) if latent.is_contiguous() else latent.unflatten(0, (b, p)), cam

    def sdf_forward(self, feature_pyramid, sdf_points, center_joint, cam_intr, sdf_scale, type="hand"):
        """upstream model.py:181-244 -> (pred_sdf (B,P,1) clamped, pred_class, pos_enc3d (B,P,30))."""
        _require_inference(self, sdf_points)
        ctx = self._ctx(feature_pyramid)
        pts = sdf_points.detach().to(torch.float32).contiguous()
        b, p, _ = pts.shape
        dev = pts.device
        _, uv = ops.project_points(pts, center_joint.contiguous(), cam_intr.contiguous(), sdf_scale, want_cam=False)
        sdfin = self.linear_sdfin.packed()
        dec = self.hand_sdf_decoder if type == "hand" else self.obj_sdf_decoder
        if ops.use_h3():
            h, rows = self._row_buffers(b * p, dev)
            ops.gather(ctx.gmaps, uv, b, mode=ops.GATHER_SUM, out=h, rows_per_sample=p, bias=sdfin[0].b,
                       act=ops.ACT_RELU, img_hw=cfg.input_img_shape)
            ops.linear(h, sdfin[1], ops.ACT_RELU, out=rows.window(0, 256))
            ops.posenc(rows, points=pts.view(b * p, 3), bins=cfg.bins_n)
            sdf = ops.sdf_decoder(dec.packed(), rows, h_a=h, clamp=cfg.ClampingDistance)
            pe = rows.window(256, 30).float().view(b, p, 30)
            return sdf.view(b, p, 1), None, pe
        h = torch.empty(b * p, 512, device=dev, dtype=torch.float32)
        ops.gather(ctx.gmaps, uv, b, mode=ops.GATHER_SUM, out=h, rows_per_sample=p, bias=sdfin[0].b,
                   act=ops.ACT_RELU, img_hw=cfg.input_img_shape)
        rows = torch.empty(b * p, ops.ROW_LD, device=dev, dtype=torch.float32)
        ops.linear(h, sdfin[1], ops.ACT_RELU, out=rows[:, :256])
        ops.posenc(rows, points=pts.view(b * p, 3), bins=cfg.bins_n)
        sdf = ops.sdf_decoder(dec.packed(), rows, h_a=h, clamp=cfg.ClampingDistance)
        pe = rows[:, 256:286].contiguous().view(b, p, 30)
        return sdf.view(b, p, 1), (None if not cfg.ClassifierBranch else None), pe

    @staticmethod
    def _row_buffers(n, dev):
        """Split-half scratch of the FP16x3 SDF chain: h (n, 512) and the decoder row buffer (n, 520)."""
        return ops.SplitRows.empty(n, 512, dev), ops.SplitRows.empty(n, ops.ROWH_LD, dev)

    def plan_candidates(self, center_joint, cam_intr, bbox, sdf_scale) -> CandidatePlan:
        return CandidatePlan(center_joint, cam_intr, bbox, sdf_scale)

    def sdf_infer(self, feature_pyramid, center_joint, cam_intr, bbox, sdf_scale, num_points, type="hand",
                  plan: Optional[CandidatePlan] = None, taps: Optional[dict] = None, level: int = 0):
        """upstream model.py:246-355 -> (pose_points (B,P,3), pose_sdf (B,P,1), pose_posenc3d (B,P,30), None).

        Candidate lattice points of the whole batch are compacted into one row buffer (sample-major, ascending
        lattice index -- the order upstream's boolean-mask indexing yields), pushed through the projected-map
        gather + linear_sdfin + SDF decoder, and the `num_points` smallest |sdf| per sample are selected.

        `level` picks the screening cascade: 0 = single-product FP16 -> FP16x3 -> exact fp32, 1 = FP16x3 -> exact,
        2 = exact fp32 FMA on every candidate.  Each cascade step is verified on the device; a direct call (no
        `taps`) reads the verdict and escalates the level itself, `Model.forward` reads it once after the whole
        forward has been queued (`_hot_path`).
        """
        ctx = self._ctx(feature_pyramid)
        if plan is None:
            plan = self.plan_candidates(center_joint, cam_intr, bbox, sdf_scale)
        b = plan.center.shape[0]
        dev = plan.center.device
        if (level == 0 and cfg.native_sdf_infer and ops.PROFILE is None and ops.use_h3() and cfg.fused_chain and cfg.gather_h16
                and not cfg.fused_gather and cfg.final_stage == "h3" and cfg.screen_single):
            # the default cascade behind ONE C entry point (csrc/sdf_infer.cu: hoisdf_sdf_infer_fwd); None = a sample has
            # no room for the screening margin -> the general path below ranks every row exactly
            res = self._sdf_infer_native(ctx, plan, int(num_points), type, taps)
            if res is not None:
                if taps is None and not bool(res[1]):              # direct call: one tiny D2H read of the verdict
                    return self.sdf_infer(ctx, center_joint, cam_intr, bbox, sdf_scale, num_points, type, plan, None, 1)
                return res[0]
        host = plan.host_offsets()
        total = int(host[-1])
        n_f = host[1:] - host[:-1]
        if int(n_f.min()) < num_points:
            # upstream fails here too (model.py:348: shape mismatch when N_f < num_points)
            raise RuntimeError("sdf_infer: sample %d has %d lattice points inside its bbox, fewer than num_points=%d"
                               % (int(n_f.argmin()), int(n_f.min()), num_points))
        cand_index, cand_uv = ops.lattice_compact(plan.center, plan.cam_intr, plan.bbox, plan.sdf_scale, cfg.bins_n,
                                                  plan.counts, plan.offsets, total)
        sdfin = self.linear_sdfin.packed()
        dec = self.hand_sdf_decoder if type == "hand" else self.obj_sdf_decoder
        packed = dec.packed()
        gmaps = ctx.gmaps
        sdf = torch.empty(total, device=dev, dtype=torch.float32)
        step = int(cfg.max_rows_per_pass)
        cap = min(step, total)
        bufs = {}
        nmin = int(n_f.min())

        def fp32_buffers(n):
            if "f32" not in bufs or bufs["f32"][0].shape[0] < n:
                bufs["f32"] = (torch.empty(n, 512, device=dev, dtype=torch.float32),
                               torch.empty(n, 512, device=dev, dtype=torch.float32),
                               torch.empty(n, ops.ROW_LD, device=dev, dtype=torch.float32))
            return bufs["f32"]

        def h3_buffers(n):
            if "h3" not in bufs or bufs["h3"][0].rows < n:
                hs, rs = self._row_buffers(n, dev)
                bufs["h3"] = (hs, rs, ops.SplitRows.empty(n, 512, dev))
            return bufs["h3"]

        def chain_h3(uv, index, out, single, row_offsets=None, rows_per_sample=0, chunk_kb=ops.SCREEN_CHUNK_KB):
            """gather -> linear_sdfin[1] -> posenc -> SDF decoder on the FP16x3 kernels (split-half rows).  These values
            only RANK candidates for the next, more accurate stage: one TMEM drain per tile, and with `single` ONE
            tensor-core product per K step instead of three."""
            n = uv.shape[0]
            if single and cfg.fused_chain and cfg.fused_gather:
                # ONE kernel: gather -> linear_sdfin.1 -> embedding -> SDF decoder; 4 bytes per row written
                ops.sdf_chain(packed, out, sdfin1=sdfin[1], gmaps16=ctx.gmaps16, uv=uv, row_offsets=row_offsets, batch=b,
                              rows_per_sample=rows_per_sample, bias0=sdfin[0].b, lattice_index=index, bins=cfg.bins_n,
                              img_hw=cfg.input_img_shape)
                return
            hs, rs, hs2 = h3_buffers(n)
            if single and cfg.fused_chain:
                if cfg.gather_h16:      # fp16 maps in, fp16 hi plane out: all the single-product chain kernel reads
                    ops.gather_h16(ctx.gmaps16, uv, b, hs.head(n), row_offsets=row_offsets, rows_per_sample=rows_per_sample,
                                   bias=sdfin[0].b, act=ops.ACT_RELU, img_hw=cfg.input_img_shape)
                else:
                    ops.gather(gmaps, uv, b, mode=ops.GATHER_SUM, out=hs.head(n), row_offsets=row_offsets,
                               rows_per_sample=rows_per_sample, bias=sdfin[0].b, act=ops.ACT_RELU,
                               img_hw=cfg.input_img_shape)
                ops.sdf_chain(packed, out, sdfin1=sdfin[1], a0=hs.head(n), lattice_index=index, bins=cfg.bins_n)
                return
            ops.gather(gmaps, uv, b, mode=ops.GATHER_SUM, out=hs.head(n), row_offsets=row_offsets,
                       rows_per_sample=rows_per_sample, bias=sdfin[0].b, act=ops.ACT_RELU, img_hw=cfg.input_img_shape)
            ops.linear(hs.head(n), sdfin[1], ops.ACT_RELU, out=rs.head(n).window(0, 256),
                       chunk_kb=chunk_kb, single=single)
            ops.posenc(rs.head(n), lattice_index=index, bins=cfg.bins_n)
            ops.sdf_decoder(packed, rs.head(n), h_a=hs.head(n), h_b=hs2.head(n), out=out,
                            chunk_kb=chunk_kb, single=single)

        def chain_f32(uv, index, out, passes, exact, row_offsets=None, rows_per_sample=0):
            """The same chain on fp32 rows: bit-faithful fp32 FMA kernels (exact), or the 3xTF32 / 1xTF32 tensor-core
            kernels of TC_MODE 'tf32'."""
            n = uv.shape[0]
            h, h2, rows = fp32_buffers(n)
            ops.gather(gmaps, uv, b, mode=ops.GATHER_SUM, out=h[:n], row_offsets=row_offsets,
                       rows_per_sample=rows_per_sample, bias=sdfin[0].b, act=ops.ACT_RELU, img_hw=cfg.input_img_shape)
            ops.linear(h[:n], ops.fma_only(sdfin[1]) if exact else sdfin[1], ops.ACT_RELU, out=rows[:n, :256],
                       passes=passes)
            ops.posenc(rows[:n], lattice_index=index, bins=cfg.bins_n)
            ops.sdf_decoder(packed, rows[:n], h_a=h[:n], h_b=h2[:n], out=out, exact=exact,
                            screening=(passes == 1 and not exact))

        def evaluate_all(passes, single=False):
            """SDF of every candidate row: FP16x3 / single-product FP16 on split-half rows by default; TC_MODE 'tf32':
            passes = 3: 3xTF32, 1: single TF32 pass; with tensor cores disabled the fp32 FMA kernels."""
            h3 = ops.use_h3()
            if h3:
                if not (single and cfg.fused_chain and cfg.fused_gather):      # the fused kernel needs no row buffers
                    h3_buffers(cap)
            else:
                fp32_buffers(cap)
            for r0 in range(0, total, step):
                n = min(step, total - r0)
                offs = plan.offsets if r0 == 0 else plan.offsets - r0
                if h3:
                    chain_h3(cand_uv[r0:r0 + n], cand_index[r0:r0 + n], sdf[r0:r0 + n], single, row_offsets=offs)
                else:
                    chain_f32(cand_uv[r0:r0 + n], cand_index[r0:r0 + n], sdf[r0:r0 + n], passes,
                              exact=not ops.USE_TENSOR_CORES, row_offsets=offs)

        def refine(src_sdf, src_offsets, src_index, src_uv, keep, target, evaluate):
            """One cascade step: keep the `keep` best rows per sample of the current ranking (lattice order) and
            re-evaluate them with the more accurate `evaluate`.  The step is loss-free for the `target` best rows
            of the accurate ranking whenever the coarse error is smaller than the |sdf| gap between the accurate
            rank-`target` value and the coarse rank-`keep` value; both are measured here on the device.  The error is
            observed on the KEPT rows only (the discarded rows are never re-evaluated), so the verdict is an estimate
            with a 3x margin, not a proof: a discarded row whose coarse error exceeded 3x the largest kept-row error
            would go unnoticed.  The configs[1] / configs[2] parity runs (profiles/r03c_parity_*.json: identical selected
            sets at full batch) are the evidence that the margin holds on this path's activations."""
            _, _, s_sdf, _, _, s_row = ops.select_points(src_sdf, src_offsets, src_index, b, keep, cfg.bins_n, 0.0,
                                                         order_by_row=True)
            s_row = s_row.view(-1).long()
            new_index = src_index.index_select(0, s_row)
            new_uv = src_uv.index_select(0, s_row)
            new_sdf = torch.empty(b * keep, device=dev, dtype=torch.float32)
            evaluate(new_uv, new_index, new_sdf, rows_per_sample=keep)
            new_offsets = torch.arange(0, (b + 1) * keep, keep, device=dev, dtype=torch.int64)
            coarse, fine = s_sdf.view(b, keep), new_sdf.view(b, keep)
            err = (coarse - fine).abs().max()                               # observed error of the coarse values
            kth = fine.abs().kthvalue(target, dim=1).values                # accurate rank-`target` |sdf|
            gap = coarse.abs().max(dim=1).values - kth                      # coarse rank-`keep` |sdf| - that
            return new_sdf, new_offsets, new_index, new_uv, dict(rows=s_row, err=err, gap=gap, keep=keep,
                                                                 verified=(gap > 3.0 * err).all())

        if cfg.final_stage == "h3" and ops.use_h3():
            # final ranking on the FP16x3 GEMM with a TMEM drain every K block (fp32-FMA-grade values)
            exact_eval = lambda uv, idx, out, rows_per_sample: chain_h3(uv, idx, out, False,      # noqa: E731
                                                                        rows_per_sample=rows_per_sample, chunk_kb=1)
        else:
            exact_eval = lambda uv, idx, out, rows_per_sample: chain_f32(uv, idx, out, 3, True,   # noqa: E731
                                                                         rows_per_sample=rows_per_sample)
        screened = pre = None
        single_used = False
        keep_exact = int(min(num_points + cfg.screen_margin_safe, nmin, 8192))
        if not ops.USE_TENSOR_CORES:
            evaluate_all(3)                                               # fp32 FMA everywhere: already exact
            sdf_sel, cand_sel, offs_sel = sdf, cand_index, plan.offsets
        elif nmin <= num_points or level >= 2:
            # no room for a screening margin: every candidate is selected anyway, rank them all exactly
            with _tensor_cores(False):
                evaluate_all(3)
            sdf_sel, cand_sel, offs_sel = sdf, cand_index, plan.offsets
        elif ops.use_h3():
            # Coarse-to-fine cascade.  (A) single-product FP16 on ALL candidates (1/3 of the tensor work) keeps the best
            # P + screen_margin_single rows; final stage: FP16x3 with a TMEM drain every K block re-evaluates those
            # and the final P are selected from its values (cfg.final_stage = "fma": (B) FP16x3 keeps
            # P + screen_margin_safe, (C) the fp32 FMA kernels re-evaluate those).
            # Equal to ranking ALL candidates with the final stage whenever each step's error is smaller than the |sdf| gap it leaves --
            # verified on the device; `screen_ok` is read by the caller after the forward has been queued, and a
            # failed check re-runs the query without stage A (or raises if the FP16x3 stage itself fails).
            keep_a = int(min(num_points + cfg.screen_margin_single, 8192))
            use_a = bool(cfg.screen_single) and level == 0 and nmin > keep_a and keep_a > keep_exact
            evaluate_all(3, single=use_a)
            cur = (sdf, plan.offsets, cand_index, cand_uv)
            single_used = use_a
            if cfg.final_stage == "h3":
                # the final stage is itself a tensor-core pass (FP16x3 draining TMEM every K block: measured as close
                # to the oracle as the fp32 FMA kernels), cheap enough to take all of stage A's survivors at once
                sdf_sel, offs_sel, cand_sel, _, screened = refine(*cur, keep_a if use_a else keep_exact, num_points,
                                                                  exact_eval)
            else:
                if use_a:
                    h3_eval = lambda uv, idx, out, rows_per_sample: chain_h3(uv, idx, out, False,    # noqa: E731
                                                                             rows_per_sample=rows_per_sample)
                    *cur, pre = refine(*cur, keep_a, keep_exact, h3_eval)
                sdf_sel, offs_sel, cand_sel, _, screened = refine(*cur, keep_exact, num_points, exact_eval)
        else:
            # TC_MODE 'tf32': 3xTF32 (or opt-in single-pass TF32, verified with one tiny D2H read) screening + exact re-rank
            passes = 1 if int(cfg.screen_passes) == 1 else 3
            evaluate_all(passes)
            margin = cfg.screen_margin if passes == 1 else cfg.screen_margin_safe
            cur = (sdf, plan.offsets, cand_index, cand_uv)
            sdf_sel, offs_sel, cand_sel, _, screened = refine(*cur, int(min(num_points + margin, nmin, 8192)),
                                                              num_points, exact_eval)
            if passes == 1 and not bool(screened["verified"]):
                evaluate_all(3)
                sdf_sel, offs_sel, cand_sel, _, screened = refine(*cur, keep_exact, num_points, exact_eval)
        sel, pts, out_sdf, pe, _flag, _ = ops.select_points(sdf_sel, offs_sel, cand_sel, b, num_points, cfg.bins_n,
                                                            cfg.ClampingDistance)
        if taps is not None:
            taps.update(index=sel, n_f=n_f.clone(), cand_index=cand_index, cand_sdf=sdf, offsets=host.clone())
            if screened is not None:
                taps["screen_gap"] = screened["gap"]
                taps["screen_err"] = screened["err"]
                taps["screen_rows"] = screened["rows"] if pre is None else pre["rows"].index_select(0, screened["rows"])
                taps["screen_verified"] = screened["verified"]
            if pre is not None:
                taps["pre_gap"], taps["pre_err"], taps["pre_verified"] = pre["gap"], pre["err"], pre["verified"]
            taps["single_pass"] = single_used
            taps["exact_sdf"], taps["exact_index"] = sdf_sel, cand_sel     # what the final selection ranked
        elif ops.use_h3() and screened is not None:
            ok = screened["verified"] if pre is None else (screened["verified"] & pre["verified"])
            if not bool(ok):                                               # direct call: one tiny D2H read
                return self.sdf_infer(ctx, center_joint, cam_intr, bbox, sdf_scale, num_points, type, plan, None,
                                      level + 1)
        return pts, out_sdf, pe, None

    def _sdf_infer_native(self, ctx, plan, num_points, type, taps):
        """`hoisdf_sdf_infer_fwd`: candidate compaction, the screening cascade, its device-side verdict and the final
        top-P in one C call with a caller-owned workspace.  Returns ((points, sdf, posenc, None), verified flag tensor) or
        None when the C side asks for the general path."""
        import ctypes as C
        from . import _capi
        b, dev = plan.center.shape[0], plan.center.device
        sdfin = self.linear_sdfin.packed()
        s1 = sdfin[1].h3
        if s1 is None or s1.scale != 1.0 or sdfin[1].n != 256 or sdfin[1].k != 512:
            return None
        dec = self.hand_sdf_decoder if type == "hand" else self.obj_sdf_decoder
        packed = dec.packed()
        if packed.struct_h3 is None:
            return None
        host = plan.host_offsets()                  # the plan's event: the B + 1 row offsets have landed in pinned memory
        total = int(host[-1])
        n_f = host[1:] - host[:-1]
        margin = int(cfg.screen_margin_single)
        keep = int(_capi.lib.hoisdf_sdf_infer_keep(num_points, margin))
        max_rows = max(ops.round_up(total, 1 << 16), 1 << 16)
        nbytes = int(_capi.lib.hoisdf_sdf_infer_workspace_bytes(b, max_rows, num_points, margin, int(cfg.bins_n)))
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        pts = torch.empty(b, num_points, 3, device=dev, dtype=torch.float32)
        sdf = torch.empty(b, num_points, 1, device=dev, dtype=torch.float32)
        pe = torch.empty(b, num_points, 30, device=dev, dtype=torch.float32)
        sel = torch.empty(b, num_points, device=dev, dtype=torch.int32)
        flag = torch.zeros(1, device=dev, dtype=torch.int32)
        err = torch.empty(1, device=dev, dtype=torch.float32)
        gap = torch.empty(b, device=dev, dtype=torch.float32)
        ok = torch.empty(1, device=dev, dtype=torch.int32)
        gm, gm16 = ops.make_pyramid(ctx.gmaps, cfg.input_img_shape), ops._pyramid_h(ctx.gmaps16, cfg.input_img_shape)
        a = _capi.SdfInferArgs()
        a.center, a.cam_intr, a.bbox = plan.center.data_ptr(), plan.cam_intr.data_ptr(), plan.bbox.data_ptr()
        a.sdf_scale, a.bins, a.batch, a.num_points, a.margin = plan.sdf_scale, int(cfg.bins_n), b, num_points, margin
        a.clamp = float(cfg.ClampingDistance)
        a.gmaps, a.gmaps16, a.bias0 = C.addressof(gm), C.addressof(gm16), sdfin[0].b.data_ptr()
        a.s1_a, a.s1_b, a.s1_c, a.ld_s1 = s1.plane_ptr(0), s1.plane_ptr(1), s1.plane_ptr(2), s1.ld
        a.b_s1, a.s1_scale = sdfin[1].b.data_ptr(), float(s1.scale)
        a.dec = C.addressof(packed.struct_h3)
        a.workspace, a.workspace_bytes, a.max_rows = ws.data_ptr(), nbytes, max_rows
        a.planned, a.chunk_counts, a.offsets, a.host_offsets = 1, plan.counts.data_ptr(), plan.offsets.data_ptr(), host.data_ptr()
        a.points, a.sdf, a.posenc, a.sel_index, a.status_flag = pts.data_ptr(), sdf.data_ptr(), pe.data_ptr(), \
            sel.data_ptr(), flag.data_ptr()
        a.screen_err, a.screen_gap, a.verified = err.data_ptr(), gap.data_ptr(), ok.data_ptr()
        diag = {}
        if taps is not None:                        # diagnostics the parity tests read
            diag = dict(cand_sdf=torch.empty(max(total, 1), device=dev, dtype=torch.float32),
                        cand_index=torch.empty(max(total, 1), device=dev, dtype=torch.int32),
                        exact_sdf=torch.empty(b * keep, device=dev, dtype=torch.float32),
                        exact_index=torch.empty(b * keep, device=dev, dtype=torch.int32),
                        screen_rows=torch.empty(b * keep, device=dev, dtype=torch.int32))
            a.cand_sdf, a.cand_index = diag["cand_sdf"].data_ptr(), diag["cand_index"].data_ptr()
            a.exact_sdf, a.exact_index = diag["exact_sdf"].data_ptr(), diag["exact_index"].data_ptr()
            a.screen_rows = diag["screen_rows"].data_ptr()
        ops._count(14)
        st = _capi.lib.hoisdf_sdf_infer_fwd(C.byref(a), ops._stream())
        if st == _capi.E_UNSUPPORTED:
            return None
        if st == _capi.E_TOO_FEW_POINTS:
            # upstream fails here too (model.py:348: shape mismatch when N_f < num_points)
            raise RuntimeError("sdf_infer: sample %d has %d lattice points inside its bbox, fewer than num_points=%d"
                               % (int(n_f.argmin()), int(n_f.min()), num_points))
        _capi.check(st, "hoisdf_sdf_infer_fwd")
        verified = ok.view(()) != 0
        if taps is not None:
            taps.update(index=sel, n_f=n_f.clone(), cand_index=diag["cand_index"][:total], cand_sdf=diag["cand_sdf"][:total],
                        offsets=host.clone(), screen_gap=gap, screen_err=err.view(()), screen_rows=diag["screen_rows"].long(),
                        screen_verified=verified, single_pass=True, exact_sdf=diag["exact_sdf"],
                        exact_index=diag["exact_index"], native=True)
        return (pts, sdf, pe, None), verified

    # ------------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------------
    def forward(self, inputs, targets, meta_info, mode, epoch_cnt=1e8, batch_ratio=0):
        dev = inputs["img"].device
        if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
            # every launch below goes to the CURRENT device's stream: make the tensors' device current (a process driving
            # several GPUs, e.g. a DataParallel replica thread that was handed cuda:1 tensors)
            with torch.cuda.device(dev):
                return self.forward(inputs, targets, meta_info, mode, epoch_cnt, batch_ratio)
        if mode == "train":
            from .train import forward_train          # upstream model.py:357-665, train branch (hoisdf_b200/train.py)
            return forward_train(self, inputs, targets, meta_info, epoch_cnt, batch_ratio)
        graphed = (self._graphs is not None and cfg.tc_backbone and cfg.tc_unet and ops.use_h3()
                   and inputs["img"].is_cuda)
        dex = cfg.dataset == "dexycb"
        with torch.no_grad():
            if graphed:
                out, decoder_out, dex_sdf = self._forward_graphed(inputs, targets, meta_info)
            else:
                img = inputs["img"]
                plans = self._plans(meta_info)
                if getattr(self, "_channels_last", False):
                    img = img.contiguous(memory_format=torch.channels_last)
                feature_pyramid, decoder_out = self.run_image_encoder(img)
                ctx = self._ctx(feature_pyramid)
                out = self._hot_path(ctx, meta_info, plans, targets["mano_param"] if dex else None)
                dex_sdf = None
                if dex:
                    # upstream model.py:370-391: SDF at the supervision points
                    root, objc, K = meta_info["mano_root"], meta_info["obj_center_cam"], meta_info["cam_intr"]
                    hand_s, _, _ = self.sdf_forward(ctx, inputs["hand_sdf_points"], root, K, cfg.hand_sdf_scale, "hand")
                    obj_s, _, _ = self.sdf_forward(ctx, inputs["obj_sdf_points"], objc, K, cfg.obj_sdf_scale, "obj")
                    dex_sdf = (hand_s, obj_s)
            taps = self.last_taps
            if dex:
                # upstream model.py:393-422,606-620: heat-map / segmentation heads, GT MANO, the extra loss entries
                out["joint_heatmap_out"] = decoder_out[:, 0]
                out["hand_seg_gt_out"] = targets["hand_seg"]
                out["hand_seg_pred_out"] = decoder_out[:, 1]
                out["obj_seg_gt_out"] = targets["obj_seg"]
                out["obj_seg_pred_out"] = decoder_out[:, 2]
                out["mano_joints_gt_out"] = taps["gt_mano"]["joints3d"].clone() if graphed else taps["gt_mano"]["joints3d"]
                out["mano_mesh_gt_out"] = taps["gt_mano"]["verts3d"].clone() if graphed else taps["gt_mano"]["verts3d"]
                if cfg.eval_losses:
                    out = {**dexycb_losses(out, taps, targets, decoder_out, dex_sdf[0], dex_sdf[1], taps["pred_mano"],
                                           taps["gt_mano"]), **out}
            if cfg.eval_losses:
                joint_gt = targets["joint_cam_no_trans"][:, 1:] if dex else None
                out = {**eval_losses(taps, targets, meta_info, joint_gt), **out}
        return out

    def _forward_graphed(self, inputs, targets, meta_info):
        """The eval forward with its two static stages replayed from CUDA graphs (see GraphedForward).  Returns
        (`*_out` dict, decoder_out, (hand, obj) SDF at the dexycb supervision points or None): copies, the graphs'
        output buffers are reused by the next forward."""
        root = meta_info["mano_root"].to(torch.float32).contiguous()
        objc = meta_info["obj_center_cam"].to(torch.float32).contiguous()
        K = meta_info["cam_intr"].to(torch.float32).contiguous()
        img = inputs["img"].to(torch.float32)
        dex = cfg.dataset == "dexycb"
        extras = None
        if dex:
            extras = {"hand_sdf_points": inputs["hand_sdf_points"].to(torch.float32).contiguous(),
                      "obj_sdf_points": inputs["obj_sdf_points"].to(torch.float32).contiguous(),
                      "mano_param": targets["mano_param"].to(torch.float32).contiguous()}
        wkey = self._weights_key()
        if self._graphs.get("weights") != wkey:
            self._graphs = {"weights": wkey}         # parameters changed: packed weights were rebuilt, recapture
        key = (tuple(img.shape), str(img.device), int(cfg.num_samp_hand), int(cfg.num_samp_obj), cfg.setting, cfg.dataset,
               cfg.final_stage, bool(cfg.screen_single), bool(cfg.tc_projection), int(cfg.backbone_chunk_kb),
               bool(cfg.fused_chain), bool(cfg.fused_gather), bool(cfg.gather_h16), None if extras is None else tuple(tuple(v.shape) for v in extras.values()))
        gs = self._graphs.get(key)
        if gs is None:
            # one eager forward first: packs the weights, sizes the workspaces, sets the kernel attributes
            pyr, _ = self.run_image_encoder(img)
            ctx = PyramidContext(pyr, self)
            self._hot_path(ctx, meta_info, None, None if extras is None else extras["mano_param"])
            if dex:
                self.sdf_forward(ctx, extras["hand_sdf_points"], root, K, cfg.hand_sdf_scale, "hand")
                self.sdf_forward(ctx, extras["obj_sdf_points"], objc, K, cfg.obj_sdf_scale, "obj")
            torch.cuda.synchronize()
            gs = self._graphs[key] = GraphedForward(self, img, root, objc, K, extras)
        plans = self._plans(meta_info)               # eager: lattice counts + the async read-back of the row counts
        gs.img.copy_(img)
        gs.root.copy_(root)
        gs.objc.copy_(objc)
        gs.K.copy_(K)
        if dex:
            for k, v in extras.items():
                gs.extras[k].copy_(v)
        gs.g1.replay()
        level = 0
        while True:
            sel, th, to, verdict = self._select_points(gs.ctx, gs.root, gs.objc, gs.K, plans, level)
            if gs.g2 is None:
                gs.capture_pose(sel)
            for k, v in sel.items():
                gs.sel[k].copy_(v)
            gs.g2.replay()
            if self._verdict_ok(verdict):
                break
            if level >= 2:
                raise RuntimeError("point-selection screening could not be verified")
            level += 1                               # rare: escalate the cascade, replay the pose stage
        self.last_taps = dict(gs.taps, hand=th, obj=to)
        out = {k: v.clone() for k, v in gs.out.items()}
        if not dex:
            return out, None, None
        return out, gs.decoder_out.clone(), tuple(t.clone() for t in gs.dex_sdf)

    def run_image_encoder(self, img):
        """ResNet-50 + U-Net -> (feature pyramid, decoder_out).  On the FP16x3 tensor-core kernels end to end when
        enabled (activations never leave the NHWC split-half format between the two networks), else cuDNN."""
        if cfg.tc_backbone and cfg.tc_unet and ops.use_h3():
            from .nets.resnet_h3 import ResNetH3
            if getattr(self, "_resnet_h3", None) is None or self._resnet_h3.net is not self.backbone_net.resnet:
                self._resnet_h3 = ResNetH3(self.backbone_net.resnet)
            unet = self._unet()
            b, _, h, w = img.shape
            cats, slots = unet.concat_slots(b, h // 32, w // 32, img.device)
            feat, skips = self._resnet_h3(img, slots)
            return unet.run(feat, skips, cats)
        img_feat, skips = self.backbone_net(img)
        return self.run_decoder(img_feat, skips)

    def _unet(self):
        if getattr(self, "_unet_h3", None) is None or self._unet_h3.dec is not self.decoder_net.resnet_decoder:
            self._unet_h3 = UNetH3(self.decoder_net.resnet_decoder)
        return self._unet_h3

    def run_decoder(self, img_feat, skips):
        """U-Net decoder: on the FP16x3 tensor-core kernels (nets/unet_h3.py) when enabled, else the cuDNN modules."""
        if cfg.tc_unet and ops.use_h3():
            return self._unet()(img_feat, skips)
        return self.decoder_net(img_feat, skips)

    def _plans(self, meta_info):
        K = meta_info["cam_intr"]
        return (self.plan_candidates(meta_info["mano_root"], K, meta_info["bbox_hand"], cfg.hand_sdf_scale),
                self.plan_candidates(meta_info["obj_center_cam"], K, meta_info["bbox_obj"], cfg.obj_sdf_scale))

    def hot_path(self, feature_pyramid, meta_info, plans=None, mano_params=None):
        """Everything of upstream Model.forward(mode='eval') after the U-Net (model.py:424-638), on our kernels.
        Returns the `*_out` entries; intermediate tensors are left in `self.last_taps`."""
        with torch.no_grad():
            return self._hot_path(feature_pyramid, meta_info, plans, mano_params)

    def _hot_path(self, feature_pyramid, meta_info, plans, mano_params=None, level: int = 0):
        root = meta_info["mano_root"].to(torch.float32).contiguous()
        objc = meta_info["obj_center_cam"].to(torch.float32).contiguous()
        K = meta_info["cam_intr"].to(torch.float32).contiguous()
        if plans is None:
            plans = self._plans(meta_info)
        ctx = self._ctx(feature_pyramid)
        sel, th, to, verdict = self._select_points(ctx, root, objc, K, plans, level)
        out, taps = self._pose_from_points(ctx, sel, root, objc, K, mano_params)
        if not self._verdict_ok(verdict):
            if level >= 2:
                raise RuntimeError("point-selection screening could not be verified")
            return self._hot_path(ctx, meta_info, plans, mano_params, level + 1)   # rare: escalate the cascade
        self.last_taps = dict(taps, hand=th, obj=to)
        return out

    def _select_points(self, ctx, root, objc, K, plans, level):
        """upstream model.py:424-481: the two `sdf_infer` calls.  Returns the selected points / SDF / positional
        encodings, the per-field diagnostics and the (event, pinned flag) of the cascade's device-side verdict."""
        Ph, Po = int(cfg.num_samp_hand), int(cfg.num_samp_obj)
        th, to = {}, {}
        hp, hsdf, hpe, _ = self.sdf_infer(ctx, root, K, None, cfg.hand_sdf_scale, Ph, "hand", plans[0], th, level)
        op, osdf, ope, _ = self.sdf_infer(ctx, objc, K, None, cfg.obj_sdf_scale, Po, "obj", plans[1], to, level)
        # verdict of the screening cascade (device-side checks): copied to pinned memory now, read after the rest of
        # the forward has been queued -- the GPU never waits for the host
        flags = [t[k] for t in (th, to) for k in ("pre_verified", "screen_verified") if k in t]
        verdict = None
        if flags:
            if getattr(self, "_ok_host", None) is None:
                self._ok_host = torch.empty(1, dtype=torch.bool, pin_memory=True)
            self._ok_host.copy_(torch.stack(flags).all().view(1), non_blocking=True)
            verdict = torch.cuda.Event()
            verdict.record()
        sel = dict(hand_points=hp, hand_sdf=hsdf, hand_posenc=hpe, obj_points=op, obj_sdf=osdf, obj_posenc=ope)
        return sel, th, to, verdict

    def _verdict_ok(self, verdict) -> bool:
        if verdict is None:
            return True
        verdict.synchronize()
        return bool(self._ok_host[0])

    def _masks(self, dev):
        """(tgt_mask on the device, memory_mask on the host) -- constants, built once (upstream rebuilds them on the
        CPU and copies them every forward, model.py:568-569)."""
        key = (str(dev), int(cfg.num_samp_hand), int(cfg.num_samp_obj))
        if getattr(self, "_mask_cache", None) is None or self._mask_cache[0] != key:
            self._mask_cache = (key, get_mano_tgt_mask().to(dev), get_mano_memory_mask())
        return self._mask_cache[1], self._mask_cache[2]

    def _pose_from_points(self, ctx, sel, root, objc, K, mano_params=None):
        """upstream model.py:483-638: point features, cross SDF queries, tokens, the two transformers, heads, MANO and
        the joint vote, from the selected points.  Static shapes, no host round trip: capturable in a CUDA graph."""
        b = ctx.batch
        dev = root.device
        Ph, Po = int(cfg.num_samp_hand), int(cfg.num_samp_obj)
        S = Ph + Po
        hs_scale, os_scale = cfg.hand_sdf_scale, cfg.obj_sdf_scale
        hand_points, hand_sdf, hand_pe = sel["hand_points"], sel["hand_sdf"], sel["hand_posenc"]
        obj_points, obj_sdf, obj_pe = sel["obj_points"], sel["obj_sdf"], sel["obj_posenc"]

        self.hand_sigmoid_beta.data.clamp_(min=2e-3)   # upstream model.py:124 (side effect on the parameter)
        self.obj_sigmoid_beta.data.clamp_(min=2e-3)
        beta_h, beta_o = self.hand_sigmoid_beta.data, self.obj_sigmoid_beta.data

        hand_fea, hand_cam = self.get_input_transformer(ctx, hand_points, root, K, hs_scale)
        obj_fea, obj_cam = self.get_input_transformer(ctx, obj_points, objc, K, os_scale)
        hand_nt = hand_cam - root[:, None, :]
        obj_nt = obj_cam - objc[:, None, :]
        hand_o_nt = hand_cam - objc[:, None, :]          # upstream model.py:498 ("bug": unscaled coords, kept)
        obj_h_nt = obj_cam - root[:, None, :]            # upstream model.py:508
        hand_o_sdf, _, hand_o_pe = self.sdf_forward(ctx, hand_o_nt * os_scale, objc, K, os_scale, "obj")
        obj_h_sdf, _, obj_h_pe = self.sdf_forward(ctx, obj_h_nt * hs_scale, root, K, hs_scale, "hand")

        hand_in = torch.empty(b, S, 256, device=dev, dtype=torch.float32)
        ops.tokens(hand_nt, hand_pe, hand_fea, hand_sdf, beta_h, hand_in, 0)
        ops.tokens(obj_h_nt, obj_h_pe, obj_fea, obj_h_sdf, beta_h, hand_in, Ph)
        obj_in = torch.empty(b, S, 256, device=dev, dtype=torch.float32)
        ops.tokens(obj_nt, obj_pe, obj_fea, obj_sdf, beta_o, obj_in, 0)
        ops.tokens(hand_o_nt, hand_o_pe, hand_fea, hand_o_sdf, beta_o, obj_in, Po)

        tgt_mask, memory_mask = self._masks(dev)      # memory_mask stays on the host: recognised as a key-range limit
        hs, memory, hand_enc = self.hand_transformer.forward_bm(hand_in, self.mano_query_embed.weight, None,
                                                                tgt_mask, memory_mask)
        _, obj_enc = self.obj_transformer.forward_bm(obj_in, None)

        Le, Lo, Ld = hand_enc.shape[0], obj_enc.shape[0], hs.shape[0]
        # FP16x3: every source tensor is split once and shared by the heads that read it
        sp = (lambda t: ops.split_rows(t.view(-1, t.shape[-1]))) if ops.use_h3() else (lambda t: None)
        # (the encoders' LayerNorm kernels already wrote split-half copies of their intermediate outputs)
        hx = getattr(self.hand_transformer.encoder, "last_inter_split", None) or sp(hand_enc)
        ox = getattr(self.obj_transformer.encoder, "last_inter_split", None) or sp(obj_enc)
        qx = sp(hs)
        hand_off = self._head_rows(self.linear_handvote, hand_enc, Le * b, Ph, S, xs=hx).view(Le, b, Ph, 60)
        hand_cls = self._head_rows(self.linear_handcls, hand_enc, Le * b, Ph, S, xs=hx).view(Le, b, Ph, 20)
        obj_rot = self._head_rows(self.linear_obj_rot, obj_enc, Lo * b, Po, S, xs=ox).view(Lo, b, Po, 3)
        obj_trans = self._head_rows(self.linear_obj_rel_trans, obj_enc, Lo * b, Po, S, xs=ox).view(Lo, b, Po, 3)
        nq = hs.shape[2]
        pose6d = self._head_rows(self.linear_pose, hs, Ld * b, cfg.mano_shape_indx, nq, xs=qx).view(
            Ld, b, cfg.mano_shape_indx, 6)
        shape = self._head_rows(self.linear_shape, hs, Ld * b, 1, nq, first=cfg.mano_shape_indx, xs=qx).view(Ld, b, 10)
        verts, joints = self.mano_head.forward_bm(pose6d, shape)
        hand_joints = ops.vote_joints(hand_nt.contiguous(), hand_off, hand_cls)
        pred_mano = gt_mano = None
        if mano_params is not None:      # dexycb eval: ground-truth MANO forward + what the pose/shape losses need
            from .nets.mano_head import rot6d2mat
            pred_mano = {"verts3d": verts, "joints3d": joints, "mano_shape": shape,
                         "mano_pose": rot6d2mat(pose6d.reshape(-1, 6)).view(Ld, b, cfg.mano_shape_indx, 3, 3)}
            gt_mano = self.mano_head.forward_gt(mano_params)

        out = {
            "mano_mesh_out": verts[-1],
            "mano_joints_out": joints[-1],
            "obj_rot_out": obj_rot[-1],
            "obj_trans_out": obj_trans[-1],
            "hand_joints_out": hand_joints[-1],
        }
        taps = dict(
            hand_points=hand_points, hand_sdf=hand_sdf, hand_posenc=hand_pe, obj_points=obj_points,
            obj_sdf=obj_sdf, obj_posenc=obj_pe, hand_fea=hand_fea, obj_fea=obj_fea, hand_o_sdf=hand_o_sdf,
            obj_h_sdf=obj_h_sdf, hand_transformer_in=hand_in, obj_transformer_in=obj_in, hs=hs, memory=memory,
            hand_encoder_out=hand_enc, obj_encoder_out=obj_enc, hand_off=hand_off, hand_cls=hand_cls, obj_rot=obj_rot,
            obj_trans=obj_trans, mano_pose6d=pose6d, mano_shape=shape, hand_joints=hand_joints, mano_verts=verts,
            mano_joints=joints, hand_points_notrans=hand_nt, pred_mano=pred_mano, gt_mano=gt_mano)
        return out, taps

    @staticmethod
    def _head_rows(mlp: MLP, x: torch.Tensor, groups: int, rows_per_group: int, group_len: int, first: int = 0,
                   xs=None):
        """Run `mlp` on tokens [first, first+rows_per_group) of every (layer, sample) group of a (L,B,T,256) tensor
        without gathering them first: the Linear kernel walks the strided row groups itself."""
        pk = mlp.packed()
        d = x.shape[-1]
        m = groups * rows_per_group
        if ops.use_h3() and all(pw.h3 is not None for pw in pk):
            # FP16x3: split the source tokens once (shared by the heads that read the same tensor), let the first
            # layer walk the strided row groups, keep the hidden activations in split-half format
            if xs is None:
                xs = ops.split_rows(x.view(-1, d))
            src = ops.SplitRows(xs.buf[first:], d)
            h = None
            for i, pw in enumerate(pk):
                last = i == len(pk) - 1
                act = ops.ACT_NONE if (last and not mlp.is_activation_last) else ops.ACT_RELU
                if i == 0:
                    # split-half output needs tiles that do not straddle two row groups; otherwise fp32 (re-split below)
                    tile_safe = groups == 1 or rows_per_group % 128 == 0
                    h = ops.linear_h3(src, pw.h3, act, split_out=(not last) and tile_safe,
                                      x_batch=(rows_per_group, group_len * xs.ld), m=m)
                    if not last and not tile_safe:
                        h = ops.split_rows(h)
                elif last and pw.n <= ops.NARROW_MAX_N and isinstance(h, ops.SplitRows):
                    # a handful of output columns: the HBM-bound narrow kernel beats a 128 x 256 tensor-core tile
                    h = ops.linear_narrow(h, pw.w, pw.b, act, m=m)
                else:
                    h = ops.linear_h3(h, pw.h3, act, split_out=not last)
            return h if h.is_contiguous() else h.contiguous()
        base = x.data_ptr() + first * d * 4
        h_ptr, ld, batch = base, d, (rows_per_group, group_len * d)
        keep = []
        for i, pw in enumerate(pk):
            last = i == len(pk) - 1
            n_ld = ops.round_up(pw.n, 4)
            alloc = torch.empty if n_ld == pw.n else torch.zeros
            y = alloc(m, n_ld, device=x.device, dtype=torch.float32)
            ops.linear_raw(h_ptr, ld, m, pw, y.data_ptr(), n_ld,
                           ops.ACT_NONE if (last and not mlp.is_activation_last) else ops.ACT_RELU, None,
                           x_batch=batch)
            keep.append(y)
            h_ptr, ld, batch = y.data_ptr(), n_ld, (0, 0)
        y = keep[-1]
        return y if y.shape[1] == pk[-1].n else y[:, :pk[-1].n].contiguous()


def get_model(mode, mano_buffers=None, mano_root="tool/mano_models"):
    """upstream main/model.py:682-766.  `mano_buffers` (dict of th_* tensors) replaces the licensed pkl."""
    backbone_net = BackboneNet(cfg.resnet_type)
    decoder_net = DecoderNet_big() if cfg.use_big_decoder else DecoderNet()
    hand_sdf_decoder = SDFDecoder(latent_size=cfg.hidden_dim, point_feat_size=cfg.PointFeatSize,
                                  use_classifier=cfg.ClassifierBranch)
    obj_sdf_decoder = SDFDecoder(latent_size=cfg.hidden_dim, point_feat_size=cfg.PointFeatSize,
                                 use_classifier=cfg.ClassifierBranch)
    hand_transformer = Transformer(d_model=cfg.hidden_dim, dropout=cfg.dropout, nhead=cfg.nheads,
                                   dim_feedforward=cfg.dim_feedforward, num_encoder_layers=cfg.enc_layers,
                                   num_decoder_layers=cfg.dec_layers, normalize_before=cfg.pre_norm,
                                   return_intermediate_dec=True)
    obj_transformer = VoteTransformer(d_model=cfg.hidden_dim, dropout=cfg.dropout, nhead=cfg.nheads,
                                      dim_feedforward=cfg.dim_feedforward, num_encoder_layers=cfg.enc_layers // 2,
                                      normalize_before=cfg.pre_norm, return_intermediate_dec=True)
    mano_layer = ManoLayer(ncomps=45, center_idx=0, flat_hand_mean=True, side="right", mano_root=mano_root,
                           use_pca=False, buffers=mano_buffers)
    return Model(backbone_net, decoder_net, hand_sdf_decoder, obj_sdf_decoder, hand_transformer, obj_transformer,
                 mano_layer)


def load_checkpoint(model: nn.Module, checkpoint, strict: bool = True, trusted_pickle: bool = False):
    """Load a released / trainer-written snapshot (upstream common/base.py:137-145 writes {"epoch", "network",
    "optimizer", ...} with the state dict of the `DataParallel` wrapper, i.e. every key prefixed `module.`;
    `Tester._make_model`, base.py:179-193, loads it into the wrapper with strict=True).  `checkpoint` is a path, the
    loaded dict, or a bare state dict; the `module.` prefix is stripped when `model` is not itself wrapped.
    Returns the checkpoint dict (epoch etc.) for the caller."""  # noqa: D401
    # weights_only=True: tensors, containers and primitives only -- a snapshot file cannot execute code on load.  Upstream's
    # trainer pickles nothing else ({"epoch", "network", "optimizer"}); `trusted_pickle=True` restores torch.load's legacy
    # behaviour for files that do carry arbitrary Python objects and come from a trusted source
    ckpt = torch.load(checkpoint, map_location="cpu", weights_only=not trusted_pickle) \
        if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__") else checkpoint
    state = ckpt["network"] if isinstance(ckpt, dict) and "network" in ckpt else ckpt
    wrapped = isinstance(model, (nn.DataParallel, nn.parallel.DistributedDataParallel))
    if not wrapped and state and all(k.startswith("module.") for k in state):
        state = {k[len("module."):]: v for k, v in state.items()}
    elif wrapped and state and not any(k.startswith("module.") for k in state):
        state = {"module." + k: v for k, v in state.items()}
    model.load_state_dict(state, strict=strict)
    return ckpt
