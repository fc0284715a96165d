#!/usr/bin/env python
"""Headline benchmark of the hoisdf_b200 hot path (contract: see the task prompt / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # our B200 path (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

Workload = BASELINE.json configs[1]: per GPU a batch of 32 synthetic 256x256 images, 2048 SDF points per
sample (1536 hand + 512 object), `ho3d` architecture (C = 3968), random weights, full
backbone + U-Net + SDF query + pose-decode eval forward.  A "step" is one such forward.  Weak scaling:
every rank owns its own 32 samples, results are all-gathered once per step.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "samples/sec (SDF query + pose decode, 2048 pts, 256² input) @1/2/4/8 B200"
UNIT = "samples/s"
P_HAND, P_OBJ = 1536, 512
ARCH = "ho3d"
WEIGHT_SEED = 0
# --config 3 = BASELINE.json configs[2] (SURVEY.md 8(d) "config 3"): global batch 128, DexYCB-shaped inputs (dexycb arch,
# C = 992, the dataset's eval branch with its SDF supervision points), 4096 points, STRONG scaling over the ranks
CONFIG3 = {"arch": "dexycb", "p_hand": 3072, "p_obj": 1024, "global_batch": 128}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "burst": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json, sustained bf16)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 100 ms.  The query process is started BEFORE the warm-up
    steps (nvidia-smi needs a few hundred ms to produce its first row, and a timed region of ten 35 ms steps is hardly
    longer than that); `begin()` / `end()` bracket the timed region on the host clock and only rows that fall inside it
    are reported.  If the region was too short for a single row, the rows of the warm-up steps that ran right before it
    (the same workload, same load) are reported instead and `window` says so."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []          # (host time, columns)
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        t0 = self.t0 if self.t0 is not None else float("-inf")
        t1 = self.t1 if self.t1 is not None else float("inf")
        valid = [(t, r) for t, r in list(self.rows) if len(r) >= 8]
        rows, window = [r for t, r in valid if t0 <= t <= t1], "timed region"
        if not rows:
            rows, window = [r for t, r in valid if t0 - 1.5 <= t <= t1 + 0.1], "warm-up steps + timed region (region too short)"
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[4:8]) if v.lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def make_inputs(seed: int, batch: int, config: int = 2):
    from hoisdf_b200 import synthetic as syn
    if config == 3:
        extra_in, targets = syn.dexycb_extras(seed, batch, CONFIG3["p_hand"], CONFIG3["p_obj"])
        return {"img": syn.image_batch(seed, batch), **extra_in}, targets, syn.camera_meta(seed, batch)
    return {"img": syn.image_batch(seed, batch)}, syn.eval_targets(batch), syn.camera_meta(seed, batch)


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the upstream algorithm on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_forward_fn(sample_batch: int):
    import torch

    from hoisdf_b200 import synthetic as syn
    from oracle import hoisdf_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    sd = syn.full_state_dict(WEIGHT_SEED, ARCH)
    ocfg = O.default_cfg(num_samp_hand=P_HAND, num_samp_obj=P_OBJ)
    inputs, _, meta = make_inputs(1000, sample_batch)

    def step():
        with torch.no_grad():
            return O.model_eval(sd, inputs["img"], meta, ocfg, ARCH)

    return step, torch.get_num_threads()


def time_cpu(sample_batch: int, steps: int, warmup: int):
    step, threads = cpu_forward_fn(sample_batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps, threads


def time_gpu_eager(dev, sample_batch: int, steps: int, warmup: int):
    """The oracle port (= upstream's algorithm, op for op) with every tensor on the GPU: stock PyTorch eager kernels
    (cuDNN convolutions, cuBLAS sgemm, ATen grid_sampler / sort), per-sample Python loop and CPU bbox mask as upstream."""
    import torch

    from hoisdf_b200 import synthetic as syn
    from oracle import hoisdf_oracle as O

    sd = {k: v.to(dev) for k, v in syn.full_state_dict(WEIGHT_SEED, ARCH).items()}
    ocfg = O.default_cfg(num_samp_hand=P_HAND, num_samp_obj=P_OBJ)
    inputs, _, meta = make_inputs(1000, sample_batch)
    img, meta = inputs["img"].to(dev), {k: v.to(dev) for k, v in meta.items()}

    def step():
        with torch.no_grad():
            return O.model_eval(sd, img, meta, ocfg, ARCH)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": sample_batch * 1000.0 / ms, "unit": UNIT, "ms_per_step": ms, "samples_per_step": sample_batch,
            "kind": "oracle port on the same GPU through stock PyTorch eager (cuDNN / cuBLAS fp32, TF32 off)",
            "sample": "%d samples/step of the configs[1] workload, %d warm-up + %d timed steps" % (sample_batch, warmup, steps)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = 2
    value, sec_per_step, threads = time_cpu(sample_batch, args.steps, args.warmup)
    sample = ("oracle port (oracle/hoisdf_oracle.py, PyTorch-CPU fp32) of the upstream Model.forward(eval): "
              "%d samples per step of the configs[1] workload" % sample_batch)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, sample_batch_note=sample_batch),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, batch=32, sample_batch_note=None, config=2):
    if config == 3:
        return {
            "workload": "BASELINE configs[2]: global batch 128 DexYCB-shaped synthetic inputs (256x256 crops, dexycb arch "
                        "C=992, dataset eval branch incl. SDF supervision points + GT MANO), 4096 SDF points/sample "
                        "(3072 hand + 1024 object), sharded by sample over the ranks (strong scaling)",
            "global_batch": batch * n_gpus, "per_gpu_batch": batch, "points_hand": CONFIG3["p_hand"],
            "points_obj": CONFIG3["p_obj"],
            "parallelism": "sample-sharded x%d, one all-gather of packed results per step" % n_gpus,
            "l2": "no explicit flush: each step streams several GB of activations/weights (>> 126 MB L2)",
            "image_encoder": "ResNet-50 + U-Net on the FP16x3 tensor-core convolution kernels (no cuDNN)",
            "cuda_graphs": "static stages replayed from CUDA graphs (--no-graphs: eager launches)",
        }
    cfg = {
        "workload": "BASELINE configs[1]: batch=32 per GPU, synthetic 256x256 images, 2048 SDF points/sample "
                    "(1536 hand + 512 object), ho3d arch (C=3968), full backbone+U-Net+SDF+decoder eval forward",
        "global_batch": batch * n_gpus, "per_gpu_batch": batch, "points_hand": P_HAND, "points_obj": P_OBJ,
        "parallelism": "sample-sharded x%d, one all-gather of packed results per step" % n_gpus,
        "l2": "no explicit flush: each step streams > 10 GB of activations/weights (>> 126 MB L2), "
              "so nothing of a previous step survives",
        "image_encoder": "ResNet-50 + U-Net on the FP16x3 tensor-core convolution kernels (no cuDNN)",
        "cuda_graphs": "static stages replayed from CUDA graphs (--no-graphs: eager launches)",
    }
    if sample_batch_note is not None:
        cfg["reference_sample_batch"] = sample_batch_note
    return cfg


# ----------------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hoisdf_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from hoisdf_b200.csrc.build import build
    build()
    from hoisdf_b200 import ops, synthetic as syn
    from hoisdf_b200.config import cfg
    from hoisdf_b200.dist import pack_outputs, packed_width
    from hoisdf_b200.model import get_model

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    c3 = args.config == 3
    arch, p_hand, p_obj = (CONFIG3["arch"], CONFIG3["p_hand"], CONFIG3["p_obj"]) if c3 else (ARCH, P_HAND, P_OBJ)
    cfg.set_setting(arch)                       # config 3: also cfg.dataset = "dexycb" (the dataset's eval branch)
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = p_hand, p_obj
    if c3:
        if CONFIG3["global_batch"] % world:
            raise SystemExit("bench.py --config 3: the global batch of 128 must divide by the number of ranks")
        B = CONFIG3["global_batch"] // world    # strong scaling: the global batch is fixed
    else:
        B = args.batch
    model = get_model("test", mano_buffers=syn.mano_buffers(WEIGHT_SEED))
    model.load_state_dict(syn.full_state_dict(WEIGHT_SEED, arch), strict=True)
    # fp32 cuDNN is ~1.6x faster in NCHW than in channels_last on B200 (36 ms vs 58 ms for this batch), and the
    # NCHW->NHWC transposes the gather needs cost 0.3 ms, so the backbone stays NCHW here
    model = model.to(dev).eval()
    if not args.no_graphs:
        # the static-shape stages (image encoder + projection; point features -> transformers -> heads) are replayed from
        # CUDA graphs: same kernels, ~400 fewer host launches per step
        model.enable_cuda_graphs()

    inputs, targets, meta = make_inputs(100 + rank, B, args.config)
    pin = lambda d: {k: v.pin_memory() for k, v in d.items()}  # noqa
    h_inputs, h_targets, h_meta = pin(inputs), pin(targets), pin(meta)
    to_dev = lambda d: {k: v.to(dev, non_blocking=True) for k, v in d.items()}  # noqa
    d_inputs, d_targets, d_meta = to_dev(h_inputs), to_dev(h_targets), to_dev(h_meta)
    width = packed_width(p_obj)
    gathered = torch.empty(world * B, width, device=dev) if world > 1 else None
    h_result = torch.empty(B, width).pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for d in (h_inputs, h_targets, h_meta) for v in d.values())
    d2h_bytes = h_result.numel() * 4

    def step_device():
        out = model(d_inputs, d_targets, d_meta, "eval")
        packed = pack_outputs(out, p_obj)
        if world > 1:
            dist.all_gather_into_tensor(gathered, packed)
        return packed

    def step_e2e():
        di, dt, dm = to_dev(h_inputs), to_dev(h_targets), to_dev(h_meta)
        out = model(di, dt, dm, "eval")
        packed = pack_outputs(out, p_obj)
        if world > 1:
            dist.all_gather_into_tensor(gathered, packed)
        h_result.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller consumes the result of every step

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        fence()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.profile_step:
        sampler.start()                      # before the warm-up: see ClockSampler
    for _ in range(max(args.warmup, 3)):
        step_device()
    fence()
    if args.profile_step:
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    n_f = None
    if model.last_taps is not None:
        n_f = (model.last_taps["hand"]["n_f"].double().mean().item(), model.last_taps["obj"]["n_f"].double().mean().item())

    sampler.begin()
    ms_total = timed(step_device, args.steps)
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B * 1000.0 / ms_step

    for _ in range(3):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e = {"value": world * B * 1000.0 / ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes}

    # the hot path alone (SURVEY.md 8(d) "Metric" (ii): sdf_infer x2 ... hand_joints, i.e. everything after the U-Net),
    # on the pyramid the image encoder produced for these inputs; eager launches
    hot_path_only = None
    if not c3:
        with torch.no_grad():
            pyr, _ = model.run_image_encoder(d_inputs["img"])
        for _ in range(2):
            model.hot_path(pyr, d_meta)
        ms_hot = timed(lambda: model.hot_path(pyr, d_meta), args.steps) / args.steps
        hot_path_only = {"value": world * B * 1000.0 / ms_hot, "unit": UNIT, "ms_per_step": ms_hot,
                         "note": "Model.hot_path (everything after the U-Net) on a resident pyramid, eager launches"}
    allgather_ms = None
    if world > 1:                               # the collective alone (one packed block per rank)
        blk = torch.zeros(B, width, device=dev)
        for _ in range(3):
            dist.all_gather_into_tensor(gathered, blk)
        allgather_ms = timed(lambda: dist.all_gather_into_tensor(gathered, blk), 20) / 20

    # roofline pass: the same steps once more, launched eagerly (no graph replay) with a CUDA-event pair on the launching
    # stream around every launch of the dominant kernel; `share_of_step` is relative to this pass's own step time
    model.enable_cuda_graphs(False)
    step_device()
    fence()
    ops.STATS["launches"] = 0
    ops.PROFILE = []
    ms_total = timed(step_device, args.steps)
    prof, ops.PROFILE = ops.PROFILE, None
    launches = ops.STATS["launches"]      # our kernels per `steps` forwards (graph replay launches the same nodes)

    # dominant kernel: the tcgen05 3xTF32 Linear kernel (stand-alone launches + the four inside every SDF-decoder
    # call); every such launch of the timed steps was bracketed by CUDA events on the launching stream
    torch.cuda.synchronize()
    tc = [p for p in prof if p[0] in ("linear_tc", "sdf_decoder", "linear_h3", "sdf_decoder_h3", "conv_h3")]
    chain = [p for p in prof if p[0] == "sdf_chain"]
    h3 = any(p[0] in ("linear_h3", "sdf_decoder_h3") for p in tc)
    fma = [p for p in prof if p[0] == "linear"]
    tc_flops = sum(p[1] for p in tc)
    tc_ms = sum(p[2].elapsed_time(p[3]) for p in tc)
    fma_flops = sum(p[1] for p in fma)
    fma_ms = sum(p[2].elapsed_time(p[3]) for p in fma)
    peaks = load_peaks()
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    kernel_name = ("hoisdf::linear_h3_kernel (tcgen05.mma kind::f16 on split-half operands, 3 products per fp32-grade "
                   "product, 1 for the candidate pre-screening; persistent, chunked TMEM accumulation drained into fp32 "
                   "registers; ALL its launches of the step: ResNet-50 + U-Net implicit-GEMM convolutions, pyramid "
                   "projection, candidate / point MLPs incl. the 4 GEMMs of every SDF-decoder call, transformer and "
                   "head linears; FLOPs counted once per product of the reference's fp32 arithmetic)") if h3 else (
        "hoisdf::linear_tf32x3_kernel (tcgen05.mma kind::tf32, 3-pass split = fp32-grade; all launches of the "
        "step incl. the 4 GEMMs of every SDF-decoder call; FLOPs counted once per fp32 product, i.e. the "
        "tensor cores execute 3x this number of TF32 MACs)")
    traffic, traffic_src = None, None
    import glob
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_h3_traffic.json")))
    if h3 and tpaths:            # newest committed ncu launch list of `bench.py --profile-step` (ncu cannot run inside a timed run)
        tpath = tpaths[-1]
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]
        traffic_src = ("profiles/%s (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of `bench.py "
                       "--profile-step`, the same workload)" % os.path.basename(tpath))
    chain_ms = sum(p[2].elapsed_time(p[3]) for p in chain)
    chain_flops = sum(p[1] for p in chain)
    chain_rows = sum(int(p[4].split("rows=")[1].split()[0]) for p in chain)
    burst = peaks.get("burst", peaks["tflops"])
    chain_stats = None
    if chain_ms > 0:
        chain_stats = {
            "kernel": "hoisdf::sdf_chain_kernel (ONE persistent tcgen05 kernel per field for the candidate chain: "
                      "linear_sdfin.1 -> NeRF embedding -> linh0..linh4 -> tanh; single-product fp16 screening "
                      "arithmetic, activations in shared / tensor memory, 4 B per row written)",
            "bound": "tensor", "achieved": chain_flops / (chain_ms * 1e-3) / 1e12, "peak": peaks["tflops"],
            "unit": "TFLOP/s", "frac": chain_flops / (chain_ms * 1e-3) / 1e12 / peaks["tflops"],
            "launches_per_step": len(chain) / args.steps, "share_of_step": chain_ms / ms_total,
            "rows_per_step": chain_rows / args.steps, "ms_per_step": chain_ms / args.steps,
            "algorithmic_bytes_per_row": 1028,
            "note": "FLOPs = executed tensor-core FLOPs (one product per K step); algorithmic HBM bytes per row: "
                    "1024 in (512 fp16) + 4 out",
        }
    roofline = {
        "kernel": kernel_name,
        "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["tflops"], "traffic": traffic, "traffic_unit": "bytes (dram read + write) per launch, "
        "average over the kernel's launches of one step", "traffic_source": traffic_src, "peak_source": peaks["source"],
        "launches_per_step": len(tc) / args.steps, "share_of_step": tc_ms / ms_total,
        "eager_pass_ms_per_step": ms_total / args.steps,
        "algorithmic_flops_per_step": tc_flops / args.steps,
        "note": "fp32-grade accuracy costs 3 tensor-core products per algorithmic product (1 in the pre-screening "
                "launches), so frac tops out near 1/3 of the fp16 peak for the three-product launches",
        "fp32_fma_linear": {"achieved_tflops": fma_flops / (fma_ms * 1e-3) / 1e12 if fma_ms > 0 else 0.0,
                            "share_of_step": fma_ms / ms_total, "launches_per_step": len(fma) / args.steps,
                            "fp32_fma_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12},
        "candidate_chain": chain_stats,
    }

    cpu_baseline = gpu_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not c3:
        v, sec, threads = time_cpu(4, 4, 1)                 # ~10-15 s of host work on the box's cores
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "oracle port of the upstream eval forward, 4 samples/step of the same workload, "
                                  "1 warm-up + 4 timed steps (%.1f s/step)" % sec}
        # "the existing GPU implementation" (BASELINE.md section 4): the same reference algorithm through stock PyTorch
        # eager on this B200 (cuDNN / cuBLAS fp32, TF32 off) -- upstream ships no CUDA of its own
        torch.cuda.empty_cache()
        gpu_eager = time_gpu_eager(dev, 8, 3, 1)
    if rank == 0:
        conf = workload_config(world, B, config=args.config)
        if n_f is not None:
            conf["mean_candidates_per_sample"] = {"hand": n_f[0], "obj": n_f[1]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": conf, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager, "hot_path_only": hot_path_only,
        }
        if c3:
            line["scaling"] = "strong"
        if allgather_ms is not None:
            line["allgather_ms"] = allgather_ms
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# --config 4 = BASELINE.json configs[3]: HO3Dv2-shaped TRAINING step (forward + backward + AdamW), batch 64 per GPU
# ----------------------------------------------------------------------------------------------------
CONFIG4 = {"arch": "ho3d", "p_hand": 600, "p_obj": 200, "batch": 64}


def run_train(args):
    """One step = hoisdf_b200.train.Trainer.step: Model.forward(mode="train") (upstream main/model.py:357-665, the
    `*_pre_points` branch of the first epochs, dropout as configured upstream) -> weighted loss sum -> backward -> AdamW.
    Weak scaling: every rank trains on its own 64 samples and the flat gradient buffer is all-reduced once per step."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hoisdf_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from hoisdf_b200.csrc.build import build
    build()
    from hoisdf_b200 import ops, synthetic as syn
    from hoisdf_b200.config import cfg
    from hoisdf_b200.model import get_model
    from hoisdf_b200.train import Trainer

    # the image encoder trains through cuDNN (hoisdf_b200/train.py): PyTorch's default conv setting (TF32 allowed), as the
    # stock upstream code would run on this GPU; every Linear / attention / gather of the hot path is ours (fp32-grade)
    torch.backends.cudnn.benchmark = True
    arch, ph, po = CONFIG4["arch"], CONFIG4["p_hand"], CONFIG4["p_obj"]
    B = args.batch if args.batch != 32 else CONFIG4["batch"]
    cfg.set_setting(arch)
    type(cfg).num_samp_hand, type(cfg).num_samp_obj = ph, po
    model = get_model("train", mano_buffers=syn.mano_buffers(WEIGHT_SEED))
    model.load_state_dict(syn.full_state_dict(WEIGHT_SEED, arch), strict=True)
    model = model.to(dev).train()
    trainer = Trainer(model, lr=1e-4)
    ins, tgt = syn.train_extras(100 + rank, B, ph, po)
    h = [{k: v.pin_memory() for k, v in d.items()} for d in ({"img": syn.image_batch(100 + rank, B), **ins}, tgt,
                                                             syn.camera_meta(100 + rank, B))]
    to_dev = lambda d: {k: v.to(dev, non_blocking=True) for k, v in d.items()}  # noqa: E731
    d_in, d_tgt, d_meta = (to_dev(d) for d in h)
    h2d = sum(v.numel() * v.element_size() for d in h for v in d.values())
    h_loss = torch.empty(1).pin_memory()

    def step_device():
        return trainer.step(d_in, d_tgt, d_meta, epoch_cnt=0, batch_ratio=0.0)[0]

    def step_e2e():
        total = trainer.step(to_dev(h[0]), to_dev(h[1]), to_dev(h[2]), epoch_cnt=0, batch_ratio=0.0)[0]
        h_loss.copy_(total.view(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        fence()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    fence()
    if args.profile_step:
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    sampler.begin()
    ops.STATS["launches"] = 0
    ms = timed(step_device, args.steps)
    launches = ops.STATS["launches"]
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(step_e2e, args.steps)
    # split of one step (events around the three phases, one extra step)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    from hoisdf_b200.train import total_loss
    trainer.zero_grad()
    ev[0].record()
    out = model(d_in, d_tgt, d_meta, "train", 0, 0.0)
    total, _ = total_loss(out)
    ev[1].record()
    total.backward()
    ev[2].record()
    torch.cuda.synchronize()
    phases = {"forward_ms": ev[0].elapsed_time(ev[1]), "backward_ms": ev[1].elapsed_time(ev[2])}
    # roofline of the dominant kernel of OURS in the step: every FP16x3 GEMM launch (forward, dX, split-K dW) timed with an
    # event pair on the launching stream during one extra step; FLOPs = 2 M N K of the fp32 product each one stands for
    roofline = None
    ops.PROFILE = []
    step_device()                       # EVERY rank takes this step: it contains the gradient all-reduce
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    if rank == 0:
        gemm = [r for r in rec if r[0] == "linear_h3"]
        if gemm:
            g_ms = sum(r[2].elapsed_time(r[3]) for r in gemm)
            g_fl = sum(r[1] for r in gemm)
            peaks = load_peaks()
            roofline = {"kernel": "hoisdf::linear_h3_kernel, all its launches of one training step (forward, dX = dZ.W, "
                                  "dW^T = X^T.dZ with split-K over the idle SMs; 3 tensor-core products per fp32-grade product)",
                        "bound": "tensor", "achieved": g_fl / g_ms / 1e9, "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": g_fl / g_ms / 1e9 / peaks["tflops"], "traffic": None, "peak_source": peaks["source"],
                        "launches_per_step": len(gemm), "ms_per_step": g_ms, "share_of_step": g_ms / (ms / args.steps),
                        "note": "per-launch event timing serialises nothing (same stream) but adds event overhead to "
                                "~330 short launches; frac tops out near 1/3 for three-product arithmetic"}
    eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.cuda.empty_cache()
        try:                                   # the same batch as ours; a smaller one if the eager graph does not fit
            eager = time_train_eager(dev, arch, B, ph, po)
        except RuntimeError:
            torch.cuda.empty_cache()
            eager = time_train_eager(dev, arch, 8, ph, po)
    if rank == 0:
        value = world * B * args.steps * 1000.0 / ms
        line = {
            "metric": "training samples/sec (forward + backward + AdamW, HO3Dv2-shaped batch, 600+200 points)",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: HO3Dv2-shaped training step (forward + backward, SDF L1 + heat-map / "
                                   "segmentation + joint-vote + MANO + object-pose losses, AdamW), batch=%d per GPU, ho3d arch "
                                   "(C=3968), 600 + 200 SDF supervision points and 600 + 200 pose points per sample "
                                   "(`*_pre_points` branch of the first epochs), dropout as upstream (0.1 / 0.2)" % B,
                       "global_batch": B * world, "per_gpu_batch": B,
                       "parallelism": "data parallel x%d, one all-reduce of the flat gradient buffer per step" % world,
                       "image_encoder": "ResNet-50 + U-Net forward / backward through cuDNN under PyTorch autograd (TF32 "
                                        "convolutions: PyTorch's default); the hot path after the pyramid on the hoisdf_b200 "
                                        "kernels (hoisdf_b200/autograd.py)",
                       "l2": "no explicit flush: a step streams tens of GB"},
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps * 1000.0 / ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "phases": phases, "gpu_eager_baseline": eager,
            "roofline": roofline, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_train_eager(dev, arch, batch, ph, po):
    """The same training step through stock PyTorch eager on this GPU: the oracle port (upstream's algorithm op for op:
    cuDNN, cuBLAS sgemm, ATen grid_sampler, autograd) + torch.optim.AdamW.  Dropout off (the port has none)."""
    import torch

    from hoisdf_b200 import synthetic as syn
    from oracle import hoisdf_oracle as O
    sd = syn.full_state_dict(WEIGHT_SEED, arch)
    p = {k: v.clone().to(dev) for k, v in sd.items()}
    names = [k for k, v in p.items() if v.is_floating_point() and "running_" not in k and "th_" not in k
             and "num_batches" not in k and "coord_change" not in k]
    for n in names:
        p[n].requires_grad_(True)
    opt = torch.optim.AdamW([p[n] for n in names], lr=1e-4)
    mv = lambda d: {k: v.to(dev) for k, v in d.items()}  # noqa: E731
    ins, tgt = syn.train_extras(1000, batch, ph, po)
    img, ins, tgt, meta = syn.image_batch(1000, batch).to(dev), mv(ins), mv(tgt), mv(syn.camera_meta(1000, batch))
    ocfg = O.default_cfg(num_samp_hand=ph, num_samp_obj=po)

    def step():
        opt.zero_grad(set_to_none=True)
        total, _ = O.train_total_loss(O.model_train(p, img, ins, tgt, meta, ocfg, arch))
        total.backward()
        opt.step()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    return {"value": batch * 1000.0 / ms, "unit": UNIT, "ms_per_step": ms, "samples_per_step": batch,
            "kind": "oracle port of the training step on the same GPU through stock PyTorch eager + torch.optim.AdamW "
                    "(cuDNN TF32 convolutions as PyTorch defaults, cuBLAS fp32 matmul)",
            "sample": "%d samples/step, 2 warm-up + 3 timed steps" % batch}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="samples per GPU per step (configs[1]: 32)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4],
                    help="2 = BASELINE configs[1] (default, the metric's config); 3 = configs[2]: global batch 128, dexycb, "
                         "4096 points, strong scaling; 4 = configs[3]: training step, batch 64 per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly (no CUDA-graph replay)")
    ap.add_argument("--profile-step", action="store_true",
                    help="for ncu --profile-from-start off: warm up, bracket ONE step with cudaProfilerStart/Stop, exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == 4:
        run_train(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
