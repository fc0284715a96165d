"""CUDA-event timing of the test-time metric kernels (csrc/metrics.cu) at the bench shape -- 32 samples, 1000-vertex
object templates, 512 pose votes per sample (configs[1]'s P_o), 21 joints -- next to the upstream formulation
(common/metrics.py:110-185,231-248 as restated in oracle/hoisdf_oracle.py) evaluated by PyTorch on the SAME GPU
(object metrics) and on the host (the per-sample numpy SVD loop upstream runs).  Developer tool, not a bench line."""
import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hoisdf_b200 import metrics as M, ops, synthetic as syn
from oracle import hoisdf_oracle as O

dev = torch.device("cuda:0")
B, N, votes = 32, 1000, 512
m = syn.metric_inputs(3, B, votes=votes, n_templates=21, n_verts=N)
templates = torch.stack([t["verts"] for t in m["templates"]]).to(dev)
ids = (m["obj_cls_ids"] - 1).to(dev)
args = [m["out"]["obj_rot"].to(dev), m["out"]["obj_trans"].to(dev), m["targets"]["obj_rot"].to(dev),
        m["targets"]["rel_obj_trans"].to(dev)]
jp, jg = m["joints_pred"].to(dev), m["joints_gt"].to(dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record()
    for _ in range(reps):
        fn()
    e[1].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]) / reps


ours = timed(lambda: ops.obj_pose_metrics(templates, ids, *args))
torch_gpu = timed(lambda: O.obj_pose_metrics(templates, ids, *args))
joints = timed(lambda: ops.hand_joint_metrics(jp, jg))
t0 = time.perf_counter()
for _ in range(5):
    O.hand_joint_metrics(jp.cpu(), jg.cpu())          # upstream: .cpu() per sample + numpy SVD (metrics.py:236-241)
host_joints = (time.perf_counter() - t0) / 5 * 1e3
a = ops.obj_pose_metrics(templates, ids, *args)
b = O.obj_pose_metrics(templates, ids, *args)
err = max(float((x - y).abs().max()) for x, y in zip(a, b))
flops = 8.0 * N * N * B                                # 3 sub, 3 mul-add, 1 min per vertex pair
res = {"shape": {"samples": B, "verts": N, "votes": votes},
       "obj_metrics_ms": ours, "obj_metrics_gflops": flops / ours / 1e6,
       "obj_metrics_torch_same_gpu_ms": torch_gpu, "speedup_vs_torch_gpu": torch_gpu / ours,
       "hand_joint_metrics_ms": joints, "hand_joint_host_numpy_ms": host_joints,
       "max_abs_diff_vs_torch_gpu": err}
print(json.dumps(res))
