"""N > 1 on real GPUs: two NCCL ranks run `dist.sharded_forward` with the REAL model, one GPU each, and
the all-gathered `*_out` must equal the single-rank forward of the whole batch bit for bit (every sample is independent
of the batch it travels in -- SURVEY.md section 8(e) -- and the collective only moves bytes).  Skipped with < 2 GPUs;
run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ok):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hoisdf_b200 import synthetic as syn
        from hoisdf_b200.config import cfg
        from hoisdf_b200.dist import sharded_forward
        from hoisdf_b200.model import get_model
        dev = torch.device("cuda", rank)
        cfg.set_setting("dexycb")
        type(cfg).dataset = "ho3d"
        type(cfg).num_samp_hand, type(cfg).num_samp_obj = 96, 40
        seed, B = 9, 5                                        # ragged: 3 + 2 samples
        model = get_model("test", mano_buffers=syn.mano_buffers(seed))
        model.load_state_dict(syn.full_state_dict(seed, "dexycb"), strict=True)
        model = model.to(dev).eval()
        to = lambda d: {k: v.to(dev) for k, v in d.items()}    # noqa: E731
        inputs, targets, meta = to({"img": syn.image_batch(seed, B)}), to(syn.eval_targets(B)), to(syn.camera_meta(seed, B))
        out = sharded_forward(model, inputs, targets, meta, 40)
        whole = model(inputs, targets, meta, "eval")          # every rank: the full batch on its own GPU
        good = all(torch.equal(out[k], whole[k]) for k in out) and out["mano_mesh_out"].shape == (B, 778, 3)
        flag = torch.tensor([int(good)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok[rank] = int(flag.item())
    finally:
        dist.destroy_process_group()


def test_sharded_forward_two_nccl_ranks(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world


def _train_worker(rank, world, port, ok):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hoisdf_b200 import synthetic as syn
        from hoisdf_b200.config import cfg
        from hoisdf_b200.model import get_model
        from hoisdf_b200.train import Trainer
        dev = torch.device("cuda", rank)
        cfg.set_setting("dexycb")
        type(cfg).dataset = "ho3d"
        type(cfg).num_samp_hand, type(cfg).num_samp_obj, type(cfg).dropout = 48, 16, 0.0
        type(cfg).random_move_dist = [0.0, 0.0, 0.0]           # no jitter: the two passes below must see the same points
        seed, B = 9, 2
        model = get_model("train", mano_buffers=syn.mano_buffers(seed))
        model.load_state_dict(syn.full_state_dict(seed, "dexycb"), strict=True)
        model = model.to(dev)
        model.hand_sdf_decoder.dropout_prob = model.obj_sdf_decoder.dropout_prob = 0.0
        to = lambda d: {k: v.to(dev) for k, v in d.items()}    # noqa: E731
        ins, tgt = syn.train_extras(seed + rank, B, 48, 16)     # every rank trains on its OWN samples
        batch = (to({"img": syn.image_batch(seed + rank, B), **ins}), to(tgt), to(syn.camera_meta(seed + rank, B)))
        # capture this rank's LOCAL flat gradient right before the exchange (recomputing it in a second pass would differ
        # by the run-to-run noise of atomics / cuDNN in the backward, measured 2e-4 of the largest gradient)
        import hoisdf_b200.train as T
        captured = {}
        real = T.allreduce_mean_

        def spy(flat, group=None):
            captured["local"] = flat.clone()
            return real(flat, group)

        T.allreduce_mean_ = spy
        tr = Trainer(model, lr=1e-4)
        tr.step(*batch, epoch_cnt=0, batch_ratio=0.0)
        T.allreduce_mean_ = real
        local = captured["local"]
        both = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(both, local)
        want = sum(both) / world
        got = tr.grad
        assert float((both[0] - both[1]).abs().max()) > 0          # the ranks really trained on different samples
        gerr = float((got - want).abs().max()) / float(want.abs().max())
        good = gerr <= 1e-6
        print("rank %d: flat gradient vs mean of local gradients: rel err %.3e (%d values)" % (rank, gerr, got.numel()), flush=True)
        # and the replicas stay identical after the update
        mine = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        theirs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(theirs, mine)
        same = all(torch.equal(t, mine) for t in theirs)
        print("rank %d: replicas identical after the step: %s" % (rank, same), flush=True)
        good = good and same
        flag = torch.tensor([int(good)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok[rank] = int(flag.item())
    finally:
        dist.destroy_process_group()


def test_training_step_two_nccl_ranks(cuda):
    """Data-parallel training step: the flat gradient buffer after Trainer.step is the mean of the two ranks' local gradients
    (one NCCL all-reduce) and the replicas' parameters stay bit-identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_train_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
