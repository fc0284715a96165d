"""N > 1 path on CPU: two gloo ranks run `sharded_forward` (with a stand-in per-sample forward) and must both end up
with the global, correctly ordered result -- including a ragged batch that does not divide by the world size."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_forward(inputs, targets, meta_info, mode):
    assert mode == "eval"
    sid = meta_info["mano_root"][:, 0]                       # sample id smuggled through the meta tensor
    b = sid.shape[0]
    po = 6
    base = sid.view(b, 1, 1)
    return {"hand_joints_out": base + torch.zeros(b, 20, 3), "mano_joints_out": base * 2 + torch.zeros(b, 21, 3),
            "mano_mesh_out": base * 3 + torch.zeros(b, 778, 3), "obj_rot_out": base * 4 + torch.zeros(b, po, 3),
            "obj_trans_out": base * 5 + inputs["img"][:, :1, :1].expand(b, po, 3)}


def _worker(rank, world, port, batch, ok):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hoisdf_b200.dist import sharded_forward
        meta = {"mano_root": torch.arange(batch, dtype=torch.float32).view(batch, 1).repeat(1, 3)}
        inputs = {"img": torch.arange(batch, dtype=torch.float32).view(batch, 1, 1) * 0.5}
        out = sharded_forward(_fake_forward, inputs, {}, meta, 6)
        ids = torch.arange(batch, dtype=torch.float32)
        good = (out["hand_joints_out"].shape == (batch, 20, 3)
                and torch.equal(out["hand_joints_out"][:, 0, 0], ids)
                and torch.equal(out["mano_mesh_out"][:, 777, 2], ids * 3)
                and torch.equal(out["obj_rot_out"][:, 5, 1], ids * 4)
                and torch.equal(out["obj_trans_out"][:, 0, 0], ids * 5 + ids * 0.5))
        ok[rank] = int(good)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(batch):
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, _free_port(), batch, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world


def test_sharded_forward_even_batch():
    _run(6)


def test_sharded_forward_ragged_batch():
    _run(5)


def test_sharded_forward_fewer_samples_than_ranks():
    _run(1)


def _grad_worker(rank, world, port, ok):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hoisdf_b200.train import allreduce_mean_
        flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)          # rank r holds (r + 1) * [0, 1, 2, ...]
        allreduce_mean_(flat)
        ok[rank] = int(torch.equal(flat, torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_mean_two_ranks():
    """The training step's only collective (hoisdf_b200.train.allreduce_mean_): one all-reduce of the flat gradient buffer,
    then the mean -- the global-batch gradient of per-rank batch-mean losses."""
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_grad_worker, args=(world, _free_port(), ok), nprocs=world, join=True)
    assert list(ok) == [1] * world


def test_gradient_allreduce_single_process_is_a_noop():
    from hoisdf_b200.train import allreduce_mean_
    flat = torch.arange(8, dtype=torch.float32)
    assert torch.equal(allreduce_mean_(flat.clone()), flat)
