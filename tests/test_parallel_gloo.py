"""CPU, world_size 2, gloo: the host logic of the N > 1 path (sharding, gradient mean all-reduce, max-over-ranks)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mde_biological_vision_systems_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
        data = torch.randn(10, 8)
        target = torch.randn(10, 1)
        idx = parallel.shard_indices(10, rank, world)
        loss = torch.nn.functional.mse_loss(model(data[idx]), target[idx])
        loss.backward()
        n = parallel.GradientAverager(model.parameters(), bucket_mb=0.0001).reduce()
        assert n == sum(p.numel() for p in model.parameters())
        torch.save([p.grad.clone() for p in model.parameters()], os.path.join(tmp, f"g{rank}.pt"))
        t = parallel.max_over_ranks(10.0 + rank)
        assert t == 10.0 + world - 1
        # --- the overlapped form: averager built first (gradient arena + hooks), replicas start from DIFFERENT weights and
        # are synchronised by broadcast_module_state; a parameter that is unused on one rank only still reduces (as zeros)
        torch.manual_seed(100 + rank)
        net = torch.nn.ModuleDict({"a": torch.nn.Linear(8, 16), "b": torch.nn.Linear(16, 1), "side": torch.nn.Linear(8, 1)})
        parallel.broadcast_module_state(net)
        torch.save({k: v.clone() for k, v in net.state_dict().items()}, os.path.join(tmp, f"w{rank}.pt"))
        avg = parallel.GradientAverager(net.parameters(), bucket_mb=0.0001)
        assert len(avg.buckets) > 1
        for step in range(2):
            avg.zero_grad()
            out = net["b"](torch.relu(net["a"](data[idx])))
            if rank == 0:  # 'side' gets a gradient on rank 0 only
                out = out + net["side"](data[idx])
            torch.nn.functional.mse_loss(out, target[idx]).backward()
            avg.reduce()
            assert all(p.grad is not None and p.grad.data_ptr() >= 0 for p in net.parameters())
        torch.save({k: p.grad.clone() for k, p in net.named_parameters()}, os.path.join(tmp, f"h{rank}.pt"))
        # --- the form a replayed CUDA graph uses (training.GraphedTrainStep): no hooks, one bucket, the arena is filled "on the
        # device" (here: by hand) and reduce() is called step after step WITHOUT a host-side zero_grad() in between
        one = parallel.GradientAverager(net.parameters(), bucket_mb=1 << 20, overlap=False)
        assert len(one.buckets) == 1 and not one._hooks
        for step in range(3):
            one.buckets[0]["flat"].fill_(float(rank + 1 + step))
            assert one.reduce() == one.buckets[0]["flat"].numel()
            assert torch.all(one.buckets[0]["flat"] == (1 + 2) / 2 + step)
    finally:
        dist.destroy_process_group()


def test_shard_indices():
    for n, world in [(10, 2), (7, 2), (16, 8), (3, 4)]:
        shards = [parallel.shard_indices(n, r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1
        assert set(sum(shards, [])) == set(range(n))
        nopad = [parallel.shard_indices(n, r, world, pad=False) for r in range(world)]
        assert sorted(sum(nopad, [])) == list(range(n))
    assert parallel.per_rank_batch(16, 8) == 2 and parallel.per_rank_batch(16, 8, use_new_batching=True) == 16


def test_gradient_mean_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(os.path.join(tmp_path, "g0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "g1.pt"))
    # reference: mean over ranks of the per-rank gradients, computed serially
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    data = torch.randn(10, 8)
    target = torch.randn(10, 1)
    grads = []
    for r in range(world):
        model.zero_grad()
        idx = parallel.shard_indices(10, r, world)
        torch.nn.functional.mse_loss(model(data[idx]), target[idx]).backward()
        grads.append([p.grad.clone() for p in model.parameters()])
    for a, b, x, y in zip(g0, g1, grads[0], grads[1]):
        torch.testing.assert_close(a, b)
        torch.testing.assert_close(a, (x + y) / 2)
    # overlapped form: identical start weights on both ranks (rank 0's), identical averaged gradients, and the one-sided
    # 'side' parameter averaged against zeros
    w0, w1 = torch.load(os.path.join(tmp_path, "w0.pt")), torch.load(os.path.join(tmp_path, "w1.pt"))
    assert all(torch.equal(w0[k], w1[k]) for k in w0)
    h0, h1 = torch.load(os.path.join(tmp_path, "h0.pt")), torch.load(os.path.join(tmp_path, "h1.pt"))
    assert all(torch.equal(h0[k], h1[k]) for k in h0)
    net = torch.nn.ModuleDict({"a": torch.nn.Linear(8, 16), "b": torch.nn.Linear(16, 1), "side": torch.nn.Linear(8, 1)})
    net.load_state_dict(w0)
    ref = {k: torch.zeros_like(p) for k, p in net.named_parameters()}
    for r in range(world):
        net.zero_grad()
        idx = parallel.shard_indices(10, r, world)
        out = net["b"](torch.relu(net["a"](data[idx])))
        if r == 0:
            out = out + net["side"](data[idx])
        torch.nn.functional.mse_loss(out, target[idx]).backward()
        for k, p in net.named_parameters():
            if p.grad is not None:
                ref[k] += p.grad / world
    for k in ref:
        torch.testing.assert_close(h0[k], ref[k])


def test_convert_sync_batchnorm_keeps_state_dict():
    """parallel.convert_sync_batchnorm swaps BatchNorm2d modules for the kernel-backed SyncBatchNorm2d without touching
    parameter identity or state_dict keys (train.py:296 semantics); eval mode on CPU is plain batch norm."""
    import torch
    from mde_biological_vision_systems_b200 import parallel
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                              torch.nn.Sequential(torch.nn.Conv2d(8, 6, 1), torch.nn.BatchNorm2d(6)))
    keys = list(net.state_dict().keys())
    w = net[1].weight
    conv = parallel.convert_sync_batchnorm(net)  # no process group here: no peer-memory arena is set up
    assert list(conv.state_dict().keys()) == keys
    assert isinstance(conv[1], parallel.SyncBatchNorm2d) and conv[1].weight is w
    assert isinstance(conv[3][1], torch.nn.SyncBatchNorm)  # 6 channels: not a multiple of 4 -> torch's module
    conv.eval()
    x = torch.randn(2, 3, 9, 9)
    ref = torch.nn.functional.batch_norm(net[0](x), w.new_zeros(8), w.new_ones(8), w, net[1].bias, False, 0.1, 1e-5)
    assert torch.allclose(conv[1](net[0](x)), ref, atol=1e-6)


def test_p2p_region_layout_and_epochs():
    """Bookkeeping of the peer-memory statistics exchange: per module [direction][epoch parity] slot sets and flag rows
    never overlap, epochs increase per direction, consecutive calls alternate parity (double buffering)."""
    from mde_biological_vision_systems_b200.parallel import _P2PRegion

    class Arena:
        world, rank = 4, 1

    c = 96
    nbytes = _P2PRegion.nbytes(Arena.world, c)
    assert nbytes % 256 == 0 and nbytes >= 4 * Arena.world * 2 * c * 8 + 4 * Arena.world * 8
    reg = _P2PRegion(Arena(), 1024, c)
    seen = {}
    for direction in (0, 1):
        for call in range(1, 5):
            arena, epoch, slot_off, flag_off = reg.next(direction)
            assert epoch == call
            seen[(direction, epoch & 1)] = (slot_off, flag_off)
            assert seen[(direction, epoch & 1)] == (slot_off, flag_off)           # same parity -> same buffers
            assert 1024 <= slot_off and slot_off + Arena.world * 2 * c * 8 <= 1024 + nbytes
            assert flag_off + Arena.world * 8 <= 1024 + nbytes
    slots = sorted(v[0] for v in seen.values())
    flags = sorted(v[1] for v in seen.values())
    assert len(set(slots)) == 4 and len(set(flags)) == 4
    assert all(b - a >= Arena.world * 2 * c * 8 for a, b in zip(slots, slots[1:]))   # slot sets do not overlap
    assert all(b - a >= Arena.world * 8 for a, b in zip(flags, flags[1:])) and flags[0] >= slots[-1] + Arena.world * 2 * c * 8
    # captured launches (training.GraphedTrainStep): the region hands out an epoch BASE and the parity-0 offsets; what the kernel
    # derives from them and the device step counter (csrc/bn_sync.cu bn_resolve_epoch) continues the eager sequence exactly
    eager = _P2PRegion(Arena(), 1024, c)
    for _ in range(3):   # three eager warm-up iterations
        eager.next(0), eager.next(1)
    cap = _P2PRegion(Arena(), 1024, c)
    cap.epochs = list(eager.epochs)
    arena = cap.arena
    arena.capture_base, arena.graph_owned = 0 + 1, False   # replay counter 0 at capture
    frozen = [cap.next(d)[1:] for d in (0, 1)]
    arena.capture_base, arena.graph_owned = None, True
    for replay in range(1, 6):
        for d in (0, 1):
            _, e_ref, slot_ref, flag_ref = eager.next(d)
            base, slot0, flag0 = frozen[d]
            e = base + replay
            par = e & 1
            assert (e, slot0 + par * Arena.world * 2 * c * 8, flag0 + par * Arena.world * 8) == (e_ref, slot_ref, flag_ref)
    with pytest.raises(RuntimeError):
        cap.next(0)   # eager call after the capture: the epochs belong to the graph
