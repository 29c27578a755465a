"""CPU, world_size 2 over gloo: the host-side logic of the N>1 paths (segment sharding + overlap-add exchange,
bucketed gradient averaging).  The CUDA model has no CPU fallback, so a deterministic stand-in ``model_fn`` plays
its part here; the sharding / exchange code under test is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tfswa_unet_b200.parallel import shard_range


def test_shard_range_matches_survey_example():
    parts = [shard_range(133, 8, r) for r in range(8)]
    assert [b - a for a, b in parts] == [17, 17, 17, 17, 17, 16, 16, 16]
    assert parts[0][0] == 0 and parts[-1][1] == 133
    assert all(parts[i][1] == parts[i + 1][0] for i in range(7))
    assert shard_range(3, 8, 7) == (3, 3)           # more ranks than items: empty shards are legal


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_model(x):
    # deterministic, segment-local stand-in for TFSWAUNet: (b,2,F,T) -> (b,2,F,T) "masks"
    return torch.sigmoid(0.3 * x + 0.1 * x.flip(1))


def _ola_worker(rank, world, port, audio, ref, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tfswa_unet_b200.separate import ShardedSeparator
    sep = ShardedSeparator(_fake_model, n_fft=256, hop_length=64, sample_rate=8000, segment_length=0.5, overlap=0.25, batch=3)
    out = sep.separate(audio)
    got = torch.cat([out["vocals"], out["other"]])
    q.put((rank, float((got - ref).abs().max())))
    dist.destroy_process_group()


def test_sharded_overlap_add_equals_sequential_reference_world2():
    from oracle.ola_oracle import separate_long
    torch.manual_seed(0)
    audio = 0.1 * torch.randn(2, 8000 * 4 + 123)        # 4.015 s stereo at 8 kHz -> 10 segments of 0.5 s, 25 % overlap
    ref = separate_long(audio, _fake_model, n_fft=256, hop=64, sr=8000, segment_length=0.5, overlap=0.25)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ola_worker, args=(r, 2, port, audio, ref, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, err in res:
        assert err < 1e-5, f"rank {rank}: sharded OLA differs from the sequential oracle by {err}"


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tfswa_unet_b200.parallel import GradAllReducer
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    red = GradAllReducer(model, bucket_bytes=256)        # tiny buckets -> several async all-reduces
    assert len(red.buckets) > 1
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(6, 8, generator=g)
    # rank-local step on this rank's half of the batch
    xs = x_all[rank * 3:(rank + 1) * 3]
    model(xs).pow(2).mean().backward()
    red.finish()
    local = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    # single-process reference: the mean over the two half-batch losses
    ref_model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    ref_model.load_state_dict(model.state_dict())
    (0.5 * ref_model(x_all[:3]).pow(2).mean() + 0.5 * ref_model(x_all[3:]).pow(2).mean()).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
    q.put((rank, float((local - ref).abs().max())))
    # a second step must start from clean bucket state
    model.zero_grad()
    model(xs).pow(2).mean().backward()
    red.finish()
    again = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    q.put((rank, float((again - ref).abs().max())))
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(4)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, err in res:
        assert err < 1e-6, f"rank {rank}: averaged gradients differ from the single-process reference by {err}"


# ---- flat arena (train_step.py): in-place bucketed exchange of contiguous gradient slices ------------------
def _tiny_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 5), torch.nn.LayerNorm(5))


def test_flat_arena_single_process_views_and_accumulation():
    from tfswa_unet_b200.train_step import FlatArena
    model, ref = _tiny_model(), _tiny_model()
    sd_before = {k: v.clone() for k, v in model.state_dict().items()}
    arena = FlatArena(model, bucket_bytes=128)
    assert len(arena.buckets) > 1 and arena.buckets[0][1] == arena.numel and arena.buckets[-1][0] == 0
    assert all(hi == arena.buckets[i - 1][0] for i, (lo, hi, _) in enumerate(arena.buckets) if i)     # contiguous, end first
    for k, v in model.state_dict().items():                      # re-homing keeps values, names and shapes
        assert torch.equal(v, sd_before[k])
    for p, o in zip(arena.params, arena.offsets):
        assert o % 128 == 0 and p.data_ptr() == arena.flat_p[o:].data_ptr() and p.grad.data_ptr() == arena.flat_g[o:].data_ptr()
    x = torch.randn(4, 8)
    for _ in range(2):                                           # second round: zero_grad is one memset, views survive
        arena.zero_grad()
        ref.zero_grad()
        model(x).pow(2).mean().backward()
        ref(x).pow(2).mean().backward()
        assert arena.finish() == 1.0
        for p, o, q in zip(arena.params, arena.offsets, ref.parameters()):
            assert p.grad.data_ptr() == arena.flat_g[o:].data_ptr(), "autograd must accumulate into the arena in place"
            assert torch.allclose(p.grad, q.grad, atol=1e-7)
    # padding between parameters stays zero
    mask = torch.ones(arena.numel, dtype=torch.bool)
    for p, o in zip(arena.params, arena.offsets):
        mask[o:o + p.numel()] = False
    assert float(arena.flat_g[mask].abs().max()) == 0.0 and float(arena.flat_p[mask].abs().max()) == 0.0
    # a stock zero_grad(set_to_none=True) detaches the views; adopt_grads brings the gradients back
    model.zero_grad(set_to_none=True)
    model(x).pow(2).mean().backward()
    arena.adopt_grads()
    for p, o, q in zip(arena.params, arena.offsets, ref.parameters()):
        assert p.grad.data_ptr() == arena.flat_g[o:].data_ptr() and torch.allclose(p.grad, q.grad, atol=1e-7)
    # in-place update of the arena is what the parameters see
    arena.flat_p.mul_(0.5)
    assert torch.allclose(model[0].weight, 0.5 * sd_before["0.weight"])


def test_fused_optimizer_refuses_cpu():
    from tfswa_unet_b200.train_step import FlatArena, FusedClipAdamW
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FusedClipAdamW(FlatArena(_tiny_model()))


def test_cosine_lr_matches_torch_scheduler():
    from tfswa_unet_b200.train_step import cosine_lr
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=50, eta_min=1e-6)
    for step in range(50):
        assert abs(cosine_lr(step, 50, 1e-3) - opt.param_groups[0]["lr"]) < 1e-9
        opt.step()
        sch.step()


def test_masked_magnitude_l1_matches_trainer_formula():
    from tfswa_unet_b200.train_step import masked_magnitude_l1
    torch.manual_seed(3)
    out, mix = torch.randn(2, 4, 9, 7), torch.rand(2, 9, 7)
    tg = [torch.rand(2, 9, 7), torch.rand(2, 9, 7)]
    ref = 0.0
    for i in range(2):                                           # trainer.py:176-186 + losses.py:265-273, restated
        sm = out[:, i * 2:(i + 1) * 2]
        mm = torch.sigmoid(torch.sqrt(sm[:, 0] ** 2 + sm[:, 1] ** 2 + 1e-8))
        ref = ref + torch.nn.functional.l1_loss(mix * mm, tg[i])
    assert abs(float(masked_magnitude_l1(out, mix, tg)) - float(ref / 2)) < 1e-7


def _arena_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tfswa_unet_b200.train_step import FlatArena
    model = _tiny_model()
    arena = FlatArena(model, bucket_bytes=128)
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(6, 8, generator=g)
    xs = x_all[rank * 3:(rank + 1) * 3]
    ref_model = _tiny_model()
    (0.5 * ref_model(x_all[:3]).pow(2).mean() + 0.5 * ref_model(x_all[3:]).pow(2).mean()).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ref_model.parameters()])
    for _ in range(2):
        arena.zero_grad()
        model(xs).pow(2).mean().backward()
        arena.average_()
        local = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        q.put((rank, float((local - ref).abs().max())))
    dist.destroy_process_group()


def test_flat_arena_inplace_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_arena_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(4)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, err in res:
        assert err < 1e-6, f"rank {rank}: arena-averaged gradients differ from the single-process reference by {err}"


def test_flat_arena_detects_parameters_that_left_it():
    from tfswa_unet_b200.train_step import FlatArena
    model = _tiny_model()
    arena = FlatArena(model)
    arena.check()
    model.double()                                   # re-allocates every parameter outside the arena
    with pytest.raises(RuntimeError, match="no longer lives in the arena"):
        arena.check()
