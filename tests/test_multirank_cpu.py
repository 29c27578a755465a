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
