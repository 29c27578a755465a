"""CPU: the overlap-add restatement (oracle/ola_oracle.py) and the product's batched ``ShardedSeparator`` (single process)
against the output of the live reference's ``SourceSeparator`` (tests/golden/make_golden_ola.py -> golden_ola_v1.pt)."""
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stand_in(x):
    return torch.sigmoid(0.3 * x + 0.1 * x.flip(1))


@pytest.fixture(scope="module")
def go():
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_ola_v1.pt"), weights_only=False)


def _audio(go):
    g = torch.Generator().manual_seed(go["audio_seed"])
    return 0.1 * torch.randn(2, go["samples"], generator=g)


@pytest.mark.parametrize("idx", [0, 1])
def test_oracle_matches_reference_separator(go, idx):
    from oracle.ola_oracle import separate_long
    case = go["cases"][idx]
    out = separate_long(_audio(go), _stand_in, n_fft=256, hop=64, sr=8000, segment_length=0.5, overlap=0.25,
                        normalize=case["normalize"])
    ref = torch.cat([case["vocals"], case["other"]])
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-7


@pytest.mark.parametrize("idx,batch", [(0, 3), (1, 8), (0, 1)])
def test_sharded_separator_matches_reference_separator(go, idx, batch):
    from tfswa_unet_b200.separate import ShardedSeparator
    case = go["cases"][idx]
    sep = ShardedSeparator(_stand_in, n_fft=256, hop_length=64, sample_rate=8000, segment_length=0.5, overlap=0.25, batch=batch,
                           normalize=case["normalize"])
    out = sep.separate(_audio(go))
    for name in ("vocals", "other"):
        ref = case[name]
        assert out[name].shape == ref.shape
        assert float((out[name] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-7, name


def test_short_clip_follows_the_single_segment_path(go):
    """audio no longer than one segment: inference.py:92-95 skips the overlap-add (no window, ISTFT's own length)"""
    from tfswa_unet_b200.separate import ShardedSeparator
    sep = ShardedSeparator(_stand_in, n_fft=256, hop_length=64, sample_rate=8000, segment_length=0.5, overlap=0.25, batch=4)
    out = sep.separate(_audio(go)[:, :go["short"]["samples"]])
    for name in ("vocals", "other"):
        ref = go["short"][name]
        assert out[name].shape == ref.shape, (out[name].shape, ref.shape)
        assert float((out[name] - ref).abs().max()) <= 1e-5 * float(ref.abs().max()) + 1e-7, name


def test_balanced_batches_and_segment_plan():
    """host logic of the separator: a rank's segments in the fewest batches of near-equal size (17 -> 6 + 6 + 5, not 8 + 8 + 1),
    covering the range exactly once and in order; the segment plan of inference.py:187-201"""
    from tfswa_unet_b200.separate import ShardedSeparator, balanced_batches
    assert balanced_batches(0, 17, 8) == [(0, 6), (6, 12), (12, 17)]
    assert balanced_batches(5, 5, 8) == [] and balanced_batches(3, 4, 8) == [(3, 4)]
    for lo, hi, mb in [(0, 133, 8), (7, 40, 8), (0, 8, 8), (0, 9, 8), (2, 67, 3)]:
        bs = balanced_batches(lo, hi, mb)
        assert bs[0][0] == lo and bs[-1][1] == hi and all(a[1] == b[0] for a, b in zip(bs, bs[1:]))
        sizes = [e - s for s, e in bs]
        assert max(sizes) <= mb and max(sizes) - min(sizes) <= 1 and len(bs) == -(-(hi - lo) // mb)
    sep = ShardedSeparator(lambda x: x, n_fft=2048, hop_length=512, sample_rate=44100, segment_length=6.0, overlap=0.25)
    starts = sep.plan(26_460_000)                       # the 10-minute mix of BASELINE configs[3]
    assert len(starts) == 133 and starts[1] - starts[0] == 198_450 and starts[-1] + 264_600 <= 26_460_000
    assert sep.plan(264_600) == [0] and sep.plan(1000) == [0]
