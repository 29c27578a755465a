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
