"""CPU: the loss terms next to the path (tfswa_unet_b200.losses) against values and gradients produced by the live
reference's ``src/training/losses.py`` (tests/golden/make_golden_losses.py -> golden_losses_v1.pt)."""
import os

import pytest
import torch

from helpers import seeded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gl():
    return torch.load(os.path.join(ROOT, "tests", "golden", "golden_losses_v1.pt"), weights_only=False)


def test_mrstft_loss_value_and_gradient(gl):
    from tfswa_unet_b200.losses import mrstft_loss
    assert len(gl["mrstft"]) == 4
    for case in gl["mrstft"]:
        pred = seeded(case["shape"], case["seeds"][0], case["scale"]).requires_grad_(True)
        tgt = seeded(case["shape"], case["seeds"][1], case["scale"])
        loss = mrstft_loss(pred, tgt, **case.get("kwargs", {}))
        loss.backward()
        assert abs(float(loss) - float(case["loss"])) <= 1e-5 * abs(float(case["loss"])), (float(loss), float(case["loss"]))
        g = pred.grad[:, :, ::case["grad_stride"]]
        assert float((g - case["grad"]).abs().max()) <= 1e-4 * float(case["grad"].abs().max()) + 1e-9
        assert abs(float(pred.grad.norm()) - float(case["grad_norm"])) <= 1e-4 * float(case["grad_norm"])


def test_source_separation_loss_dictionary(gl):
    from tfswa_unet_b200.losses import source_separation_loss
    stems = ("vocals", "other")
    ps = {s: seeded((2, 33, 20), 700 + j).abs() for j, s in enumerate(stems)}
    ts_ = {s: seeded((2, 33, 20), 710 + j).abs() for j, s in enumerate(stems)}
    pa = {s: seeded((2, 2, 4500), 720 + j, 0.1) for j, s in enumerate(stems)}
    ta = {s: seeded((2, 2, 4500), 730 + j, 0.1) for j, s in enumerate(stems)}
    for got, ref in ((source_separation_loss(ps, ts_, pa, ta), gl["combined"][0]),
                     (source_separation_loss(ps, ts_, use_mrstft=False), gl["combined"][1])):
        assert set(got) == set(ref), set(got) ^ set(ref)
        for k in ref:
            assert abs(float(got[k]) - float(ref[k])) <= 1e-5 * abs(float(ref[k])) + 1e-8, k


def test_mrstft_rejects_mismatched_resolution_lists():
    from tfswa_unet_b200.losses import mrstft_loss
    with pytest.raises(ValueError):
        mrstft_loss(torch.zeros(1, 1, 4096), torch.zeros(1, 1, 4096), fft_sizes=(512,), hop_sizes=(128, 64))
