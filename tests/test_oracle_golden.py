"""CPU: pin the oracle restatement against outputs of the live reference (tests/golden)."""
import pytest
import torch

from oracle import tfswa_oracle as O
from helpers import seeded, check_grad, assert_close

TOL = 2e-5   # fp32 reference vs fp32 oracle: different op order only


def _module_state(kind, C, seed, **kw):
    """Build the state dict a reference module of this kind would have, filled by the shared recipe."""
    from tfswa_unet_b200 import layout
    sd = layout.empty_state(kind, C, **kw)
    O.randomize_state_(sd, seed, kw.get("gain", 1.0))
    return sd


BRANCH_CASES = [("tsa_c32", "tsa", 32, 0), ("fsa_c32", "fsa", 32, 0), ("swa_c32_s0", "swa", 32, 0),
                ("swa_c32_s4", "swa", 32, 4), ("tsa_c64", "tsa", 64, 0), ("fsa_c128", "fsa", 128, 0),
                ("swa_c256_s4", "swa", 256, 4)]


def _run(fn, x, params, seed):
    x = x.clone().requires_grad_(True)
    params = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k and "attn_mask" not in k else v)
              for k, v in params.items()}
    y = fn(x, params)
    w = seeded(y.shape, seed + 7)
    (y * w).sum().backward()
    return y, x.grad, {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}


@pytest.mark.parametrize("name,kind,C,shift", BRANCH_CASES)
def test_branch_matches_reference(golden, name, kind, C, shift):
    case = golden["cases"][name]
    sd = _module_state(kind, C, case["seed"], shift=shift)
    x = seeded(case["shape"], case["seed"] + 100)
    fn = {"tsa": lambda x, p: O.tsa(x, p), "fsa": lambda x, p: O.fsa(x, p),
          "swa": lambda x, p: O.swa(x, p, 8, shift)}[kind]
    y, dx, grads = _run(fn, x, sd, case["seed"])
    assert_close(name + ".y", y, case["y"], TOL)
    assert_close(name + ".dx", dx, case["dx"], TOL)
    assert set(grads) == set(case["grads"])
    for k, g in grads.items():
        check_grad(f"{name}.{k}", g, case["grads"][k], 5e-5)


BLOCK_CASES = ["block_c32_s0_eval", "block_c32_s4_skip_eval", "block_c32_s4_train", "block_c64_s4_skip_train"]


@pytest.mark.parametrize("name", BLOCK_CASES)
def test_block_matches_reference(golden, name):
    case = golden["cases"][name]
    C = case["shape"][1]
    sd = _module_state("block", C, case["seed"], shift=case["shift"])
    x = seeded(case["shape"], case["seed"] + 100)
    skip = seeded(case["shape"], case["seed"] + 200) if case["with_skip"] else None
    stats = {}
    y, dx, grads = _run(lambda x, p: O.tfswa_block(x, p, case["shift"], skip=skip, training=case["train"],
                                                    new_stats=stats), x, sd, case["seed"])
    assert_close(name + ".y", y, case["y"], TOL)
    assert_close(name + ".dx", dx, case["dx"], 5e-5)
    for k, g in grads.items():
        check_grad(f"{name}.{k}", g, case["grads"][k], 1e-4, atol=3e-4 if case["train"] else None)
    if case["train"]:
        for k, v in case["buffers"].items():
            assert_close(f"{name}.{k}", stats[k].float(), v.float(), 1e-5)


@pytest.mark.parametrize("name,kind,cin,cout", [("down_32_64_eval", "down", 32, 64), ("down_64_128_train", "down", 64, 128),
                                                ("up_64_32_eval", "up", 64, 32), ("up_128_64_train", "up", 128, 64)])
def test_resample_matches_reference(golden, name, kind, cin, cout):
    case = golden["cases"][name]
    sd = _module_state(kind, cin, case["seed"], cout=cout)
    x = seeded(case["shape"], case["seed"] + 100)
    stats = {}
    fn = O.downsample if kind == "down" else O.upsample
    y, dx, grads = _run(lambda x, p: fn(x, p, case["train"], stats), x, sd, case["seed"])
    assert_close(name + ".y", y, case["y"], TOL)
    assert_close(name + ".dx", dx, case["dx"], 5e-5)
    for k, g in grads.items():
        check_grad(f"{name}.{k}", g, case["grads"][k], 1e-4, atol=3e-4 if case["train"] else None)
    if case["train"]:
        for k, v in case["buffers"].items():
            assert_close(f"{name}.{k}", stats[k].float(), v.float(), 1e-5)


@pytest.mark.parametrize("name", ["unet_65x41_eval", "unet_64x96_train"])
def test_unet_matches_reference(golden, name):
    case = golden["cases"][name]
    sd = _module_state("unet", 32, case["seed"], cin=case["cin"], cout=case["cout"], gain=case["gain"])
    x = seeded(case["shape"], case["seed"] + 100)
    stats = {}
    taps = {}
    y, dx, grads = _run(lambda x, p: O.unet_forward(x, p, training=case["train"], new_stats=stats, taps=taps),
                        x, sd, case["seed"])
    assert_close(name + ".logits", taps["logits"], case["logits"], 1e-4)
    assert_close(name + ".y", y, case["y"], 2e-5)
    assert_close(name + ".dx", dx, case["dx"], 2e-4)
    assert set(grads) == set(case["grads"])
    for k, g in grads.items():
        check_grad(f"{name}.{k}", g, case["grads"][k], 5e-4, atol=3e-4 if case["train"] else None)
    if case["train"]:
        for k, v in case["buffers"].items():
            assert_close(f"{name}.{k}", stats[k].float(), v.float(), 2e-5)


def test_state_layout_and_param_count(golden):
    from tfswa_unet_b200 import layout
    sd = layout.empty_state("unet", 32, cin=2, cout=2)
    ref = golden["state_layout"]
    assert list(sd.keys()) == list(ref.keys())
    for k, shp in ref.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    n = sum(v.numel() for k, v in sd.items() if "running" not in k and "num_batches" not in k and "attn_mask" not in k)
    assert n == golden["num_parameters"] == 15404834


def test_attn_mask_buffer_matches_reference(golden):
    m = O.reference_attn_mask_buffer(8, 4)
    assert torch.equal(m.to(torch.int8), golden["attn_mask_ws8_s4"])


def test_swin_mask_feature_is_consistent_with_buffer_on_64x64():
    # the reference buffer is the Swin mask of a 64x64 map; our general builder must agree there
    assert torch.equal(O.swin_shift_mask(64, 64, 8, 4), O.reference_attn_mask_buffer(8, 4))


def test_flop_model_matches_survey():
    assert O.count_block_flops(1, 32, 32, 48) == 154_140_672
    assert O.count_block_flops(2, 64, 20, 28) == 446_824_448
    assert O.count_model_flops(1, 2, 2, 64, 96) == 13_127_385_088
    assert abs(O.count_model_flops(8, 2, 2, 1025, 517) / 1e12 - 13.38) < 0.01


def test_product_flop_model_equals_the_checker():
    """bench.py divides by tfswa_unet_b200.flops (product side); it must agree with the oracle's independent count."""
    from tfswa_unet_b200 import flops
    for args in [(1, 32, 32, 48), (2, 64, 20, 28), (8, 32, 1025, 517)]:
        assert flops.block_flops(*args) == O.count_block_flops(*args)
    assert flops.block_flops(1, 32, 32, 48) == 154140672                   # FlopCounterMode on the reference
    for args in [(1, 2, 2, 64, 96), (8, 2, 2, 1025, 517), (8, 4, 4, 1025, 517)]:
        assert flops.model_flops(*args) == O.count_model_flops(*args)
    assert flops.model_flops(1, 2, 2, 64, 96) == 13127385088


def test_product_window_helpers_and_mask_buffer_match_reference():
    """the PRODUCT's window_partition / window_reverse and the attn_mask buffer its ShiftedWindowAttention
    registers, against outputs of the live reference (tests/golden/make_golden_windows.py) - not only the oracle's."""
    import os
    import tfswa_unet_b200 as T
    from tfswa_unet_b200.attention import window_partition, window_reverse
    from helpers import seeded
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_windows_v1.pt"), weights_only=False)
    for case in g["cases"]:
        x = seeded(case["shape"], case["seed"])
        win = window_partition(x, case["ws"])
        assert torch.equal(win, case["windows"]), case["shape"]
        assert torch.equal(window_reverse(win, case["ws"], case["shape"][2], case["shape"][3]), x)
        assert torch.equal(O.window_partition(x, case["ws"]), case["windows"])          # and the checker's own
    for (ws, shift), ref in g["masks"].items():
        m = T.ShiftedWindowAttention(32, ws, 8, shift_size=shift)
        if ref is None:
            assert m.attn_mask is None and "attn_mask" not in m.state_dict()
        else:
            assert m.attn_mask.dtype == torch.float32 and tuple(m.attn_mask.shape) == tuple(ref.shape)
            assert torch.equal(m.attn_mask.to(torch.int8), ref), (ws, shift)
            assert "attn_mask" in m.state_dict()


def test_product_attn_mask_matches_main_golden(golden):
    import tfswa_unet_b200 as T
    m = T.ShiftedWindowAttention(32, 8, 8, shift_size=4)
    assert torch.equal(m.attn_mask.to(torch.int8), golden["attn_mask_ws8_s4"])
