"""CPU: the fitted GELU used for the MLP hidden activations of the fused bf16 tail kernels (csrc/common.cuh::gelu_hidden)
restated in numpy and held against the erf form the reference uses (nn.GELU(), attention.py:123)."""
import math

import numpy as np


def gelu_erf(x):
    return 0.5 * x * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


def gelu_hidden_model(x, tanh=np.tanh):
    x = np.maximum(x, -6.0)
    x2 = np.minimum(x * x, 51.6)
    g = (-3.58732361e-4 * x2 + 0.0370503451) * x2 + 0.797458471
    hx = 0.5 * x
    return hx * tanh(x * g) + hx


def test_fit_error_against_erf_gelu():
    x = np.linspace(-40.0, 40.0, 400001)
    err = np.abs(gelu_hidden_model(x) - gelu_erf(x))
    assert err.max() < 3.2e-5, err.max()


def test_error_with_hardware_tanh_bound():
    """MUFU.TANH: 2^-11 relative error on tanh; worst case in either direction stays below 1.5e-3 absolute and below the
    bf16 half-ulp of the result for x > -1.15"""
    x = np.linspace(-40.0, 12.0, 520001)
    ref = gelu_erf(x)
    for s in (+1.0, -1.0):
        got = gelu_hidden_model(x, tanh=lambda u: np.tanh(u) * (1.0 + s * 2.0 ** -11))
        err = np.abs(got - ref)
        assert err[x <= 0].max() < 1.5e-3                       # absolute where the result is small
        assert np.all(err[x > 0] <= np.abs(ref[x > 0]) * 2.6e-4 + 4e-5)      # relative where it is ~x
        sel = x > -1.15
        assert np.all(err[sel] <= np.abs(ref[sel]) * 2.0 ** -9 + 4e-5)


def test_saturation_and_monotone_argument():
    x = np.array([-1e4, -100.0, -8.0, 8.0, 100.0, 1e4])
    y = gelu_hidden_model(x)
    assert np.allclose(y[:3], 0.0, atol=1e-6) and np.allclose(y[3:], x[3:], rtol=1e-6)
