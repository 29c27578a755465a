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


def test_fast_gelu_gradient_formula_matches_exact():
    """common.cuh::gelu_erf_grad_fast restated in float64: A&S 7.1.26 erf and the Gaussian density share exp(-x^2/2)"""
    import numpy as np
    from math import erf, sqrt, pi
    x = np.linspace(-12, 12, 200001)
    z = np.abs(x) / sqrt(2.0)
    t = 1.0 / (1.0 + 0.3275911 * z)
    e = np.exp(-z * z)
    p = ((((1.061405429 * t - 1.453152027) * t + 1.421413741) * t - 0.284496736) * t + 0.254829592)
    erfv = np.sign(x) * (1.0 - p * t * e)
    fast = 0.5 + 0.5 * erfv + x * (e / sqrt(2.0 * pi))
    exact = np.array([0.5 * (1.0 + erf(v / sqrt(2.0))) for v in x]) + x * np.exp(-0.5 * x * x) / sqrt(2.0 * pi)
    assert float(np.abs(fast - exact).max()) < 2e-7          # bf16 resolution of the gradient is 4e-3
