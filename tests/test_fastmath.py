"""Accuracy of the branch-free FP64 math layer (ilqc_b200/csrc/fastmath.cuh) against 40-digit mpmath values.
CPU: the same templates in the test-only host build; GPU (`-m gpu`): the CUDA library (hardware reciprocal /
reciprocal-square-root seeds).  Bound: 2 ulp, the documented class of the CUDA math library's sin/cos/atan2."""
import math

import mpmath as mp
import numpy as np
import pytest
import torch

mp.mp.dps = 40
ULP_BOUND = 2.0


def ulp_err(got, exact):
    """max |got - exact| / ulp(exact) with exact given as mpmath numbers"""
    worst = 0.0
    for g, e in zip(got, exact):
        if e == 0:
            assert g == 0.0
            continue
        u = math.ulp(float(e))
        worst = max(worst, float(abs(mp.mpf(float(g)) - e) / u))
    return worst


def samples(seed, n=4000):
    g = np.random.default_rng(seed)
    return g


def run_checks(eng):
    g = samples(0)
    # sin / cos: small angles, a few turns, and the whole supported range
    x = np.concatenate([g.uniform(-math.pi, math.pi, 3000), g.uniform(-40, 40, 2000), g.uniform(-1e5, 1e5, 1500),
                        g.uniform(-1e-3, 1e-3, 500), [0.0, math.pi / 4, -math.pi / 4, math.pi / 2, 1e-300, 105614.9]])
    s, c = eng.fastmath_eval("sincos", torch.tensor(x))
    es = ulp_err(s.cpu().numpy(), [mp.sin(mp.mpf(float(v))) for v in x])
    ec = ulp_err(c.cpu().numpy(), [mp.cos(mp.mpf(float(v))) for v in x])
    assert es <= ULP_BOUND and ec <= ULP_BOUND, (es, ec)
    worst = {'sin': es, 'cos': ec}
    # atan2 over all octants and magnitudes, incl. axis cases
    y = np.concatenate([g.normal(size=3000), g.normal(size=1000) * 1e-8, g.normal(size=1000) * 1e8,
                        [0.0, 0.0, 1.0, -1.0, 1.0, -1.0, 1e-200, -0.0]])
    xx = np.concatenate([g.normal(size=3000), g.normal(size=1000), g.normal(size=1000) * 1e-3,
                         [1.0, -1.0, 0.0, 0.0, 1.0, -1.0, 1e-200, -2.0]])
    a, _ = eng.fastmath_eval("atan2", torch.tensor(y), torch.tensor(xx))
    a = a.cpu().numpy()
    ea = ulp_err(a[:-1], [mp.atan2(mp.mpf(float(p)), mp.mpf(float(q))) for p, q in zip(y[:-1], xx[:-1])])
    assert ea <= ULP_BOUND, ea
    worst['atan2'] = ea
    assert math.copysign(1.0, a[-1]) == -1.0 and abs(a[-1]) == math.pi      # atan2(-0, -2) = -pi
    # atan
    t = np.concatenate([g.normal(size=3000) * 3, g.uniform(-0.5, 0.5, 1500), g.normal(size=500) * 1e6, [0.0, 1.0, -1.0, 1e300]])
    a, _ = eng.fastmath_eval("atan", torch.tensor(t))
    ea = ulp_err(a.cpu().numpy(), [mp.atan(mp.mpf(float(v))) for v in t])
    assert ea <= ULP_BOUND, ea
    worst['atan'] = ea
    print("fastmath max ulp errors:", worst)
    # division / reciprocal: correctly rounded almost always -> compare with IEEE
    p, q = g.normal(size=20000) * np.exp(g.uniform(-50, 50, 20000)), g.normal(size=20000) * np.exp(g.uniform(-50, 50, 20000))
    d, _ = eng.fastmath_eval("div", torch.tensor(p), torch.tensor(q))
    assert np.mean(d.cpu().numpy() == p / q) > 0.999
    assert np.max(np.abs(d.cpu().numpy() - p / q) / np.abs(p / q)) < 2.3e-16
    r, _ = eng.fastmath_eval("rcp", torch.tensor(q))
    assert np.mean(r.cpu().numpy() == 1.0 / q) > 0.999
    w = np.abs(p) + 1e-200
    sq, isq = eng.fastmath_eval("sqrt_rsqrt", torch.tensor(w))
    assert np.mean(sq.cpu().numpy() == np.sqrt(w)) > 0.999
    assert np.max(np.abs(isq.cpu().numpy() * np.sqrt(w) - 1.0)) < 4.5e-16


def test_fastmath_accuracy_host_build(sim_engine):
    run_checks(sim_engine)


@pytest.mark.gpu
def test_fastmath_accuracy_cuda(cuda_engine):
    run_checks(cuda_engine)
