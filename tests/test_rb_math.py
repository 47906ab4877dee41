"""Accuracy of the shared elementary layer (rb_math.h) against float64 libm. These routines replace GLSL built-ins whose
precision is implementation-defined; they must be at least as accurate as a GPU driver's (a few ulp)."""
import numpy as np

SIN, COS, LOG, EXP, EXP2, ACOS = range(6)


def ulp_err(y, ref):
    ref32 = ref.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    ulp = np.where(ulp == 0, np.finfo(np.float32).tiny, ulp)
    return np.abs(y.astype(np.float64) - ref) / ulp


def test_sin_cos(ol):
    x = np.concatenate([np.linspace(0, 2 * np.pi, 200001), np.linspace(-50, 50, 100001)]).astype(np.float32)
    # ulp error relative to max(|ref|, 2^-6): near zeros of sin/cos absolute error is what matters downstream
    for fn, f in ((SIN, np.sin), (COS, np.cos)):
        y = ol.rb_math(fn, x)
        ref = f(x.astype(np.float64))
        err = np.abs(y - ref) / np.maximum(np.abs(ref), 2.0 ** -6)
        assert err.max() < 4 * 2.0 ** -23, err.max()
    s, c = ol.rb_math(SIN, x), ol.rb_math(COS, x)
    assert np.abs(s.astype(np.float64) ** 2 + c.astype(np.float64) ** 2 - 1).max() < 4e-7


def test_log(ol):
    x = np.concatenate([np.logspace(-30, 30, 200001), np.linspace(1e-5, 1.0, 100001)]).astype(np.float32)
    y = ol.rb_math(LOG, x)
    assert ulp_err(y, np.log(x.astype(np.float64))).max() <= 2.0
    sp = ol.rb_math(LOG, np.array([0.0, -1.0, np.inf, 1.0], np.float32))
    assert sp[0] == -np.inf and np.isnan(sp[1]) and sp[2] == np.inf and sp[3] == 0.0


def test_exp(ol):
    x = np.linspace(-87.0, 88.0, 400001).astype(np.float32)
    y = ol.rb_math(EXP, x)
    assert ulp_err(y, np.exp(x.astype(np.float64))).max() <= 2.0
    sp = ol.rb_math(EXP, np.array([-200.0, -87.4, 0.0, 100.0], np.float32))
    assert sp[0] == 0.0 and sp[1] == 0.0 and sp[2] == 1.0 and sp[3] == np.inf     # subnormal results are flushed


def test_exp2_exact_on_integers(ol):
    k = np.arange(-100, 100).astype(np.float32)
    assert (ol.rb_math(EXP2, k) == np.exp2(k.astype(np.float64)).astype(np.float32)).all()
    x = np.linspace(-20, 20, 100001).astype(np.float32)
    assert ulp_err(ol.rb_math(EXP2, x), np.exp2(x.astype(np.float64))).max() <= 3.0


def test_acos(ol):
    x = np.linspace(-1, 1, 200001).astype(np.float32)
    y = ol.rb_math(ACOS, x)
    ref = np.arccos(x.astype(np.float64))
    assert (np.abs(y - ref) / np.maximum(ref, 2.0 ** -6)).max() < 4 * 2.0 ** -23
    assert np.isnan(ol.rb_math(ACOS, np.array([1.0000001, -1.5], np.float32))).all()
