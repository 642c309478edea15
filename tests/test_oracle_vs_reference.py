"""CPU suite: oracle restatement against the live compiled reference (oracle/_ref), when present.
Also checks the primitive conv/linear ops in the reference's own channel-major layouts."""
import ctypes as C

import numpy as np
import pytest

from conftest import max_rel

import pyoracle

pytestmark = pytest.mark.skipif(not pyoracle.REF_SO.exists(), reason="oracle/_ref not built")
fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def ref(model_dir):
    r = pyoracle.Reference(model_dir)
    yield r
    r.close()


def _p(a):
    return a.ctypes.data_as(fp)


@pytest.mark.parametrize("cin,cout,T,k,s,g", [(8, 6, 17, 7, 1, 1), (6, 4, 9, 3, 1, 1), (4, 4, 5, 1, 1, 1), (4, 4, 12, 4, 2, 2)])
def test_conv1d_primitive(oracle, ref, cin, cout, T, k, s, g):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((cin, T)).astype(np.float32)
    w = rng.standard_normal((cout, cin // g, k)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    y0 = np.zeros((cout, T // s), np.float32)
    y1 = np.zeros_like(y0)
    ref.lib.ptts_conv1d_forward(_p(y0), _p(x), _p(w), _p(b), cin, cout, T, k, s, g)
    oracle.lib.orc_conv1d(_p(y1), _p(x), _p(w), _p(b), cin, cout, T, k, s, g)
    assert max_rel(y1, y0) < 1e-5


@pytest.mark.parametrize("cin,cout,T,k,s,g", [(8, 6, 7, 12, 6, 1), (6, 4, 5, 10, 5, 1), (4, 4, 1, 8, 4, 1), (4, 4, 6, 32, 16, 4)])
def test_convtr1d_primitive(oracle, ref, cin, cout, T, k, s, g):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((cin, T)).astype(np.float32)
    w = rng.standard_normal((cin, cout // g, k)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    y0 = np.zeros((cout, T * s), np.float32)
    y1 = np.zeros_like(y0)
    ref.lib.ptts_convtr1d_forward(_p(y0), _p(x), _p(w), _p(b), cin, cout, T, k, s, g)
    oracle.lib.orc_convtr1d(_p(y1), _p(x), _p(w), _p(b), cin, cout, T, k, s, g)
    assert max_rel(y1, y0) < 1e-5


def test_generate_and_decode_live(oracle, ref, voice, golden):
    ids = golden["p2_ids"]
    r = ref.flowlm_generate(ids, voice[:16], 3, lsd_steps=1, temp=0.0)
    o = oracle.flowlm_generate(ids, voice[:16], 3, lsd_steps=1, temp=0.0)
    assert max_rel(o["latents"], r["latents"]) < 5e-5
    sl = ref.scale_latents(r["latents"])
    assert max_rel(oracle.mimi_decode(sl), ref.mimi_decode(sl)) < 5e-5
