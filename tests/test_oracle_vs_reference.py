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


@pytest.mark.parametrize("cin,cout,T,k,s,g", [(8, 6, 17, 7, 1, 1), (6, 4, 9, 3, 1, 1), (4, 4, 5, 1, 1, 1), (4, 4, 12, 4, 2, 2),
                                              # the strided causal convs of the voice-cloning encoder (k = 2 * stride, ptts.c:936, 914)
                                              (8, 16, 24, 8, 4, 1), (6, 12, 20, 10, 5, 1), (4, 8, 24, 12, 6, 1), (8, 8, 64, 32, 16, 1)])
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


def test_voice_encoder_composition(weights, ref):
    """orc_voice_encode against the same module order (ptts.c:919-944, 914) composed in numpy from the REFERENCE's
    own Conv1d primitive. The encoder transformer is switched off for this comparison (layer scales = 0 make it the identity; its
    arithmetic is the decoder transformer's, pinned by test_generate_and_decode_live)."""
    W = dict(weights)
    for l in range(2):
        for k in ("layer_scale_1.scale", "layer_scale_2.scale"):
            W[f"mimi.encoder_transformer.transformer.layers.{l}.{k}"] = np.zeros(512, np.float32)
    o = pyoracle.Oracle(W)
    rng = np.random.default_rng(3)
    n = 1920 * 3 - 700                       # not a whole number of frames: zero padded on the right
    pcm = (0.3 * rng.standard_normal(n)).astype(np.float32)
    got = o.voice_encode(pcm)

    def conv(x, name, cin, cout, k, s, bias=True):
        w = np.ascontiguousarray(W["mimi." + name + ".weight"], np.float32)
        b = np.ascontiguousarray(W["mimi." + name + ".bias"], np.float32) if bias else None
        T = x.shape[1]
        y = np.zeros((cout, T // s), np.float32)
        ref.lib.ptts_conv1d_forward(_p(y), _p(np.ascontiguousarray(x)), _p(w), _p(b) if bias else None, cin, cout, T, k, s, 1)
        return y

    elu = lambda v: np.where(v >= 0, v, np.expm1(np.minimum(v, 0))).astype(np.float32)
    x = np.zeros((1, 1920 * 3), np.float32)
    x[0, :n] = pcm
    x = conv(x, "encoder.model.0.conv", 1, 64, 7, 1)
    ch = 64
    for idx, ratio in ((1, 4), (4, 5), (7, 6)):
        h = conv(elu(x), f"encoder.model.{idx}.block.1.conv", ch, ch // 2, 3, 1)
        x = x + conv(elu(h), f"encoder.model.{idx}.block.3.conv", ch // 2, ch, 1, 1)
        x = conv(elu(x), f"encoder.model.{idx + 2}.conv", ch, 2 * ch, 2 * ratio, ratio)
        ch *= 2
    x = conv(elu(x), "encoder.model.11.conv", 512, 512, 3, 1)
    lat = conv(x, "downsample.conv.conv", 512, 512, 32, 16, bias=False)          # [512, frames]
    want = lat.T @ W["flow_lm.speaker_proj_weight"].T                             # [frames, 1024]
    assert got.shape == (3, 1024)
    assert max_rel(got, want) < 2e-5
