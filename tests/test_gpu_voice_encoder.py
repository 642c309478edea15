"""GPU suite: voice-cloning encoder through the C ABI against the oracle's orc_voice_encode.
The reference has no forward pass for these tensors (it only shape-checks them, ptts.c:914-944), so the oracle composition is
pinned on its primitives (tests/test_oracle_vs_reference.py::test_voice_encoder_composition), not end to end. Tolerance: the
device path carries fp16 activations like the decoder (PCM 63 dB there): conditioning rows within 5e-3 max-relative."""
import numpy as np
import pytest

from conftest import max_rel, max_rel_per_frame

pytestmark = pytest.mark.gpu
TOL = 5e-3


@pytest.fixture(scope="module")
def eng(pkg, weights):
    e = pkg.Engine(weights, 0)
    yield e
    e.close()


def _pcm(n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 24000.0
    return (0.25 * np.sin(2 * np.pi * 220.0 * t) + 0.1 * rng.standard_normal(n)).astype(np.float32)


@pytest.mark.parametrize("n_samples", [1920, 1920 * 3 - 700, 1920 * 20])
def test_encoder_vs_oracle(eng, oracle, n_samples):
    """1 frame, a ragged tail (zero padded to whole frames), and 20 frames = 320 transformer tokens (> the 250-token window)."""
    pcm = _pcm(n_samples, n_samples)
    got = eng.voice_encode(pcm)
    want = oracle.voice_encode(pcm)
    assert got.shape == want.shape == ((n_samples + 1919) // 1920, 1024)
    print(f"voice encoder, {got.shape[0]} frames: max-rel {max_rel(got, want):.2e} (per-frame {max_rel_per_frame(got, want):.2e})")
    assert max_rel(got, want) < TOL


def test_encoded_voice_drives_generation(eng, oracle, golden):
    """The encoder's rows are a cond_prefix like a voice file's audio_prompt (ptts.c:293-364): generation from them matches the
    oracle's generation from the same rows."""
    cond = eng.voice_encode(_pcm(1920 * 6, 7))
    ids = golden["p2_ids"]
    g = eng.flowlm_generate_latents(ids, cond, 6, lsd_steps=1, temp=0.0)
    o = oracle.flowlm_generate(ids, cond, 6, lsd_steps=1, temp=0.0)
    assert max_rel(g["latents"], o["latents"]) < 1e-3


def test_missing_encoder_tensors_fail_loudly(pkg, weights):
    W = {k: v for k, v in weights.items() if "encoder" not in k and "downsample" not in k}
    e = pkg.Engine(W, 0)
    with pytest.raises(pkg.EngineError, match="no encoder tensors"):
        e.voice_encode(np.zeros(1920, np.float32))
    e.close()
