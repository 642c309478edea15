"""GPU suite: drop-in proof. The reference's own CLI (main.c + ptts.c + tokenizer + safetensors + WAV
writer, UNMODIFIED) linked against host/ptts_b200_models.c + libptts_b200.so must produce the same
WAV as the reference CPU build for the same command line (north star: >= 40 dB SNR)."""
import subprocess
import wave
from pathlib import Path

import numpy as np
import pytest

from conftest import snr_db

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "pocket-tts.c_b200" / "build" / "ptts_b200_cli"
REF = ROOT / "oracle" / "_ref" / "ptts_ref"


def _wav(path):
    with wave.open(str(path), "rb") as w:
        assert w.getframerate() == 24000 and w.getnchannels() == 1 and w.getsampwidth() == 2
        return np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).astype(np.float64)


@pytest.mark.skipif(not (CLI.exists() and REF.exists()), reason="prebuilt CLI binaries not present")
def test_reference_cli_on_b200_engine_matches_reference_cpu(model_dir, tmp_path):
    args = ["-d", str(model_dir), "-p", "Hello world!", "--voice", "alba", "-t", "0.7", "-S", "123", "--frames", "5",
            "--eos-threshold", "1e9", "-q"]
    a, b = tmp_path / "b200.wav", tmp_path / "ref.wav"
    subprocess.run([str(CLI)] + args + ["-o", str(a)], check=True, timeout=300)
    subprocess.run([str(REF)] + args + ["-o", str(b)], check=True, timeout=300)
    x, y = _wav(a), _wav(b)
    assert x.shape == y.shape == (5 * 1920,)
    assert snr_db(x, y) > 40.0


@pytest.mark.skipif(not (CLI.exists() and REF.exists()), reason="prebuilt CLI binaries not present")
def test_reference_cli_mimi_test_hook_on_b200_engine(model_dir):
    """--mimi-test (main.c:413-466) calls ptts_mimi_forward_one: the statistic line printed by the drop-in binary must
    agree with the reference CPU binary's."""
    import re
    args = ["-d", str(model_dir), "-p", "Hello world!", "--voice", "alba", "-t", "0", "--mimi-test", "--frames", "3"]
    a = subprocess.run([str(CLI)] + args, check=True, timeout=300, capture_output=True, text=True).stdout
    b = subprocess.run([str(REF)] + args, check=True, timeout=300, capture_output=True, text=True).stdout
    pat = re.compile(r"Mimi decode \(transformer\) stats: mean=(\S+) min=(\S+) max=(\S+)")
    ma, mb = pat.search(a), pat.search(b)
    assert ma and mb, (a, b)
    va, vb = np.array([float(x) for x in ma.groups()]), np.array([float(x) for x in mb.groups()])
    assert np.abs(va - vb).max() < 2e-3 * max(1.0, np.abs(vb).max())


RELOAD = ROOT / "pocket-tts.c_b200" / "build" / "ptts_b200_reload_check"


@pytest.mark.skipif(not (RELOAD.exists() and REF.exists()), reason="prebuilt host-layer binaries not present")
def test_host_layer_engine_lifetime_across_free_and_reload(synth, tmp_path):
    """ptts_free + ptts_load_dir of ANOTHER model (possibly at a recycled safetensors_file_t address) must never reuse the
    first model's engine (the stale-pointer cache hazard of ptts_cuda.c:617-644), repeated ptts_generate calls on one
    context must reuse one engine, and reloads must not accumulate engines (host/ptts_b200_models.c)."""
    dir_a, dir_b = synth.model_dir(0), synth.model_dir(1)
    out = subprocess.run([str(RELOAD), str(dir_a), str(dir_b), str(tmp_path)], check=True, timeout=600, capture_output=True, text=True).stdout
    assert "final engines alive" in out, out
    args = ["-p", "Hello world!", "--voice", "alba", "-t", "0.7", "-S", "123", "--frames", "5", "--eos-threshold", "1e9", "-q"]
    subprocess.run([str(REF), "-d", str(dir_a)] + args + ["-o", str(tmp_path / "ref_a.wav")], check=True, timeout=300)
    subprocess.run([str(REF), "-d", str(dir_b)] + args + ["-o", str(tmp_path / "ref_b.wav")], check=True, timeout=300)
    a, b, a2 = _wav(tmp_path / "a.wav"), _wav(tmp_path / "b.wav"), _wav(tmp_path / "a2.wav")
    ra, rb = _wav(tmp_path / "ref_a.wav"), _wav(tmp_path / "ref_b.wav")
    assert snr_db(a, ra) > 40.0 and snr_db(b, rb) > 40.0      # each context ran on ITS OWN weights
    assert snr_db(b, ra) < 10.0                                # and the two models really differ
    assert np.array_equal(a, a2)                               # reloading A reproduces A bit for bit


BATCH = ROOT / "pocket-tts.c_b200" / "build" / "ptts_b200_batch_check"


@pytest.mark.skipif(not (BATCH.exists() and REF.exists()), reason="prebuilt host-layer binaries not present")
def test_host_layer_generate_batch_matches_reference_cli_per_text(model_dir, synth, tmp_path):
    """ptts_generate_batch (the additive ptts.h-level entry of host/ptts_b200_models.c): texts of different lengths are prepared and
    tokenised by the reference's own host code, bucketed by ptts_estimate_frames, run through the engine pool and returned in the
    ORIGINAL order; each WAV must match the reference CPU CLI run on that text alone (same seed, frames from the same estimate)."""
    texts = ["Hello world!", synth.make_prompts(1, seed=5)[0], "One two three four five six seven.", "Short one.",
             synth.make_prompts(1, seed=6, lo=40, hi=50)[0]]
    subprocess.run([str(BATCH), str(model_dir), str(tmp_path)] + texts, check=True, timeout=600)
    lengths = set()
    for i, t in enumerate(texts):
        ref = tmp_path / f"ref{i}.wav"
        subprocess.run([str(REF), "-d", str(model_dir), "-p", t, "--voice", "alba", "-t", "0.7", "-S", str(123 + i), "--eos-threshold", "1e9",
                        "-q", "-o", str(ref)], check=True, timeout=600)
        x, y = _wav(tmp_path / f"b{i}.wav"), _wav(ref)
        assert x.shape == y.shape, f"utterance {i}: {x.shape} vs {y.shape}"
        assert snr_db(x, y) > 40.0, f"utterance {i}"
        lengths.add(x.shape[0])
    assert len(lengths) >= 3      # the batch really was ragged


@pytest.mark.skipif(not (CLI.exists() and REF.exists()), reason="prebuilt CLIs not present")
def test_cloned_voice_file_works_with_the_unmodified_reference(pkg, model_dir, synth, weights, tmp_path):
    """Voice cloning end to end: PCM -> ptts_b200_voice_encode -> a voice file in the reference's own layout
    ("audio_prompt" [1, T, 1024], ptts.c:293-364). The UNMODIFIED reference CPU CLI and the reference CLI on the B200 engine both
    load it through the reference's voice loader (--voice <path>) and must produce the same audio."""
    import shutil
    md = tmp_path / "model"
    shutil.copytree(model_dir, md)
    t = np.arange(24000 * 2) / 24000.0
    pcm = (0.3 * np.sin(2 * np.pi * 180.0 * t) * (1.0 + 0.5 * np.sin(2 * np.pi * 3.0 * t))).astype(np.float32)
    eng = pkg.Engine(weights, 0)
    cond = eng.voice_encode(pcm)
    eng.close()
    assert cond.shape == (25, 1024) and np.isfinite(cond).all()
    synth.write_safetensors(md / "embeddings" / "cloned.safetensors", {"audio_prompt": cond[None]}, "F32")
    args = ["-d", str(md), "-p", "Hello cloned world!", "--voice", "cloned", "-t", "0.7", "-S", "5", "--frames", "4", "--eos-threshold", "1e9", "-q"]
    subprocess.run([str(CLI)] + args + ["-o", str(tmp_path / "b200.wav")], check=True, timeout=600)
    subprocess.run([str(REF)] + args + ["-o", str(tmp_path / "ref.wav")], check=True, timeout=600)
    x, y = _wav(tmp_path / "b200.wav"), _wav(tmp_path / "ref.wav")
    assert x.shape == y.shape and snr_db(x, y) > 40.0
