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
