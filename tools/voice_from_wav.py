"""Turn a 24 kHz mono WAV into a Pocket-TTS voice file (voice cloning):
    python tools/voice_from_wav.py MODEL_DIR voice.wav NAME
runs the engine's voice-cloning encoder (ptts_b200_voice_encode) and writes MODEL_DIR/embeddings/NAME.safetensors with the tensor
"audio_prompt" [1, T, 1024] - the layout the reference's voice loader reads (ptts.c:293-364), so the unmodified reference CLI can
use it as `--voice NAME`. 16-bit PCM WAV only; other sample rates are refused (the encoder runs at 24 kHz)."""
import sys
import wave
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402


def read_wav(path) -> np.ndarray:
    with wave.open(str(path), "rb") as w:
        if w.getframerate() != 24000 or w.getsampwidth() != 2:
            raise SystemExit(f"{path}: need 24 kHz 16-bit PCM (got {w.getframerate()} Hz, {8 * w.getsampwidth()} bit)")
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.float32) / 32768.0
        if w.getnchannels() > 1:
            pcm = pcm.reshape(-1, w.getnchannels()).mean(axis=1)
    return pcm


def main():
    if len(sys.argv) != 4:
        raise SystemExit(__doc__)
    model_dir, wav, name = Path(sys.argv[1]), sys.argv[2], sys.argv[3]
    pkg = load_package()
    eng = pkg.Engine.from_dir(model_dir, 0)
    cond = eng.voice_encode(read_wav(wav))
    eng.close()
    out = model_dir / "embeddings" / f"{name}.safetensors"
    out.parent.mkdir(parents=True, exist_ok=True)
    pkg.synth.write_safetensors(out, {"audio_prompt": cond[None]}, "F32")
    print(f"{out}: audio_prompt [1, {cond.shape[0]}, 1024] ({cond.shape[0] / 12.5:.2f} s)")


if __name__ == "__main__":
    main()
