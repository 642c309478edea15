"""Generate tests/golden/reference_vectors.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists and `make -C oracle ref`
has produced oracle/_ref/libptts_ref.so):

    python tests/make_golden.py

The reference repo stores no numeric vectors for this path, so the
golden data are outputs of the reference itself on the synthetic checkpoint
(pocket-tts.c_b200/synth.py, seed 0): token ids, FlowLM latents / first cond / first
flow / first EOS logit at temp 0 (LSD steps 1 and 4) and at temp 0.7 seed 123, and
Mimi PCM for the first frames.  The vectors travel to the GPU box; the reference
sources do not.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "pocket-tts.c_b200"))
import pyoracle  # noqa: E402
import synth  # noqa: E402

PROMPTS = ["Hello world!", synth.make_prompts(2, seed=123)[0], synth.make_prompts(2, seed=123)[1]]


def main():
    md = synth.model_dir(0)
    ref = pyoracle.Reference(md)
    voice = ref.voice("alba")
    out = {}
    for i, text in enumerate(PROMPTS):
        prep, words, eos_after = ref.prepare_text(text)
        ids = ref.tokenize(prep)
        out[f"p{i}_ids"] = ids
        out[f"p{i}_meta"] = np.array([words, eos_after, ref.estimate_frames(words)], np.int32)
    ids = out["p0_ids"]
    for steps in (1, 4):
        r = ref.flowlm_generate(ids, voice, 8, lsd_steps=steps, temp=0.0)
        out[f"hello_s{steps}_latents"] = r["latents"]
        out[f"hello_s{steps}_cond0"] = r["cond0"]
        out[f"hello_s{steps}_flow0"] = r["flow0"]
        out[f"hello_s{steps}_eos0"] = np.float32(r["eos0"])
    r = ref.flowlm_generate(ids, voice, 6, lsd_steps=1, temp=0.7, seed=123)
    out["hello_t07_latents"] = r["latents"]
    # EOS policy: threshold chosen between the observed logits so the stop rule fires mid-utterance
    r = ref.flowlm_generate(ids, voice, 12, lsd_steps=1, temp=0.0, eos_enabled=1, eos_threshold=-1.2,
                            eos_min_frames=1, eos_after=2)
    out["hello_eos_frames"] = np.int32(r["frames"])
    out["hello_eos_latents"] = r["latents"]
    # second prompt (about 100 chars), no voice prefix
    r = ref.flowlm_generate(out["p1_ids"], np.zeros((0, 1024), np.float32), 4, lsd_steps=1, temp=0.0)
    out["p1_novoice_latents"] = r["latents"]
    # Mimi: decode 4 frames of the hello latents (scaled); keep all 7680 samples
    sl = ref.scale_latents(out["hello_s1_latents"][:4])
    out["hello_scaled"] = sl
    out["hello_pcm"] = ref.mimi_decode(sl)
    # ptts_generate end to end (public API), 3 frames, temp 0.7 seed 123
    out["e2e_pcm"] = ref.generate("Hello world!", "alba", num_frames=3, seed=123, temp=0.7)
    ref.close()
    np.savez_compressed(ROOT / "tests" / "golden" / "reference_vectors.npz", **out)
    for k, v in out.items():
        print(k, np.asarray(v).shape)


if __name__ == "__main__":
    main()
