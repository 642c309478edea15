"""Synthetic Pocket-TTS checkpoint / voice / tokenizer / prompt generator.

There is no network, so the Hugging Face weights the reference downloads
(download_model.py:16-17) are replaced by random-init weights of the named
architecture, written in the exact on-disk layout the reference loads
(ptts_flowlm.c:866-957, ptts_mimi.c:378-517, ptts.c:293-364).

Every weight is rounded to a bf16-representable value so that the bf16 device
copy used by the B200 engine is lossless against the f32 oracle.

Everything here is deterministic in (seed) and is regenerated on the GPU box;
only the tokenizer model is a committed fixture (tests/golden/tokenizer.model).
"""
from __future__ import annotations

import json
import os
import struct
from pathlib import Path

import numpy as np

D_MODEL = 1024
N_HEADS = 16
N_LAYERS = 6
FFN = 4096
LATENT = 32
FLOW_DIM = 512
FLOW_DEPTH = 6
VOCAB = 4000
MIMI_D = 512
MIMI_LAYERS = 2
MIMI_FFN = 2048
FRAME_SAMPLES = 1920
FRAME_RATE = 12.5

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round f32 to the nearest bf16-representable f32 (round-to-nearest-even)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return r.astype(np.uint16).reshape(x.shape)


def make_weights(seed: int = 0, full_mantissa: bool = False) -> "dict[str, np.ndarray]":
    """Return {tensor name: f32 ndarray} for FlowLM + Mimi decoder.

    Default: every value bf16-representable. full_mantissa=True keeps all 24 significand bits (an F32
    checkpoint as the reference accepts it, ptts_safetensors.c:294-334): the engine's bf16 / fp16 weight
    rounding is then visible against an oracle that runs on the unrounded values."""
    rng = np.random.default_rng(seed)
    W: dict[str, np.ndarray] = {}
    bf16_round = (lambda x: np.ascontiguousarray(x, dtype=np.float32)) if full_mantissa else globals()["bf16_round"]

    def lin(name, out, inn, scale=1.0):
        W[name] = bf16_round(rng.standard_normal((out, inn), dtype=np.float32) * (scale / np.sqrt(inn)))

    def vec(name, n, mean=0.0, std=0.02):
        W[name] = bf16_round(mean + std * rng.standard_normal(n, dtype=np.float32))

    p = "flow_lm."
    W[p + "conditioner.embed.weight"] = bf16_round(
        rng.standard_normal((VOCAB + 1, D_MODEL), dtype=np.float32) * 0.7)
    lin(p + "speaker_proj_weight", D_MODEL, 512)
    vec(p + "emb_std", LATENT, 1.0, 0.1)
    vec(p + "emb_mean", LATENT, 0.0, 0.1)
    vec(p + "bos_emb", LATENT, 0.0, 1.0)
    lin(p + "input_linear.weight", D_MODEL, LATENT)
    vec(p + "out_norm.weight", D_MODEL, 1.0, 0.05)
    vec(p + "out_norm.bias", D_MODEL, 0.0, 0.05)
    lin(p + "out_eos.weight", 1, D_MODEL)
    vec(p + "out_eos.bias", 1, 0.0, 0.5)
    for l in range(N_LAYERS):
        q = f"{p}transformer.layers.{l}."
        lin(q + "self_attn.in_proj.weight", 3 * D_MODEL, D_MODEL)
        lin(q + "self_attn.out_proj.weight", D_MODEL, D_MODEL, 0.5)
        vec(q + "norm1.weight", D_MODEL, 1.0, 0.05)
        vec(q + "norm1.bias", D_MODEL, 0.0, 0.05)
        vec(q + "norm2.weight", D_MODEL, 1.0, 0.05)
        vec(q + "norm2.bias", D_MODEL, 0.0, 0.05)
        lin(q + "linear1.weight", FFN, D_MODEL)
        lin(q + "linear2.weight", D_MODEL, FFN, 0.5)
    f = p + "flow_net."
    lin(f + "cond_embed.weight", FLOW_DIM, D_MODEL)
    vec(f + "cond_embed.bias", FLOW_DIM)
    lin(f + "input_proj.weight", FLOW_DIM, LATENT)
    vec(f + "input_proj.bias", FLOW_DIM)
    for t in range(2):
        q = f"{f}time_embed.{t}."
        lin(q + "mlp.0.weight", FLOW_DIM, 256)
        vec(q + "mlp.0.bias", FLOW_DIM)
        lin(q + "mlp.2.weight", FLOW_DIM, FLOW_DIM)
        vec(q + "mlp.2.bias", FLOW_DIM)
        vec(q + "mlp.3.alpha", FLOW_DIM, 1.0, 0.05)
        W[q + "freqs"] = bf16_round(
            np.exp(-np.log(10000.0) * np.arange(128, dtype=np.float32) / 128.0).astype(np.float32))
    for b in range(FLOW_DEPTH):
        q = f"{f}res_blocks.{b}."
        vec(q + "in_ln.weight", FLOW_DIM, 1.0, 0.05)
        vec(q + "in_ln.bias", FLOW_DIM, 0.0, 0.05)
        lin(q + "mlp.0.weight", FLOW_DIM, FLOW_DIM)
        vec(q + "mlp.0.bias", FLOW_DIM)
        lin(q + "mlp.2.weight", FLOW_DIM, FLOW_DIM)
        vec(q + "mlp.2.bias", FLOW_DIM)
        lin(q + "adaLN_modulation.1.weight", 3 * FLOW_DIM, FLOW_DIM, 0.5)
        vec(q + "adaLN_modulation.1.bias", 3 * FLOW_DIM, 0.0, 0.1)
    lin(f + "final_layer.linear.weight", LATENT, FLOW_DIM)
    vec(f + "final_layer.linear.bias", LATENT)
    lin(f + "final_layer.adaLN_modulation.1.weight", 2 * FLOW_DIM, FLOW_DIM, 0.5)
    vec(f + "final_layer.adaLN_modulation.1.bias", 2 * FLOW_DIM, 0.0, 0.1)

    m = "mimi."
    W[m + "quantizer.output_proj.weight"] = bf16_round(
        rng.standard_normal((MIMI_D, LATENT, 1), dtype=np.float32) / np.sqrt(LATENT))
    W[m + "upsample.convtr.weight"] = bf16_round(
        rng.standard_normal((MIMI_D, 1, 32), dtype=np.float32) * 0.7)

    def conv(name, out, inn, k, scale=1.0):
        W[name + ".weight"] = bf16_round(
            rng.standard_normal((out, inn, k), dtype=np.float32) * (scale / np.sqrt(inn * k)))
        vec(name + ".bias", out)

    def convtr(name, inn, out, k, stride):
        W[name + ".weight"] = bf16_round(
            rng.standard_normal((inn, out, k), dtype=np.float32) * (1.0 / np.sqrt(inn * k / stride)))
        vec(name + ".bias", out)

    d = m + "decoder.model."
    conv(d + "0.conv", 512, 512, 7)
    convtr(d + "2.convtr", 512, 256, 12, 6)
    conv(d + "3.block.1.conv", 128, 256, 3)
    conv(d + "3.block.3.conv", 256, 128, 1, 0.5)
    convtr(d + "5.convtr", 256, 128, 10, 5)
    conv(d + "6.block.1.conv", 64, 128, 3)
    conv(d + "6.block.3.conv", 128, 64, 1, 0.5)
    convtr(d + "8.convtr", 128, 64, 8, 4)
    conv(d + "9.block.1.conv", 32, 64, 3)
    conv(d + "9.block.3.conv", 64, 32, 1, 0.5)
    # final conv scaled down so |pcm| < 1 (the int16 WAV must not clip, SURVEY.md §7 trap b)
    conv(d + "11.conv", 1, 64, 3, 0.08)
    for l in range(MIMI_LAYERS):
        q = f"{m}decoder_transformer.transformer.layers.{l}."
        lin(q + "self_attn.in_proj.weight", 3 * MIMI_D, MIMI_D)
        lin(q + "self_attn.out_proj.weight", MIMI_D, MIMI_D)
        vec(q + "norm1.weight", MIMI_D, 1.0, 0.05)
        vec(q + "norm1.bias", MIMI_D, 0.0, 0.05)
        vec(q + "norm2.weight", MIMI_D, 1.0, 0.05)
        vec(q + "norm2.bias", MIMI_D, 0.0, 0.05)
        lin(q + "linear1.weight", MIMI_FFN, MIMI_D)
        lin(q + "linear2.weight", MIMI_D, MIMI_FFN)
        vec(q + "layer_scale_1.scale", MIMI_D, 0.1, 0.01)
        vec(q + "layer_scale_2.scale", MIMI_D, 0.1, 0.01)
    # Voice-cloning side (shapes from ptts.c:914-944, 973): SEANet encoder, encoder transformer,
    # learnt stride-16 downsampler. Appended AFTER everything else so the tensors above keep their values.
    en = m + "encoder.model."
    conv(en + "0.conv", 64, 1, 7)
    idx, ch = 1, 64
    for ratio in (4, 5, 6):
        conv(f"{en}{idx}.block.1.conv", ch // 2, ch, 3)
        conv(f"{en}{idx}.block.3.conv", ch, ch // 2, 1, 0.5)
        idx += 2
        conv(f"{en}{idx}.conv", 2 * ch, ch, 2 * ratio)
        idx += 1
        ch *= 2
    idx += 1
    conv(f"{en}{idx}.conv", 512, 512, 3)
    for l in range(MIMI_LAYERS):
        q = f"{m}encoder_transformer.transformer.layers.{l}."
        lin(q + "self_attn.in_proj.weight", 3 * MIMI_D, MIMI_D)
        lin(q + "self_attn.out_proj.weight", MIMI_D, MIMI_D)
        vec(q + "norm1.weight", MIMI_D, 1.0, 0.05)
        vec(q + "norm1.bias", MIMI_D, 0.0, 0.05)
        vec(q + "norm2.weight", MIMI_D, 1.0, 0.05)
        vec(q + "norm2.bias", MIMI_D, 0.0, 0.05)
        lin(q + "linear1.weight", MIMI_FFN, MIMI_D)
        lin(q + "linear2.weight", MIMI_D, MIMI_FFN)
        vec(q + "layer_scale_1.scale", MIMI_D, 0.1, 0.01)
        vec(q + "layer_scale_2.scale", MIMI_D, 0.1, 0.01)
    W[m + "downsample.conv.conv.weight"] = bf16_round(
        rng.standard_normal((MIMI_D, MIMI_D, 32), dtype=np.float32) / np.sqrt(MIMI_D * 32))
    return W


def make_voice(seed: int = 0, frames: int = 125) -> np.ndarray:
    rng = np.random.default_rng(seed + 7919)
    return bf16_round(0.5 * rng.standard_normal((1, frames, D_MODEL), dtype=np.float32))


def write_safetensors(path, tensors: "dict[str, np.ndarray]", dtype: str = "BF16") -> None:
    """Minimal safetensors writer (8-byte LE header size, JSON header, raw data)."""
    header = {}
    off = 0
    blobs = []
    for name, arr in tensors.items():
        if dtype == "BF16":
            raw = f32_to_bf16_bits(arr).tobytes()
        else:
            raw = np.ascontiguousarray(arr, dtype=np.float32).tobytes()
        header[name] = {"dtype": dtype, "shape": list(arr.shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
        blobs.append(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as fh:
        fh.write(struct.pack("<Q", len(hj)))
        fh.write(hj)
        for b in blobs:
            fh.write(b)


def read_safetensors(path) -> "dict[str, np.ndarray]":
    """Read F32/BF16/F16 safetensors into f32 arrays (mirrors ptts_safetensors.c:279-337)."""
    with open(path, "rb") as fh:
        n = struct.unpack("<Q", fh.read(8))[0]
        header = json.loads(fh.read(n))
        base = 8 + n
        out = {}
        for name, meta in header.items():
            if name == "__metadata__":
                continue
            s, e = meta["data_offsets"]
            fh.seek(base + s)
            raw = fh.read(e - s)
            dt = meta["dtype"]
            if dt == "F32":
                a = np.frombuffer(raw, dtype=np.float32).copy()
            elif dt == "BF16":
                a = (np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
            elif dt == "F16":
                a = np.frombuffer(raw, dtype=np.float16).astype(np.float32)
            else:
                raise ValueError(f"unsupported dtype {dt} for {name}")
            out[name] = a.reshape(meta["shape"])
        return out


_SYL_A = ["b", "br", "c", "ch", "d", "dr", "f", "fl", "g", "gr", "h", "j", "k", "l", "m", "n", "p",
          "pl", "qu", "r", "s", "sh", "st", "t", "th", "tr", "v", "w", "y", "z"]
_SYL_B = ["a", "e", "i", "o", "u", "ai", "ea", "ou", "io", "ee"]
_SYL_C = ["", "", "n", "r", "s", "t", "l", "m", "nd", "st", "ng", "ck"]


def word_list(n: int = 1800, seed: int = 11) -> "list[str]":
    rng = np.random.default_rng(seed)
    words = []
    seen = set()
    while len(words) < n:
        k = int(rng.integers(1, 4))
        w = "".join(_SYL_A[int(rng.integers(len(_SYL_A)))] + _SYL_B[int(rng.integers(len(_SYL_B)))] +
                    _SYL_C[int(rng.integers(len(_SYL_C)))] for _ in range(k))
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


def make_prompts(n: int, seed: int = 123, lo: int = 95, hi: int = 105) -> "list[str]":
    """Seeded word-salad sentences of lo..hi characters."""
    rng = np.random.default_rng(seed)
    words = word_list()
    out = []
    for _ in range(n):
        target = int(rng.integers(lo, hi + 1))
        s = ""
        while True:
            w = words[int(rng.integers(len(words)))]
            cand = (s + " " + w) if s else w.capitalize()
            if len(cand) + 1 > target:
                break
            s = cand
        out.append(s + ".")
    return out


def train_tokenizer(out_path) -> None:
    """Train the 4000-piece unigram tokenizer fixture (run once; the result is committed)."""
    import sentencepiece as spm
    import tempfile
    words = word_list(8000)
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as td:
        corpus = Path(td) / "corpus.txt"
        with open(corpus, "w") as fh:
            for _ in range(40000):
                k = int(rng.integers(6, 20))
                ws = [words[int(rng.integers(len(words)))] for _ in range(k)]
                fh.write(" ".join(ws).capitalize() + ".\n")
            fh.write("Hello world!\nHello world.\n")
        prefix = str(Path(td) / "tok")
        spm.SentencePieceTrainer.train(
            input=str(corpus), model_prefix=prefix, vocab_size=VOCAB, model_type="unigram",
            normalization_rule_name="nmt_nfkc", add_dummy_prefix=True, num_threads=1,
            character_coverage=1.0, input_sentence_size=40000, shuffle_input_sentence=False,
            minloglevel=2)
        data = Path(prefix + ".model").read_bytes()
    Path(out_path).write_bytes(data)


def model_dir(seed: int = 0, voice_frames: int = 125, root=None) -> Path:
    """Materialise (once) the synthetic model directory and return its path.

    Layout (ptts.c:82-145, 154-213): tts_b6369a24.safetensors, tokenizer.model,
    embeddings/alba.safetensors.
    """
    root = Path(root or os.environ.get("PTTS_B200_MODEL_CACHE", "/tmp/ptts_b200_models"))
    d = root / f"synth2_s{seed}_v{voice_frames}"   # synth2: + voice-cloning encoder tensors
    done = d / ".complete"
    if done.exists():
        return d
    (d / "embeddings").mkdir(parents=True, exist_ok=True)
    write_safetensors(d / "tts_b6369a24.safetensors", make_weights(seed), "BF16")
    write_safetensors(d / "embeddings" / "alba.safetensors",
                      {"audio_prompt": make_voice(seed, voice_frames)}, "BF16")
    tok = GOLDEN_DIR / "tokenizer.model"
    if not tok.exists():
        train_tokenizer(tok)
    (d / "tokenizer.model").write_bytes(tok.read_bytes())
    done.write_text("ok\n")
    return d


if __name__ == "__main__":
    import sys
    print(model_dir(int(sys.argv[1]) if len(sys.argv) > 1 else 0))
