"""CPU suite: the C-ABI library loads and exports every symbol include/ptts_b200.h declares
(no compute calls without a GPU), and the product refuses to run without a device."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(pkg):
    header = (ROOT / "include" / "ptts_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(ptts_b200_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 18
    lib = pkg.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ptts_b200.h but not exported"
    assert sorted(pkg.EXPORTS) == declared


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", str(ROOT / "pocket-tts.c_b200" / "libptts_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ptts_ref" not in out and "cublas" not in out.lower() and "nvrtc" not in out.lower()


def test_no_cpu_fallback_without_gpu(pkg, weights):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.EngineError, match="no CUDA device|CUDA"):
        pkg.Engine(weights, 0)


def test_prepare_text_and_frames_match_reference(pkg, golden):
    s, words, eos_after = pkg.prepare_text("Hello world!")
    assert (s, words, eos_after) == ("        Hello world!", 2, 5)
    assert pkg.estimate_frames(words) == int(golden["p0_meta"][2])
    for i, text in enumerate(pkg.synth.make_prompts(2, seed=123), start=1):
        s, words, eos_after = pkg.prepare_text(text)
        assert [words, eos_after, pkg.estimate_frames(words)] == list(golden[f"p{i}_meta"])
    s, w, e = pkg.prepare_text("  tabs\tand\nnewlines  collapse ")
    assert s == "        Tabs and newlines collapse." and w == 4 and e == 5
