import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name is not a Python identifier, so load it by path)."""
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def model_dir(synth):
    return synth.model_dir(0)


@pytest.fixture(scope="session")
def weights(synth, model_dir):
    return synth.read_safetensors(model_dir / "tts_b6369a24.safetensors")


@pytest.fixture(scope="session")
def voice(synth, model_dir):
    return synth.read_safetensors(model_dir / "embeddings" / "alba.safetensors")["audio_prompt"].reshape(-1, 1024)


@pytest.fixture(scope="session")
def oracle(weights):
    import pyoracle
    pyoracle.build(ref=False)
    return pyoracle.Oracle(weights)


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(ROOT / "tests" / "golden" / "reference_vectors.npz"))


def max_rel(a, b):
    """max |a-b| / max |b|  - the vector-normalised 'max-relative' of SURVEY.md section 7 (hard parts)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def max_rel_per_frame(a, b):
    """max over frames f of max |a_f - b_f| / max |b_f|: each [32]-latent normalised by ITS OWN peak (tighter than max_rel)."""
    a = np.asarray(a, np.float64).reshape(-1, 32)
    b = np.asarray(b, np.float64).reshape(-1, 32)
    return float((np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-30)).max())


def snr_db(x, ref):
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(10 * np.log10((ref ** 2).sum() / max(((x - ref) ** 2).sum(), 1e-300)))
