"""GPU suite: FlowLM generation through the C ABI against the oracle and the reference's golden
vectors. Tolerance (north star): latents within 1e-3 max-relative, normalised by the vector max
(SURVEY.md section 7 'parity definition'); fp32 accumulate. Token ids are bit-exact by construction
(the reference tokenizer is reused unchanged; ids come from the golden file)."""
import numpy as np
import pytest

from conftest import max_rel

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def eng(pkg, weights):
    e = pkg.Engine(weights, 0)
    yield e
    e.close()


@pytest.mark.parametrize("steps", [1, 4])
def test_batch1_latents_vs_reference_golden(eng, voice, golden, steps):
    g = eng.flowlm_generate_latents(golden["p0_ids"], voice, 8, lsd_steps=steps, temp=0.0)
    assert g["frames"] == 8
    assert max_rel(g["latents"], golden[f"hello_s{steps}_latents"]) < TOL
    assert max_rel(g["cond0"], golden[f"hello_s{steps}_cond0"]) < TOL
    assert max_rel(g["flow0"], golden[f"hello_s{steps}_flow0"]) < TOL
    assert abs(g["eos0"] - float(golden[f"hello_s{steps}_eos0"])) < 1e-3


def test_batch1_noise_path(eng, voice, golden):
    g = eng.flowlm_generate_latents(golden["p0_ids"], voice, 6, lsd_steps=1, temp=0.7, seed=123)
    assert max_rel(g["latents"], golden["hello_t07_latents"]) < TOL


def test_batch1_eos_policy(eng, voice, golden):
    g = eng.flowlm_generate_latents(golden["p0_ids"], voice, 12, lsd_steps=1, temp=0.0, eos_enabled=1,
                                    eos_threshold=-1.2, eos_min_frames=1, eos_after=2)
    assert g["frames"] == int(golden["hello_eos_frames"])
    assert max_rel(g["latents"], golden["hello_eos_latents"]) < TOL


def test_batch1_no_voice_prefix(eng, golden):
    g = eng.flowlm_generate_latents(golden["p1_ids"], np.zeros((0, 1024), np.float32), 4, lsd_steps=1, temp=0.0)
    assert max_rel(g["latents"], golden["p1_novoice_latents"]) < TOL


def test_batch1_vs_oracle_long(eng, oracle, voice, golden):
    ids = golden["p2_ids"]
    o = oracle.flowlm_generate(ids, voice, 24, lsd_steps=1, temp=0.7, seed=7)
    g = eng.flowlm_generate_latents(ids, voice, 24, lsd_steps=1, temp=0.7, seed=7)
    assert g["frames"] == o["frames"] == 24
    assert max_rel(g["latents"], o["latents"]) < TOL


def test_voice_cache_reuse_and_change(eng, voice, golden):
    a = eng.flowlm_generate_latents(golden["p0_ids"], voice, 3, temp=0.0)
    b = eng.flowlm_generate_latents(golden["p0_ids"], voice[:50], 3, temp=0.0)
    c = eng.flowlm_generate_latents(golden["p0_ids"], voice, 3, temp=0.0)
    assert np.array_equal(a["latents"], c["latents"])          # deterministic, cache refreshed correctly
    assert not np.allclose(a["latents"], b["latents"])


def test_bad_arguments_fail_loudly(eng, pkg, voice):
    with pytest.raises(pkg.EngineError):
        eng.flowlm_generate_latents(np.zeros(0, np.int32), voice, 3)
    with pytest.raises(pkg.EngineError):
        eng.flowlm_generate_latents(np.array([1, 2], np.int32), voice, 0)


@pytest.mark.parametrize("precise,tol", [(1, 1e-3), (0, 3e-2)])
def test_batched_tensor_core_path_vs_oracle(eng, pkg, oracle, voice, golden, precise, tol):
    """Ragged batch (different token counts and frame counts, one padded slot group) through the
    tcgen05 path; every utterance must match the oracle run one at a time."""
    toks = [golden["p0_ids"], golden["p1_ids"], golden["p2_ids"], golden["p1_ids"][:7], golden["p2_ids"][3:20]]
    frames = [6, 4, 5, 3, 6]
    seeds = [11, 12, 13, 14, 15]
    out = eng.generate_batch(toks, voice, frames, seeds=seeds, params=pkg.default_params(temp=0.7, precise=precise),
                             want_audio=False)
    assert list(out["frames"]) == frames
    for i, (t, f, s) in enumerate(zip(toks, frames, seeds)):
        o = oracle.flowlm_generate(t, voice, f, lsd_steps=1, temp=0.7, seed=s)
        assert max_rel(out["latents"][i, :f], o["latents"]) < tol, f"utterance {i}"
        assert not out["latents"][i, f:].any()


def test_batched_eos_stops_each_utterance_independently(eng, pkg, oracle, voice, golden):
    toks = [golden["p0_ids"], golden["p1_ids"], golden["p2_ids"]]
    p = pkg.default_params(temp=0.0, eos_enabled=1, eos_threshold=-1.2, eos_after=2)
    out = eng.generate_batch(toks, voice, 12, seeds=1, params=p, want_audio=False)
    for i, t in enumerate(toks):
        o = oracle.flowlm_generate(t, voice, 12, lsd_steps=1, temp=0.0, eos_enabled=1, eos_threshold=-1.2, eos_after=2)
        assert int(out["frames"][i]) == o["frames"], f"utterance {i}"
        assert max_rel(out["latents"][i, :o["frames"]], o["latents"]) < TOL


def test_batched_multi_step_lsd_and_wider_batch(eng, pkg, oracle, voice, golden):
    """40 utterances (attention groups of 2, one partial row tile) with 4 Euler steps per frame: covers
    the MID phase of the persistent step kernel and the split-K folds at a second batch size."""
    base = [golden["p0_ids"], golden["p1_ids"], golden["p2_ids"], golden["p1_ids"][:9], golden["p2_ids"][2:17]]
    toks = [base[i % 5] for i in range(40)]
    frames = [3 + (i % 3) for i in range(40)]
    seeds = [100 + i for i in range(40)]
    out = eng.generate_batch(toks, voice, frames, seeds=seeds, params=pkg.default_params(lsd_steps=4, temp=0.7), want_audio=False)
    assert list(out["frames"]) == frames
    for i in (0, 1, 2, 3, 4, 17, 39):
        o = oracle.flowlm_generate(toks[i], voice, frames[i], lsd_steps=4, temp=0.7, seed=seeds[i])
        assert max_rel(out["latents"][i, :frames[i]], o["latents"]) < TOL, f"utterance {i}"


def test_batched_run_is_repeatable(eng, pkg, voice, golden):
    """The persistent kernel keeps no state between launches: two runs of one staged batch agree bit for bit."""
    toks = [golden["p0_ids"], golden["p1_ids"], golden["p2_ids"]] * 6
    a = eng.generate_batch(toks, voice, 5, seeds=3, params=pkg.default_params(temp=0.7), want_audio=False)
    b = eng.generate_batch(toks, voice, 5, seeds=3, params=pkg.default_params(temp=0.7), want_audio=False)
    assert np.array_equal(a["latents"], b["latents"])


def test_batched_long_voice_prefix_beyond_the_shared_memory_tile(eng, pkg, oracle, voice, golden):
    """Voice prefix of 185 frames: the first 128 prefix keys are served from the shared-memory tile of the
    persistent kernel's attention phase, the remaining 57 straight from the prefix cache in L2."""
    cond = np.concatenate([voice, voice[:60] * 0.7], axis=0).astype(np.float32)
    toks = [golden["p0_ids"], golden["p1_ids"], golden["p2_ids"], golden["p2_ids"][2:15]]
    out = eng.generate_batch(toks, cond, 5, seeds=[21, 22, 23, 24], params=pkg.default_params(temp=0.7), want_audio=False)
    for i, t in enumerate(toks):
        o = oracle.flowlm_generate(t, cond, 5, lsd_steps=1, temp=0.7, seed=21 + i)
        assert max_rel(out["latents"][i, :5], o["latents"]) < TOL, f"utterance {i}"
    g = eng.flowlm_generate_latents(golden["p0_ids"], cond, 5, lsd_steps=1, temp=0.7, seed=21)   # batch-1 path, same prefix
    assert max_rel(g["latents"], out["latents"][0, :5]) < TOL


def test_full_batch_of_256_spot_check(eng, pkg, oracle, voice, golden):
    """The bench-size batch (two row tiles of 128, 9 CTAs per head): a few utterances against the oracle."""
    base = [golden["p0_ids"], golden["p1_ids"], golden["p2_ids"], golden["p1_ids"][:9], golden["p2_ids"][2:17]]
    toks = [base[i % 5] for i in range(256)]
    seeds = [1000 + i for i in range(256)]
    out = eng.generate_batch(toks, voice, 4, seeds=seeds, params=pkg.default_params(temp=0.7), want_audio=False)
    for i in (0, 127, 128, 201, 255):
        o = oracle.flowlm_generate(toks[i], voice, 4, lsd_steps=1, temp=0.7, seed=seeds[i])
        assert max_rel(out["latents"][i, :4], o["latents"]) < TOL, f"utterance {i}"


def test_batch_of_512_spot_check(eng, pkg, oracle, voice, golden):
    """A batch beyond the bench size (four row tiles of 128, two attention pairs per warp): a few utterances against the
    oracle, and the padded tail (n = 500 -> 512 slots) must not disturb them; more than 1024 utterances are refused."""
    base = [golden["p0_ids"], golden["p1_ids"], golden["p2_ids"], golden["p1_ids"][:9], golden["p2_ids"][2:17]]
    toks = [base[i % 5] for i in range(500)]
    seeds = [3000 + i for i in range(500)]
    out = eng.generate_batch(toks, voice, 3, seeds=seeds, params=pkg.default_params(temp=0.7), want_audio=False)
    for i in (0, 255, 256, 383, 499):
        o = oracle.flowlm_generate(toks[i], voice, 3, lsd_steps=1, temp=0.7, seed=seeds[i])
        assert max_rel(out["latents"][i, :3], o["latents"]) < TOL, f"utterance {i}"
    with pytest.raises(pkg.EngineError):
        eng.generate_batch(toks + toks + toks[:40], voice, 2, seeds=1, params=pkg.default_params(temp=0.7), want_audio=False)


def test_batched_long_prompts_span_several_prefill_tiles(eng, pkg, oracle, voice, golden):
    """Prompts of 70 and 141 tokens: the tensor-core prefill attention works on 64-row query tiles (two and three tiles here,
    each appending its own rows to the cache), next to a short prompt in the same batch."""
    long1 = np.concatenate([golden["p2_ids"]] * 3)[:70].astype(np.int32)
    long2 = np.concatenate([golden["p0_ids"], golden["p1_ids"], golden["p2_ids"]] * 6)[:141].astype(np.int32)
    toks = [long1, golden["p1_ids"], long2]
    out = eng.generate_batch(toks, voice, 4, seeds=[41, 42, 43], params=pkg.default_params(temp=0.7), want_audio=False)
    for i, t in enumerate(toks):
        o = oracle.flowlm_generate(t, voice, 4, lsd_steps=1, temp=0.7, seed=41 + i)
        assert max_rel(out["latents"][i, :4], o["latents"]) < TOL, f"utterance {i} ({len(t)} tokens)"
