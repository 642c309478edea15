"""GPU suite: the tcgen05/TMEM/TMA contraction kernel (csrc/gemm_tc.cuh) against numpy, through the
C ABI test hook. Operands are rounded to the 16-bit storage type first, so the comparison isolates
the kernel (fp32 accumulation order is the only difference): tolerance 1e-5 of max|ref| for exactly
representable operands, 2e-5 for fp32 operands carried as hi+lo 16-bit pairs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(pkg, weights):
    e = pkg.Engine(weights, 0)
    yield e
    e.close()


def _round(x, bf16):
    import torch
    t = torch.from_numpy(x)
    return (t.bfloat16() if bf16 else t.half()).float().numpy()


def _ref_conv(A, B, batches, rows, N, K, taps, shift0, step):
    C = K // taps
    A3 = A.reshape(batches, rows, C).astype(np.float64)
    out = np.zeros((batches, rows, N))
    for tap in range(taps):
        sh = shift0 + tap * step
        Wt = B[:, tap * C:(tap + 1) * C].astype(np.float64)
        src = np.zeros_like(A3)
        if sh < 0:
            src[:, -sh:, :] = A3[:, :rows + sh, :]
        elif sh == 0:
            src[:] = A3
        else:
            src[:, :rows - sh, :] = A3[:, sh:, :]
        out += src @ Wt.T
    return out.reshape(batches * rows, N)


@pytest.mark.parametrize("rows,N,K,bn,bf16", [
    (128, 128, 64, 128, True), (128, 256, 512, 256, True), (300, 96, 256, 32, False), (257, 640, 128, 128, False),
    (64, 64, 1024, 64, True), (1000, 1536, 512, 256, False)])
def test_plain_gemm(eng, rows, N, K, bn, bf16):
    rng = np.random.default_rng(0)
    A = _round(rng.standard_normal((rows, K)).astype(np.float32), bf16)
    B = _round(rng.standard_normal((N, K)).astype(np.float32), bf16)
    out = eng.test_gemm(A, B, 1, rows, N, K, bf16=bf16, bn=bn)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-5


@pytest.mark.parametrize("split", [1, 2])
def test_split_operand_keeps_fp32_inputs(eng, split):
    rng = np.random.default_rng(1)
    rows, N, K = 256, 128, 512
    A = rng.standard_normal((rows, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    if split == 1:
        B = _round(B, True)     # weights exactly bf16, activations fp32 carried as hi+lo
    else:
        A = _round(A, True)
    out = eng.test_gemm(A, B, 1, rows, N, K, bf16=True, split=split, bn=128)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.abs(out - ref).max() / np.abs(ref).max() < 2e-5


def test_transposed_store_and_split_k(eng):
    rng = np.random.default_rng(2)
    rows, N, K = 3072, 48, 1024     # weights as A (M), 48 utterances as N
    A = _round(rng.standard_normal((rows, K)).astype(np.float32), True)
    B = rng.standard_normal((N, K)).astype(np.float32)
    out = eng.test_gemm(A, B, 1, rows, N, K, bf16=True, split=2, transposed=True, bn=64, split_k=4)
    ref = (A.astype(np.float64) @ B.astype(np.float64).T).T
    assert out.shape == (N, rows)
    assert np.abs(out - ref).max() / np.abs(ref).max() < 2e-5


@pytest.mark.parametrize("batches,rows,C,N,taps,shift0,step,bn", [
    (2, 200, 128, 64, 3, -2, 1, 64),      # causal Conv1d k=3
    (3, 130, 512, 512, 7, -6, 1, 256),    # dec_in geometry
    (2, 77, 256, 640, 2, 0, -1, 128),     # ConvTranspose1d 256->128 stride 5 as 2-tap phase GEMM
    (1, 16, 64, 64, 1, 0, 0, 64)])
def test_implicit_gemm_conv_taps(eng, batches, rows, C, N, taps, shift0, step, bn):
    rng = np.random.default_rng(3)
    K = C * taps
    A = _round(rng.standard_normal((batches * rows, C)).astype(np.float32), False)
    B = _round(rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K), False)
    out = eng.test_gemm(A, B, batches, rows, N, K, taps=taps, shift0=shift0, step=step, bf16=False, bn=bn)
    ref = _ref_conv(A, B, batches, rows, N, K, taps, shift0, step)
    assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-5
