"""pocket-tts B200 engine - Python host mirror of the C ABI (include/ptts_b200.h).

The compute path is libptts_b200.so (hand-written sm_100a CUDA behind a C ABI).  This module is
only the ctypes binding plus the host-side glue the reference keeps in C (text preparation,
tokenisation, frame estimate: ptts.c:219-291), so that tests and bench.py read like the
reference's own call sequence (ptts.c:1011-1161).

There is NO fallback: if the shared library is missing or no B200 is present, construction
raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import synth  # noqa: F401

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libptts_b200.so"

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)

FRAME_SAMPLES = 1920
SAMPLE_RATE = 24000


class _Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("dtype", C.c_int), ("numel", C.c_int64)]


class GenParams(C.Structure):
    _fields_ = [("lsd_steps", C.c_int), ("temp", C.c_float), ("noise_clamp", C.c_float), ("eos_enabled", C.c_int),
                ("eos_threshold", C.c_float), ("eos_min_frames", C.c_int), ("eos_after", C.c_int), ("precise", C.c_int)]


EXPORTS = [
    "ptts_b200_create", "ptts_b200_destroy", "ptts_b200_last_error", "ptts_b200_flowlm_generate_latents",
    "ptts_b200_flowlm_scale_latents", "ptts_b200_mimi_decode", "ptts_b200_mimi_decode_batch",
    "ptts_b200_generate_batch", "ptts_b200_batch_stage", "ptts_b200_batch_run", "ptts_b200_batch_fetch",
    "ptts_b200_sync", "ptts_b200_timer_start", "ptts_b200_timer_stop", "ptts_b200_profile_enable",
    "ptts_b200_profile_read", "ptts_b200_profile_work", "ptts_b200_profile_class_name", "ptts_b200_launch_count", "ptts_b200_test_gemm",
    "ptts_b200_stream_begin", "ptts_b200_stream_next", "ptts_b200_mimi_forward_one", "ptts_b200_mimi_stage",
    "ptts_b200_mimi_run", "ptts_b200_mimi_fetch", "ptts_b200_step_kernel_ms", "ptts_b200_pool_create", "ptts_b200_pool_destroy",
    "ptts_b200_pool_size", "ptts_b200_pool_engine", "ptts_b200_pool_generate", "ptts_b200_voice_encode",
]


def load_library():
    """dlopen the product library; raise (never fall back) when it is missing."""
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} not built: run `python __graft_entry__.py` (make -C pocket-tts.c_b200)")
    L = C.CDLL(str(LIB_PATH))
    vp = C.c_void_p
    L.ptts_b200_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(_Tensor), C.c_int]
    L.ptts_b200_destroy.argtypes = [vp]
    L.ptts_b200_destroy.restype = None
    L.ptts_b200_last_error.restype = C.c_char_p
    L.ptts_b200_flowlm_generate_latents.argtypes = [
        vp, _ip, C.c_int, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int64, C.c_int, C.c_float,
        C.c_int, C.c_int, _fp, _ip, _fp, _fp, _fp]
    L.ptts_b200_flowlm_scale_latents.argtypes = [vp, _fp, C.c_int, _fp]
    L.ptts_b200_mimi_decode.argtypes = [vp, _fp, C.c_int, _fp, _ip]
    L.ptts_b200_mimi_decode_batch.argtypes = [vp, _fp, C.c_int, C.c_int, _fp]
    L.ptts_b200_voice_encode.argtypes = [vp, _fp, C.c_int, _fp, C.c_int, _ip]
    stage_args = [vp, C.c_int, _ip, _ip, _fp, C.c_int, _ip, _lp, _ip, C.POINTER(GenParams)]
    L.ptts_b200_generate_batch.argtypes = stage_args + [_fp, _ip, _fp]
    L.ptts_b200_batch_stage.argtypes = stage_args
    L.ptts_b200_batch_run.argtypes = [vp, C.c_int]
    L.ptts_b200_batch_fetch.argtypes = [vp, _fp, _ip, _fp]
    L.ptts_b200_sync.argtypes = [vp]
    L.ptts_b200_timer_start.argtypes = [vp]
    L.ptts_b200_timer_stop.argtypes = [vp, _fp]
    L.ptts_b200_profile_enable.argtypes = [vp, C.c_int]
    L.ptts_b200_profile_read.argtypes = [vp, C.c_int, _fp, _lp, _ip]
    L.ptts_b200_profile_work.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ptts_b200_profile_class_name.argtypes = [C.c_int]
    L.ptts_b200_profile_class_name.restype = C.c_char_p
    L.ptts_b200_launch_count.argtypes = [vp]
    L.ptts_b200_launch_count.restype = C.c_int64
    L.ptts_b200_test_gemm.argtypes = [vp, _fp, _fp, _fp] + [C.c_int] * 12
    L.ptts_b200_stream_begin.argtypes = [vp, _ip, C.c_int, _fp, C.c_int, C.c_int, C.c_int64, C.POINTER(GenParams)]
    L.ptts_b200_stream_next.argtypes = [vp, C.c_int, _fp, _ip, _ip]
    L.ptts_b200_mimi_forward_one.argtypes = [vp, _fp, _fp]
    L.ptts_b200_mimi_stage.argtypes = [vp, _fp, C.c_int, C.c_int]
    L.ptts_b200_mimi_run.argtypes = [vp]
    L.ptts_b200_mimi_fetch.argtypes = [vp, _fp]
    L.ptts_b200_step_kernel_ms.argtypes = [vp, _fp, _ip]
    L.ptts_b200_pool_create.argtypes = [C.POINTER(vp), _ip, C.c_int, C.POINTER(_Tensor), C.c_int]
    L.ptts_b200_pool_destroy.argtypes = [vp]
    L.ptts_b200_pool_destroy.restype = None
    L.ptts_b200_pool_size.argtypes = [vp]
    L.ptts_b200_pool_engine.argtypes = [vp, C.c_int]
    L.ptts_b200_pool_engine.restype = vp
    L.ptts_b200_pool_generate.argtypes = [vp] + stage_args[1:] + [C.c_int, _fp, _ip, _fp]
    return L


def _f(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


class EngineError(RuntimeError):
    pass


class Engine:
    """One engine per GPU (weights replicated; utterances shard across engines, no collective)."""

    def __init__(self, weights: "dict[str, np.ndarray]", device: int = 0):
        self.h = None
        self.lib = load_library()
        self._keep = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
        names = [k.encode() for k in self._keep]
        tab = (_Tensor * len(names))()
        for i, (k, v) in enumerate(self._keep.items()):
            tab[i] = _Tensor(names[i], v.ctypes.data, 0, v.size)
        self._names = names
        h = C.c_void_p()
        if self.lib.ptts_b200_create(C.byref(h), device, tab, len(names)) != 0:
            raise EngineError("ptts_b200_create: " + self.error())
        self.h = h
        self._keep = None  # the engine owns device copies; host weights are no longer needed

    @classmethod
    def from_dir(cls, model_dir, device: int = 0):
        """Mirror of ptts_load_dir (ptts.c:370-412): weights file in the reference's on-disk layout."""
        w = synth.read_safetensors(Path(model_dir) / "tts_b6369a24.safetensors")
        return cls(w, device)

    def error(self) -> str:
        return (self.lib.ptts_b200_last_error() or b"").decode()

    def close(self):
        if getattr(self, "h", None):
            self.lib.ptts_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise EngineError(f"{what}: {self.error()}")

    # ---- drop-in single-utterance entries (same argument meaning as the reference)
    def flowlm_generate_latents(self, tokens, cond, max_frames, lsd_steps=1, temp=0.0, noise_clamp=0.0, seed=123,
                                eos_enabled=0, eos_threshold=-4.0, eos_min_frames=1, eos_after=0):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        cond = np.ascontiguousarray(cond, dtype=np.float32).reshape(-1, 1024)
        lat = np.zeros((max_frames, 32), np.float32)
        used = C.c_int(0)
        eos0 = C.c_float(0)
        cond0 = np.zeros(1024, np.float32)
        flow0 = np.zeros(32, np.float32)
        self._ck(self.lib.ptts_b200_flowlm_generate_latents(
            self.h, _i(tokens), len(tokens), _f(cond) if cond.shape[0] else None, cond.shape[0], max_frames, lsd_steps,
            temp, noise_clamp, seed, eos_enabled, eos_threshold, eos_min_frames, eos_after, _f(lat), C.byref(used),
            C.byref(eos0), _f(cond0), _f(flow0)), "flowlm_generate_latents")
        n = used.value
        return dict(latents=lat[:n], frames=n, eos0=eos0.value, cond0=cond0, flow0=flow0)

    def scale_latents(self, lat):
        lat = np.ascontiguousarray(lat, np.float32).reshape(-1, 32)
        out = np.empty_like(lat)
        self._ck(self.lib.ptts_b200_flowlm_scale_latents(self.h, _f(lat), lat.shape[0], _f(out)), "scale_latents")
        return out

    def mimi_decode(self, latents):
        latents = np.ascontiguousarray(latents, np.float32).reshape(-1, 32)
        F = latents.shape[0]
        out = np.zeros(F * FRAME_SAMPLES, np.float32)
        n = C.c_int(0)
        self._ck(self.lib.ptts_b200_mimi_decode(self.h, _f(latents), F, _f(out), C.byref(n)), "mimi_decode")
        return out[: n.value]

    def mimi_forward_one(self, latent):
        """ptts_mimi_forward_one (ptts_mimi.c:550-564): one scaled latent[32] -> transformer embedding[512]."""
        latent = np.ascontiguousarray(latent, np.float32).reshape(32)
        out = np.zeros(512, np.float32)
        self._ck(self.lib.ptts_b200_mimi_forward_one(self.h, _f(latent), _f(out)), "mimi_forward_one")
        return out

    # ---- additive batched entries
    def voice_encode(self, pcm):
        """24 kHz mono PCM -> conditioning rows [frames, 1024] (ptts_b200_voice_encode): a cond_prefix for generate calls."""
        pcm = np.ascontiguousarray(pcm, np.float32).reshape(-1)
        F = (pcm.size + FRAME_SAMPLES - 1) // FRAME_SAMPLES
        out = np.zeros((F, 1024), np.float32)
        n = C.c_int(0)
        self._ck(self.lib.ptts_b200_voice_encode(self.h, _f(pcm), pcm.size, _f(out), F, C.byref(n)), "voice_encode")
        return out[:n.value]

    def mimi_stage(self, latents):
        latents = np.ascontiguousarray(latents, np.float32)
        self._mimi_staged = (latents.shape[0], latents.shape[1])
        self._ck(self.lib.ptts_b200_mimi_stage(self.h, _f(latents), latents.shape[0], latents.shape[1]), "mimi_stage")

    def mimi_run(self):
        self._ck(self.lib.ptts_b200_mimi_run(self.h), "mimi_run")

    def mimi_fetch(self, out=None):
        n, F = self._mimi_staged
        if out is None:
            out = np.zeros((n, F * FRAME_SAMPLES), np.float32)
        self._ck(self.lib.ptts_b200_mimi_fetch(self.h, _f(out)), "mimi_fetch")
        return out

    def mimi_decode_batch(self, latents, out=None):
        latents = np.ascontiguousarray(latents, np.float32)
        n, F = latents.shape[0], latents.shape[1]
        if out is None:
            out = np.zeros((n, F * FRAME_SAMPLES), np.float32)
        self._ck(self.lib.ptts_b200_mimi_decode_batch(self.h, _f(latents), n, F, _f(out)), "mimi_decode_batch")
        return out

    @staticmethod
    def _pack(token_lists, max_frames, seeds, eos_after):
        n = len(token_lists)
        offs = np.zeros(n + 1, np.int32)
        offs[1:] = np.cumsum([len(t) for t in token_lists])
        flat = np.concatenate([np.asarray(t, np.int32) for t in token_lists]).astype(np.int32)
        mf = np.ascontiguousarray(np.broadcast_to(np.asarray(max_frames, np.int32), (n,)))
        sd = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, np.int64), (n,)))
        ea = None if eos_after is None else np.ascontiguousarray(np.broadcast_to(np.asarray(eos_after, np.int32), (n,)))
        return n, flat, offs, mf, sd, ea

    def stage(self, token_lists, cond, max_frames, seeds=123, eos_after=None, params: GenParams | None = None):
        params = params or default_params()
        cond = np.ascontiguousarray(cond, dtype=np.float32).reshape(-1, 1024)
        n, flat, offs, mf, sd, ea = self._pack(token_lists, max_frames, seeds, eos_after)
        self._staged = (n, int(mf.max()))
        self._ck(self.lib.ptts_b200_batch_stage(self.h, n, _i(flat), _i(offs), _f(cond) if cond.shape[0] else None,
                                                cond.shape[0], _i(mf), sd.ctypes.data_as(_lp), _i(ea), C.byref(params)),
                 "batch_stage")

    def run(self, with_mimi=True):
        self._ck(self.lib.ptts_b200_batch_run(self.h, 1 if with_mimi else 0), "batch_run")

    def fetch(self, want_latents=True, want_audio=True, audio_out=None):
        n, F = self._staged
        lat = np.zeros((n, F, 32), np.float32) if want_latents else None
        used = np.zeros(n, np.int32)
        pcm = audio_out if audio_out is not None else (np.zeros((n, F * FRAME_SAMPLES), np.float32) if want_audio else None)
        self._ck(self.lib.ptts_b200_batch_fetch(self.h, _f(lat), _i(used), _f(pcm)), "batch_fetch")
        return dict(latents=lat, frames=used, audio=pcm)

    def generate_batch(self, token_lists, cond, max_frames, seeds=123, eos_after=None, params: GenParams | None = None,
                       want_audio=True, audio_out=None):
        """ptts_b200_generate_batch: host buffers in, host buffers out (H2D, kernels and the chunk-pipelined D2H
        of the PCM all inside the one C call)."""
        params = params or default_params()
        cond = np.ascontiguousarray(cond, dtype=np.float32).reshape(-1, 1024)
        n, flat, offs, mf, sd, ea = self._pack(token_lists, max_frames, seeds, eos_after)
        F = int(mf.max())
        self._staged = (n, F)
        lat = np.zeros((n, F, 32), np.float32)
        used = np.zeros(n, np.int32)
        pcm = audio_out if audio_out is not None else (np.zeros((n, F * FRAME_SAMPLES), np.float32) if want_audio else None)
        self._ck(self.lib.ptts_b200_generate_batch(self.h, n, _i(flat), _i(offs), _f(cond) if cond.shape[0] else None, cond.shape[0],
                                                   _i(mf), sd.ctypes.data_as(_lp), _i(ea), C.byref(params), _f(lat), _i(used), _f(pcm)),
                 "generate_batch")
        return dict(latents=lat, frames=used, audio=pcm)

    # ---- additive streaming entry (single utterance)
    def stream_begin(self, tokens, cond, max_frames, seed=123, params: GenParams | None = None):
        params = params or default_params()
        tokens = np.ascontiguousarray(tokens, np.int32)
        cond = np.ascontiguousarray(cond, dtype=np.float32).reshape(-1, 1024)
        self._ck(self.lib.ptts_b200_stream_begin(self.h, _i(tokens), len(tokens), _f(cond) if cond.shape[0] else None, cond.shape[0],
                                                 int(max_frames), int(seed), C.byref(params)), "stream_begin")

    def stream_next(self, want_frames=1):
        """-> (pcm [frames*1920], done)"""
        out = np.zeros(want_frames * FRAME_SAMPLES, np.float32)
        nf, done = C.c_int(0), C.c_int(0)
        self._ck(self.lib.ptts_b200_stream_next(self.h, int(want_frames), _f(out), C.byref(nf), C.byref(done)), "stream_next")
        return out[: nf.value * FRAME_SAMPLES], bool(done.value)

    def stream(self, tokens, cond, max_frames, seed=123, params: GenParams | None = None, chunk=4):
        """Generator of PCM chunks (first chunk = 1 frame, for first-audio latency)."""
        self.stream_begin(tokens, cond, max_frames, seed, params)
        want = 1
        while True:
            pcm, done = self.stream_next(want)
            if len(pcm):
                yield pcm
            if done:
                return
            want = chunk

    # ---- measurement hooks
    def sync(self):
        self._ck(self.lib.ptts_b200_sync(self.h), "sync")

    def timer_start(self):
        self._ck(self.lib.ptts_b200_timer_start(self.h), "timer_start")

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._ck(self.lib.ptts_b200_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return ms.value

    def step_kernel_ms(self):
        """(ms, frame steps) of the most recent launch of the persistent FlowLM step kernel, events around that launch only."""
        ms, st = C.c_float(0), C.c_int(0)
        self._ck(self.lib.ptts_b200_step_kernel_ms(self.h, C.byref(ms), C.byref(st)), "step_kernel_ms")
        return ms.value, st.value

    def profile(self, on: bool):
        self._ck(self.lib.ptts_b200_profile_enable(self.h, 1 if on else 0), "profile_enable")

    def profile_read(self):
        ms = np.zeros(16, np.float32)
        cnt = np.zeros(16, np.int64)
        nc = C.c_int(0)
        self._ck(self.lib.ptts_b200_profile_read(self.h, 16, _f(ms), cnt.ctypes.data_as(_lp), C.byref(nc)), "profile_read")
        fl = np.zeros(16, np.float64)
        by = np.zeros(16, np.float64)
        dp = C.POINTER(C.c_double)
        self._ck(self.lib.ptts_b200_profile_work(self.h, 16, fl.ctypes.data_as(dp), by.ctypes.data_as(dp)), "profile_work")
        return {self.lib.ptts_b200_profile_class_name(i).decode():
                dict(ms=float(ms[i]), launches=int(cnt[i]), flops=float(fl[i]), bytes=float(by[i])) for i in range(nc.value)}

    def launch_count(self) -> int:
        return int(self.lib.ptts_b200_launch_count(self.h))

    def test_gemm(self, A, B, batches, rows, N, K, taps=1, shift0=0, step=0, bf16=True, split=0, transposed=False, bn=128,
                  split_k=1):
        A = np.ascontiguousarray(A, np.float32)
        B = np.ascontiguousarray(B, np.float32)
        out = np.zeros(batches * rows * N, np.float32)
        self._ck(self.lib.ptts_b200_test_gemm(self.h, _f(A), _f(B), _f(out), batches, rows, N, K, taps, shift0, step,
                                              1 if bf16 else 0, split, 1 if transposed else 0, bn, split_k), "test_gemm")
        return out.reshape(N, batches * rows) if transposed else out.reshape(batches * rows, N)


class Pool:
    """ptts_b200_pool: one engine per device, utterances bucketed by length and sharded by the C library (one host thread per
    device), results gathered in the original order. devices=None: every visible device."""

    def __init__(self, weights: "dict[str, np.ndarray]", devices=None):
        self.h = None
        self.lib = load_library()
        keep = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
        names = [k.encode() for k in keep]
        tab = (_Tensor * len(names))()
        for i, (k, v) in enumerate(keep.items()):
            tab[i] = _Tensor(names[i], v.ctypes.data, 0, v.size)
        dv = None if devices is None else np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        if self.lib.ptts_b200_pool_create(C.byref(h), _i(dv), 0 if dv is None else len(dv), tab, len(names)) != 0:
            raise EngineError("ptts_b200_pool_create: " + (self.lib.ptts_b200_last_error() or b"").decode())
        self.h = h

    @property
    def size(self) -> int:
        return int(self.lib.ptts_b200_pool_size(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.ptts_b200_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def generate(self, token_lists, cond, max_frames, seeds=123, eos_after=None, params: "GenParams | None" = None, max_batch=0,
                 want_audio=True, audio_out=None, want_latents=True):
        params = params or default_params()
        cond = np.ascontiguousarray(cond, dtype=np.float32).reshape(-1, 1024)
        n, flat, offs, mf, sd, ea = Engine._pack(token_lists, max_frames, seeds, eos_after)
        F = int(mf.max())
        lat = np.zeros((n, F, 32), np.float32) if want_latents else None
        used = np.zeros(n, np.int32)
        pcm = audio_out if audio_out is not None else (np.zeros((n, F * FRAME_SAMPLES), np.float32) if want_audio else None)
        rc = self.lib.ptts_b200_pool_generate(self.h, n, _i(flat), _i(offs), _f(cond) if cond.shape[0] else None, cond.shape[0], _i(mf),
                                              sd.ctypes.data_as(_lp), _i(ea), C.byref(params), int(max_batch), _f(lat), _i(used), _f(pcm))
        if rc != 0:
            raise EngineError("pool_generate: " + (self.lib.ptts_b200_last_error() or b"").decode())
        return dict(latents=lat, frames=used, audio=pcm)


def default_params(lsd_steps=1, temp=0.0, noise_clamp=0.0, eos_enabled=0, eos_threshold=-4.0, eos_min_frames=1,
                   eos_after=0, precise=1) -> GenParams:
    return GenParams(lsd_steps, temp, noise_clamp, eos_enabled, eos_threshold, eos_min_frames, eos_after, precise)


# ---------------------------------------------------------------- host glue kept from the reference
def prepare_text(text: str):
    """ptts_prepare_text (ptts.c:219-283): returns (prepared, word_count, eos_after)."""
    out = []
    in_space = True
    words = 0
    for ch in text:
        if ch in "\n\r\t":
            ch = " "
        if ch == " ":
            if not in_space:
                out.append(" ")
                in_space = True
            continue
        if in_space:
            words += 1
        in_space = False
        out.append(ch)
    s = "".join(out).rstrip(" ")
    if not s:
        raise ValueError("Text prompt cannot be empty")
    for i, ch in enumerate(s):
        if ch.isascii() and ch.isalpha():
            s = s[:i] + ch.upper() + s[i + 1:]
            break
    if s[-1].isascii() and s[-1].isalnum():
        s += "."
    eos_after = 5 if words <= 4 else 3
    if words < 5:
        s = " " * 8 + s
    return s, words, eos_after


def estimate_frames(words: int) -> int:
    """ptts_estimate_frames (ptts.c:285-291)."""
    words = max(words, 1)
    return max(int((np.float32(words) * np.float32(1.0) + np.float32(2.0)) * np.float32(12.5)), 1)
