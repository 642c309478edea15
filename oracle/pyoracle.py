"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``    - oracle/liboracle.so, our C restatement (oracle/ptts_oracle.c).
* ``Reference`` - oracle/_ref/libptts_ref.so, the UNMODIFIED reference compiled by
  oracle/Makefile from /root/reference (the binary travels to the GPU box; the
  sources do not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "liboracle.so"
REF_SO = HERE / "_ref" / "libptts_ref.so"
REF_CLI = HERE / "_ref" / "ptts_ref"

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", str(HERE)] + targets, check=True)


def _f(a):
    return a.ctypes.data_as(_fp)


class _Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", _fp), ("numel", C.c_int64)]


class Oracle:
    def __init__(self, weights: "dict[str, np.ndarray]"):
        if not ORACLE_SO.exists():
            build(ref=False)
        self.lib = L = C.CDLL(str(ORACLE_SO))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(_Tensor), C.c_int]
        L.orc_last_error.restype = C.c_char_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_noise.argtypes = [C.c_int64, C.c_float, C.c_float, C.c_int, _fp]
        L.orc_flowlm_generate.restype = C.c_int
        L.orc_flowlm_generate.argtypes = [
            C.c_void_p, _ip, C.c_int, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int64,
            C.c_int, C.c_float, C.c_int, C.c_int, _fp, _ip, _fp, _fp, _fp, _fp, _fp, _fp]
        L.orc_flow_net.argtypes = [C.c_void_p, _fp, C.c_float, C.c_float, _fp, _fp]
        L.orc_scale_latents.argtypes = [C.c_void_p, _fp, C.c_int, _fp]
        L.orc_mimi_decode.restype = C.c_int
        L.orc_mimi_decode.argtypes = [C.c_void_p, _fp, C.c_int, _fp, _ip]
        L.orc_voice_encode.restype = C.c_int
        L.orc_voice_encode.argtypes = [C.c_void_p, _fp, C.c_int, _fp, C.c_int, _ip]
        L.orc_linear.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
        L.orc_conv1d.argtypes = [_fp, _fp, _fp, _fp] + [C.c_int] * 6
        L.orc_convtr1d.argtypes = [_fp, _fp, _fp, _fp] + [C.c_int] * 6
        L.orc_silu.restype = C.c_float
        L.orc_silu.argtypes = [C.c_float]
        self._keep = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
        self._names = [k.encode() for k in self._keep]
        tab = (_Tensor * len(self._keep))()
        for i, (k, v) in enumerate(self._keep.items()):
            tab[i] = _Tensor(self._names[i], _f(v), v.size)
        self._tab = tab
        self.h = L.orc_create(tab, len(self._keep))
        if not self.h:
            raise RuntimeError("orc_create: " + L.orc_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def noise(self, seed, temp, clamp, frames):
        out = np.zeros((frames, 32), np.float32)
        self.lib.orc_noise(seed, temp, clamp, frames, _f(out))
        return out

    def flowlm_generate(self, tokens, cond, max_frames, lsd_steps=1, temp=0.0, noise_clamp=0.0, seed=123,
                        eos_enabled=0, eos_threshold=-4.0, eos_min_frames=1, eos_after=0, noise=None):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        cond = np.ascontiguousarray(cond, dtype=np.float32).reshape(-1, 1024)
        lat = np.zeros((max_frames, 32), np.float32)
        used = C.c_int(0)
        eos0 = C.c_float(0)
        cond0 = np.zeros(1024, np.float32)
        flow0 = np.zeros(32, np.float32)
        allc = np.zeros((max_frames, 1024), np.float32)
        alle = np.zeros(max_frames, np.float32)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float32)
        rc = self.lib.orc_flowlm_generate(
            self.h, tokens.ctypes.data_as(_ip), len(tokens), _f(cond), cond.shape[0], max_frames, lsd_steps,
            temp, noise_clamp, seed, eos_enabled, eos_threshold, eos_min_frames, eos_after,
            _f(lat), C.byref(used), C.byref(eos0), _f(cond0), _f(flow0), _f(allc), _f(alle),
            None if nz is None else _f(nz))
        if rc != 0:
            raise RuntimeError("orc_flowlm_generate failed")
        n = used.value
        return dict(latents=lat[:n], frames=n, eos0=eos0.value, cond0=cond0, flow0=flow0,
                    conds=allc[:n], eos=alle[:n])

    def flow_net(self, cond, s, t, x):
        cond = np.ascontiguousarray(cond, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(32, np.float32)
        self.lib.orc_flow_net(self.h, _f(cond), s, t, _f(x), _f(out))
        return out

    def scale_latents(self, lat):
        lat = np.ascontiguousarray(lat, np.float32)
        out = np.empty_like(lat)
        self.lib.orc_scale_latents(self.h, _f(lat), lat.shape[0], _f(out))
        return out

    def mimi_decode(self, latents):
        latents = np.ascontiguousarray(latents, np.float32).reshape(-1, 32)
        F = latents.shape[0]
        out = np.zeros(F * 1920, np.float32)
        n = C.c_int(0)
        rc = self.lib.orc_mimi_decode(self.h, _f(latents), F, _f(out), C.byref(n))
        if rc != 0 or n.value != F * 1920:
            raise RuntimeError("orc_mimi_decode failed")
        return out


    def voice_encode(self, pcm):
        """24 kHz mono PCM -> conditioning rows [frames, 1024] (orc_voice_encode)."""
        pcm = np.ascontiguousarray(pcm, np.float32).reshape(-1)
        F = (pcm.size + 1919) // 1920
        out = np.zeros((F, 1024), np.float32)
        n = C.c_int(0)
        if self.lib.orc_voice_encode(self.h, _f(pcm), pcm.size, _f(out), F, C.byref(n)) != 0 or n.value != F:
            raise RuntimeError("orc_voice_encode: " + self.lib.orc_last_error().decode())
        return out


class _Params(C.Structure):
    # ptts.h:27-38
    _fields_ = [("sample_rate", C.c_int), ("num_steps", C.c_int), ("num_frames", C.c_int),
                ("seed", C.c_int64), ("temp", C.c_float), ("noise_clamp", C.c_float),
                ("eos_enabled", C.c_int), ("eos_threshold", C.c_float), ("eos_min_frames", C.c_int),
                ("eos_after", C.c_int)]


class _Audio(C.Structure):
    # ptts_audio.h:10-15
    _fields_ = [("sample_rate", C.c_int), ("channels", C.c_int), ("num_samples", C.c_int), ("samples", _fp)]


class Reference:
    """The compiled, unmodified reference driven through its own public symbols."""

    def __init__(self, model_dir, threads: int | None = None):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        if threads:
            os.environ["OMP_NUM_THREADS"] = str(threads)
        self.lib = L = C.CDLL(str(REF_SO))
        L.ptts_load_dir.restype = C.c_void_p
        L.ptts_load_dir.argtypes = [C.c_char_p]
        L.ptts_free.argtypes = [C.c_void_p]
        L.ptts_get_error.restype = C.c_char_p
        L.ptts_tokenize.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(_ip), _ip]
        L.ptts_prepare_text.restype = C.c_void_p
        L.ptts_prepare_text.argtypes = [C.c_char_p, _ip, _ip]
        L.ptts_estimate_frames.argtypes = [C.c_int]
        L.ptts_load_voice_conditioning.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(_fp), _ip]
        L.ptts_flowlm_load.restype = C.c_void_p
        L.ptts_flowlm_load.argtypes = [C.c_void_p]
        L.ptts_flowlm_free.argtypes = [C.c_void_p]
        L.ptts_flowlm_generate_latents.argtypes = [
            C.c_void_p, _ip, C.c_int, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int64,
            C.c_int, C.c_float, C.c_int, C.c_int, _fp, _ip, _fp, _fp, _fp]
        L.ptts_flowlm_scale_latents.argtypes = [C.c_void_p, _fp, C.c_int, _fp]
        L.ptts_mimi_load.restype = C.c_void_p
        L.ptts_mimi_load.argtypes = [C.c_void_p]
        L.ptts_mimi_free.argtypes = [C.c_void_p]
        L.ptts_mimi_decode.argtypes = [C.c_void_p, _fp, C.c_int, _fp, _ip]
        L.ptts_generate.restype = C.POINTER(_Audio)
        L.ptts_generate.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(_Params)]
        L.ptts_audio_free.argtypes = [C.POINTER(_Audio)]
        L.ptts_audio_save_wav.argtypes = [C.POINTER(_Audio), C.c_char_p]
        L.ptts_linear_forward.argtypes = [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int]
        L.ptts_conv1d_forward.argtypes = [_fp, _fp, _fp, _fp] + [C.c_int] * 6
        L.ptts_convtr1d_forward.argtypes = [_fp, _fp, _fp, _fp] + [C.c_int] * 6
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]
        self.ctx = L.ptts_load_dir(str(model_dir).encode())
        if not self.ctx:
            raise RuntimeError("ptts_load_dir: " + (L.ptts_get_error() or b"").decode())
        self._fm = None
        self._mm = None

    def close(self):
        if self._fm:
            self.lib.ptts_flowlm_free(self._fm)
        if self._mm:
            self.lib.ptts_mimi_free(self._mm)
        if self.ctx:
            self.lib.ptts_free(self.ctx)
        self._fm = self._mm = self.ctx = None

    @property
    def fm(self):
        if not self._fm:
            self._fm = self.lib.ptts_flowlm_load(self.ctx)
            assert self._fm
        return self._fm

    @property
    def mm(self):
        if not self._mm:
            self._mm = self.lib.ptts_mimi_load(self.ctx)
            assert self._mm
        return self._mm

    def prepare_text(self, text):
        wc, ea = C.c_int(0), C.c_int(0)
        p = self.lib.ptts_prepare_text(text.encode(), C.byref(wc), C.byref(ea))
        if not p:
            raise RuntimeError("ptts_prepare_text failed")
        s = C.string_at(p).decode()
        self.libc.free(p)
        return s, wc.value, ea.value

    def tokenize(self, prepared):
        ids = _ip()
        n = C.c_int(0)
        if self.lib.ptts_tokenize(self.ctx, prepared.encode(), C.byref(ids), C.byref(n)) != 0:
            raise RuntimeError("ptts_tokenize failed")
        out = np.array([ids[i] for i in range(n.value)], dtype=np.int32)
        self.libc.free(ids)
        return out

    def estimate_frames(self, words):
        return self.lib.ptts_estimate_frames(words)

    def voice(self, name=None):
        p = _fp()
        n = C.c_int(0)
        if self.lib.ptts_load_voice_conditioning(self.ctx, name.encode() if name else None, C.byref(p), C.byref(n)) != 0:
            raise RuntimeError("voice load failed")
        out = np.ctypeslib.as_array(p, shape=(n.value, 1024)).copy()
        self.libc.free(p)
        return out

    def flowlm_generate(self, tokens, cond, max_frames, lsd_steps=1, temp=0.0, noise_clamp=0.0, seed=123,
                        eos_enabled=0, eos_threshold=-4.0, eos_min_frames=1, eos_after=0):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        cond = np.ascontiguousarray(cond, dtype=np.float32).reshape(-1, 1024)
        lat = np.zeros((max_frames, 32), np.float32)
        used = C.c_int(0)
        eos0 = C.c_float(0)
        cond0 = np.zeros(1024, np.float32)
        flow0 = np.zeros(32, np.float32)
        rc = self.lib.ptts_flowlm_generate_latents(
            self.fm, tokens.ctypes.data_as(_ip), len(tokens), _f(cond), cond.shape[0], max_frames, lsd_steps,
            temp, noise_clamp, seed, eos_enabled, eos_threshold, eos_min_frames, eos_after,
            _f(lat), C.byref(used), C.byref(eos0), _f(cond0), _f(flow0))
        if rc != 0:
            raise RuntimeError("ptts_flowlm_generate_latents failed")
        n = used.value
        return dict(latents=lat[:n], frames=n, eos0=eos0.value, cond0=cond0, flow0=flow0)

    def scale_latents(self, lat):
        lat = np.ascontiguousarray(lat, np.float32)
        out = np.empty_like(lat)
        self.lib.ptts_flowlm_scale_latents(self.fm, _f(lat), lat.shape[0], _f(out))
        return out

    def mimi_decode(self, latents):
        latents = np.ascontiguousarray(latents, np.float32).reshape(-1, 32)
        F = latents.shape[0]
        out = np.zeros(F * 1920, np.float32)
        n = C.c_int(0)
        if self.lib.ptts_mimi_decode(self.mm, _f(latents), F, _f(out), C.byref(n)) != 0:
            raise RuntimeError("ptts_mimi_decode failed")
        return out[: n.value]

    def generate(self, text, voice="alba", num_frames=0, num_steps=1, seed=123, temp=0.7,
                 eos_enabled=0, eos_threshold=1e9):
        """ptts_generate (ptts.c:1011): the reference's own end-to-end public call."""
        p = _Params(24000, num_steps, num_frames, seed, temp, 0.0, eos_enabled, eos_threshold, 1, 0)
        a = self.lib.ptts_generate(self.ctx, text.encode(), voice.encode() if voice else None, C.byref(p))
        if not a:
            raise RuntimeError("ptts_generate: " + (self.lib.ptts_get_error() or b"").decode())
        out = np.ctypeslib.as_array(a.contents.samples, shape=(a.contents.num_samples,)).copy()
        self.lib.ptts_audio_free(a)
        return out
