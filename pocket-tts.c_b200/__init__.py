"""pocket-tts B200 engine: Python mirror of the C ABI (include/ptts_b200.h)."""
from . import synth  # noqa: F401
