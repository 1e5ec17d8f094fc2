"""Stand-in for librosa: only `librosa.filters.mel` (called at vietTTS/nat/dsp.py:109-111)."""
from . import filters  # noqa: F401
