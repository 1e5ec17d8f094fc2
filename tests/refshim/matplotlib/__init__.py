"""import-only stub (vietTTS/nat/text2mel.py:8 imports matplotlib.pyplot at module level)."""
