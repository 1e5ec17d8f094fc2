"""import-only stub (vietTTS/nat/data_loader.py:5)."""
