"""Why the tensor-core mode is bf16x3 and not plain BF16 / TF32: CPU emulation of the operand rounding
of each mode on the whole generator (scripts/precision_study.py), against the stated waveform tolerance."""
import importlib.util
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
WAV_LINF, WAV_RMS = 1e-4, 1e-5        # the tolerance of tests/test_gpu_hifigan.py / test_gpu_tc_conv.py


def test_only_split_arithmetic_meets_the_tolerance():
    spec = importlib.util.spec_from_file_location("precision_study", REPO / "scripts" / "precision_study.py")
    ps = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ps)
    r = ps.study(T=6, modes=("fp32", "bf16x1", "tf32x1", "bf16x3"))
    assert r["fp32"][0] < 5e-6
    assert r["bf16x3"][0] < WAV_LINF and r["bf16x3"][1] < WAV_RMS            # the chosen mode: inside
    assert r["bf16x1"][0] > 10 * WAV_LINF and r["tf32x1"][0] > 3 * WAV_LINF  # single-product modes: outside
    assert r["bf16x3"][0] < r["tf32x1"][0] / 20
