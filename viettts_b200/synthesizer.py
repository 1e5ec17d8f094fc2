"""Command line front end, mirroring vietTTS/synthesizer.py (`python -m viettts_b200.synthesizer`).

Same flags as the reference (--text --output --sample-rate --silence-duration --lexicon-file) and the same
pipeline (synthesizer.py:34-39): normalise -> text2mel -> mel2wave -> 16-bit PCM WAV.  Two additions that the
B200 path makes worthwhile: `--text-file` synthesises one utterance per input line as ragged batches through a
single library call per batch (`Engine.tts`), and the WAV writer is built in (the reference needs `soundfile`).
"""
from __future__ import annotations

import re
import struct
import unicodedata
from argparse import ArgumentParser
from pathlib import Path

import numpy as np

from . import config

_SIL = config.SPECIAL_PHONEMES[config.SIL_INDEX]

# synthesizer.py:21-32 as an ordered rewrite table (applied after NFKC + lower + strip)
_REWRITES = (
    (re.compile(r"[\n.,:]+"), f" {_SIL} "),        # sentence punctuation and newlines become a silence token
    (re.compile(r'"'), " "),
    (re.compile(r"\s+"), " "),
    (re.compile(r"[.,:;?!]+"), f" {_SIL} "),
    (re.compile(r"[ ]+"), " "),
    (re.compile(f"( {_SIL}+)+ "), f" {_SIL} "),    # runs of silence tokens collapse to one
)


def nat_normalize_text(text: str) -> str:
    text = unicodedata.normalize("NFKC", text).lower().strip()
    for pattern, repl in _REWRITES:
        text = pattern.sub(repl, text)
    return text.strip()


def float_to_pcm16(wave) -> np.ndarray:
    """libsndfile's float -> PCM_16 conversion (what `soundfile.write(path, float_array, sr)` stores in a .wav):
    round-to-nearest-even of x * 0x7FFF; samples are clipped to the int16 range instead of wrapping."""
    x = np.asarray(wave, np.float32).astype(np.float64) * 32767.0
    return np.clip(np.rint(x), -32768, 32767).astype("<i2")


def write_wav(path, wave, sample_rate: int = config.SAMPLE_RATE) -> None:
    """Mono 16-bit PCM RIFF/WAVE file with the canonical 44-byte header (synthesizer.py:39)."""
    pcm = float_to_pcm16(np.ravel(wave))
    data = pcm.tobytes()
    header = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE"
    header += b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, int(sample_rate), int(sample_rate) * 2, 2, 16)
    header += b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(header + data)


def read_wav(path):
    """Inverse of write_wav for tests: (float32 wave in [-1,1), sample_rate)."""
    raw = Path(path).read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE" and raw[12:16] == b"fmt "
    fmt, ch, sr, _, _, bits = struct.unpack("<HHIIHH", raw[20:36])
    assert (fmt, ch, bits) == (1, 1, 16) and raw[36:40] == b"data"
    n = struct.unpack("<I", raw[40:44])[0]
    return np.frombuffer(raw[44:44 + n], "<i2").astype(np.float32) / 32767.0, sr


def synthesize_lines(lines, lexicon_file, silence_duration=-1.0, seed=None, max_rows=32, engine=None):
    """Batched text -> list of waveforms (input order).  Lines are sorted by token count and cut into batches of
    at most `max_rows` rows so the padding inside a batch stays small; every batch is one `Engine.tts` call."""
    from .nat import text2mel as t2m
    from .hifigan.mel2wave import load_generator
    engine = t2m.load_duration(engine)
    _, ck_seed = t2m.load_acoustic(engine)
    if seed is None:
        seed = ck_seed          # default stream key: the checkpoint's rng words, same as the single --text path
    load_generator(engine)
    toks = [t2m.text2tokens(nat_normalize_text(line), lexicon_file) for line in lines]
    order = sorted(range(len(toks)), key=lambda i: len(toks[i]))
    out = [None] * len(toks)
    for s in range(0, len(order), max_rows):
        idx = order[s:s + max_rows]
        L = max(len(toks[i]) for i in idx)
        tok = np.zeros((len(idx), L), np.int32)
        lens = np.zeros(len(idx), np.int32)
        for r, i in enumerate(idx):
            tok[r, : len(toks[i])] = toks[i]
            lens[r] = len(toks[i])
        waves, _ = engine.tts(tok, lens, silence_duration=silence_duration, seed=seed)
        for r, i in enumerate(idx):
            out[i] = waves[r]
    return out


def main(argv=None) -> int:
    parser = ArgumentParser(description="B200-native vietTTS synthesizer")
    parser.add_argument("--text", type=str)
    parser.add_argument("--text-file", type=Path, default=None, help="one utterance per line; outputs <output stem>_NNNN.wav")
    parser.add_argument("--output", default="clip.wav", type=Path)
    parser.add_argument("--sample-rate", default=16000, type=int)
    parser.add_argument("--silence-duration", default=-1, type=float)
    parser.add_argument("--lexicon-file", default=None)
    parser.add_argument("--seed", default=None, type=int,
                        help="prenet dropout: key of the on-device counter stream; default = the checkpoint's rng "
                             "(--text: the reference's own JAX/Haiku mask stream; --text-file: the device stream keyed by those words)")
    args = parser.parse_args(argv)
    lexicon = args.lexicon_file if args.lexicon_file is not None else config.LEXICON_FILE

    if args.text_file is not None:
        lines = [ln for ln in args.text_file.read_text().splitlines() if ln.strip()]
        waves = synthesize_lines(lines, lexicon, args.silence_duration, seed=args.seed)
        for i, w in enumerate(waves):
            fn = args.output.with_name(f"{args.output.stem}_{i:04d}{args.output.suffix or '.wav'}")
            print("writing output to file", fn)
            write_wav(fn, w, args.sample_rate)
        return 0

    if args.text is None:
        parser.error("--text or --text-file is required")
    from .hifigan.mel2wave import mel2wave
    from .nat.text2mel import text2mel
    text = nat_normalize_text(args.text)
    print("Normalized text input:", text)
    mel = text2mel(text, lexicon, args.silence_duration, seed=args.seed)
    wave = mel2wave(mel)
    print("writing output to file", args.output)
    write_wav(args.output, wave, args.sample_rate)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
