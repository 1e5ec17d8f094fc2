"""CPU oracle for the vietTTS hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package, and there only as the checker
(or the timed CPU baseline), never as the thing shipped.  Nothing under
`viettts_b200/` imports it; the product path fails loudly when the CUDA library
is missing.

What each module restates (file:line into /root/reference):

  hifigan_oracle.py  vietTTS/hifigan/model.py:8-125 (Generator, ResBlock1, get_padding)
                     PINNED: checked against the reference's own importable torch
                     implementation vietTTS/hifigan/torch_model.py:156-209 through the
                     reference's own converter convert_torch_model_to_haiku.py:27-62
                     (tests/golden/make_golden.py, fixtures in tests/golden/).
  nat_oracle.py      vietTTS/nat/model.py:9-47,76-144 and nat/text2mel.py:61-82; also the callers of the
                     path: DurationModel (model.py:49-70), the text2mel duration fix-ups
                     (text2mel.py:85-103), the teacher-forced pass with zoneout
                     (model.py:146-169) and the GTA forward (gta.py:28-41).
                     PARITY UNPINNED: jax / dm-haiku are not installable here, the
                     reference's tests hold no golden vectors for this path
                     (tests/test_nat_acoustic.py is a stale shape test), so this is a
                     restatement of the published dm-haiku semantics (hk.LSTM,
                     hk.deep_rnn_with_skip_connections, hk.BatchNorm, hk.Conv1D,
                     hk.dropout), arbitrated by its own float64 mode.  Each building block is
                     cross-checked against torch's independent operator
                     (tests/test_oracle_crosschecks.py); the two reference shape tests
                     (tests/test_nat_duration.py, tests/test_nat_acoustic.py) are reproduced.
  mel_oracle.py      vietTTS/nat/dsp.py:11-25,65-128 (rolling_window, batched_stft,
                     MelFilter).  PARITY UNPINNED for the same reason (jax + librosa
                     absent); cross-checked against torch.stft and torchaudio's
                     Slaney filterbank, which are independent implementations.
"""
