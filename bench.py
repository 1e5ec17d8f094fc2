#!/usr/bin/env python
"""Benchmark of the vietTTS hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the one the metric is quoted on): per GPU a batch of 32
synthetic 100-phoneme / 5 s utterances (N = 312 mel frames, 79 872 samples each) goes through
the NAT acoustic model (encoder, Gaussian upsampling, autoregressive decoder with prenet
dropout, postnet) and the HiFiGAN generator.  One "step" = one such batch.  Weak scaling: every
rank gets its own 32 utterances; the weights are broadcast once from rank 0 over NCCL.

  value  : audio samples / s, whole job, inputs resident in HBM (device-pointer C ABI), CUDA events
  e2e    : same metric through the host-buffer C ABI call (vtts_synthesize_host) with numpy inputs:
           H2D of tokens/durations and D2H of the waveform inside the timed region
  roofline: HiFiGAN generator (98 % of the FLOPs): algorithmic 614.1 MFLOP per mel frame / measured
           stage time (CUDA events recorded around the stage inside the timed region)
  cpu_baseline: the oracle port (torch CPU restatement of the reference) on the host cores,
           bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

from viettts_b200 import config as C  # noqa: E402
from viettts_b200 import synthetic  # noqa: E402

METRIC = "audio_samples_per_sec"
UNIT = "samples/s"


def peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_tflops=d.get("bf16_tflops", 1590.0),
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def make_batch(batch: int, phonemes: int, seconds: float, seed0: int):
    toks, durs, nfs = [], [], []
    for b in range(batch):
        tk, d = synthetic.utterance(seed0 + b, phonemes, seconds)
        d = (np.asarray(d, np.float32) * np.float32(C.SAMPLE_RATE)) / np.float32(C.HOP)
        toks.append(np.asarray(tk, np.int32))
        durs.append(d[0])
        nfs.append(int(np.sum(d, dtype=np.float32)))
    return np.stack(toks), np.stack(durs).astype(np.float32), np.asarray(nfs, np.int32)


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running beside the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle port) -- the only place bench.py executes oracle/
# ---------------------------------------------------------------------------------------------
_CPU_THREADS = None


def cpu_pick_threads(hp, ck):
    """The oracle port is a torch CPU program: the autoregressive part is a Python loop over small
    matmuls (best with few threads), the generator is conv-bound (best with many).  Pick, per stage,
    the fastest thread count on a tiny probe so that the CPU arm is not handicapped by
    oversubscription on a 100+-thread host."""
    global _CPU_THREADS
    if _CPU_THREADS is not None:
        return _CPU_THREADS
    import torch
    from oracle import hifigan_oracle, nat_oracle
    cores = os.cpu_count() or 1
    tokens, durs, nfs = make_batch(1, 20, 0.4, 7)
    mel = synthetic.mel_input(1, 1, 24)

    def t_nat():
        with torch.no_grad():
            nat_oracle.inference(ck, tokens, durs, int(nfs[0]), None)

    def t_hg():
        with torch.no_grad():
            hifigan_oracle.generator_forward(hp, mel)

    best = {}
    for name, fn, cands in (("nat", t_nat, [1, 2, 4, 8, 16]), ("hifigan", t_hg, [4, 8, 16, 32, 64, 128, cores])):
        res = []
        for n in sorted(set(c for c in cands if c <= cores)):
            torch.set_num_threads(n)
            fn()
            t0 = time.perf_counter()
            fn()
            res.append((time.perf_counter() - t0, n))
        best[name] = min(res)[1]
    _CPU_THREADS = best
    return best


def cpu_port_step(hp, ck, tokens, durs, nfs, masks):
    """One pass of the reference algorithm (torch CPU restatement) over the given utterances."""
    import torch
    from oracle import hifigan_oracle, nat_oracle
    th = cpu_pick_threads(hp, ck)
    n = int(nfs[0])
    with torch.no_grad():
        torch.set_num_threads(th["nat"])
        mel = nat_oracle.inference(ck, tokens, durs, n, masks)
        torch.set_num_threads(th["hifigan"])
        wav = hifigan_oracle.generator_forward(hp, mel.numpy())
    return int(wav.numel())


def cpu_baseline(hp, ck, phonemes, seconds, budget_s=12.0, rows=1):
    th = cpu_pick_threads(hp, ck)
    tokens, durs, nfs = make_batch(rows, phonemes, seconds, 9000)
    masks = synthetic.dropout_masks(3, rows, int(nfs[0]))
    cpu_port_step(hp, ck, tokens[:1], durs[:1], nfs[:1], masks[:1])  # warm-up
    t0 = time.perf_counter()
    samples, it = 0, 0
    while True:
        samples += cpu_port_step(hp, ck, tokens, durs, nfs, masks)
        it += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return dict(value=samples / dt, unit=UNIT, cores=max(th.values()), kind="port",
                sample=f"{it} passes of {rows} utterance(s) ({phonemes} phonemes, {int(nfs[0])} frames) through oracle/ (torch CPU; "
                       f"threads: acoustic {th['nat']}, generator {th['hifigan']} of {os.cpu_count()} host threads, picked by a probe), {dt:.1f} s")


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation (oracle port; the
    reference's JAX/Haiku path cannot be installed offline), best host thread counts."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hp = synthetic.hifigan_params(1234)
    ck = synthetic.acoustic_ckpt(1234)
    th = cpu_pick_threads(hp, ck)
    rows = args.ref_rows
    tokens, durs, nfs = make_batch(rows, args.phonemes, args.seconds, 0)
    masks = synthetic.dropout_masks(3, rows, int(nfs[0]))
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_port_step(hp, ck, tokens, durs, nfs, masks)
    t0 = time.perf_counter()
    samples = 0
    for _ in range(args.steps):
        samples += cpu_port_step(hp, ck, tokens, durs, nfs, masks)
    dt = time.perf_counter() - t0
    val = samples / dt
    desc = (f"{args.steps} steps x {rows} utterance(s) through oracle/ (torch CPU restatement; threads: acoustic {th['nat']}, "
            f"generator {th['hifigan']} of {os.cpu_count()} host threads)")
    out = dict(metric=METRIC, value=val, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
               rtf=(dt / (samples / C.SAMPLE_RATE)),
               config=dict(workload=f"NAT acoustic + HiFiGAN, {args.phonemes}-phoneme / {args.seconds:g} s utterances, batch {args.batch} per GPU",
                           sample=f"{rows} utterance(s) per step (bounded sample of the batch-{args.batch} workload)"),
               cpu_baseline=dict(value=val, unit=UNIT, cores=max(th.values()), kind="port", sample=desc),
               e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from viettts_b200 import parallel
    from viettts_b200.engine import Engine

    eng = Engine(local)
    eng.set_precision(args.precision)
    hp = synthetic.hifigan_params(1234) if rank == 0 else None
    ck = synthetic.acoustic_ckpt(1234) if rank == 0 else None
    t_w = time.perf_counter()
    wbytes = parallel.load_weights_distributed(eng, hp, ck, dev)
    t_w = time.perf_counter() - t_w

    B = args.batch
    tokens, durs, nfs = make_batch(B, args.phonemes, args.seconds, 1000 * rank)
    N = int(nfs.max())
    L = tokens.shape[1]
    seed = 0xC0FFEE + rank
    tok_t = torch.from_numpy(tokens).to(dev)
    dur_t = torch.from_numpy(durs).to(dev)
    nf_t = torch.from_numpy(nfs).to(dev)
    mel_t = torch.empty((B, N, C.MEL_DIM), dtype=torch.float32, device=dev)
    wav_t = torch.empty((B, N * C.HOP), dtype=torch.float32, device=dev)
    samples_step = int(nfs.sum()) * C.HOP

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(marks=None):
        if marks is not None:
            marks[0].record()
        eng.acoustic_forward(tok_t, dur_t, N, n_frames_t=nf_t, seed=seed, out=mel_t)
        if marks is not None:
            marks[1].record()
        eng.hifigan_forward(mel_t, nf_t, out=wav_t)
        if marks is not None:
            marks[2].record()

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    marks = [[ev(), ev(), ev()] for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    e0.record()
    for k in range(args.steps):
        step(marks[k])
    e1.record()
    torch.cuda.synchronize()
    barrier()
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    ac_ms = float(np.mean([m[0].elapsed_time(m[1]) for m in marks]))
    hg_ms = float(np.mean([m[1].elapsed_time(m[2]) for m in marks]))
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * samples_step / (ms_step / 1e3)

    # ---- e2e through the host-buffer C ABI ----
    wav_pinned = Engine.pinned_empty((B, N * C.HOP))          # page-locked result buffer, reused every step
    for _ in range(2):
        eng.synthesize(tokens, durs, n_frames=nfs, seed=seed, out=wav_pinned)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wav_h = eng.synthesize(tokens, durs, n_frames=nfs, seed=seed, out=wav_pinned)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_val = world * samples_step * args.steps / dt
    h2d = tokens.nbytes + durs.nbytes + nfs.nbytes
    d2h = wav_h.nbytes

    # ---- STFT/log-mel kernel (MelFilter, nat/dsp.py:104-128): HBM-bound streaming kernel, measured separately ----
    mel_info = None
    if rank == 0:
        S = 79872
        MB = 512                                     # 512 x 5 s = 164 MB of samples: larger than L2
        wav_m = torch.rand((MB, S), dtype=torch.float32, device=dev) - 0.5
        mel_m = torch.empty((MB, S // C.HOP, C.MEL_DIM), dtype=torch.float32, device=dev)
        for _ in range(3):
            eng.melspec_forward(wav_m, out=mel_m)
        m0, m1 = ev(), ev()
        m0.record()
        for _ in range(5):
            eng.melspec_forward(wav_m, out=mel_m)
        m1.record()
        torch.cuda.synchronize()
        mms = m0.elapsed_time(m1) / 5
        mbytes = MB * S * 5.25                      # 4 B in + 1.25 B out per sample
        pkm = peaks()
        mflop = MB * (S // C.HOP) * 29704.0         # sparse-filterbank FLOP count per frame
        mel_info = dict(samples_per_s=MB * S / (mms / 1e3), ms=mms, achieved_gbs=mbytes / (mms / 1e3) / 1e9,
                        peak_gbs=pkm["hbm_gbs"], frac_hbm=mbytes / (mms / 1e3) / 1e9 / pkm["hbm_gbs"],
                        achieved_tflops_fp32=mflop / (mms / 1e3) / 1e12, batch=MB, samples_per_row=S,
                        fp32_peak_tflops=74.4, frac_fp32=mflop / (mms / 1e3) / 1e12 / 74.4,
                        note="one warp per frame pair, FFT-1024 = 32 x 32 four-step with register-resident 32-point transforms; arithmetic intensity "
                             "22 FLOP/B sits above the FP32 ridge (11 FLOP/B): the kernel is FP32-issue bound, not HBM bound")
        del wav_m, mel_m

    # ---- callers of the path: duration model + one-call token->wav, chunked vocoding latency ----
    callers = None
    if rank == 0 and world == 1 and not args.no_callers:     # side measurements at N=1 only, like cpu_baseline
        eng.load_duration(synthetic.duration_ckpt(1234))
        for _ in range(2):
            waves, _ = eng.tts(tokens, silence_duration=0.05, seed=seed)
        t0 = time.perf_counter()
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            waves, _ = eng.tts(tokens, silence_duration=0.05, seed=seed)
        dt_tts = (time.perf_counter() - t0) / reps
        tts_samples = int(sum(w.size for w in waves))
        dur_ms = eng.last_stage_ms(3)
        mel1 = synthetic.mel_input(3, 1, N)
        for _ in range(2):
            next(eng.mel2wave_stream(mel1, chunk_frames=32))
        t0 = time.perf_counter()
        for _ in range(10):
            next(eng.mel2wave_stream(mel1, chunk_frames=32))
        first_ms = (time.perf_counter() - t0) / 10 * 1e3
        t0 = time.perf_counter()
        n_stream = sum(p.size for p in eng.mel2wave_stream(mel1, chunk_frames=32))
        all_ms = (time.perf_counter() - t0) * 1e3
        # GTA forward (gta.py:28-41): int16 audio -> MelFilter -> teacher-forced acoustic model with zoneout
        eng.load_mel_filterbank()
        S_g = N * C.HOP
        wav_i16 = (np.random.default_rng(1).standard_normal((B, S_g)) * 3000).astype(np.int16)
        dur_sec = durs * np.float32(C.HOP / C.SAMPLE_RATE)
        for _ in range(2):
            eng.gta(wav_i16, tokens, dur_sec, seed=seed)
        t0 = time.perf_counter()
        for _ in range(reps):
            eng.gta(wav_i16, tokens, dur_sec, seed=seed)
        dt_gta = (time.perf_counter() - t0) / reps
        gta_dev_ms = eng.last_stage_ms(1)
        callers = dict(
            gta=dict(api="vtts_gta_host (int16 audio -> log-mel -> shift -> teacher-forced acoustic model, zoneout + dropout on), host buffers",
                     batch=B, frames_per_s=B * N / dt_gta, ms_per_call=dt_gta * 1e3, teacher_forced_model_ms=gta_dev_ms,
                     autoregressive_model_ms=ac_ms),
            text_to_wav=dict(api="vtts_tts_host (duration model -> duration fix-ups -> acoustic -> trailing-silence trim -> generator), host buffers",
                             batch=B, samples_per_s=tts_samples / dt_tts, ms_per_call=dt_tts * 1e3, samples_per_call=tts_samples,
                             duration_model_ms=dur_ms),
            streaming_vocoder=dict(api="Engine.mel2wave_stream, B=1, 32-frame chunks + 16-frame recomputed halo, host buffers",
                                   first_chunk_ms=first_ms, audio_ms_per_chunk=32 * C.HOP / C.SAMPLE_RATE * 1e3,
                                   whole_utterance_ms=all_ms, samples=n_stream))

    if rank == 0:
        pk = peaks()
        frames = int(nfs.sum())
        flops = frames * C.HIFIGAN_FLOP_PER_FRAME
        ach = flops / (hg_ms / 1e3) / 1e12
        traffic = None
        tp = REPO / "profiles" / "r1_traffic.json"
        if tp.exists():
            tj = json.loads(tp.read_text())
            w = tj.get("workload", {})
            if w.get("batch") == B and w.get("mel_frames") == N and w.get("precision") == args.precision:
                traffic = tj["generator_dram_bytes_per_step"]
        ceiling = pk["bf16_tflops_sustained"] / 3.0 if args.precision != "fp32" else 74.4
        out = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_step,
            higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f32" if args.precision == "fp32" else "f32 (bf16x3 split products on tcgen05, fp32 accumulate/storage)", data="synthetic",
            rtf=(ms_step / 1e3) / (world * samples_step / C.SAMPLE_RATE),
            config=dict(workload=f"NAT acoustic + HiFiGAN, {args.phonemes}-phoneme / {args.seconds:g} s utterances, batch {B} per GPU (BASELINE configs[2])",
                        batch_per_gpu=B, phonemes=L, mel_frames=N, samples_per_utterance=N * C.HOP, parallelism=f"utterance-sharded x{world}",
                        l2="activations per step (>4 GB) exceed the 126 MB L2; no flush needed", dropout="on-device threefry keep-masks"),
            stages_ms=dict(acoustic=ac_ms, hifigan=hg_ms),
            e2e=dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), ms_per_step=1e3 * dt / args.steps,
                     api="vtts_synthesize_host via viettts_b200.Engine.synthesize (numpy in; numpy out in a page-locked buffer the D2H copy lands in)"),
            gpu_launches=int(launches),
            roofline=dict(bound="tensor",
                          kernel=("tc_conv_kernel: the 29 generator launches of one step (+ conv_post, 1 % of the stage time)" if args.precision != "fp32"
                                  else "conv1d_nwc_kernel: the 29 generator launches of one step (+ conv_post)"),
                          achieved=ach, peak=pk["bf16_tflops_sustained"], unit="TFLOP/s", frac=ach / pk["bf16_tflops_sustained"],
                          traffic=traffic, traffic_unit="bytes of DRAM traffic per step, all launches of the kernel (ncu, profiles/r1_traffic.json)",
                          algorithmic_flops_per_step=flops, launch_ms=hg_ms,
                          frac_of_mode_ceiling=ach / ceiling,
                          mode_ceiling=("1/3 of the bf16 peak: bf16x3 issues three bf16 MMAs per algorithmic product" if args.precision != "fp32"
                                        else "FP32 FMA pipe, nominal 74.4 TFLOP/s"),
                          peak_source=pk["source"] + ", sustained bf16 dense",
                          note=("algorithmic fp32 FLOPs; the bf16x3 path issues 3 bf16 MMAs per algorithmic product, so 1/3 of the bf16 peak is its ceiling"
                                if args.precision != "fp32" else "strict-fp32 path runs on the FP32 FMA pipe (nominal 74 TFLOP/s)")),
            clocks=clocks, weights=dict(bytes=wbytes, broadcast_s=t_w), melspec=mel_info, callers=callers,
        )
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(synthetic.hifigan_params(1234), synthetic.acoustic_ckpt(1234), args.phonemes, args.seconds)
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--phonemes", type=int, default=100)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--ref-rows", type=int, default=1, help="utterances per step of the CPU reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-callers", action="store_true", help="skip the duration/tts/gta/streaming side measurements (profiling runs)")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp32"],
                    help="conv arithmetic: bf16x3 = tcgen05 split-bf16 with fp32 accumulate (default), fp32 = FMA pipe")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
