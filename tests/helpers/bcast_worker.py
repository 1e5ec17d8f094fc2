"""Worker of tests/test_gpu_broadcast.py: launched with torch.distributed.run, one rank per GPU.
Rank 0 loads the synthetic checkpoints from host memory, every other rank receives them through
vtts_broadcast_weights; then all ranks synthesise the same utterance and compare waveforms bit for bit."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from viettts_b200 import parallel, synthetic  # noqa: E402
from viettts_b200.engine import Engine  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    hp = synthetic.hifigan_params(1234) if rank == 0 else None
    ck = synthetic.acoustic_ckpt(1234) if rank == 0 else None
    dk = synthetic.duration_ckpt(1234) if rank == 0 else None
    comm = parallel.load_weights_via_abi(eng, hp, ck, dk)
    tokens, dur = synthetic.utterance(3, 30, 1.0)
    d = (np.asarray(dur, np.float32) * np.float32(16000)) / np.float32(256)
    wav = eng.synthesize(np.asarray(tokens, np.int32)[None], d, seed=7)
    dsec = eng.predict_duration(np.asarray(tokens, np.int32)[None])
    t = torch.from_numpy(np.concatenate([wav.ravel(), dsec.ravel()])).cuda()
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    ok = all(torch.equal(outs[0], o) for o in outs) and bool(torch.isfinite(t).all()) and float(t.abs().max()) > 1e-3
    comm.close()
    eng.close()
    if rank == 0:
        print("BCAST_OK" if ok else "BCAST_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
