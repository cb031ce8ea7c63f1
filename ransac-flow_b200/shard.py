"""Pair-level data parallelism: one process per GPU, pair ``i`` goes to
rank ``i % world_size`` (replacing the reference's manual ``--beginIndex/--endIndex`` shards,
evaluation/evalCorr/evaluation.py:99-100), and ONE all-gather of fixed-size per-pair records
collects the results.  No collective inside the data path."""
import os

import numpy as np
import torch
import torch.distributed as dist

RECORD_FLOATS = 16       # pair_id, nH, status, nbInlier, H[9], 3 spare


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def my_pairs(n_pairs, rank, world):
    """Round-robin shard: the pair ids this rank processes, in order."""
    return list(range(rank, n_pairs, world))


def pack_record(pair_id, H, nbInlier=0, status=0, nH=1):
    r = np.zeros(RECORD_FLOATS, dtype=np.float32)
    r[0], r[1], r[2], r[3] = pair_id, nH, status, nbInlier
    if H is not None:
        r[4:13] = np.asarray(H, dtype=np.float32).reshape(-1)[:9]
    return r


def gather_records(records, n_pairs, world, device=None):
    """All-gather every rank's [n_local, RECORD_FLOATS] records; returns [n_pairs, RECORD_FLOATS]
    ordered by pair id on every rank.  Ranks with fewer pairs pad with pair_id = -1."""
    per = (n_pairs + world - 1) // world
    buf = np.full((per, RECORD_FLOATS), -1, dtype=np.float32)
    if len(records):
        buf[:len(records)] = np.stack(records)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    if world > 1:
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        allr = torch.cat(out, dim=0).cpu().numpy()
    else:
        allr = t.cpu().numpy()
    allr = allr[allr[:, 0] >= 0]
    return allr[np.argsort(allr[:, 0], kind="stable")]
