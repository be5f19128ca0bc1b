"""CPU tier: the N>1 plumbing (frame / sequence sharding, keyframe-state gather, max-over-ranks timing) with
world_size 2 on the gloo backend."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam3_detailed_comments_b200 import replay
from orb_slam3_detailed_comments_b200._native import KP_DTYPE


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(rank)
        frames = replay.shard_frames(11, rank, world)
        seqs = replay.shard_sequences(8, rank, world)
        n = 40 + 10 * rank
        kps = np.zeros(n, KP_DTYPE)
        kps["x"], kps["y"], kps["octave"] = rng.random(n) * 640, rng.random(n) * 480, rng.integers(0, 8, n)
        ur, desc = rng.random(n).astype(np.float32), rng.integers(0, 256, (n, 32), dtype=np.uint8)
        pose = np.array([0, 0, 0, 1, rank, 0, 0], np.float32)
        # rank 1 inserts a keyframe on the second round only (ragged / empty contributions)
        blocks0 = replay.gather_keyframes(replay.pack_keyframe(pose, kps, ur, desc))
        mine = replay.pack_keyframe(pose, kps, ur, desc) if rank == 0 else torch.zeros(0, dtype=torch.uint8)
        blocks1 = replay.gather_keyframes(mine)
        t = replay.max_over_ranks(10.0 + rank, torch.device("cpu"))
        ok = (len(blocks0) == world and all(replay.unpack_keyframe(b)["pose"][4] == r for r, b in enumerate(blocks0))
              and replay.unpack_keyframe(blocks0[rank])["desc"].tobytes() == desc.tobytes()
              and [b.numel() > 0 for b in blocks1] == [True, False] and t == 10.0 + world - 1)
        other = replay.unpack_keyframe(blocks0[1 - rank])
        ok = ok and len(other["octave"]) == 40 + 10 * (1 - rank)
        ret[rank] = (ok, frames.tolist(), seqs)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_gather():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] and ret[1][0]
    assert sorted(ret[0][1] + ret[1][1]) == list(range(11))          # every frame processed exactly once
    assert sorted(ret[0][2] + ret[1][2]) == list(range(8)) and len(ret[0][2]) == 4
