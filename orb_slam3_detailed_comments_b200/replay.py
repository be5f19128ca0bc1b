"""Multi-GPU sequence replay plumbing: one process per GPU, frames / sequences sharded across
ranks with NO data-path collective; the only exchange is the gather of new-keyframe / map-point state when
several replicas share one map (Atlas multi-session), done with torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests).  Payloads are ~60 KB per keyframe: latency-bound, off the per-frame path."""
import numpy as np
import torch
import torch.distributed as dist

KF_FIELDS = (("pose", np.float32, 7), ("kp_xy", np.float32, 2), ("kp_octave", np.int32, 1), ("uright", np.float32, 1),
             ("desc", np.uint8, 32))


def shard_frames(n_frames, rank, world):
    """Independent-frame replay: frame i goes to rank i % world (extraction + stereo are stateless)."""
    return np.arange(rank, n_frames, world)


def shard_sequences(n_sequences, rank, world):
    """Batched sequence replay (config 5): whole sequences stay on one GPU (tracking is temporal)."""
    return [s for s in range(n_sequences) if s % world == rank]


def pack_keyframe(pose7, kps, uright, desc):
    """SoA block of one new keyframe -> one contiguous uint8 tensor: [n:int32][pose 7 f32][n x (x,y f32)][n x octave i32]
    [n x uright f32][n x 32 B descriptors]."""
    n = len(kps)
    parts = [np.array([n], np.int32).view(np.uint8), np.asarray(pose7, np.float32).view(np.uint8),
             np.stack([kps["x"], kps["y"]], 1).astype(np.float32).reshape(-1).view(np.uint8),
             kps["octave"].astype(np.int32).view(np.uint8), np.asarray(uright, np.float32).view(np.uint8),
             np.ascontiguousarray(desc, np.uint8).reshape(-1)]
    return torch.from_numpy(np.concatenate(parts))


def unpack_keyframe(buf):
    b = buf.cpu().numpy() if isinstance(buf, torch.Tensor) else buf
    n = int(b[:4].view(np.int32)[0])
    o = 4
    pose = b[o:o + 28].view(np.float32).copy(); o += 28
    xy = b[o:o + 8 * n].view(np.float32).reshape(n, 2).copy(); o += 8 * n
    octave = b[o:o + 4 * n].view(np.int32).copy(); o += 4 * n
    ur = b[o:o + 4 * n].view(np.float32).copy(); o += 4 * n
    desc = b[o:o + 32 * n].reshape(n, 32).copy()
    return dict(pose=pose, xy=xy, octave=octave, uright=ur, desc=desc)


def gather_keyframes(block, device=None, group=None):
    """All-gather one (possibly empty) keyframe block per rank.  Ragged sizes are handled with a size exchange and
    padding to the maximum.  Returns the list of per-rank blocks (uint8 tensors, empty where a rank had none)."""
    world = dist.get_world_size(group)
    dev = device or block.device
    block = block.to(dev)
    size = torch.tensor([block.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    mx = int(max(int(s.item()) for s in sizes))
    if mx == 0:
        return [torch.zeros(0, dtype=torch.uint8) for _ in range(world)]
    pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
    pad[:block.numel()] = block
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return [o[:int(s.item())].cpu() for o, s in zip(out, sizes)]


def max_over_ranks(ms, device):
    """Timing rule of bench.py: device-measured milliseconds, maximum over ranks."""
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
