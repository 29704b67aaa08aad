"""Frame-level data parallelism: one process per GPU, frames are independent units
(SURVEY.md §8e), frame i -> rank i mod W (the shape of the reference's dead
DistEvalHook, mmdet/core/evaluation/eval_hooks.py:72), one NCCL all_gather of the
fixed-size result tensors at the end of a shard.  No collective inside a frame."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_frames(n_frames, rank, world):
    """Indices of the frames this rank owns."""
    return list(range(rank, n_frames, world))


def gather_detections(det, ndet):
    """det [F_local, cap, 9] f32, ndet [F_local] i32 (same F_local on every rank) ->
    on every rank: det [W, F_local, cap, 9], ndet [W, F_local].  Frame (r, j) is global frame j*W + r."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return det.unsqueeze(0), ndet.unsqueeze(0)
    W = dist.get_world_size()
    dets = [torch.empty_like(det) for _ in range(W)]
    nds = [torch.empty_like(ndet) for _ in range(W)]
    dist.all_gather(dets, det.contiguous())
    dist.all_gather(nds, ndet.contiguous())
    return torch.stack(dets, 0), torch.stack(nds, 0)


def interleave(det_all, ndet_all):
    """[W, F, cap, 9] -> global frame order [W*F, cap, 9] (frame j*W + r)."""
    W, Fl = det_all.shape[0], det_all.shape[1]
    return (det_all.permute(1, 0, 2, 3).reshape(W * Fl, *det_all.shape[2:]),
            ndet_all.permute(1, 0).reshape(W * Fl))


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
