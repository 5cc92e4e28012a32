"""Utterance sharding across the GPUs of one box (BASELINE config 4).

The path shards by utterance: every utterance is independent given the weights (eval-mode BN, no
cross-utterance op in models.py:201-250), so rank r of N decodes utterances
[r*ceil(B/N), (r+1)*ceil(B/N)) with its own replica and the only collective is one all_gather of
the waveforms at the end (NCCL on GPUs; the same code runs over gloo in the CPU tests).  The
reference has no equivalent: infer.py always calls the vocoder with batch 1."""
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) per rank; the first n_items % world_size ranks get one extra item, so any
    batch size (including fewer utterances than ranks) is covered exactly once."""
    if n_items < 0 or world_size < 1:
        raise ValueError("shard_bounds: need n_items >= 0 and world_size >= 1")
    base, extra = divmod(n_items, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def decode_sharded(decode_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], latent: torch.Tensor,
                   mel_ref: torch.Tensor, group: Optional["dist.ProcessGroup"] = None,
                   gather: bool = True) -> torch.Tensor:
    """Every rank passes the SAME full batch (latent [B,T0,D], mel_ref [B or 1,Tm,M]); each decodes its
    shard with `decode_fn(latent_shard, mel_shard) -> wav [b,1,L]` (fp32) or `[b,L]` (int16 PCM, the fused
    epilogue of infer.py:206-212,234: half the bytes on the wire) and, when `gather`, all ranks return the full
    waveform tensor of that layout (one all_gather; ragged shards are padded to the largest shard for the
    collective and trimmed afterwards)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = latent.shape[0]
    bounds = shard_bounds(B, world)
    lo, hi = bounds[rank]
    mel_shard = mel_ref if mel_ref.shape[0] == 1 else mel_ref[lo:hi]
    if hi > lo:
        wav = decode_fn(latent[lo:hi], mel_shard)
        L = wav.shape[-1]
    else:
        wav, L = None, None
    if world == 1 or not gather:
        return wav
    # agree on L (ranks with an empty shard do not know it) and on the padded shard size
    dev = latent.device if wav is None else wav.device
    meta = torch.tensor([L if L is not None else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX, group=group)
    L = int(meta.item())
    cap = max(h - l for l, h in bounds)
    # ranks with an empty shard learn the layout (dtype, rank of the tensor) from the others
    lay = torch.tensor([0 if wav is None else (2 if wav.dtype == torch.int16 else 1)], dtype=torch.int64, device=dev)
    dist.all_reduce(lay, op=dist.ReduceOp.MAX, group=group)
    pcm = int(lay.item()) == 2
    dtype = torch.int16 if pcm else torch.float32
    shape = (L,) if pcm else (1, L)
    buf = torch.zeros(cap, *shape, dtype=dtype, device=dev)
    if wav is not None:
        buf[: hi - lo] = wav
    out = torch.empty(world * cap, *shape, dtype=dtype, device=dev)
    if pcm:      # neither NCCL nor gloo has an int16 datatype: the PCM bytes travel as uint8
        dist.all_gather_into_tensor(out.view(torch.uint8), buf.view(torch.uint8), group=group)
    else:
        dist.all_gather_into_tensor(out, buf, group=group)
    parts = [out[r * cap: r * cap + (h - l)] for r, (l, h) in enumerate(bounds)]
    return torch.cat(parts, dim=0)
