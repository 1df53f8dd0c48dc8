"""Batch-shard plumbing for multi-GPU inference: one process per GPU, independent stereo pairs per rank, and the
single collective of the path -- an all_gather of the per-image EPE partial sums -- mirroring the reference's eval
loop (stereo/modeling/trainer_template.py:313-329: all_gather of per-image metric values, then a mean over images;
stereo/datasets/__init__.py:64-65: DistributedSampler sharding).

Backend-agnostic on purpose: NCCL over NVLink on the GPUs, gloo in the CPU tests.  The payload is <= 8 bytes per
image, so the collective is latency-only; no custom transport is justified.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """DistributedSampler(shuffle=False) semantics: pad to a multiple of world by wrapping, then stride by world."""
    if n_items <= 0:
        return []
    per_rank = (n_items + world - 1) // world
    padded = list(range(n_items)) + [i % n_items for i in range(per_rank * world - n_items)]
    return padded[rank:per_rank * world:world]


def gather_epe_partials(partials, indexes=None):
    """partials: (B_local, 2) = {sum |err|, #valid} per image.  Returns (all_partials, all_indexes) on every rank,
    concatenated in rank order.  One all_gather (two when image indexes are supplied)."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return partials, indexes
    parts = [torch.empty_like(partials) for _ in range(world)]
    dist.all_gather(parts, partials.contiguous())
    idx = None
    if indexes is not None:
        idxs = [torch.empty_like(indexes) for _ in range(world)]
        dist.all_gather(idxs, indexes.contiguous())
        idx = torch.cat(idxs, 0)
    return torch.cat(parts, 0), idx


def mean_epe(all_partials, all_indexes=None):
    """Per-image EPE = sum/valid (0 when no valid pixel, metric_per_image.py:38-39), de-duplicated by image index like
    the unique_dict pass of trainer_template.py:323-329, then averaged over images."""
    per_image = torch.where(all_partials[:, 1] > 0, all_partials[:, 0] / all_partials[:, 1].clamp(min=1),
                            torch.zeros_like(all_partials[:, 0]))
    if all_indexes is not None:
        seen = {}
        for i, v in zip(all_indexes.tolist(), per_image.tolist()):
            seen.setdefault(int(i), v)
        vals = list(seen.values())
        return sum(vals) / max(len(vals), 1)
    return per_image.mean().item()
