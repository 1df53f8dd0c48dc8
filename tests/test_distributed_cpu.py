"""CPU: the N>1 plumbing (batch sharding + the single all_gather of per-image EPE) with gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openstereo_b200 import distributed as osd
    idx = osd.shard_indices(n_images, rank, world)
    # synthetic per-image partial sums: image i has error sum 10*i+1 over i+1 valid pixels; image 3 has no valid pixel
    part = torch.tensor([[0.0, 0.0] if i == 3 else [10.0 * i + 1.0, float(i + 1)] for i in idx])
    allp, alli = osd.gather_epe_partials(part, torch.tensor(idx, dtype=torch.int64))
    mean = osd.mean_epe(allp, alli)
    torch.save({"idx": idx, "all_idx": alli, "all_part": allp, "mean": mean}, os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [8, 7])
def test_gather_world2(tmp_path, n_images):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    want = [0.0 if i == 3 else (10.0 * i + 1.0) / (i + 1) for i in range(n_images)]
    want_mean = sum(want) / n_images
    for r in res:
        assert sorted(set(r["all_idx"].tolist())) == list(range(n_images))     # every image covered, padding de-duplicated
        assert r["mean"] == pytest.approx(want_mean, rel=1e-6)
        assert r["all_part"].shape == (2 * ((n_images + 1) // 2), 2)
    assert torch.equal(res[0]["all_part"], res[1]["all_part"])                  # identical on every rank


def test_shard_indices():
    from openstereo_b200 import distributed as osd
    assert osd.shard_indices(8, 0, 2) == [0, 2, 4, 6] and osd.shard_indices(8, 1, 2) == [1, 3, 5, 7]
    assert osd.shard_indices(7, 1, 2) == [1, 3, 5, 0]                           # wraps like DistributedSampler
    assert osd.shard_indices(32, 3, 8) == [3, 11, 19, 27]
    assert osd.shard_indices(0, 0, 2) == []
    assert osd.gather_epe_partials(torch.ones(2, 2))[0].shape == (2, 2)         # world 1: no collective
