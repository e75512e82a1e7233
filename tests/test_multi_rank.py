"""The N>1 path on CPU: two ranks over gloo shard a batch (`shard_bounds`) and gather their analyzer outputs on
rank 0 (`gather_results` = the one collective of a multi-GPU job; NCCL on the GPU box, gloo here)."""
import os
import socket

import numpy as np
import pytest

from maxent_b200 import batched


def test_shard_bounds_cover_the_batch_exactly():
    for n, w in ((65536, 8), (10, 3), (5, 8), (0, 4), (4096, 1)):
        cuts = [batched.shard_bounds(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[r][1] == cuts[r + 1][0] for r in range(w - 1))
        sizes = [hi - lo for lo, hi in cuts]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, tmp):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = batched.shard_bounds(n_total, world, rank)
    out = batched.BatchedMaxEntResult()
    idx = np.arange(lo, hi)
    out.alpha_index = np.stack([idx % 7, idx % 5, idx % 3, -np.ones_like(idx), -np.ones_like(idx)], 1).astype(np.int32)
    out.chi2 = idx[:, None] + np.linspace(0, 1, 6)[None, :]
    out.A_out = np.broadcast_to(idx[:, None, None].astype(float), (hi - lo, 5, 4)).copy()
    got = batched.gather_results(out, dst=0)
    if rank == 0:
        np.savez(os.path.join(tmp, "gathered.npz"), **got)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [11, 1, 8])
def test_two_rank_gather_over_gloo(tmp_path, n_total):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    idx = np.arange(n_total)
    np.testing.assert_array_equal(got["alpha_index"][:, 0], idx % 7)
    np.testing.assert_array_equal(got["chi2"][:, 0], idx.astype(float))
    assert got["A_out"].shape == (n_total, 5, 4)
    np.testing.assert_array_equal(got["A_out"][:, 2, 1], idx.astype(float))
