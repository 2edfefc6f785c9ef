"""CPU test of the N > 1 path's host logic: world_size 2 and 3 over gloo, the oracle standing in
for the device kernels.  Checks that the sharded orchestration (rank-0 values, broadcast,
per-rank column blocks, all-gather) reproduces the single-process result exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util

ROOT = util.ROOT


def test_shard_range_partitions():
    from ddc_svd_b200.sharding import shard_range
    for mn in (1, 2, 7, 64, 100, 4096):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                blk, i0, ns = shard_range(mn, world, r)
                assert blk * world >= mn and 0 <= ns <= blk
                cover += list(range(i0, i0 + ns))
            assert cover == list(range(mn))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util as u
    from ddc_svd_b200.sharding import shard_range, sharded_svd_step
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = u.p
    orc = u.oracle()
    A = u.rand_matrix(n, n)
    A_mod = torch.zeros((n, n), dtype=torch.float64)          # column-major image: row j = column j
    alpha = torch.zeros(n, dtype=torch.float64); beta = torch.zeros(n + 1, dtype=torch.float64)
    sigma = torch.zeros(n, dtype=torch.float64)
    blk, _, _ = shard_range(n, world, rank)
    Ublk = torch.zeros((blk, n), dtype=torch.float64); Vblk = torch.zeros((blk, n), dtype=torch.float64)
    Ufull = torch.zeros((blk * world, n), dtype=torch.float64); Vfull = torch.zeros((blk * world, n), dtype=torch.float64)

    def values_fn():
        Af, al, be = u.oracle_bidiag(A)
        A_mod.copy_(torch.from_numpy(np.ascontiguousarray(Af.T)))
        alpha.copy_(torch.from_numpy(al)); beta[: n - 1].copy_(torch.from_numpy(be))
        s = np.zeros(n)
        orc.orc_ddc_values(n, p(alpha.numpy()), p(beta.numpy()), p(s))
        sigma.copy_(torch.from_numpy(s))

    def vectors_fn(i0, ns):
        al, be, sg = alpha.numpy(), beta.numpy(), sigma.numpy()
        X = np.zeros(n * n); Y = np.zeros(n * n)
        orc.orc_right_vectors(n, n, p(al), p(be), p(sg), p(X))
        orc.orc_left_vectors(n, n, p(al), p(be), p(sg), p(X), p(Y))
        Af = np.asfortranarray(A_mod.numpy().T); AT = np.asfortranarray(Af.T)
        for t in range(ns):
            uo = np.zeros(n); vo = np.zeros(n)
            orc.orc_apply_left(n, n, i0 + t, p(Af), p(Y), p(uo))
            orc.orc_apply_right(n, n, i0 + t, p(AT), p(X), p(vo))
            Ublk[t].copy_(torch.from_numpy(uo)); Vblk[t].copy_(torch.from_numpy(vo))

    sharded_svd_step(dist, rank, world, n, values_fn, vectors_fn, [A_mod, alpha, beta, sigma],
                     [(Ufull, Ublk), (Vfull, Vblk)])
    dist.barrier()
    if rank == world - 1:                      # a non-root rank holds the complete result too
        np.savez(os.path.join(out_dir, "sharded.npz"), sigma=sigma.numpy(), U=Ufull.numpy()[:n].T,
                 V=Vfull.numpy()[:n].T)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_path_matches_single_process(tmp_path, world):
    n = 45
    port = 29500 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    s, U, V, _ = util.oracle_svd(util.rand_matrix(n, n))
    assert np.array_equal(got["sigma"], s)
    assert np.array_equal(got["U"], U) and np.array_equal(got["V"], V)
