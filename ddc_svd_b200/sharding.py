"""CPU MODEL of the multi-GPU host logic — SURVEY.md 8e.  The product's multi-GPU path lives in the library
(ddc_svd_b200/host/svd_gpu.c: groups of ranks, NCCL, compact-WY panels broadcast while the factorization runs;
bench.py and the GPU tests call it through the C ABI).  This module restates its partitioning and its sequence of
exchanges over torch.distributed so that the orchestration can be exercised on CPU with gloo
(tests/test_sharding_gloo.py, the oracle standing in for the kernels); shard_range() is pinned to the library's
svdgpu_shard_range() and svdgpu_plan_chunks() by tests/test_host_logic.py.

The bidiagonalization and the dDC singular values do not shard (n dependent steps with global
reductions): they run on rank 0.  The twisted-factorization vector solves and the back-transform
are independent per singular value / column, so they shard by contiguous singular-value blocks.
The only exchanges are
    broadcast(rank 0 -> all): reflector matrix A_mod, alpha, beta, sigma     ("all-gather the bidiagonal")
    all_gather: U and V column blocks (and the polished singular values)
over torch.distributed (NCCL on the GPUs, gloo in the CPU tests).  The compute steps are passed in
as callables so that the same orchestration is exercised on CPU (tests/test_sharding_gloo.py, with
the oracle standing in) and on the GPUs (bench.py, with the C-ABI device entry points).
"""
from typing import Callable, Tuple


def shard_range(mn: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(block, i0, ns): equal contiguous blocks of ceil(mn/world) singular values; the last ranks
    may get fewer (or none).  block is the padded per-rank column count used by all_gather."""
    block = (mn + world - 1) // world
    i0 = min(rank * block, mn)
    ns = max(0, min(block, mn - i0))
    return block, i0, ns


def sharded_svd_step(dist, rank: int, world: int, mn: int,
                     values_fn: Callable[[], None], vectors_fn: Callable[[int, int], None],
                     bcast_tensors, gather_pairs) -> None:
    """One svd_gpu step on `world` ranks.

    values_fn()          rank 0 only: bidiagonalization + dDC, fills bcast_tensors in place
    vectors_fn(i0, ns)   every rank: columns [i0, i0+ns) of U and V into its block tensors
    bcast_tensors        tensors broadcast from rank 0 (A_mod, alpha, beta, sigma)
    gather_pairs         [(full, block), ...] for all_gather_into_tensor
    """
    if rank == 0:
        values_fn()
    if world > 1:
        for t in bcast_tensors:
            dist.broadcast(t, 0)
    _, i0, ns = shard_range(mn, world, rank)
    if ns > 0:
        vectors_fn(i0, ns)
    if world > 1:
        for full, blk in gather_pairs:
            dist.all_gather_into_tensor(full, blk)
