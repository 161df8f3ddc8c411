"""CPU tests of the multi-GPU host logic: pair sharding, batch planning (sequential-folder frame
reuse) and the world_size-2 plumbing over gloo (barrier, max-over-ranks timing, host-side gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchpiv_b200.dataset import FrameBatch, plan_batches, shard_range


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (8, 8), (4001, 8), (4000, 3)])
def test_shard_range_is_a_contiguous_partition(n, world):
    blocks = [shard_range(n, r, world) for r in range(world)]
    flat = [i for b in blocks for i in b]
    assert flat == list(range(n))                                  # contiguous, ordered, complete
    sizes = [len(b) for b in blocks]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(n, world, world)


def test_plan_batches_sequential_uploads_each_frame_once():
    files = [f"f{i:04d}.bmp" for i in range(12)]
    seq = list(zip(files[:-1], files[1:]))                         # 11 pairs, pair i = (f_i, f_i+1)
    batches = plan_batches(seq, 4)
    assert [len(b) for b in batches] == [4, 4, 3]
    assert all(b.chained for b in batches)
    assert batches[0].files == files[0:5] and batches[2].files == files[8:12]
    for b in batches:
        for i, (fa, fb) in enumerate(b.pairs):
            assert b.files[b.index_a[i]] == fa and b.files[b.index_b[i]] == fb
        assert b.index_b == [i + 1 for i in b.index_a]             # overlapping views [0:K], [1:K+1]
    assert sum(len(b.files) for b in batches) == 11 + 3            # K + 1 frames per batch, not 2 K
    # a shard starts where its block starts
    shard = plan_batches(seq, 4, shard_range(len(seq), 1, 2))
    assert shard[0].first_pair == 6 and sum(len(b) for b in shard) == 5


def test_plan_batches_pairs_mode_keeps_two_stacks():
    files = [f"f{i}.bmp" for i in range(10)]
    pairs = list(zip(files[::2], files[1::2]))
    (b0, b1) = plan_batches(pairs, 3)
    assert not b0.chained and len(b0) == 3 and len(b1) == 2
    assert b0.files == [files[0], files[2], files[4], files[1], files[3], files[5]]
    assert b0.index_a == [0, 1, 2] and b0.index_b == [3, 4, 5]
    single = FrameBatch(0, [pairs[0]])
    assert single.chained and single.files == list(pairs[0])       # one pair: 2 frames either way
    with pytest.raises(ValueError):
        plan_batches(pairs, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pairs, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from torchpiv_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.dist_env() == (rank, world, rank)
        block = sharding.shard_range(n_pairs, rank, world)
        # every rank "processes" its block: the payload encodes the pair index
        local = [(i, np.full((3, 4), float(i)), np.full((3, 4), -float(i))) for i in block]
        sharding.barrier()
        slowest = sharding.max_over_ranks(10.0 + 5.0 * rank)        # rank 1 is the slow one
        assert slowest == 10.0 + 5.0 * (world - 1)
        merged = sharding.gather_results(local, dst=0)
        if rank == 0:
            assert [m[0] for m in merged] == list(range(n_pairs))
            assert all(np.all(m[1] == m[0]) and np.all(m[2] == -m[0]) for m in merged)
            value = n_pairs / (slowest * 1e-3)                       # whole-job units / max-over-ranks time
            open(os.path.join(out_dir, "ok"), "w").write(f"{value:.3f}")
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sharding_and_gather(tmp_path):
    n_pairs, world = 11, 2
    mp.spawn(_worker, args=(world, _free_port(), n_pairs, str(tmp_path)), nprocs=world, join=True)
    assert float(open(tmp_path / "ok").read()) == pytest.approx(n_pairs / 15e-3, rel=1e-6)


def test_single_process_helpers_need_no_process_group():
    from torchpiv_b200 import sharding
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.gather_results([(2, "b"), (0, "a")]) == [(0, "a"), (2, "b")]
    sharding.barrier()
