"""World-size-2 `gloo` tests of the multi-GPU plumbing (pair sharding, max-over-ranks timing, counters,
result gather) on CPU.  The data path itself has no collective: ranks own disjoint pairs."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from pyfocusr_b200 import dist as fdist

    r, l, w = fdist.init("gloo")
    assert (r, w) == (rank, world)
    mine = fdist.pair_shard(11, rank, world)
    fdist.barrier()
    t_max = fdist.all_reduce_max(10.0 + rank)           # slowest rank's device time
    n_sum = fdist.all_reduce_sum(len(mine))             # units processed by all ranks
    gathered = fdist.gather_objects({"rank": rank, "pairs": mine})
    fdist.finalize()
    ret[rank] = (mine, t_max, n_sum, gathered)


def test_pair_sharding_and_reductions_world2():
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] == [0, 2, 4, 6, 8, 10] and ret[1][0] == [1, 3, 5, 7, 9]
    for r in range(world):
        mine, t_max, n_sum, gathered = ret[r]
        assert t_max == 11.0 and n_sum == 11.0
        assert sorted(sum((g["pairs"] for g in gathered), [])) == list(range(11))


def test_pair_shard_single_process_and_errors():
    sys.path.insert(0, ROOT)
    from pyfocusr_b200 import dist as fdist

    assert fdist.pair_shard(5, 0, 1) == [0, 1, 2, 3, 4]
    assert fdist.pair_shard(3, 3, 4) == [] or fdist.pair_shard(3, 3, 4) == []
    covered = sorted(sum((fdist.pair_shard(1024, r, 8) for r in range(8)), []))
    assert covered == list(range(1024)) and all(len(fdist.pair_shard(1024, r, 8)) == 128 for r in range(8))
    with pytest.raises(ValueError):
        fdist.pair_shard(4, 2, 2)
    assert fdist.all_reduce_max(3.5) == 3.5 and fdist.gather_objects("x") == ["x"]


def test_sub_batches_of_a_rank():
    sys.path.insert(0, ROOT)
    from pyfocusr_b200 import dist as fdist

    mine = fdist.pair_shard(1024, 3, 8)
    for n_sub in (1, 2, 3, 4, 128, 500):
        groups = fdist.sub_batches(mine, n_sub)
        assert len(groups) == min(n_sub, 128) and sum(groups, []) == mine          # contiguous, in order, complete
        assert min(map(len, groups)) >= 1 and max(map(len, groups)) - min(map(len, groups)) <= 1
    assert fdist.sub_batches([7], 2) == [[7]] and fdist.sub_batches([1, 2, 3], 0) == [[1, 2, 3]]
    # a rank without pairs (pair_shard(3, 3, 4) == []) has no sub-batch at all, not one empty one
    assert fdist.pair_shard(3, 3, 4) == [] and fdist.sub_batches([], 2) == []
