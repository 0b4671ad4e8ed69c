"""The N > 1 path on CPU: world_size-2 gloo run of the host-side sharding logic (problem assignment,
max-over-ranks timing, split-vector gathering) with the CPU oracle standing in for the per-rank solve."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    import chainb200 as cp
    import pyoracle as ref
    from chainb200 import parallel, synth

    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_problems, K = 4, 8
    mine = parallel.my_problems(n_problems, rank, world)
    mtd = cp.LazyBisectCostBottleneckSplitter(cp.AffineConnectivityModel(0, 10, 1, 100), 0.01)
    local = {i: ref.partition_stripe(synth.erdos_renyi(400 + 100 * i, 6), K, mtd).spl for i in mine}
    table = parallel.gather_split_vectors(local, n_problems, K)
    t = parallel.max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        out.put((mine, table.tolist(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    mine, table, t = out.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert mine == [0, 2]
    assert t == [2.0, 10.0]
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import chainb200 as cp
    import pyoracle as ref
    from chainb200 import synth

    mtd = cp.LazyBisectCostBottleneckSplitter(cp.AffineConnectivityModel(0, 10, 1, 100), 0.01)
    for i in range(4):
        exp = ref.partition_stripe(synth.erdos_renyi(400 + 100 * i, 6), 8, mtd).spl
        assert table[i] == exp.tolist()


def test_row_blocks_cover_all_rows():
    import sys

    for p in (ROOT,):
        if p not in sys.path:
            sys.path.insert(0, p)
    from chainb200 import parallel

    for m in (1, 7, 100, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [parallel.row_block(m, r, world) for r in range(world)]
            assert blocks[0][0] == 1 and blocks[-1][1] == m + 1
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
