"""Multi-process sharding logic on CPU: gloo backend, world_size 2, rendezvous on 127.0.0.1.
The compute function here is the CPU oracle (test infrastructure) standing in for Engine.count_reads;
the partition, the per-rank slicing and the host-side gather are the code under test."""
import os
import socket

import numpy as np
import pytest

from strkit_b200.sharding import count_reads_sharded, estimated_cost, partition_catalog


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    cost = rng.integers(1, 100, 1000).astype(np.int64)
    cost[10] = 50_000  # one pathologically expensive locus
    for n in (1, 2, 3, 8):
        b = partition_catalog(cost, n)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all() and len(b) == n + 1
        shares = np.array([cost[b[i]:b[i + 1]].sum() for i in range(n)])
        assert shares.sum() == cost.sum()
        if n == 2:
            assert abs(int(shares[0]) - int(shares[1])) <= 50_000
    assert partition_catalog(np.zeros(0, dtype=np.int64), 2).tolist() == [0, 0, 0]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    import torch.distributed as dist

    from strkit_b200 import synth
    from tests import oracle_lib

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = synth.generate(synth.CONFIGS[1], 24, seed=5).to_host()
        orc = oracle_lib.load()

        def compute(b):
            return orc.count_loci(b.arena, b.seq_off, b.lens, b.est_cn, b.read_begin, b.motif_off, b.motif_len)[0]

        res = count_reads_sharded(batch, compute, rank, world)
        if rank == 0:
            q.put((res, compute(batch)))
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_count_matches_single_process_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded, single = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sharded.shape == single.shape and np.array_equal(sharded, single)


def test_estimated_cost_follows_read_begin():
    from strkit_b200 import LocusReads, pack_loci

    b = pack_loci([LocusReads("CAG", [2, 2], ["CAGCAG", "CAGCAG"], ["AA", "AA"], ["TT", "TT"]),
                   LocusReads("AT", [1], ["AT"], ["A"], ["G"])])
    assert estimated_cost(b).tolist() == [2 * 100, 16]
