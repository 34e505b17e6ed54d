"""N>1 host logic on CPU: world_size-2 gloo processes shard a query set, each "aligns" its block (the oracle stands in
for the GPU here -- this test is about sharding and the ordered gather, not about kernels) and rank 0 must end up with
exactly the unsharded result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outfile):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from sina_b200 import shard, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, m, c, o = synth.synth_msa(200, W=700, L=200, seed=3)
    msa = O.MSA(m, c, o, 700)
    qm, qo = synth.synth_queries(tree, 13, "full", seed=4)   # 13 queries over 2 ranks: uneven blocks
    orc = O.Oracle()
    ix = orc.index_build(msa, 6, 0)
    fp = O.FamParams(fs_min=10, fs_max=10, fs_min_len=50, fs_full_len=180, fs_req_gaps=3)
    lm, lo_off, lo, hi = shard.shard_queries(qm, qo, world, rank)
    res, oc, om, cells, posts, nt = orc.run_batch(ix, msa, lm, lo_off, fp, O.AlignParams(), nthreads=1)
    status = np.array([r.status for r in res], np.int32)
    t = shard.max_over_ranks(dist, 1.0 + rank)
    got = shard.gather_ordered(dist, oc, om, status, lo_off)
    if rank == 0:
        fres, foc, fom, _, _, _ = orc.run_batch(ix, msa, qm, qo, fp, O.AlignParams(), nthreads=1)
        cols, masks, st, off = got
        ok = (t == float(world) and (off == qo).all() and (cols == foc[:len(cols)]).all() and (masks == fom[:len(masks)]).all()
              and (st == np.array([r.status for r in fres], np.int32)).all() and len(cols) == int(qo[-1]))
        open(outfile, "w").write("ok" if ok else "mismatch")
    else:
        assert got is None
    orc.index_free(ix)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from sina_b200 import shard
    for n in (0, 1, 7, 10, 4096, 10001):
        for world in (1, 2, 3, 8):
            blocks = [shard.shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_gather(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"
