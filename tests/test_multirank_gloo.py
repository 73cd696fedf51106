"""world_size-2 gloo test of the N > 1 path on CPU: reads shard by rank with no data-path collective, the per-read
records are gathered on rank 0 in input order, and the step time is the max over ranks.  (The per-rank compute is the
CUDA library on a GPU box; here each rank fills its records with a deterministic function of the read so that the
plumbing -- bounds, gather order, completeness, max-reduce -- is what is being tested.)"""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_reads, out_path):
    sys.path.insert(0, os.path.join(ROOT, "telomere-analyzer_b200"))
    import torch.distributed as dist
    from nanotel_b200._lib import RESULT_DTYPE
    from nanotel_b200.shard import gather_records, max_over_ranks, shard_bounds
    from nanotel_b200.synth import synth_reads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buf, off, meta = synth_reads(n_reads, seed=99, median_len=800, min_len=50, max_len=5000)
    bounds = shard_bounds(meta["lengths"], world)
    b, e = bounds[rank]
    local = np.zeros(e - b, RESULT_DTYPE)
    for k, i in enumerate(range(b, e)):
        seq = buf[int(off[i]):int(off[i + 1])]
        local["n_win"][k] = len(seq)
        local["win_offset"][k] = int(seq.astype(np.int64).sum()) * 31 + i
    full = gather_records(local, bounds, rank, world)
    tmax = max_over_ranks(1.0 + rank, world)
    assert tmax == float(world)
    if rank == 0:
        exp_len = meta["lengths"]
        assert np.array_equal(full["n_win"], exp_len)
        exp = np.array([int(buf[int(off[i]):int(off[i + 1])].astype(np.int64).sum()) * 31 + i for i in range(n_reads)])
        assert np.array_equal(full["win_offset"], exp)
        open(out_path, "w").write("ok %d" % len(full))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "done.txt")
    mp.spawn(_worker, args=(2, _free_port(), 301, out), nprocs=2, join=True)
    assert open(out).read() == "ok 301"
