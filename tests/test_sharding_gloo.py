"""The multi-GPU path on CPU: world_size-2 gloo processes each take the pair-index shard
the library would give their GPU (jlp_shard_range) and generate it with the oracle; the
rank-ordered concatenation must equal the unsharded output byte for byte -- duplicate
chains that straddle the shard boundary included.  No data-path collective exists; gloo
is used only to gather the results for the comparison (as bench.py uses NCCL only for its
barrier and max-over-ranks timing)."""
import ctypes as C
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import jackalope_b200 as J
from jackalope_b200 import _lib
from oracle.compare import oracle_run

ARGS = dict(n_reads=3000, read_length=100, paired=True, seed=18, prob_dup=0.45, read_pool_size=1000)


def _genome():
    return J.random_genome(3, [4000, 9000, 2000], seed=10)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = _genome()
    n_pairs = ARGS["n_reads"] // 2
    lo, hi = C.c_uint64(), C.c_uint64()
    assert _lib.lib().jlp_shard_range(0, n_pairs, rank, world, C.byref(lo), C.byref(hi)) == 0
    o = oracle_run(g, ARGS["n_reads"], ARGS["read_length"], True, ARGS["seed"], lo=lo.value, hi=hi.value,
                   prob_dup=ARGS["prob_dup"], read_pool_size=ARGS["read_pool_size"])
    parts = [None] * world
    dist.all_gather_object(parts, (lo.value, hi.value, o["r1"], o["r2"]))
    if rank == 0:
        import pickle
        with open(out_path, "wb") as fh:
            pickle.dump(parts, fh)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_concatenate_to_the_unsharded_output(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    import pickle
    out_path = str(tmp_path / "parts.pkl")
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    parts = pickle.load(open(out_path, "rb"))
    full = oracle_run(_genome(), ARGS["n_reads"], ARGS["read_length"], True, ARGS["seed"], prob_dup=ARGS["prob_dup"],
                      read_pool_size=ARGS["read_pool_size"], want_ledger=True)
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == ARGS["n_reads"] // 2
    assert b"".join(p[2] for p in parts) == full["r1"]
    assert b"".join(p[3] for p in parts) == full["r2"]
    # the boundary really is inside a duplicate chain for this seed, or at least chains exist near it
    plan = np.concatenate(full["plan"])
    leaders = plan[:, 3]
    assert np.sum(leaders != np.arange(leaders.size)) > 300
