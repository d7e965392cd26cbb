"""N>1 host logic on CPU: two gloo ranks cut their shards exactly as bench.py does (own text +
halo from the next rank), scan them with the oracle and all-gather the results; rank 0 checks the
concatenation against one scan of the whole input.  No collective is on the data path itself."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import bench
    import phfpfac_b200 as pf
    from _oracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pats = pf.synth_patterns(1, 400, 3, 4, 64) + b"GET /\nHost: www.\n"
    o = Oracle(pats, 1, 256)
    mpl = o.max_pat_len
    buf, n_valid = bench.make_shard(pf, pats, mpl, 1, 4, n, rank, world)
    # force a match that straddles the rank boundary: the tail of rank 0 + the head of rank 1
    straddle = b"Host: www."
    if rank == 0:
        buf[n - 4:n] = np.frombuffer(straddle[:4], dtype=np.uint8)
        buf[n:n + 6] = np.frombuffer(straddle[4:], dtype=np.uint8)
    pos, ids = o.scan(buf[:n_valid])
    keep = pos < n
    pos, ids = pos[keep] + rank * n, ids[keep]
    cnt = torch.tensor([len(pos)], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, cnt)
    m = int(max(c.item() for c in counts))
    pad = torch.full((2, m), -1, dtype=torch.int64)
    pad[0, :len(pos)] = torch.from_numpy(pos)
    pad[1, :len(ids)] = torch.from_numpy(ids.astype(np.int64))
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    ok = True
    if rank == 0:
        whole = np.concatenate([bench.make_shard(pf, pats, mpl, 1, 4, n, r, world)[0][:n] for r in range(world)])
        whole[n - 4:n + 6] = np.frombuffer(straddle, dtype=np.uint8)
        wpos, wids = o.scan(whole)
        gp = np.concatenate([g[0, :int(c.item())].numpy() for g, c in zip(gathered, counts)])
        gi = np.concatenate([g[1, :int(c.item())].numpy() for g, c in zip(gathered, counts)])
        ok = bool(np.array_equal(gp, wpos) and np.array_equal(gi, wids.astype(np.int64)))
        ok = ok and bool(((wpos == n - 4)).any())
        q.put((ok, len(wpos)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3 * 65536 + 123, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok and n >= 6
