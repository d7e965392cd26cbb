"""N>1 host logic on CPU: two gloo ranks cut ONE input with the product's own sharding rule
(pfac_job_plan through pf.plan_shard -- what pfac_job_run and bench.py --scaling strong apply: a
contiguous range of start positions per rank plus a halo of max_pat_len-1 readable bytes), scan
their shards with the oracle and all-gather the results; rank 0 checks the concatenation against
one scan of the whole input.  No collective is on the data path itself."""
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
    for p in (ROOT, HERE, os.path.join(ROOT, "tools")):
        sys.path.insert(0, p)
    import pfac_synth as synth
    import phfpfac_b200 as pf
    from _oracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pats = synth.synth_patterns(1, 400, 3, 4, 64) + b"GET /\nHost: www.\n"
    o = Oracle(pats, 1, 256)
    mpl = o.max_pat_len
    whole = synth.synth_text(1, 4, n, patterns=pats)          # every rank generates the same seeded input
    start, n_starts, n_valid = pf.plan_shard(n, world, mpl, rank)
    # force a match that straddles the rank boundary: it starts in rank 0's range, ends in rank 1's
    b0 = pf.plan_shard(n, world, mpl, 0)[1]
    straddle = b"Host: www."
    whole[b0 - 4:b0 + 6] = np.frombuffer(straddle, dtype=np.uint8)
    pos, ids = o.scan(whole[start:start + n_valid])
    keep = pos < n_starts                                       # a shard reports the matches that START in it
    pos, ids = pos[keep] + start, ids[keep]
    cnt = torch.tensor([len(pos)], dtype=torch.int64)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, cnt)
    m = int(max(c.item() for c in counts))
    pad = torch.full((2, m), -1, dtype=torch.int64)
    pad[0, :len(pos)] = torch.from_numpy(pos)
    pad[1, :len(ids)] = torch.from_numpy(ids.astype(np.int64))
    gathered = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad)
    if rank == 0:
        wpos, wids = o.scan(whole)
        gp = np.concatenate([g[0, :int(c.item())].numpy() for g, c in zip(gathered, counts)])
        gi = np.concatenate([g[1, :int(c.item())].numpy() for g, c in zip(gathered, counts)])
        ok = bool(np.array_equal(gp, wpos) and np.array_equal(gi, wids.astype(np.int64)))
        ok = ok and bool((wpos == b0 - 4).any())
        # the plan itself: the shards tile [0, n) and every halo is max_pat_len-1 bytes (clipped to n)
        plan = [pf.plan_shard(n, world, mpl, r) for r in range(world)]
        ok = ok and plan[0][0] == 0 and all(plan[r][0] + plan[r][1] == plan[r + 1][0] for r in range(world - 1))
        ok = ok and plan[-1][0] + plan[-1][1] == n
        ok = ok and all(p[2] == min(p[1] + mpl - 1, n - p[0]) for p in plan)
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
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5 * 65536 + 123, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok and n >= 6
