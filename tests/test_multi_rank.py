"""World-size-2 `gloo` test of the host-side multi-GPU logic: every rank computes the SAME deterministic
LPT partition through the C ABI (no GPU needed), the shards are a disjoint cover with balanced frame counts, and
per-utterance results gathered by rank reassemble in the caller's order.  The data path itself has no collective."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
    import qwen3tts_cuda as q
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1003)                       # BASELINE config 3: 512 utterances of 25..750 frames
    lens = rng.integers(25, 751, size=512)
    part = q.partition_lpt(lens, world)
    parts = [None] * world
    dist.all_gather_object(parts, part.tolist())
    assert all(p == parts[0] for p in parts)                # identical map on every rank, no communication needed
    mine = [i for i in range(len(lens)) if part[i] == rank]
    # each rank "decodes" its shard: the host-side length arithmetic of the boundary (ST.swift:831-833, Q3.swift:746-752)
    results = {i: (q.trim_length(int(lens[i]) * 1920, int(lens[i]) * 1920 - 7), int(lens[i])) for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    merged = {}
    for g in gathered:
        assert not (merged.keys() & g.keys())               # disjoint
        merged.update(g)
    assert sorted(merged) == list(range(len(lens)))         # cover
    assert all(merged[i] == (int(lens[i]) * 1920 - 7, int(lens[i])) for i in merged)
    load = torch.tensor([float(sum(lens[i] for i in mine))])
    loads = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(loads, load)
    total = float(sum(l.item() for l in loads))
    assert total == float(lens.sum())
    assert max(l.item() for l in loads) <= 1.001 * total / world      # LPT on 512 items: within 0.1 % of perfect
    # the bench's timing reduction: MAX over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == 10.0 + world - 1
    if rank == 0:
        ret.put("ok")
    dist.destroy_process_group()


def test_lpt_sharding_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) == "ok"


def test_lpt_edge_cases():
    sys.path.insert(0, os.path.join(ROOT, "swift-qwen3-tts_b200", "python"))
    import qwen3tts_cuda as q
    assert q.partition_lpt([], 4).tolist() == []
    assert q.partition_lpt([5], 8).tolist() == [0]
    p = q.partition_lpt([10, 10, 10, 10], 2).tolist()
    assert sorted(p) == [0, 0, 1, 1]
    p = q.partition_lpt([100, 1, 1, 1, 1], 2).tolist()       # the long utterance alone, the rest together
    assert p[0] != p[1] and len(set(p[1:])) == 1
