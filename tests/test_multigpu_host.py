"""N > 1 host logic on CPU: world_size-2 gloo processes shard a pair batch with the InferenceSampler rule and
exchange their flat result buffers with the pipeline's single all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from probenb200.pipeline import FusedOutput, all_gather_flat, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total_pairs, rank, world)
    B, M = hi - lo, 2
    out = FusedOutput(B, M, "cpu")
    # fake fused results: image g (global index) has (g % 3) detections whose score encodes g
    row = 0
    for b in range(B):
        g = lo + b
        out.offsets[b * M] = row
        out.offsets[b * M + 1] = row
        n = g % 3
        out.counts[b] = n
        for i in range(n):
            out.scores[row + i] = float(g) + 0.1 * i
            out.classes[row + i] = g % 2
            out.boxes[row + i] = torch.tensor([g, i, g + 10, i + 10], dtype=torch.float32)
        row += n
    out.offsets[B * M] = row
    full = all_gather_flat(out.flat)
    assert full.shape[0] == world
    merged = []
    for r in range(world):
        rlo, rhi = shard_range(total_pairs, r, world)
        merged += FusedOutput.split(full[r], rhi - rlo, M)
    q.put((rank, [(None if m is None else (m[1].tolist(), m[2].tolist(), m[0][:, 0].tolist())) for m in merged]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_gather_reassembles_global_order():
    world, total = 2, 8  # equal shards: the all-gather needs same-sized buffers per rank
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, merged in results:
        assert len(merged) == total
        for g, m in enumerate(merged):
            if g % 3 == 0:
                assert m is None
            else:
                scores, classes, x1 = m
                assert len(scores) == g % 3 and abs(scores[0] - g) < 1e-6 and classes[0] == float(g % 2) and x1[0] == float(g)
