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


def _ragged_worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from probenb200.pipeline import all_gather_ragged
    lo, hi = shard_range(total, rank, world)
    rows = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 5)
    parts = all_gather_ragged(rows)
    q.put((rank, [p.shape[0] for p in parts], torch.cat(parts)[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_ragged_last_shard_is_gathered_in_global_order():
    """A validation set that does not divide by the rank count (7 rows over 2 ranks -> 4 + 3; FLIR's 1013 pairs over 8 GPUs):
    the padded gather returns every rank's rows, unpadded, in InferenceSampler order."""
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ragged_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, sizes, col in results:
        assert sizes == [4, 3] and col == [float(i) for i in range(total)]


def test_detection_rows_round_trip():
    """The fixed-stride per-image rows the multi-GPU save_predictions CLI gathers decode back to the same Instances."""
    import importlib.util
    from probenb200 import detector
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("cli_save_host", os.path.join(root, "demo", "FLIR", "demo_FLIR_save_predictions.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    B, K = 3, 3
    buf = detector.DetectionBuffers(B, K, "cpu")
    g = torch.Generator().manual_seed(0)
    buf.counts[:] = torch.tensor([2, 0, 100], dtype=torch.int32)
    for name in ("boxes", "scores", "class_logits", "probs", "vars"):
        getattr(buf, name).copy_(torch.rand(getattr(buf, name).shape, generator=g))
    buf.classes.copy_(torch.randint(0, K, buf.classes.shape, generator=g, dtype=torch.int32))
    rows = cli.detection_rows(buf, B)
    assert rows.shape == (B, cli.row_width(K))
    want = buf.to_instances([(512, 640)] * B)
    got = cli.rows_to_instances(rows, K, (512, 640))
    for w, g_ in zip(want, got):
        assert len(w) == len(g_)
        assert torch.equal(w.pred_boxes.tensor, g_.pred_boxes.tensor) and torch.equal(w.scores, g_.scores)
        assert torch.equal(w.pred_classes, g_.pred_classes) and torch.equal(w.class_logits, g_.class_logits)
        assert torch.equal(w.prob_score, g_.prob_score) and torch.equal(w.vars, g_.vars)
