"""world_size-2 gloo test of the evaluation reduction: per-rank integer state summed with
one all-reduce reproduces the single-process reference metrics bit for bit, including the dict insertion
order that fixes the float summation order of the mIoU (Q11)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rangeclip_oracle as O

INT32_MAX = 2**31 - 1


def _batch_state(seg, topk, E, cmap, C):
    """Numpy statement of rc_eval_hist + rc_eval_fold for one batch (test helper, CPU)."""
    gt = seg.reshape(-1)
    k = topk.shape[1]
    tk = np.transpose(topk, (0, 2, 3, 1)).reshape(-1, k)
    t1 = tk[:, 0]
    ge, p1 = cmap[gt], cmap[t1]
    hit = (cmap[tk] == ge[:, None]).any(1)
    orc = np.where(hit, ge, t1)
    h = [np.bincount(ge, minlength=C), np.bincount(p1, minlength=C), np.bincount(ge[ge == p1], minlength=C),
         np.bincount(orc, minlength=C), np.bincount(ge[orc == ge], minlength=C)]
    present = (h[0] + h[1]) > 0
    acc = np.stack([h[2], h[0] + h[1] - h[2], h[4], h[0] + h[3] - h[4]]) * present[None]
    counters = np.array([E[gt, t1].sum(), E[gt[:, None], tk].any(1).sum(), gt.size])
    return acc.astype(np.int64), counters.astype(np.int64), present


def _make_data(seed=0, C=40, k=5, n_batches=5, B=2, H=8, W=8):
    rng = np.random.default_rng(seed)
    eq = {i: {i} for i in range(C)}
    for a, b in [(3, 7), (7, 11), (11, 2), (20, 21), (5, 30)]:
        eq[a].add(b); eq[b].add(a)
    E = O.build_equivalence_tensor(eq, C); cmap = O.build_equivalence_class_map(E)
    segs = [rng.integers(0, 10 + 6 * i, (B, H, W)) for i in range(n_batches)]
    topks = []
    for s in segs:
        t = rng.integers(0, C, (B, k, H, W))
        hit = rng.random((B, H, W)) < 0.6
        t[:, 0][hit] = s[hit]
        topks.append(t)
    return C, E, cmap, segs, topks


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rangeclip_b200.distributed import reduce_state_tensors, shard_batches
    from rangeclip_b200.evaluation import finalize_metrics
    C, E, cmap, segs, topks = _make_data()
    acc = torch.zeros(4, C, dtype=torch.int64); counters = torch.zeros(3, dtype=torch.int64)
    first_seen = torch.full((C,), INT32_MAX, dtype=torch.int32)
    for gb in shard_batches(len(segs), rank, world):
        a, c, present = _batch_state(segs[gb], topks[gb], E, cmap, C)
        acc += torch.from_numpy(a); counters += torch.from_numpy(c)
        fs = torch.where(torch.from_numpy(present), torch.tensor(gb, dtype=torch.int32), torch.tensor(INT32_MAX, dtype=torch.int32))
        first_seen = torch.minimum(first_seen, fs)
    reduce_state_tensors(acc, counters, first_seen)
    valid = set(cmap[segs[-1].reshape(-1)].tolist())
    fin = finalize_metrics(acc, counters, first_seen, valid)
    if rank == 0:
        out_q.put({k_: fin[k_] for k_ in ("mIoU_t1", "mIoU_tk", "pixel_accuracy_t1", "pixel_accuracy_tk", "intersection_topk",
                                         "union_topk", "intersection_top1", "union_top1")})
    dist.destroy_process_group()


def test_two_rank_metric_reduction_is_bit_exact():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C, E, cmap, segs, topks = _make_data()
    st = O.MetricState()
    for seg, topk in zip(segs, topks):
        O.metrics_accumulate(st, seg.reshape(-1), np.transpose(topk, (0, 2, 3, 1)).reshape(-1, topk.shape[1]), E, cmap)
    ref = O.metrics_finalize(st, segs[-1], cmap)
    for key in ("mIoU_t1", "mIoU_tk", "pixel_accuracy_t1", "pixel_accuracy_tk"):
        assert got[key] == ref[key], key
    for nm in ("intersection_top1", "union_top1", "intersection_topk", "union_topk"):
        assert list(got[nm].items()) == list(getattr(st, nm).items()), nm


def test_shard_batches_partition():
    from rangeclip_b200.distributed import shard_batches
    for n in (0, 1, 7, 79):
        for world in (1, 2, 8):
            seen = sorted(b for r in range(world) for b in shard_batches(n, r, world))
            assert seen == list(range(n))
