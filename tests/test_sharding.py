"""Host-side logic of the multi-GPU path, on CPU: frame-range plans, the compact wire format and a
world_size-2 `gloo` run in which every rank processes its frame range with a one-frame halo (the
per-rank engine here is the CPU oracle — the CUDA library needs a GPU) and rank 0 gathers.  The
gathered result must equal the single-process run over the whole video byte for byte."""
import os
import socket

import numpy as np
import pytest

from locomouse_cpp_b200 import sharding, synth
from locomouse_cpp_b200.types import Results, diff_results


def test_frame_range_partitions():
    for n in (0, 1, 7, 10, 1000, 10001):
        for world in (1, 2, 3, 4, 8):
            got = [sharding.frame_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.frame_range(10, 2, 2)


def test_plan_whole_videos_or_ranges():
    p = sharding.plan(64, 10000, 8)             # SURVEY config 4: whole videos, no halo
    assert all(len(x) == 8 for x in p) and not any(s.needs_halo for x in p for s in x)
    assert sorted(s.video for x in p for s in x) == list(range(64))
    p = sharding.plan(1, 1000, 4)               # one video cut into ranges with halos
    assert [(s.f0, s.f1) for x in p for s in x] == [(0, 250), (250, 500), (500, 750), (750, 1000)]
    assert [s.needs_halo for x in p for s in x] == [False, True, True, True]


SMALL = dict(n_rows=160, n_cols=420, side_h=64, bb_w=128, bb_h_side_tm=56, tsize=12, mouse_scale=0.32, cand_cap=32,
             det_cap=2048, match_cap=128)


def _small_problem(n=10):
    spec = synth.SynthSpec(**SMALL)
    return synth.make_problem(spec, n, seed=77)


def test_pack_unpack_roundtrip(oracle):
    cfg, model, bkg, calib, frames, bx, bs, bb = _small_problem(6)
    res = oracle.detect(cfg, model, bkg, calib, frames.numpy(), bx, bs, bb)
    buf = sharding.pack(res)
    back = sharding.unpack(buf)
    assert diff_results(back, res) == [] and back.checksum() == res.checksum()
    full = sum(getattr(res, a).nbytes for a in Results.ARRAYS)
    assert buf.nbytes < full / 4            # only live entries travel
    empty = sharding.unpack(sharding.pack(Results(0, 32, 128)))
    assert empty.n == 0


def test_weighted_counts_follow_the_weights():
    for total in (0, 5, 1000, 80000):
        for w in ([1.0], [1, 1], [23.3] * 4 + [35.6] * 4, [3, 1, 2]):
            c = sharding.weighted_counts(total, w)
            assert sum(c) == total and all(x >= 0 for x in c)
            if total >= 100 * len(w):
                ideal = [total * x / sum(w) for x in w]
                assert all(abs(a - b) <= 1 for a, b in zip(c, ideal))
    assert sharding.weighted_counts(20000, [55.5, 55.5]) == [10000, 10000]
    c = sharding.weighted_counts(80000, [23.3] * 4 + [35.6] * 4, quantum=512)
    assert sum(c) == 80000 and all(x % 512 == 0 for x in c[:4]) and min(c) >= 512
    r = sharding.weighted_ranges(10, [1, 1], 4)
    assert r[0][0] == 0 and r[-1][1] == 10 and r[0][1] == r[1][0]
    with pytest.raises(ValueError):
        sharding.weighted_counts(10, [1, 0])


def _worker(rank, world, port, n, out_path, weights=None, shared=False):
    import torch.distributed as dist

    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, model, bkg, calib, frames, bx, bs, bb = _small_problem(n)
    frames = frames.numpy()
    if weights is None:
        (shard,) = sharding.plan(1, n, world)[rank]
        f0, f1 = shard.f0, shard.f1
    else:   # shards sized by per-rank weights (measured host-to-device rates on the GPU box): unequal counts, ragged gather
        f0, f1 = sharding.weighted_ranges(n, weights)[rank]
    prev = frames[f0 - 1] if f0 > 0 else None
    local = oracle.detect(cfg, model, bkg, calib, frames[f0:f1], bx[f0:f1], bs[f0:f1], bb[f0:f1], prev_frame=prev,
                          first_frame_index=f0)
    if shared:   # result buffers in named shared memory: rank 0 reads the peers' records in place
        sh = sharding.shared_results(local.n, local.cand_cap, local.match_cap, local.n_tail_points, rank, tag=str(port), pin=False)
        sh.raw[:] = local.raw
        local = sh
    parts = sharding.gather_to_rank0(local)
    if rank == 0:
        whole = sharding.concat(parts)
        np.save(out_path, np.array([whole.checksum(), whole.n], np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("weights,shared", [(None, False), ((2.0, 1.0), False), (None, True), ((1.0, 2.0), True)])
def test_two_rank_gloo_frame_sharding_equals_single_process(oracle, tmp_path, weights, shared):
    import torch.multiprocessing as mp

    n = 9
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, n, out, weights, shared), nprocs=2, join=True)
    cfg, model, bkg, calib, frames, bx, bs, bb = _small_problem(n)
    ref = oracle.detect(cfg, model, bkg, calib, frames.numpy(), bx, bs, bb)
    got = np.load(out)
    assert int(got[1]) == n and int(got[0]) == ref.checksum()
    assert int(ref.match_n.sum()) > 0   # the halo frame mattered for something pairable
