"""Host-side logic of the row-band sharding (mrflow_b200/sharding.py) with world_size 2 on CPU (gloo):
band plan, halo planning, the distributed exact-median protocol and the statistics all-reduces.  The
per-shard device steps are replaced by numpy twins here; the GPU tests cover the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _np_steps(keys):
    from mrflow_b200.sharding import order_key
    k = order_key(keys).astype(np.uint64)

    def hist_fn(prefix, p):
        shift = 56 - 8 * p
        sel = k if p == 0 else k[(k >> np.uint64(shift + 8)) == np.uint64(prefix >> (shift + 8))]
        return np.bincount(((sel >> np.uint64(shift)) & np.uint64(0xFF)).astype(np.int64), minlength=256)

    def succ_fn(kth):
        le = k <= np.uint64(kth)
        gt = k[~le]
        return int(le.sum()), int(np.isnan(keys).sum()), int(gt.min()) if gt.size else 0xFFFFFFFFFFFFFFFF
    return hist_fn, succ_fn


def _np_steps13(keys):
    """numpy twins of k_s13_hist / k_s13_successor (csrc/select13.cuh)"""
    from mrflow_b200.sharding import DistSelect13, order_key
    k = order_key(keys).astype(np.uint64)

    def hist_fn(prefix, p):
        shift = DistSelect13.SHIFTS[p]
        sel = k if p == 0 else k[(k >> np.uint64(shift + 13)) == np.uint64(prefix >> (shift + 13))]
        return np.bincount(((sel >> np.uint64(shift)) & np.uint64(8191)).astype(np.int64), minlength=8192)

    def succ_fn(kth):
        le = k <= np.uint64(kth)
        gt = k[~le]
        gt = gt[gt != np.uint64(0xFFFFFFFFFFFFFFFF)]
        return int(np.isnan(keys).sum()), int(le.sum()), (~int(gt.min()) & 0xFFFFFFFFFFFFFFFF) if gt.size else 0
    return hist_fn, succ_fn


def _worker(rank, world, port, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mrflow_b200 import sharding
    out = {}
    try:
        stats = sharding.BandStats()
        rs = np.random.RandomState(5)
        for case, n in (("odd", 20001), ("even", 4000), ("dup", 999), ("tiny", 3)):
            full = rs.standard_normal(n) * 10.0 ** rs.randint(-3, 4, n)
            if case == "dup":
                full[rs.randint(0, n, n // 2)] = 1.25
                full[:4] = [np.inf, -np.inf, 0.0, -0.0]
            cut = n // 3                                   # uneven shards
            mine = full[:cut] if rank == 0 else full[cut:]
            h, s = _np_steps(mine)
            sel = sharding.DistSelect(h, s, stats._allreduce)
            out[case] = (sel.median(stats.total(mine.size)), float(np.median(full)))
            ks = sorted({0, n // 2, n - 1})
            out[case + "_kth"] = [(sel.kth(k)[0], float(np.sort(full)[k])) for k in ks]
            h13, s13 = _np_steps13(mine)
            out[case + "_13"] = sharding.DistSelect13(h13, s13, stats._allreduce, stats._allgather).median()
        full = rs.standard_normal(101); full[7] = np.nan
        mine = full[:50] if rank == 0 else full[50:]
        h, s = _np_steps(mine)
        out["nan"] = sharding.DistSelect(h, s, stats._allreduce).median(101)
        h13, s13 = _np_steps13(mine)
        out["nan13"] = sharding.DistSelect13(h13, s13, stats._allreduce, stats._allgather).median()
        h13, s13 = _np_steps13(np.zeros(0))
        out["empty13"] = sharding.DistSelect13(h13, s13, stats._allreduce, stats._allgather).median()
        out["total"] = stats.total(rank + 5)
        mm = stats.minmax(torch.tensor([-1.0 - rank, 3.0 + rank], dtype=torch.float64))
        out["minmax"] = mm
        out["sum"] = float(stats.device_sum(torch.tensor([1.5 + rank], dtype=torch.float64))[0])

        # the 21x21 epipole window of a 100-row frame cut into bands 0..48 / 48..100: a window inside rank 1's
        # band and one that straddles the boundary both add up to the whole-frame window on both ranks, and
        # both ranks then find the same epipole (compute_structure.py:453-511)
        from mrflow_b200 import planeparallax as pp
        rs2 = np.random.RandomState(11)
        y0, y1 = (0, 48) if rank == 0 else (48, 100)
        yy, xx = np.mgrid[:100, :200].astype(np.float64)
        for name, qe in (("epi_inside", np.array([90.3, 70.6])), ("epi_straddle", np.array([90.0, 50.0]))):
            qt = qe + np.array([1.5, -0.8])                # the objective's minimum: flow perpendicular to (qt - p)
            u = torch.from_numpy((-(yy - qt[1]) * 0.01 + 0.002 * rs2.standard_normal(xx.shape)).astype(np.float32))
            v = torch.from_numpy(((xx - qt[0]) * 0.01 + 0.002 * rs2.standard_normal(xx.shape)).astype(np.float32))
            rigid = torch.ones((100, 200), dtype=torch.uint8); rigid[65:70, 85:90] = 0; rigid[45:50, 95:99] = 0
            whole = pp.correct_epipole(u, v, qe, rigid)
            mine_q = pp.correct_epipole(u[y0:y1], v[y0:y1], qe, rigid[y0:y1], y0=y0, frame_h=100, stats=stats)
            out[name] = (mine_q.tolist(), whole.tolist(), float(np.linalg.norm(whole - qe)))
        out["epi_outside"] = pp.correct_epipole(u[y0:y1], v[y0:y1], np.array([500.0, 70.0]), rigid[y0:y1], y0=y0,
                                                frame_h=100, stats=stats).tolist()
    finally:
        q_out = out
        dist.destroy_process_group()
    queue.put((rank, q_out))


def test_distributed_statistics_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        o = res[r]
        for case in ("odd", "even", "dup", "tiny"):
            assert o[case][0] == o[case][1], (case, o[case])
            assert o[case + "_13"] == o[case][1], (case, o[case + "_13"], o[case][1])
            for got, want in o[case + "_kth"]:
                assert got == want
        assert np.isnan(o["nan"]) and np.isnan(o["nan13"]) and np.isnan(o["empty13"])
        assert o["total"] == 11 and o["minmax"] == (-2.0, 4.0) and o["sum"] == 4.0
        for name in ("epi_inside", "epi_straddle"):
            got, whole, moved = o[name]
            assert got == whole, (name, got, whole)         # bands == whole frame, bit for bit, on both ranks
            assert 0 < moved < 10                           # and the BFGS did run and move the epipole
        assert o["epi_outside"] == [500.0, 70.0]


def test_band_plan_and_halo():
    from mrflow_b200 import sharding, synth
    for h, world in ((436, 2), (2160, 8), (375, 4), (50, 8)):
        plan = sharding.band_plan(h, world)
        assert len(plan) == world and plan[0][0] == 0 and plan[-1][1] == h
        assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
        assert all(y0 % 8 == 0 for y0, y1 in plan if y1 > y0)
    t = synth.make_triplet((120, 160), seed=3, n_waves=4)
    labels = np.linspace(0.5, 9.0, 51)
    up, down = sharding.required_halo((40, 80), 160, 120, t["homographies"][1], t["epipoles"][1], t["mu"][1], 1.0, labels)
    assert 3 <= up <= 60 and 3 <= down <= 60
    # zero parallax, identity homography: only the taps and the margin
    up0, down0 = sharding.required_halo((40, 80), 160, 120, np.eye(3), t["epipoles"][1], t["mu"][1], 1.0, np.zeros(3))
    assert up0 == down0 == 7


def test_order_key_roundtrip():
    from mrflow_b200.sharding import key_value, order_key
    v = np.array([-np.inf, -3.5, -0.0, 0.0, 1e-300, 2.0, np.inf])
    k = order_key(v)
    assert np.all(np.diff(k.astype(np.float64)) >= 0) and k[2] < k[3]
    assert [key_value(x) for x in k] == v.tolist()
    assert np.isnan(key_value(order_key(np.array([np.nan]))[0]))
