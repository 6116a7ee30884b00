"""Row bands over 2 GPUs (NCCL, one process per GPU): the banded, peer-gathered cost volume equals
the whole-frame single-GPU volume bit for bit (sharding must not change arithmetic), and the
whole-frame statistics (mu, B, labels) agree with the single-GPU pass."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, queue):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    out = {}
    try:
        import traceback
        from mrflow_b200 import device, planeparallax as pp, sharding, synth
        from mrflow_b200.pipeline.build_cost_volume import cost_volume_pass
        t = synth.make_triplet((200, 256), seed=31, epipole="inside", n_waves=8)
        t["rigidity"][20:60, 30:90] = 0.0
        h, w = t["shape"]
        stats = sharding.BandStats()
        peers = (sharding.PeerVolume(51, h, w), sharding.PeerVolume(51, h, w))
        # deliberately small first halo: the halo_miss retry must kick in
        res = sharding.banded_cost_volume(t, stats, rank, world, peer=peers, halo=2)
        torch.cuda.synchronize(); dist.barrier()
        # the host-driven statistics protocol (what the gloo tests cover) gives the same bits
        res_h = sharding.banded_cost_volume(t, stats, rank, world, halo=res["halo"], resident=False)
        out["host_equals_resident"] = bool(res_h["mu"] == res["mu"] and res_h["B"] == res["B"]
                                           and np.array_equal(res_h["labels"], res["labels"])
                                           and all(torch.equal(res_h["argmins"][f], res["argmins"][f]) for f in range(2)))
        torch.cuda.synchronize(); dist.barrier()
        # small all-reduces through NVLink peer memory (one kernel per rank), a building block next to NCCL: the
        # ops themselves, slot wrap-around included; a rank that arrives 1.5 s late is waited for (the time-out is
        # 20 s of wall clock, after which the result is poisoned and check() raises)
        import time
        mb = sharding.PeerMailbox()
        ok = True
        for it in range(40):
            if it == 3 and rank == 1:
                torch.cuda.synchronize(); time.sleep(1.5)
            a = (torch.arange(256, dtype=torch.int64, device="cuda") + it) * (rank + 1)
            mb.allreduce(a, "sum_u64")
            ok &= bool(torch.equal(a, (torch.arange(256, dtype=torch.int64, device="cuda") + it) * 3))
            b = torch.tensor([5 + rank, 7 - rank, -1 if rank else 3], dtype=torch.int64, device="cuda")
            mb.allreduce(b, "min_u64")
            ok &= b.tolist() == [5, 6, 3]                   # unsigned order: -1 is the largest key
            c = torch.tensor([1.5 - rank, -2.0 * rank, float("nan") if rank == it % 2 else 0.0, 4.0], dtype=torch.float64,
                             device="cuda")
            mb.allreduce(c, "min_f64")
            cl = c.tolist()
            ok &= cl[0] == 0.5 and cl[1] == -2.0 and cl[2] != cl[2] and cl[3] == 4.0
        out["mailbox_ops"] = bool(ok)
        mb.check()
        out["mailbox_status"] = mb.status()
        mb.close()
        # the pre-pass WITHOUT collective launches: histogram reductions, successor and range gathers over NVLink
        # peer memory from inside the kernels (csrc/peersync.cuh) -- same bits as the NCCL route, step after step
        # (the histogram sets alternate with the step's parity), also when one rank arrives late
        ps = sharding.PeerSync()
        same = True
        for it in range(5):
            if it == 2 and rank == 0:
                torch.cuda.synchronize(); time.sleep(0.7)
            res_p = sharding.banded_cost_volume(t, stats, rank, world, halo=res["halo"], peer_sync=ps)
            same &= bool(res_p["mu"] == res["mu"] and res_p["B"] == res["B"] and np.array_equal(res_p["labels"], res["labels"])
                         and all(torch.equal(res_p["argmins"][f], res["argmins"][f]) for f in range(2))
                         and all(torch.equal(res_p["cost"][f], res_h["cost"][f]) for f in range(2)))
        ps.check()
        out["peer_sync_equals_nccl"] = same
        out["peer_sync_status"] = ps.status()
        ps.close()
        # an epipole estimate whose 21x21 window straddles the band boundary (row 104): the bands assemble the
        # window and every rank solves the same 441-pixel problem -- same bits as one GPU
        t2 = dict(t, epipoles=[t["epipoles"][0] + np.array([0.0, 11.0]), t["epipoles"][1] + np.array([0.0, 12.0])])
        res2 = sharding.banded_cost_volume(t2, stats, rank, world, halo=res["halo"])
        out["straddle_q"] = [q.tolist() for q in res2["q"]]
        out["halo"] = res["halo"]
        y0, y1 = res["rows"]
        if rank == 0:
            band = pp.Band.from_host(t["images"], t["flow"], t["rigidity"], None)
            one = cost_volume_pass(band, t["homographies"], t["epipoles"], t["rigidity"] > 0, pp.LocalStats(),
                                   range_mask="rigid_only")
            out["mu"] = (res["mu"], one[2]); out["B"] = (res["B"], one[3])
            out["labels_close"] = bool(np.allclose(res["labels"], one[5], rtol=1e-11, atol=0))
            # whole-frame kernel with the banded run's scalars: must match the gathered volume bitwise
            ref64, tex = pp.channels(band)
            nonrigid = torch.from_numpy((t["rigidity"] == 0).astype(np.uint8)).cuda()
            whole = pp.cost_volumes(band, ref64, tex, torch.from_numpy(res["labels"]).cuda(), t["homographies"], res["q"],
                                    res["mu"], res["B"], nonrigid)
            for f in range(2):
                out["equal%d" % f] = bool(torch.equal(peers[f].tensor(), whole[f][0]))
                out["argmin_equal%d" % f] = bool(torch.equal(res["argmins"][f], whole[f][2][y0:y1]))
            # the statistics of the banded pre-pass are the one-GPU pre-pass's, bit for bit (mu: same summation
            # tree; B: exact median; labels: exact min / max)
            from mrflow_b200.pipeline.build_cost_volume import cost_volume_pass_resident
            one_r = cost_volume_pass_resident(band, t["homographies"], t["epipoles"], torch.from_numpy(
                (t["rigidity"] > 0).astype(np.uint8)).cuda(), range_mask="rigid_only", want_cost=False)
            mu1, B1, lab1 = pp.state_to_host(one_r["buffers"].state)
            out["stats_bit_equal"] = bool(mu1 == res["mu"] and B1 == res["B"] and np.array_equal(lab1, res["labels"])
                                          and all(np.array_equal(a, b) for a, b in zip(one_r["q"], res["q"])))
            band2 = pp.Band.from_host(t2["images"], t2["flow"], t2["rigidity"], None)
            one2 = cost_volume_pass_resident(band2, t2["homographies"], t2["epipoles"], torch.from_numpy(
                (t2["rigidity"] > 0).astype(np.uint8)).cuda(), range_mask="rigid_only", want_cost=False)
            mu2, B2, lab2 = pp.state_to_host(one2["buffers"].state)
            out["straddle_equal"] = bool(all(np.array_equal(a, b) for a, b in zip(one2["q"], res2["q"])) and mu2 == res2["mu"]
                                         and B2 == res2["B"] and np.array_equal(lab2, res2["labels"])
                                         and all(torch.equal(res2["argmins"][f], one2["argmins"][f][y0:y1]) for f in range(2)))
            out["straddle_moved"] = bool(any(not np.array_equal(a, b) for a, b in zip(one2["q"], t2["epipoles"])))
        torch.cuda.synchronize()
    except Exception:                       # noqa: BLE001  -- report instead of leaving the peer rank in a barrier
        out["error"] = traceback.format_exc()
    queue.put((rank, out))
    queue.close(); queue.join_thread()      # flush the feeder thread before the hard exit
    os._exit(0)                             # no collective teardown: a failed rank must not hang the other


def _worker_large(rank, world, port, queue):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    out = {}
    try:
        import traceback
        from mrflow_b200 import sharding
        # 2.65 M pixels per band: the select kernels then run at their largest grid (4 blocks per SM), more blocks
        # than fit beside the concurrent channel kernels -- the peer-memory exchange must not depend on a whole
        # grid being resident (a wait inside the pushing grid once dead-locked exactly here)
        rec = sharding.bench_bands((1728, 3072), 3, 1, n_waves=4)
        out["rec"] = rec
    except Exception:                       # noqa: BLE001
        out["error"] = traceback.format_exc()
    queue.put((rank, out))
    queue.close(); queue.join_thread()
    os._exit(0)


def test_large_bands_peer_memory_exchange():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_large, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue as _queue
    res = {}
    try:
        for _ in procs:
            try:
                k, v = q.get(timeout=150 if not res else 30)
                res[k] = v
            except _queue.Empty:
                break
    finally:
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()
    for r in res.values():
        assert "error" not in r, r["error"]
    assert len(res) == 2, "a rank did not report: got %s" % sorted(res)
    rec = res[0]["rec"]
    assert rec["bands_bit_exact"] is True and rec["n_miss"] == 0
    # no wait ran into its time-out: a step takes milliseconds, not 20 s
    assert rec["compute_only"]["ms_per_step"] < 50 and rec["compute_only_nccl_collectives"]["ms_per_step"] < 50


def test_two_band_volume_equals_whole_frame():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue as _queue
    res = {}
    try:
        for _ in procs:
            try:
                k, v = q.get(timeout=120 if not res else 20)
                res[k] = v
            except _queue.Empty:
                break
    finally:
        for p in procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()
    for r in res.values():
        assert "error" not in r, r["error"]
    assert len(res) == 2, "a rank did not report (hung in a collective?): got %s" % sorted(res)
    o = res[0]
    assert o["halo"] > 2                                   # the retry widened the halo
    np.testing.assert_allclose(o["mu"][0], o["mu"][1], rtol=1e-14)
    np.testing.assert_allclose(o["B"][0], o["B"][1], rtol=1e-11)
    assert o["labels_close"]
    assert res[0]["host_equals_resident"] and res[1]["host_equals_resident"]
    for r in (0, 1):
        assert res[r]["mailbox_ops"] and res[r]["mailbox_status"] == 0
        assert res[r]["peer_sync_equals_nccl"] and res[r]["peer_sync_status"][1] == 0 and res[r]["peer_sync_status"][0] == 5
    assert res[0]["straddle_q"] == res[1]["straddle_q"]
    assert o["stats_bit_equal"] and o["straddle_equal"]
    for f in range(2):
        assert o["equal%d" % f] and o["argmin_equal%d" % f]
