#!/usr/bin/env python
"""PCIe ceiling of the box for the end-to-end path: every rank moves what one triplet moves through the host API
(193 MB device -> pinned host, 43 MB pinned host -> device), all ranks at once, copies only -- no kernels.

    python tools/micro/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/micro/pcie_probe.py

Prints one JSON line (rank 0): per-rank and aggregate GB/s for D2H alone, H2D alone and both directions at once
(time = max over ranks, barrier on both sides), and the triplets/s ceiling that follows for the e2e metric."""
import json
import os
import time

import torch
import torch.distributed as dist

D2H, H2D = 192874624, 42860544          # bytes per triplet (bench.py: e2e.d2h_bytes_per_triplet / h2d_bytes_per_triplet)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev_out = torch.empty(D2H, dtype=torch.uint8, device="cuda")
    dev_in = torch.empty(H2D, dtype=torch.uint8, device="cuda")
    host_out = torch.empty(D2H, dtype=torch.uint8, pin_memory=True)
    host_in = torch.empty(H2D, dtype=torch.uint8, pin_memory=True)
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(mode, reps=20):
        for _ in range(3):
            if mode != "h2d":
                host_out.copy_(dev_out, non_blocking=True)
            if mode != "d2h":
                dev_in.copy_(host_in, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if mode != "h2d":
                with torch.cuda.stream(s_down):
                    host_out.copy_(dev_out, non_blocking=True)
            if mode != "d2h":
                with torch.cuda.stream(s_up):
                    dev_in.copy_(host_in, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.cpu()[0]) / reps

    res = {"n_gpus": world, "d2h_bytes": D2H, "h2d_bytes": H2D}
    for mode in ("d2h", "h2d", "both"):
        s = run(mode)
        nbytes = (D2H if mode != "h2d" else 0) + (H2D if mode != "d2h" else 0)
        res[mode] = {"ms_per_triplet": s * 1e3, "gbs_per_rank": nbytes / s / 1e9, "gbs_aggregate": world * nbytes / s / 1e9}
    res["e2e_ceiling_pixel_labels_per_s"] = world * 2 * 436 * 1024 * 51 / (res["both"]["ms_per_triplet"] * 1e-3)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
