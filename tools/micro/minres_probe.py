"""Time the persistent MINRES kernel alone on a Sintel-size system (CUDA events), for the blocks-per-SM knob."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mrflow_b200 import variational

h, w = (436, 1024) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
n = h * w
rs = np.random.RandomState(0)
planes = {k: torch.from_numpy(rs.uniform(0.1, 1.0, n)).cuda() for k in ("D", "p1x", "p1y", "p2x", "p2y")}
rigid = torch.from_numpy((rs.rand(n) < 0.9).astype(np.uint8)).cuda()
b = torch.from_numpy(rs.standard_normal(n)).cuda() * rigid
op = variational.Operator(h, w, rigid, planes["D"], planes["p1x"], planes["p1y"], planes["p2x"], planes["p2y"], None, 0.5, 5.0, 0.0)
ws = variational.MinresWorkspace(h, w)
for rtol in (1e-5, 1e-9):
    x, info = variational.minres(op, b, rtol=rtol, ws=ws)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        x, info = variational.minres(op, b, rtol=rtol, ws=ws)
    e.record(); e.synchronize()
    ms = a.elapsed_time(e) / 5
    st = ws.state.get()
    print("   block 0 per iteration: phase A %.2f us, phases B+C %.2f us, barrier waits %.2f us" % tuple(st[k] / 1e3 / info["itn"] for k in (29, 30, 31)))
    print("blocks/SM=%s rtol=%g itn=%d  %.3f ms/solve  %.2f us/iteration" % (os.environ.get("MRF_MINRES_BLOCKS_PER_SM", "default"), rtol,
                                                                         info["itn"], ms, ms * 1e3 / info["itn"]))
