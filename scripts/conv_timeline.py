#!/usr/bin/env python
"""Per-role timeline of CTA 0 of the resident-weight convolution kernels during one SimpleCNN training step at the
benchmark configuration (10 clients x batch 32, TF32 path): flb_debug_trace_set + one eager flb_train_step.

    FLB_TRACE=1 python -m flb200.build --force && python scripts/conv_timeline.py > gpurun_out/conv_timeline.txt
(the trace points are compiled in only with -DFLB_TRACE=1; rebuild without it afterwards)
Events: 1 role start (tile = warp), 10 weight load issued, 11 activation box issued, 20 weights landed, 21 accumulator free,
22 activation box landed, 23 k-block MMAs issued, 30 accumulator complete, 31 epilogue done, 2 role done (tile = warp)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import flb200  # noqa: E402,F401
from flb200 import _lib as L  # noqa: E402
from flb200.models_pytorch import ModelFactory  # noqa: E402
from flb200.simulation import FederatedRoundEngine  # noqa: E402

NAMES = {1: "start", 2: "done", 10: "W issued", 11: "A issued", 20: "W landed", 21: "acc free", 22: "A landed", 23: "MMAs issued",
         30: "acc complete", 31: "epilogue done", 24: "commit issued", 27: "next tile found", 28: "acc wait (2t+acc)", 29: "acc wait over", 25: "lookahead done", 26: "tile_setup done"}
dev = torch.device("cuda:0")
eng = FederatedRoundEngine("simple_cnn", 10, dev, dp_mode="update", precision="tf32", seed=1)
torch.manual_seed(0)
eng.set_global_weights(ModelFactory.create_model("simple_cnn").get_model_weights())
eng.load_synthetic()
tr = eng.trainer
tr.set_global_row(eng.global_row)
tr._fill_args(1e-3, "adam", train=True)
st = L.stream_ptr(dev)
ap = C.byref(tr.args)
L.call("flb_train_begin_epoch", ap, st)
for _ in range(3):
    L.call("flb_train_step", ap, st)
torch.cuda.synchronize()
cap = 4096
buf = torch.zeros(8 + 24 * cap, dtype=torch.uint8, device=dev)
L.call("flb_debug_trace_set", L.ptr(buf), cap)
L.call("flb_train_step", ap, st)
torch.cuda.synchronize()
L.call("flb_debug_trace_set", None, 0)
hdr = buf[:8].view(torch.int32).tolist()
allrec = buf[8:8 + 24 * cap].view(torch.int64).view(-1, 3).cpu().tolist()
per = cap // 8
rec = []
for w in range(8):
    n = allrec[w * per + per - 1][0]
    rec += [(ev, tile, clk, w) for ev, tile, clk in allrec[w * per: w * per + n]]
print(f"# {len(rec)} events from CTA 0 of the resident conv kernels of one step; clocks relative to each kernel's first event")
rec.sort(key=lambda r: r[2])
t0 = None
last_start = None
for ev, tile, clk, w in rec:
    if ev == 1 and (last_start is None or clk - last_start > 4000):
        t0 = clk
        print("---- kernel ----")
    if ev == 1:
        last_start = clk
    print(f"{clk - t0:8d} cyc  w{w} {NAMES.get(ev, ev):15s} {'' if ev in (1, 2) else 'tile ' + str(tile)}")
