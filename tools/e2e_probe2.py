"""Repeat the idx-only pipeline leg to catch intermittent slow runs; per-batch GPU durations (python tools/e2e_probe2.py)."""
import os, sys, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet2_b200 import workloads as W
from pointnet2_b200.host import SetAbstractionPipeline
dev = torch.device("cuda:0")
b, n, m, r, s = 32, 4096, 1024, 0.1, 32
xyz = W.cloud_uniform(b, n, 100)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(dev)
use_flush = os.environ.get("PROBE_FLUSH", "1") == "1"
depth = int(os.environ.get("PROBE_DEPTH", "3"))
for rep in range(int(os.environ.get("PROBE_REPS", "8"))):
    pipe = SetAbstractionPipeline(b, n, m, r, s, depth=depth, device=dev, want_grouped=False)
    for sl in pipe.slots:
        sl.h_xyz.numpy()[...] = xyz
    marks = []
    def run(steps, rec=False):
        for _ in range(steps):
            if pipe.full():
                pipe.collect()
            i = pipe._next
            with torch.cuda.stream(pipe.streams[i]):
                if use_flush:
                    flush.zero_()
                if rec:
                    e0 = torch.cuda.Event(enable_timing=True); e0.record(pipe.streams[i])
            pipe.submit()
            if rec:
                e1 = torch.cuda.Event(enable_timing=True); e1.record(pipe.streams[i]); marks.append((e0, e1))
        while pipe.pending():
            pipe.collect()
    run(6)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for ps in pipe.streams:
        ps.wait_event(a)
    run(50, rec=True)
    for ev in pipe.done:
        st.wait_event(ev)
    z.record(st)
    torch.cuda.synchronize()
    per = [e0.elapsed_time(e1) for e0, e1 in marks]
    print(json.dumps(dict(rep=rep, excl=os.environ.get("PN2_SA_EXCLUSIVE", "1"), flush=use_flush, depth=depth, ms_per_step=round(a.elapsed_time(z) / 50, 4),
                          batch_ms_median=round(statistics.median(per), 3), batch_ms_max=round(max(per), 3), batch_ms_min=round(min(per), 3))), flush=True)
    del pipe
