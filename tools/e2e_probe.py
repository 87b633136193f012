"""e2e pipeline probe: ms/step of SetAbstractionPipeline for depth x want_grouped (python tools/e2e_probe.py)."""
import os, sys, json
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
for want in (True, False):
    for depth in (1, 2, 3, 4):
        pipe = SetAbstractionPipeline(b, n, m, r, s, depth=depth, device=dev, want_grouped=want)
        for sl in pipe.slots:
            sl.h_xyz.numpy()[...] = xyz
        def run(steps):
            for _ in range(steps):
                if pipe.full():
                    pipe.collect()
                if use_flush:
                    with torch.cuda.stream(pipe.streams[pipe._next]):
                        flush.zero_()
                pipe.submit()
            while pipe.pending():
                pipe.collect()
        run(6)
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for ps in pipe.streams:
            ps.wait_event(a)
        run(40)
        for ev in pipe.done:
            st.wait_event(ev)
        z.record(st)
        torch.cuda.synchronize()
        print(json.dumps(dict(exclusive=os.environ.get("PN2_SA_EXCLUSIVE", "1"), flush=use_flush, want_grouped=want, depth=depth, ms_per_step=a.elapsed_time(z) / 40)), flush=True)
