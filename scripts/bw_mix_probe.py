#!/usr/bin/env python3
"""HBM bandwidth of simple streams with different read : write mixes (torch elementwise kernels, 8 GB per array, CUDA events).
Context for the roofline fractions of the Cartesian Local kernels, whose stage-0 traffic is 42 % reads / 58 % writes: the `peak` of
MEASURED_PEAKS.json is a COPY (1 : 1). One JSON line per mix."""
import json
import torch

n = 1 << 30  # doubles: 8 GB
dev = torch.device("cuda:0")
a = torch.ones(n, dtype=torch.float64, device=dev)
b = torch.ones(n, dtype=torch.float64, device=dev)
c = torch.empty(n, dtype=torch.float64, device=dev)
d = torch.empty(n, dtype=torch.float64, device=dev)


def timed(fn, bytes_moved, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return bytes_moved*reps/(e0.elapsed_time(e1)*1e-3)/1e9


cases = {
    "read only (sum)": (lambda: a.sum(), 8*n),
    "2 reads : 1 write (c = a + b)": (lambda: torch.add(a, b, out=c), 24*n),
    "1 read : 1 write (copy)": (lambda: c.copy_(a), 16*n),
    "1 read : 2 writes (c = a; d = a, two kernels)": (lambda: (c.copy_(a), d.copy_(a)), 32*n),
    "write only (fill)": (lambda: c.fill_(1.5), 8*n),
}
out = {}
for name, (fn, nbytes) in cases.items():
    out[name] = round(timed(fn, nbytes), 1)
print(json.dumps({"probe": "hbm read/write mix", "unit": "GB/s", "array_gb": 8*n/1e9, "results": out}))
