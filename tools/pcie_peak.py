"""PCIe ceiling of the box for the e2e figure: pinned host <-> device copies of the e2e message size, each direction alone and
both at once (two streams), CUDA events."""
import torch

n = 448 * 1024 * 1024
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_event(a)
    s2.wait_event(a)
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    e1, e2 = torch.cuda.Event(), torch.cuda.Event()
    e1.record(s1)
    e2.record(s2)
    torch.cuda.current_stream().wait_event(e1)
    torch.cuda.current_stream().wait_event(e2)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


run(True, True, 2)
for name, u, d in (("H2D alone", True, False), ("D2H alone", False, True), ("both directions at once", True, True)):
    ms = run(u, d)
    print(f"{name}: {ms:.2f} ms per 448 MiB message = {n / ms / 1e6:.1f} GB/s per direction")
