"""PCIe copy rates of this box: H2D alone, D2H alone, both directions at once (pinned host memory, copy engines).
Context for the `e2e` number: pbn_step_host moves 2 B up + 4 B down per env-step.  Development aid."""
import torch

dev = torch.device("cuda:0")


def rate(nbytes_up, nbytes_dn, reps=20):
    hu = torch.empty(max(nbytes_up, 1), dtype=torch.uint8).pin_memory()
    hd = torch.empty(max(nbytes_dn, 1), dtype=torch.uint8).pin_memory()
    du = torch.empty(max(nbytes_up, 1), dtype=torch.uint8, device=dev)
    dd = torch.empty(max(nbytes_dn, 1), dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    best = 1e9
    for trial in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        for _ in range(reps):
            if nbytes_up:
                with torch.cuda.stream(s1):
                    du.copy_(hu, non_blocking=True)
            if nbytes_dn:
                with torch.cuda.stream(s2):
                    hd.copy_(dd, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


for up, dn, label in ((2 << 20, 0, "H2D 2 MB"), (0, 4 << 20, "D2H 4 MB"), (2 << 20, 4 << 20, "H2D 2 MB + D2H 4 MB at once"),
                      (64 << 20, 0, "H2D 64 MB"), (0, 64 << 20, "D2H 64 MB"), (64 << 20, 64 << 20, "H2D 64 MB + D2H 64 MB at once"),
                      (32 << 20, 64 << 20, "H2D 32 MB + D2H 64 MB at once")):
    ms = rate(up, dn)
    print("%-32s %8.1f us per round  %6.1f GB/s total (up %.1f, down %.1f)" % (label, ms * 1e3, (up + dn) / ms / 1e6, up / ms / 1e6, dn / ms / 1e6), flush=True)
