"""PCIe floor of the e2e step: D2H of 280 MB and H2D of 72 MB, alone and together (pinned memory, two streams)."""
import time, torch
d2h_n, h2d_n = 280 << 20, 72 << 20
dev_out = torch.empty(d2h_n, dtype=torch.uint8, device="cuda"); host_out = torch.empty(d2h_n, dtype=torch.uint8).pin_memory()
dev_in = torch.empty(h2d_n, dtype=torch.uint8, device="cuda"); host_in = torch.empty(h2d_n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(do_d2h, do_h2d, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        if do_d2h:
            with torch.cuda.stream(s1): host_out.copy_(dev_out, non_blocking=True)
        if do_h2d:
            with torch.cuda.stream(s2): dev_in.copy_(host_in, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for _ in range(2): run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print("D2H 280 MiB %.2f ms (%.1f GB/s) ; H2D 72 MiB %.2f ms (%.1f GB/s) ; both %.2f ms" % (a, d2h_n / a / 1e6, b, h2d_n / b / 1e6, c))
