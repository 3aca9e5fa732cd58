"""DMA time to / from different places of pinned host allocations (the ends of a large pinned allocation were seen to
be ~45 us slower per 256 KB row than its middle in bench.py's e2e leg)."""
import time
import torch

dev = torch.device("cuda", 0)
N = 65536
d_small = torch.empty((N,), dtype=torch.int32, device=dev)
d_big = torch.empty((5 * 1024 * 1024 // 4 + N,), dtype=torch.int32, device=dev)


def timed(fn, reps=200):
    for i in range(10):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
        torch.cuda.current_stream().synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


small = torch.empty((N,), dtype=torch.int32).pin_memory()
big = torch.empty((64, N), dtype=torch.int32).pin_memory()           # 16 MB
print("pinned base addresses mod 2 MB: small %d KB, big %d KB" % (small.data_ptr() % (2 << 20) // 1024, big.data_ptr() % (2 << 20) // 1024))
print("H2D 256 KB, one small pinned tensor reused      : %6.1f us" % timed(lambda i: d_small.copy_(small, non_blocking=True)))
print("H2D 256 KB, rows 0-3 of a 16 MB pinned tensor    : %6.1f us" % timed(lambda i: d_small.copy_(big[i % 4], non_blocking=True)))
print("H2D 256 KB, rows 28-35 (middle)                  : %6.1f us" % timed(lambda i: d_small.copy_(big[28 + i % 8], non_blocking=True)))
print("H2D 256 KB, rows 60-63 (end)                     : %6.1f us" % timed(lambda i: d_small.copy_(big[60 + i % 4], non_blocking=True)))
print("H2D 256 KB, all 64 rows in turn                  : %6.1f us" % timed(lambda i: d_small.copy_(big[i % 64], non_blocking=True)))
for r in range(0, 64, 4):
    print("   row %2d alone (reused): %6.1f us" % (r, timed(lambda i: d_small.copy_(big[r], non_blocking=True), reps=60)), end="")
    if r % 16 == 12:
        print()
stage = torch.empty((5 * 1024 * 1024 // 4,), dtype=torch.int32).pin_memory()                # 5 MB on its own
arena = torch.empty((9 * 1024 * 1024 // 4,), dtype=torch.int32).pin_memory()                # 9 MB: 2 MB pad | 5 MB | 2 MB pad
mid = arena[(2 << 20) // 4:(7 << 20) // 4]
src = d_big[:5 * 1024 * 1024 // 4]
print("D2H 5 MB into its own pinned tensor              : %6.1f us" % timed(lambda i: stage.copy_(src, non_blocking=True)))
print("D2H 5 MB into the middle of a 9 MB pinned tensor : %6.1f us" % timed(lambda i: mid.copy_(src, non_blocking=True)))
blob = torch.empty((N * 5,), dtype=torch.uint8).pin_memory()
dblob = torch.empty((N * 5,), dtype=torch.uint8, device=dev)
ab = torch.empty((4 << 20) + N * 5, dtype=torch.uint8).pin_memory()
midb = ab[2 << 20:(2 << 20) + N * 5]
print("D2H 320 KB into its own pinned tensor            : %6.1f us" % timed(lambda i: blob.copy_(dblob, non_blocking=True)))
print("D2H 320 KB into the middle of a 4.3 MB tensor    : %6.1f us" % timed(lambda i: midb.copy_(dblob, non_blocking=True)))
