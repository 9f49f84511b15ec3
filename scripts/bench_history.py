"""Times crb_build_history on synthetic interactions already on the device (profiles/ notes)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cleverrec_b200.engine import Engine
eng = Engine(0)
for n, U, I in ((10_000_000, 1_000_000, 200_000), (100_000_000, 10_000_000, 2_000_000), (400_000_000, 10_000_000, 2_000_000)):
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.randint(0, U, (n,), device="cuda", generator=g, dtype=torch.int32)
    i = torch.randint(0, I, (n,), device="cuda", generator=g, dtype=torch.int32)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.build_history(u, i, U, I)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("crb_build_history n=%d users=%d items=%d: %.1f ms (%.2e interactions/s, %.0f GB/s of the 8n input bytes)" % (n, U, I, dt * 1e3, n / dt, 8 * n / dt / 1e9))
    del u, i
    eng._hist = None; eng._lists = None
    torch.cuda.empty_cache()
