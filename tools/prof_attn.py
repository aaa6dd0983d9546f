"""Timing driver: NARRE attention (K3) at configs[2] size, both sides — tensor-core pair kernels vs the per-side kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rbr_b200
from rbr_b200 import ops

B, R, H, A = int(os.environ.get("B", 4096)), 10, 150, 32
gen = torch.Generator().manual_seed(0)
def side(n_ids):
    feat = (torch.randn(B, R, H, generator=gen).abs() * 0.5).cuda().requires_grad_(True)
    oid = torch.randint(0, n_ids, (B, R), generator=gen).cuda()
    prm = [((torch.rand(H, A, generator=gen) * 2 - 1) * 0.1), ((torch.rand(A, A, generator=gen) * 2 - 1) * 0.1),
           ((torch.rand(A, 1, generator=gen) * 2 - 1) * 0.1), torch.full((A,), 0.1), torch.full((1,), 0.1), torch.randn(n_ids, A, generator=gen)]
    return feat, oid, [p.cuda().requires_grad_(True) for p in prm]
su, si = side(12000), side(20000)
params = su[2] + si[2]
go = [torch.randn(B, H, device="cuda") for _ in range(2)]

def run_pair():
    ou, scu, oi, sci = ops.NarreAttnPairFn.apply(su[0], su[1], si[0], si[1], *params, (0, 0), None, params)
    return ou, oi
def run_old():
    ou, _ = ops.NarreAttnFn.apply(su[0], su[1], *su[2], 0, None, su[2])
    oi, _ = ops.NarreAttnFn.apply(si[0], si[1], *si[2], 0, None, si[2])
    return ou, oi

def timeit(fn, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for name, f in (("pair-tc", run_pair), ("per-side", run_old)):
    with torch.no_grad():
        t_f = timeit(f)
    def fb():
        ou, oi = f()
        torch.autograd.backward([ou, oi], go)
    t_fb = timeit(fb)
    print(f"{name}: fwd {t_f:.1f} us, fwd+bwd {t_fb:.1f} us (bwd ~{t_fb - t_f:.1f} us)")
