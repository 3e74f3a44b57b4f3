"""Ad-hoc GPU bring-up check for the RCPS kernels (golden parity + a first timing). Not part of the test suite."""
import glob, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from im2im_uq_b200 import rcps
from oracle import rcps_oracle as orc

dev = torch.device("cuda:0")
ok = True
for f in sorted(glob.glob("tests/golden/rcps_*.npz")):
    g = np.load(f)
    out = torch.from_numpy(g["outputs"]).to(dev); lab = torch.from_numpy(g["labels"]).to(dev)
    for key_l, key_c in (("lam_prime", "counts_prime"), ("lambdas", "counts_grid")):
        lam = torch.from_numpy(g[key_l]).to(dev)
        for generic in (False, True):
            counts, totals = rcps.miss_counts(out, lab, lam, force_generic=generic)
            torch.cuda.synchronize()
            good = np.array_equal(counts.cpu().numpy(), g[key_c]) and np.array_equal(totals.cpu().numpy(), g[key_c].sum(0, dtype=np.int64))
            ok &= good
            print(f"{os.path.basename(f):28s} {key_l:10s} generic={generic!s:5s} {'OK' if good else 'MISMATCH'}")
            if not good:
                d = counts.cpu().numpy() != g[key_c]
                print("   mismatches:", d.sum(), "first:", np.argwhere(d)[:5].tolist())
    lo, p, up = rcps.quantile_nested_sets(out.clone(), float(g["lhat"]))
    good = np.array_equal(lo.cpu().numpy(), g["lower_at_lhat"], equal_nan=True) and np.array_equal(up.cpu().numpy(), g["upper_at_lhat"], equal_nan=True)
    ok &= good
    print(f"{os.path.basename(f):28s} nested_sets {'OK' if good else 'MISMATCH'}")
    mm = rcps.miss_map(out, lab, float(g["lam_prime"][len(g["lam_prime"]) // 2]))
    good = np.array_equal(mm.cpu().numpy(), orc.c_miss_map(g["outputs"], g["labels"], float(g["lam_prime"][len(g["lam_prime"]) // 2])))
    ok &= good
    print(f"{os.path.basename(f):28s} miss_map {'OK' if good else 'MISMATCH'}")

# bigger seeded case vs oracle: 24 x 320x320, fastmri grid
torch.manual_seed(0)
def synth(n, h, w):
    pred = torch.rand(n, 1, h, w, device=dev); sig = 0.02 + 0.1 * torch.rand(n, 1, h, w, device=dev)
    lower = pred - sig * (0.5 + torch.rand_like(pred)); upper = pred + sig * (0.5 + torch.rand_like(pred))
    label = pred + sig * torch.randn_like(pred)
    return torch.stack([lower, pred, upper], dim=1).contiguous(), label
lambdas = torch.linspace(0, 6, 1000); lam_prime = (lambdas - (lambdas[1] - lambdas[0])).to(dev)
out, lab = synth(24, 320, 320)
counts, totals = rcps.miss_counts(out, lab, lam_prime)
t0 = time.time(); ref = orc.c_miss_table(out.cpu().numpy(), lab.cpu().numpy(), lam_prime.cpu().numpy()); t1 = time.time()
good = np.array_equal(counts.cpu().numpy(), ref)
ok &= good
print(f"24x320x320 L=1000 vs C oracle ({t1 - t0:.1f}s, {orc.num_threads()} threads): {'OK' if good else 'MISMATCH'}")
counts_g, _ = rcps.miss_counts(out, lab, lam_prime, force_generic=True)
print("generic == staged:", torch.equal(counts, counts_g))

# timing
for n in (1000, 4000):
    out, lab = synth(n, 320, 320)
    c = torch.empty((n, 1000), dtype=torch.int32, device=dev); t = torch.empty(1000, dtype=torch.int64, device=dev)
    for generic in (False, True):
        for _ in range(3): rcps.miss_counts(out, lab, lam_prime, counts=c, totals=t, force_generic=generic)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): rcps.miss_counts(out, lab, lam_prime, counts=c, totals=t, force_generic=generic)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        gb = (n * 102400 * 16 + n * 1000 * 4) / 1e9
        print(f"N={n} generic={generic}: {ms:.3f} ms/pass  {gb / ms * 1e3:.0f} GB/s  {n / ms * 1e3:.0f} img/s")
    del out, lab
print("ALL OK" if ok else "FAILURES")
