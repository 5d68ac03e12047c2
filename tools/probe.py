"""Ad-hoc phase timing on one GPU (not the bench contract): python tools/probe.py [n] [ppc] [order]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smilei_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ppc = (4, 2, 2)
order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
T = 10. / 511.
dx = 0.5 * T ** 0.5
dt = 0.95 * dx / 3 ** 0.5
p = smilei_b200.Patch((n,) * 3, (dx,) * 3, dt, interp_order=order, n_species=2)
N = n ** 3 * 16
t0 = time.time()
for s, (m, q) in enumerate(((1836., 1), (1., -1))):
    p.species_config(s, m, "boris", int(N * 1.1))
    p.species_init_thermal(s, ppc, 1.0, q, T, seed=0)   # same seed -> same positions, rho = 0
    p.sort(s)
p.synchronize()
print("init+sort %.2fs, particles/species %d" % (time.time() - t0, N), flush=True)
L = n * dx
buf = torch.zeros(8 * max(N // 20, 1 << 16), dtype=torch.float64, device="cuda")

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

names = ["zeroJ", "dyn0", "dyn1", "exch", "sumJ", "maxwell", "exchB", "center", "sort0", "sort1"]
acc = {k: 0. for k in names}
for it in range(steps):
    e = [ev()]
    p.restart_rhoJ(); e.append(ev())
    p.dynamics(0); e.append(ev())
    p.dynamics(1); e.append(ev())
    if it == steps - 1: print('flags after dynamics', p.debug_flags(), flush=True)
    for s in range(2):
        for dim in range(3):
            for side in (0, 1):
                k = p.leaving_pack(s, dim, side, L if side == 0 else -L, buf.data_ptr(), buf.numel() // 8)
                p.arriving_unpack(s, buf.data_ptr(), k)
    e.append(ev())
    for dim in range(3):
        for f in ("Jx", "Jy", "Jz"):
            p.halo_sum_self(f, dim)
    e.append(ev())
    p.maxwell(); e.append(ev())
    for dim, comps in ((0, ("By", "Bz")), (1, ("Bx", "Bz")), (2, ("Bx", "By"))):
        for f in comps:
            p.halo_exchange_self(f, dim)
    e.append(ev())
    p.center_B(); e.append(ev())
    p.sort(0); e.append(ev())
    p.sort(1); e.append(ev())
    torch.cuda.synchronize()
    if it > 0:
        for i, k in enumerate(names):
            acc[k] += e[i].elapsed_time(e[i + 1])
    if it == 0 or it == steps - 1:
        uk, ue = p.energy()
        print("step", it, "Ukin", uk, "Uelm", ue, "n", p.species_count(0), p.species_count(1), "flags", p.debug_flags(), flush=True)
m = steps - 1
tot = sum(acc.values()) / m
print({k: round(v / m, 3) for k, v in acc.items()}, "ms; total %.2f ms/step" % tot)
dyn = (acc["dyn0"] + acc["dyn1"]) / m
print("pushes/s (dynamics only) %.3f G ; whole step %.3f G ; yee cell-updates/s %.3f G" % (2 * N / dyn / 1e6, 2 * N / tot / 1e6, n ** 3 / (acc["maxwell"] / m) / 1e6))
