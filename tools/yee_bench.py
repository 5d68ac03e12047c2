"""Time sb200_maxwell (+ center_B) alone on a 256^3 patch: python tools/yee_bench.py [n] [order]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, smilei_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dx = 0.07; dt = 0.95 * dx / 3 ** 0.5
p = smilei_b200.Patch((n,) * 3, (dx,) * 3, dt, interp_order=order, n_species=0)
rng = np.random.default_rng(0)
for k in ("Ex", "Ey", "Ez", "Bx", "By", "Bz", "Jx", "Jy", "Jz"):
    p.field_set(k, 1e-3 * rng.standard_normal(p.field_dims(k)))
for _ in range(5):
    p.maxwell(); p.center_B()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
K = 50
tm = tc = 0.
for _ in range(K):
    e0.record(); p.maxwell(); e1.record(); p.center_B(); e2.record()
    torch.cuda.synchronize()
    tm += e0.elapsed_time(e1); tc += e1.elapsed_time(e2)
tm /= K; tc /= K
peak = 6456.8
print("maxwell %.3f ms  center_shell %.3f ms  -> %.2f G cell-updates/s (maxwell only), %.1f%% of %.0f GB/s at 192 B/cell ; incl. shell %.2f G"
      % (tm, tc, n ** 3 / tm / 1e6, 100 * 192 * n ** 3 / (tm * 1e-3) / 1e9 / peak, peak, n ** 3 / (tm + tc) / 1e6))
