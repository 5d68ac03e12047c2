#!/usr/bin/env python
"""bench.py — the hot-path benchmark (contract in the task statement, §④).

    python bench.py --gpus N --steps K --warmup W            # this build, N GPUs of one node
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU operators

A "step" is one full PIC time step of the hot path on synthetic uniform thermal plasma:
restartRhoJ, gather+push+BC+deposit for both species, particle migration, J halo sum, Yee
(Ampere, Faraday, B centring), B halo, cell sort.  Workload at N=1 = BASELINE.json configs[1]:
256^3 cells, 16 ppc, 2 species, order 2, Boris, periodic; at N>1 the same box PER GPU on a
2x1x1 / 2x2x1 / 2x2x2 grid (configs[4], weak scaling).  Inputs are resident in HBM before the
timed region for `value`; `e2e` repeats the metric through the C ABI with host buffers.

metric = particle pushes/s over the WHOLE step (all species, all ranks); the Yee cell-updates/s
(the second half of BASELINE.json's metric) and the kernel-only pushes/s are reported beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Libraries print banners on fd 1 (NCCL's "NCCL version ..." line comes
# from C code), so fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved fd.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

METRIC = "particle pushes/s (gather+push+deposit, whole PIC step incl. Yee + sort + exchange)"
UNIT = "pushes/s"
T_KEV = 10.
PPC = (4, 2, 2)


def plasma_constants():
    T = T_KEV / 511.
    dx = 0.5 * T ** 0.5
    dt = 0.95 * dx / 3 ** 0.5
    return T, dx, dt


def rank_grid_for(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


def namelist_source(ncell, order, pusher):
    T, dx, dt = plasma_constants()
    return f"""
Main(geometry="3Dcartesian", interpolation_order={order}, timestep={dt!r}, number_of_timesteps=1,
     cell_length=[{dx!r}]*3, number_of_cells={list(ncell)!r}, number_of_patches=[1,1,1],
     EM_boundary_conditions=[["periodic"]], gpu_computing=True)
Species(name="proton", position_initialization="regular", regular_number=[4,2,2], momentum_initialization="mj",
        particles_per_cell=16, mass=1836.0, charge=1.0, charge_density=1., temperature=[{T!r}], pusher="{pusher}",
        boundary_conditions=[["periodic","periodic"]]*3)
Species(name="electron", position_initialization="proton", momentum_initialization="mj",
        particles_per_cell=16, mass=1.0, charge=-1.0, charge_density=1., temperature=[{T!r}], pusher="{pusher}",
        boundary_conditions=[["periodic","periodic"]]*3)
"""


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's OWN operator classes on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(order, pusher_id, steps, warmup, target_seconds=12., min_particles=32 * 1024 * 1024):
    """The reference's own CPU implementation of the path on the host cores, all threads, one OpenMP thread per patch
    at a time (VectorPatch.cpp:4777), on DISTINCT patches of 16^3 cells x 16 ppc x 2 species whose total
    (>= min_particles) leaves the caches:

      * order 2: the VECTORISED path a production CPU run uses (SpeciesV::dynamics with vectorization_mode "on":
        Interpolator3D2OrderV, the pusher, computeParticleCellKeys, Projector3D2OrderV, then SpeciesV::sortParticles
        with the leavers coming back through the receive buffer) + solveMaxwell  -> kind "reference-V";
        the scalar operator classes (no sort) are timed beside it on the same sample for comparison;
      * order 4 (no V classes in the reference for 3D order 4): the scalar operators, kind "reference"."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_lib as ol
    if not ol.have_ref():
        return None
    ref = ol.Reference(fast=True)
    orc = ol.Oracle()
    T, dx, dt = plasma_constants()
    n = (16, 16, 16)
    g = ol.make_grid(n, order, (dx, dx, dx), dt)
    rng = np.random.default_rng(0)
    F = ol.random_fields(g, rng, scale=1e-3)
    cores = os.cpu_count() or 1
    nper = 16 ** 3 * 16
    npatch = max(cores, -(-min_particles // (2 * nper)))
    npatch = -(-npatch // cores) * cores                 # whole rounds of the thread team
    parts = []
    for mass, q in ((1836., 1), (1., -1)):
        P = ol.random_particles(g, rng, nper, p_scale=(T / mass) ** 0.5, charge=q)
        P["w"][:] = dx ** 3 / 16.
        keys = orc.cell_keys(g, P["x"], P["y"], P["z"])
        first, perm = orc.counting_sort_perm(keys, 17 ** 3)
        parts.append((mass, {k: np.ascontiguousarray(v[perm]) for k, v in P.items()}, first))
    vector = order == 2

    def one(nsteps, use_vector):
        t = 0.
        for mass, P, first in parts:
            if use_vector:
                tt, _ = ref.time_dynamics_V(g, pusher_id, mass, F, P, first, npatch, nsteps, cores, with_sort=True)
            else:
                tt, _ = ref.time_dynamics(g, order, pusher_id, mass, F, P, npatch, nsteps, cores)
            t += tt
        t += ref.time_maxwell(g, npatch, nsteps, cores)
        return t
    one(1, vector)                           # warm-up (page faults, OpenMP team)
    t1 = one(1, vector)
    per_call = max(1, min(50, int(target_seconds / max(t1, 1e-3) / max(steps, 1))))
    for _ in range(max(warmup - 1, 0)):
        one(1, vector)
    times = [one(per_call, vector) / per_call for _ in range(steps)]
    t_step = sum(times) / len(times)
    pushes = 2 * nper * npatch
    out = {"value": pushes / t_step, "t_step": t_step, "cores": cores, "kind": "reference-V" if vector else "reference",
           "lib": os.path.basename(ref.path),
           "sample": f"{npatch} distinct patches of 16^3 cells x 16 ppc x 2 species ({pushes} particles/step), the "
                     f"reference's own classes (oracle/_ref, -O3 -march=native), " +
                     ("vectorised species path: Interpolator3D2OrderV + pusher + computeParticleCellKeys + "
                      "Projector3D2OrderV + sortParticles (leavers return through the receive buffer)" if vector else
                      "scalar operators: gather + push + BC tag + deposit, no sort") +
                     f" + solveMaxwell, {per_call} step(s) per timing, {cores} OpenMP threads"}
    if vector:
        ts = one(1, False)
        out["scalar_value"] = pushes / ts
        out["scalar_what"] = "the scalar operator classes (Interpolator3D2Order, Projector3D2Order, no sort) on the same sample"
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(2, 0, args.steps, args.warmup)
    if r is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libsmilei_ref.so not built"})
        return
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["t_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 3D thermal plasma, 16 ppc, 2 species, order 2, Boris (bounded CPU sample of configs[1], >= 32 M particles)",
                       "sample": r["sample"]},
            "cpu_baseline": {k: r[k] for k in ("value", "cores", "kind", "sample", "scalar_value", "scalar_what") if k in r} | {"unit": UNIT},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def other_ceilings(args):
    """What the committed ncu capture of the dominant kernel says about the units that bind it instead of HBM (shared-memory
    wavefronts, FP64 pipe, issue slots); reported next to the HBM roofline, never re-measured here."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)["k_dynamics_o2"]
        if t["config"]["ncell"] == args.ncell and t["config"]["order"] == args.order:
            return t.get("other_ceilings_from_the_same_capture")
    except Exception:
        pass
    return None


def measured_traffic(args):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, when this run is the captured
    configuration (it is not re-measured here: a number taken under a profiler is never a bench value)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)["k_dynamics_o2"]
        if t["config"]["ncell"] == args.ncell and t["config"]["order"] == args.order:
            return t["mean_bytes_per_launch"]
    except Exception:
        pass
    return None



def order4_run(args, local):
    """BASELINE.json configs[2] (tst3d_v_o4_thermal_plasma scaled to 256^3): the same box, 16 ppc, order-4 interpolation
    and projection, Boris, one GPU; the dominant kernel timed with CUDA events exactly like the order-2 line."""
    import torch
    from smilei_b200 import namelist
    from smilei_b200.simulation import Simulation
    n = args.ncell
    params = namelist.load_namelist(namelist_source([n, n, n], 4, args.pusher), is_source=True)
    sim = Simulation(params, rank_grid=(1, 1, 1), rank=0, device=f"cuda:{local}", capacity_factor=1.08)
    T, dx, dt = plasma_constants()
    sim.init_thermal(PPC, density=1.0, temperature=T, seed=0)
    for _ in range(3):
        sim.step()
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    steps = max(3, min(args.steps, 5))
    e0, e1, dyn_ev = ev(), ev(), []
    e0.record()
    for _ in range(steps):
        sim.EMfields.restartRhoJ()
        for sp in sim.vecSpecies:
            a, b = ev(), ev()
            a.record()
            sp.dynamics(sim.EMfields, sim.smpi)
            b.record()
            dyn_ev.append((a, b))
        sim.exchange_and_solve()
        for sp in sim.vecSpecies:
            sim.patch.sort(sp.ispec)
        sim.EMfields.centerMagneticFields()
        sim.itime += 1
    e1.record()
    torch.cuda.synchronize()
    npart = sum(sim.patch.species_count(s) for s in range(2))
    ms_step = e0.elapsed_time(e1) / steps
    dyn_ms = sum(a.elapsed_time(b) for a, b in dyn_ev) / len(dyn_ev)
    dyn_bytes = 110. * (npart / 2.) + 96. * n ** 3
    peak, _ = measured_peak()
    uk, ue = sim.scalars()
    sim.close()
    return {"workload": f"synthetic 3D thermal plasma {n}^3 cells, 16 ppc, 2 species, order 4, {args.pusher} "
                        "(BASELINE.json configs[2] scaled to the bench box)",
            "steps": steps, "ms_per_step": ms_step, "value": npart / (ms_step * 1e-3), "unit": UNIT,
            "ms_per_launch": dyn_ms, "pushes_per_s_dynamics_kernel": (npart / 2.) / (dyn_ms * 1e-3),
            "roofline_frac": dyn_bytes / (dyn_ms * 1e-3) / 1e9 / peak, "kernel": "k_dynamics_o4 (one species)",
            "energies": {"Ukin": [float(v) for v in uk], "Uelm": ue}}


def multi_gpu_parity(world, rank, local, grid, steps=6, nloc=24, ppc=8):
    """The N-rank run against ONE rank from the same global state (driver-visible correctness of the exchange path):
    a box of nloc^3 cells per rank, thermal plasma with fast electrons (plenty cross rank faces, edges and corners),
    `steps` steps on the N ranks and on rank 0 alone; Ukin per species, Uelm and the particle counts are compared
    after every step."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from smilei_b200 import namelist
    from smilei_b200.simulation import Simulation
    gsize = [nloc * g for g in grid]
    T, dx, dt = plasma_constants()
    rng = np.random.default_rng(2024)
    N = gsize[0] * gsize[1] * gsize[2] * ppc
    pos = [rng.random(N) * gsize[d] * dx for d in range(3)]
    state = {}
    for name, mass, q in (("proton", 1836., 1), ("electron", 1., -1)):
        s = (T / mass) ** 0.5 * 3.0
        state[name] = dict(x=pos[0].copy(), y=pos[1].copy(), z=pos[2].copy(), px=s * rng.standard_normal(N),
                           py=s * rng.standard_normal(N), pz=s * rng.standard_normal(N), w=np.full(N, dx ** 3 / ppc),
                           q=np.full(N, q, dtype=np.int16))

    def run(rank_grid, r):
        params = namelist.load_namelist(namelist_source(gsize, 2, "boris"), is_source=True)
        sim = Simulation(params, rank_grid=rank_grid, rank=r, device=f"cuda:{local}")
        mn = [sim.pcoord[d] * sim.n[d] * dx for d in range(3)]
        mx = [(sim.pcoord[d] + 1) * sim.n[d] * dx for d in range(3)]
        for sp in sim.vecSpecies:
            a = state[sp.name]
            inside = np.ones(N, bool)
            for d, c in enumerate("xyz"):
                inside &= (a[c] >= mn[d]) & (a[c] < mx[d])
            sim.set_particles(sp.ispec, **{k: np.ascontiguousarray(v[inside]) for k, v in a.items()})
        hist = []
        for _ in range(steps):
            sim.step()
            uk, ue = sim.scalars()
            hist.append((list(uk), ue, sim.n_particles()))
        sim.close()
        return hist
    many = run(grid, rank)
    out = None
    if rank == 0:
        import torch.distributed as d2
        # rank 0 alone: a world of one (no collectives are issued by a 1x1x1 rank grid)
        one = run((1, 1, 1), 0)
        ukin_rel = max(abs(a - b) / abs(b) for (ua, _, _), (ub, _, _) in zip(many, one) for a, b in zip(ua, ub))
        uelm_rel = max(abs(ea - eb) / abs(eb) for (_, ea, _), (_, eb, _) in zip(many, one))
        out = {"ranks": world, "rank_grid": list(grid), "cells_per_rank": nloc ** 3, "particles": 2 * N, "steps": steps,
               "ukin_rel": ukin_rel, "uelm_rel": uelm_rel,
               "n_particles_equal": all(na == nb for (_, _, na), (_, _, nb) in zip(many, one)),
               "what": "N ranks vs rank 0 alone from the same global particles, max over steps and species"}
    dist.barrier()
    return out


LASER_WAKE_SRC = """
dx, dtrans, dt = 0.2, 3., 0.19
Main(geometry="3Dcartesian", interpolation_order=2, timestep=dt, number_of_timesteps=1000000,
     cell_length=[dx, dtrans, dtrans], number_of_cells={gsize!r}, number_of_patches=[{npx}, 1, 1],
     EM_boundary_conditions=[["silver-muller"]],
     EM_boundary_conditions_k=[[1., 0., 0.], [-1., 0., 0.], [1., 0.005, 0.], [1., -0.005, 0.], [1., 0., 0.005], [1., 0., -0.005]],
     solve_poisson=False)
MovingWindow(time_start={tstart!r}, velocity_x=0.9997)
Species(name="electron", position_initialization="regular", momentum_initialization="cold", particles_per_cell=1,
        mass=1.0, charge=-1.0, charge_density=0.000494, pusher="vay", boundary_conditions=[["remove", "remove"]]*3)
LaserGaussian3D(box_side="xmin", a0=2., focus=[0., {gsize[1]}*dtrans/2., {gsize[2]}*dtrans/2.], waist=10.,
                time_envelope=tgaussian(center=2**0.5*19.80, fwhm=19.80))
"""


def laser_wake_run(args, world, rank, local):
    """BASELINE.json configs[3] scaled (benchmarks/tst3d_s_o2_laser_wake_yee_vay.py: Vay pusher, cold plasma at
    1 particle per cell, Gaussian laser through the Silver-Mueller xmin side with the benchmark's oblique absorption
    vectors, `remove` particles, moving window sliding by one 8-cell patch): ncell^3 cells per GPU on the same rank
    grids as the thermal workload.  Warm-up in the fixed box, then K timed steps WHILE THE WINDOW MOVES (field slide,
    particles handed to the -x neighbour, new particles created on the last rank along x, re-sort)."""
    import torch
    import torch.distributed as dist
    from smilei_b200 import capi, namelist
    from smilei_b200.simulation import Simulation
    grid = rank_grid_for(world)
    n = args.ncell
    gsize = [n * g for g in grid]
    dt = 0.19
    warm = max(args.warmup, 3)
    params = namelist.load_namelist(LASER_WAKE_SRC.format(gsize=gsize, npx=gsize[0] // 8, tstart=(warm + 0.25) * dt),
                                    is_source=True)
    sim = Simulation(params, rank_grid=grid, rank=rank, device=f"cuda:{local}")
    sim.create_particles()
    npart = float(sum(sim.n_particles()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sim.run(warm)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    sim.run(args.steps)
    e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms = max(e0.elapsed_time(e1), 0.)
    t = torch.tensor([ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0].item())
    clocks = sampler.stop() if rank == 0 else None
    uk, ue = sim.scalars()
    npart_end = float(sum(sim.n_particles()))
    line = {
        "metric": METRIC, "value": npart * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"laser wake (BASELINE.json configs[3] scaled): {n}^3 cells per GPU, 1 ppc cold electrons, "
                               f"Vay, order 2, Silver-Mueller + laser at xmin, remove, moving window (stride 8 cells), "
                               f"rank grid {grid[0]}x{grid[1]}x{grid[2]}",
                   "cells_per_gpu": n ** 3, "particles_total": int(npart), "particles_at_end": int(npart_end),
                   "window_cells_moved": int(sim.simWindow.n_moved),
                   "l2": "fields of 13 x 150 MB and particle columns larger than L2; no flush needed"},
        "wall_ms_per_step": float(t[1].item()) / args.steps,
        "yee_cell_updates_per_s_incl_everything": n ** 3 * world * args.steps / (ms * 1e-3),
        "energies": {"Ukin": [float(v) for v in uk], "Uelm": ue},
        "gpu_launches": int(capi.launch_count() - launches0),
    }
    if clocks is not None:
        line["clocks"] = clocks
    if rank == 0:
        emit(line)
    sim.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncell", type=int, default=256, help="cells per dimension PER GPU")
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--pusher", default="boris")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--workload", default="thermal", choices=["thermal", "laser_wake"],
                    help="thermal = BASELINE.json configs[1]/[4] (the bench contract); laser_wake = configs[3] scaled")
    ap.add_argument("--no-order4", action="store_true", help="skip the order-4 sub-measurement (configs[2])")
    ap.add_argument("--no-parity", action="store_true", help="skip the N-rank vs 1-rank parity run at N > 1")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from smilei_b200 import capi, namelist
    from smilei_b200.simulation import Simulation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.workload == "laser_wake":
        return laser_wake_run(args, world, rank, local)
    grid = rank_grid_for(world)
    nloc = args.ncell
    gsize = [nloc * g for g in grid]
    params = namelist.load_namelist(namelist_source(gsize, args.order, args.pusher), is_source=True)
    sim = Simulation(params, rank_grid=grid, rank=rank, device=f"cuda:{local}", capacity_factor=1.08)
    T, dx, dt = plasma_constants()
    sim.init_thermal(PPC, density=1.0, temperature=T, seed=0)
    torch.cuda.synchronize()
    npart_local = sum(sim.patch.species_count(s) for s in range(2))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sim.step()
    barrier()

    # ---- timed region: EXACTLY K steps, device events, phase events for the roofline of the top kernel
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    dyn_ev, mw_ev, ctr_ev = [], [], []
    p = sim.patch
    launches0 = capi.launch_count()
    barrier()
    e0.record()
    for _ in range(args.steps):
        # the same call sequence as Simulation.step(), with events around the two hot kernels
        sim.EMfields.restartRhoJ()
        for sp in sim.vecSpecies:
            a, b = ev(), ev()
            a.record()
            sp.dynamics(sim.EMfields, sim.smpi)
            b.record()
            dyn_ev.append((a, b))
        a, b = ev(), ev()
        sim.exchange_and_solve(maxwell_events=(a, b))     # particle exchange (overlapped at N > 1), J sum, Yee, B exchange
        mw_ev.append((a, b))
        for sp in sim.vecSpecies:
            p.sort(sp.ispec)
        a, b = ev(), ev()
        a.record()
        sim.EMfields.centerMagneticFields()
        b.record()
        ctr_ev.append((a, b))
        sim.itime += 1
    e1.record()
    barrier()
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(npart_local)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms = float(t.item())
    npart = float(tot.item())
    ms_per_step = ms / args.steps
    value = npart * args.steps / (ms * 1e-3)
    dyn_ms = sum(a.elapsed_time(b) for a, b in dyn_ev) / len(dyn_ev)          # per launch (one species)
    mw_ms = sum(a.elapsed_time(b) for a, b in mw_ev) / len(mw_ev)
    ctr_ms = sum(a.elapsed_time(b) for a, b in ctr_ev) / len(ctr_ev)
    ncell_local = nloc ** 3
    # algorithmic bytes of one k_dynamics launch: 110 B per particle
    # (read 7 doubles + 1 short, write 6 doubles + 1 int) + 96 B per cell (6 field reads + 3 J read-modify-writes)
    dyn_bytes = 110. * (npart_local / 2.) + 96. * ncell_local
    peak, peak_src = measured_peak()
    achieved = dyn_bytes / (dyn_ms * 1e-3) / 1e9
    uk, ue = sim.scalars()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"synthetic 3D thermal plasma {nloc}^3 cells per GPU, 16 ppc, 2 species, order {args.order}, "
                               f"{args.pusher}, periodic, rank grid {grid[0]}x{grid[1]}x{grid[2]} (BASELINE.json configs[1]/[4])",
                   "cells_per_gpu": ncell_local, "particles_total": int(npart), "T_keV": T_KEV, "dx": dx, "dt": dt,
                   "l2": "inputs larger than L2 (SoA columns of 2.1 GB each, fields 150 MB each; no flush needed)"},
        "pushes_per_s_dynamics_kernel": (npart_local / 2.) / (dyn_ms * 1e-3) * world,
        # Yee = E update + B update + B_m centring: solver sweeps + the ghost-shell centring kernel
        "yee_cell_updates_per_s": ncell_local / ((mw_ms + ctr_ms) * 1e-3) * world,
        "yee_roofline_frac": (192. * ncell_local / ((mw_ms + ctr_ms) * 1e-3) / 1e9) / peak,
        "yee_ms": {"ampere_faraday_center": mw_ms, "center_shell": ctr_ms},
        "roofline": {"bound": "hbm", "kernel": "k_dynamics (gather+push+BC+deposit, one species)", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(args), "other_ceilings": other_ceilings(args), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dyn_bytes, "ms_per_launch": dyn_ms,
                     "note": "bound by the shared-memory pipe (66 / 78 % of its wavefront peak) and FP64 latency, not by HBM (DESIGN.md §4.2, profiles/r2_final_dynamics_256.txt, r2_lds_pattern.txt); "
                             "traffic = DRAM bytes per launch from the ncu --set full capture recorded in profiles/r2_traffic.json"},
        "energies": {"Ukin": [float(v) for v in uk], "Uelm": ue},
        "gpu_launches": int(launches),
    }
    if clocks is not None:
        line["clocks"] = clocks

    # ---- e2e: the same step through the C ABI with HOST buffers, H2D of the inputs and D2H of the
    #      step's result (energy scalars) inside the timed region
    if not args.no_e2e:
        line["e2e"] = e2e_run(sim, args, world, rank, npart, barrier)
    # ---- CPU baseline on rank 0 at N=1
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.order, capi.PUSHERS[args.pusher], 3, 1, target_seconds=10.)
        if r is not None:
            line["cpu_baseline"] = {k: r[k] for k in ("value", "cores", "kind", "sample", "scalar_value", "scalar_what")
                                    if k in r} | {"unit": UNIT}
    sim.close()
    # ---- configs[2]: the order-4 kernel on the same box (N = 1 only; a sub-object of the same line)
    if world == 1 and args.order == 2 and not args.no_order4:
        line["order4"] = order4_run(args, local)
    # ---- N > 1: the exchange path checked against one rank, where the driver sees it
    if world > 1 and not args.no_parity:
        par = multi_gpu_parity(world, rank, local, grid)
        if rank == 0:
            line["multi_gpu_parity"] = par
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_run(sim, args, world, rank, npart, barrier):
    import numpy as np
    import torch
    import torch.distributed as dist
    p = sim.patch
    host = []
    pinned = True
    for s in range(2):
        cols = p.species_get(s)                      # D2H once, outside the timed region
        cols.pop("key")
        if pinned:
            try:
                pin = {}
                for k, v in cols.items():
                    t = torch.empty(v.shape, dtype=torch.float64 if v.dtype == np.float64 else torch.int16, pin_memory=True)
                    t.numpy()[...] = v
                    pin[k] = t
                cols = {k: t.numpy() for k, t in pin.items()}
                host.append((cols, pin))
                continue
            except Exception:
                pinned = False
        host.append((cols, None))
    h2d = sum(v.nbytes for cols, _ in host for v in cols.values())
    d2h = 8 * 3

    def step():
        for s, (c, _) in enumerate(host):
            p.species_set(s, c["x"], c["y"], c["z"], c["px"], c["py"], c["pz"], c["w"], c["q"])    # H2D (host buffers)
            p.sort(s)
        sim.step()
        return sim.patch.energy()                                                                   # D2H of the step's result
    step()
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        step()
    e1.record()
    barrier()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # host-side copies count: the slower of device and wall clock
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"value": npart * args.e2e_steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "pinned_host_memory": pinned,
            "what": "per step: sb200_species_set of every SoA column from host memory (H2D), sb200_sort, the full "
                    "PIC step, sb200_energy (D2H)"}


if __name__ == "__main__":
    main()
