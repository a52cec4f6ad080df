#!/usr/bin/env python
"""bench.py -- particle-steps/s of the full sheath PIC step (PIC_L_DD physics) on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU port of the reference path

One "step" = one full timestep (re-injection + the whole Picard loop: fused
gather/push/absorb/deposit kernel, [all-reduce], field update, until converged) over
every particle of the workload.  A particle-step = one particle through one step.

N=1 workload = BASELINE.json configs[1]: 1e8 particles/species (2e8 total), 4097-node
(4096-cell) grid, PIC_L_DD physics.  N>1: weak scaling, 2e8 particles per GPU, particle
decomposition with one fp64 all-reduce of [jh|j1|absorbed counts] per Picard iteration.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

E_CH, ME, MP, KB = 1.602E-19, 9.11E-31, 1.67E-27, 1.38E-23
METRIC = "particle-steps/sec (full PIC step)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--particles-per-gpu", type=float, default=2e8)
    ap.add_argument("--total-particles", type=float, default=0, help="strong scaling: fixed total (e.g. 1e9)")
    ap.add_argument("--cells", type=int, default=4096)
    ap.add_argument("--sort-every", type=int, default=None,
                    help="steps between cell sorts (default: 12 for the sheath on grids the wide-window kernel serves, "
                         "<= 4352 nodes, with the particle decomposition and the default deposit; 8 otherwise)")
    ap.add_argument("--heavy-sort-every", type=int, default=1, help="every n-th sort also re-sorts the ions (1 = always)")
    ap.add_argument("--deposit", default="window", choices=["window", "window-det", "window-blocked", "window-big", "window-ldg", "warp", "atomic"])
    ap.add_argument("--workload", default="sheath", choices=["sheath", "explicit", "pypic", "boris"],
                    help="sheath = BASELINE configs[1] (default, the driver's bench); explicit / pypic / boris = "
                         "the other movers of SURVEY.md 8(d) at the same size (single GPU, device-resident)")
    ap.add_argument("--decomposition", default="particle", choices=["particle", "slab"],
                    help="sheath workload only: particle decomposition (default, all-reduce of the grid) or spatial "
                         "slabs (halo exchange + particle migration, BASELINE config 5)")
    ap.add_argument("--reduce", default="nccl", choices=["nccl", "p2p"],
                    help="N > 1: sum of the grid accumulators over ranks by NCCL all-reduce (default) or inside the field "
                         "kernel over NVLink peer memory (off by default: parity run pending, DESIGN.md 6)")
    ap.add_argument("--strong-total", type=float, default=1e9,
                    help="second timed region: the north-star strong-scaling case, this many particles in TOTAL over the "
                         "N GPUs (BASELINE config 4), reported as the sub-record `strong_scaling`; 0 disables it")
    ap.add_argument("--strong-steps", type=int, default=12)
    ap.add_argument("--boris-full-store", action="store_true",
                    help="--workload boris: carry y, z and the per-particle clock through the push (112 B per particle-step) "
                         "instead of the lean store (x, vx, vy, vz: the 64 B row of SURVEY.md 8d)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-slab-leg", action="store_true",
                    help="N > 1: skip the `slab_decomposition` sub-record (the same workload on spatial slabs: halo exchange + "
                         "particle migration, BASELINE config 5's decomposition)")
    ap.add_argument("--slab-steps", type=int, default=16)
    ap.add_argument("--slab-field", default="distributed", choices=["distributed", "replicated"],
                    help="slab decomposition: field update on every rank's own nodes + guard nodes with two small all-gathers "
                         "per Picard iteration (default), or the whole grid gathered and updated on every rank (A/B)")
    ap.add_argument("--no-api-leg", action="store_true",
                    help="skip the `reference_api` record (the step as PIC_L_DD.main_i drives it: host MT19937 draws, carried v,w)")
    ap.add_argument("--api-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=float, default=4e6, help="particles in the CPU-baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=16)
    args = ap.parse_args()
    if args.sort_every is None:
        wide = (args.workload == "sheath" and args.decomposition == "particle" and args.cells + 1 <= 4352
                and args.deposit in ("window", "window-det", "window-blocked") and os.environ.get("PIC_V6_NARROW") != "1")
        args.sort_every = 12 if wide else 8
        if args.workload == "explicit" and os.environ.get("PIC_S_NARROW") != "1" and args.cells + 1 <= 8192:
            args.sort_every = 16          # 15-node windows of l_push_deposit_v2_k (profiles/r2_periodic_wide.txt)
    return args


def workload(args, world):
    Ng = args.cells + 1
    dx, dt = 1e-5, 1e-12
    L = dx * (Ng - 1)
    if args.total_particles:
        N = int(args.total_particles); scaling = "strong"
    else:
        N = int(args.particles_per_gpu) * world; scaling = "weak"
    N -= N % 2
    Te = Ti = 10.0 * 11600.
    density = 1e19
    return dict(N=N, Ng=Ng, dx=dx, dt=dt, L=L, Te=Te, Ti=Ti, density=density, p2c=L * density / N,
                kBTe=KB * Te, kBTi=KB * Ti, scaling=scaling, tol=1e-5, maxiter=20)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons DURING the timed region: NVML in-process (a sample
    costs ~0.1 ms, so even a 100 ms region sees dozens), nvidia-smi subprocess as fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    PERIOD = float(os.environ.get("PIC_CLOCK_PERIOD_MS", "5")) * 1e-3     # seconds between NVML samples

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.src = index, [], False, "nvidia-smi"
        self.h = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.h, self.nv, self.src = nv.nvmlDeviceGetHandleByIndex(phys), nv, "nvml"
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def sample(self):
        if self.h is not None:
            nv = self.nv
            sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                    nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap]
            return [sm, self.smax, pw] + [bool(rs & b) for b in bits]
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        c = [t.strip() for t in out.strip().split("\n")[0].split(",")]
        return [float(c[0]), float(c[1]), float(c[2])] + [t.lower().startswith("active") for t in c[3:7]]

    def run(self):
        while not self.stop_flag:
            try:
                self.rows.append(self.sample())
            except Exception:
                pass
            time.sleep(self.PERIOD if self.h is not None else 0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.src}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for k, n in enumerate(self.NAMES) if any(r[3 + k] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.rows[0][1],
                "power_w_max": max(r[2] for r in self.rows), "reasons": reasons, "samples": len(self.rows),
                "source": self.src}


def bind_to_gpu_numa_node(index):
    """Pins this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) so that the
    pinned staging buffers of the e2e leg are first-touched on the GPU's own NUMA node; with 8
    ranks on one socket's memory the host<->device copies are limited by the inter-socket link."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = index
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            phys = int(vis.split(",")[index])
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(phys))
        return True
    except Exception:
        return False


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to every rank, which would
    make the CPU arm a one-thread run: the thread count is taken from the affinity mask instead and
    passed to the port explicitly (its OpenMP regions carry a num_threads clause)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- CPU arm
def cpu_port_run(w, n_sample, steps, warmup, threads=None):
    """Times the C port of the reference's sheath timestep (oracle/c/dd_oracle.c) on the
    host cores.  Bounded sample of the same workload: same grid, same physics, n_sample
    particles (p2c rescaled so the plasma density is unchanged)."""
    from oracle import c_oracle
    threads = c_oracle.team_size(threads or host_threads())      # what OpenMP grants, not what was asked for
    n = int(n_sample); n -= n % 2
    rs = np.random.RandomState(1)
    h = n // 2
    x0 = rs.uniform(0, w["L"], n)
    u0 = np.concatenate([rs.normal(0, np.sqrt(w["kBTe"] / ME), h), rs.normal(0, np.sqrt(w["kBTi"] / MP), h)])
    E0 = np.zeros(w["Ng"])
    p2c = w["L"] * w["density"] / n
    times, iters = [], []
    for s in range(warmup + steps):
        act = np.ones(n)
        t0 = time.perf_counter()
        x1, u1, E1, j1, k, r = c_oracle.dd_picard_step(x0, u0, [-E_CH, E_CH], [ME, MP], h, act, E0, p2c, w["Ng"],
                                                       w["dx"], w["dt"], w["L"], w["tol"], w["maxiter"], threads)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt); iters.append(k)
        # next step: survivors keep going, absorbed slots are re-drawn (host RNG)
        dead = act != 1
        x0 = np.where(dead, rs.uniform(0, w["L"], n), x1)
        sig = np.concatenate([np.full(h, np.sqrt(w["kBTe"] / ME)), np.full(h, np.sqrt(w["kBTi"] / MP))])
        u0 = np.where(dead, rs.normal(0, 1, n) * sig, u1)
        E0 = E1
    tot = sum(times)
    return dict(value=n * len(times) / tot, ms_per_step=1e3 * tot / len(times), cores=threads,
                sample="%d particles (%.3g of the workload), %d-node grid, %d timed steps, mean %.1f Picard iterations"
                       % (n, n / w["N"], w["Ng"], len(times), float(np.mean(iters))), iters=float(np.mean(iters)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args, max(1, args.gpus))
    threads = host_threads()
    # size the sample so the whole run ends within a few minutes
    res = cpu_port_run(w, args.cpu_sample, max(1, min(args.steps, 10)), max(1, min(args.warmup, 2)), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "1D sheath PIC_L_DD physics, %d particles, %d-node grid" % (w["N"], w["Ng"]),
                   "note": "CPU port of the reference loop (oracle/c/dd_oracle.c, OpenMP) on a bounded sample; the "
                           "reference itself is pure Python and cannot travel to the GPU box"},
        "cpu_baseline": {"value": res["value"], "unit": "particle-steps/s", "cores": res["cores"], "kind": "port",
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ----------------------------------------------------------------------------- GPU arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import ctypes as C
    from pypic_b200 import _lib, device as D
    from pypic_b200.dist import Comm
    from pypic_b200.sheath import SheathSim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:                      # the 1-GPU run keeps every host core for the CPU-baseline leg
        bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    w = workload(args, world)
    dev = torch.device("cuda", local)

    def make_sim(wl):
        sm = SheathSim(wl["N"], wl["Ng"], wl["dx"], wl["dt"], wl["p2c"], tol=wl["tol"], maxiter=wl["maxiter"],
                       kBT=(wl["kBTe"], wl["kBTi"]), carry_vw=False, deposit=args.deposit, rng="philox", seed=1,
                       comm=comm, device=dev, sort_every=args.sort_every, reduce=args.reduce)
        sm.heavy_sort_every = max(1, args.heavy_sort_every)
        # synthetic initial state, generated on the device (x~U(0,L), u~N(0,sqrt(kT/m)))
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
        ns = sm.n_split
        sm.x0.uniform_(0.0, 1.0, generator=gen).mul_(wl["L"]).clamp_(1e-12, wl["L"] * (1 - 1e-12))
        sm.u0.normal_(0.0, 1.0, generator=gen)
        sm.u0[:ns].mul_(float(np.sqrt(wl["kBTe"] / ME)))
        sm.u0[ns:].mul_(float(np.sqrt(wl["kBTi"] / MP)))
        torch.cuda.synchronize()
        return sm

    def timed_steps(sm, steps, warmup, sampler=None):
        """W untimed steps, then exactly K steps between barrier + synchronize, CUDA events, max over ranks."""
        for _ in range(warmup):
            sm.step()
        sm.check()
        torch.cuda.synchronize()
        comm.barrier()
        if sampler:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        events = sm.iter_events = []
        l0 = sm.kernel_launches
        its = []
        torch.cuda.synchronize()
        comm.barrier()
        ev0.record()
        for _ in range(steps):
            k, r = sm.step()
            its.append(k)
        ev1.record()
        torch.cuda.synchronize()
        comm.barrier()
        sm.iter_events = None
        if sampler:
            sampler.stop_flag = True
        t_ms = comm.max_float(ev0.elapsed_time(ev1), device=dev)
        sm.check()
        return dict(ms=t_ms, iters=its, events=events, launches=sm.kernel_launches - l0)

    sim = make_sim(w)
    sampler = ClockSampler(local) if rank == 0 else None
    tr = timed_steps(sim, args.steps, args.warmup, sampler)
    ms, iters, iter_events, launches = tr["ms"], tr["iters"], tr["events"], tr["launches"]
    kernel_ms = [e[0].elapsed_time(e[1]) for e in iter_events]
    ms_with_u = [e[0].elapsed_time(e[1]) for e in iter_events if e[2] and not e[3]]
    ms_without_u = [e[0].elapsed_time(e[1]) for e in iter_events if not e[2] and not e[3]]
    ms_first = [e[0].elapsed_time(e[1]) for e in iter_events if e[3]]
    mean_iter_ms = float(np.mean(kernel_ms))
    kbar = float(np.mean(iters))
    value = w["N"] * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (fused gather+push+absorb+deposit), HBM bound
    # algorithmic bytes per particle-step = 32*k + 16 -> per launch N_local*(32 + 16/k)
    alg_bytes_launch = sim.N * (32.0 + 16.0 / kbar)
    achieved = alg_bytes_launch / (mean_iter_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                "traffic": None, "kernel": "dd_picard_iter_v6_k", "peak_source": peak_src,
                "kernel_ms_mean": mean_iter_ms, "kernel_share_of_step": float(np.sum(kernel_ms) / ms),
                "algorithmic_bytes_per_launch": alg_bytes_launch, "mean_picard_iterations": kbar,
                "kernel_ms_by_kind": {"first_iteration": float(np.mean(ms_first)) if ms_first else None,
                                      "later_without_u1_store": float(np.mean(ms_without_u)) if ms_without_u else None,
                                      "later_with_u1_store": float(np.mean(ms_with_u)) if ms_with_u else None},
                "u1_repair_passes": int(sim.u_repairs),
                "note": "achieved = ALGORITHMIC bytes (SURVEY 8d: 32 B per Picard iteration + 16 B commit per particle-step) / "
                        "measured launch time; the kernel moves fewer DRAM bytes than that (traffic: first iteration 24 B, "
                        "light iterations 32 B, last iteration 40 B per particle -> 160 B instead of 176 B per 5-iteration "
                        "particle-step), so frac can approach 1 while the actual DRAM rate is ~0.88 of the copy peak"}
    if world > 1:
        # the coupled step runs at the pace of the slowest rank in every Picard iteration: the spread of the
        # per-rank kernel times (clocks differ from GPU to GPU under the power cap) is what the weak-scaling
        # efficiency loses, whatever the reduction costs
        t = torch.tensor([mean_iter_ms], dtype=torch.float64, device=dev)
        allk = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allk, t)
        per_rank = [float(v.item()) for v in allk]
        roofline["kernel_ms_mean_per_rank"] = per_rank
        roofline["step_ms_if_every_iteration_waits_for_the_slowest_rank"] = float(kbar * max(per_rank) + (ms - kbar * mean_iter_ms))
    tf = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tf):
        try:
            roofline["traffic"] = json.load(open(tf))["dram_bytes_per_particle"] * sim.N   # ncu, per launch
        except Exception:
            pass

    # ---- e2e: the same step through the C ABI with HOST (pinned) buffers, every copy timed.
    # pic_host_dd_step_batches advances a queue of independent batches (here: the same synthetic
    # state re-sent every step, results into two alternating pinned output sets) and pipelines
    # them through two device slots: upload of batch b+1 and download of batch b-1 overlap the
    # Picard loop of batch b.  Per step 16 B/particle go up and 17 B/particle come down.
    e2e = None
    if not args.no_e2e:
        Nl, Ng = sim.N, w["Ng"]
        nb = max(1, args.e2e_steps)
        pin = lambda n, dt=torch.float64: torch.empty(n, dtype=dt, pin_memory=True)
        hx0, hu0, hE0 = pin(Nl), pin(Nl), pin(Ng)
        outs = [dict(x1=pin(Nl), u1=pin(Nl), act=pin(Nl, torch.int8), E1=pin(Ng), j1=pin(Ng)) for _ in range(min(nb, 2))]
        hx0.copy_(sim.x0); hu0.copy_(sim.u0); hE0.copy_(sim.E0)
        torch.cuda.synchronize()
        h2d = Nl * 16 + Ng * 8
        d2h = Nl * 17 + Ng * 16
        if world == 1:
            P = _lib.DDParams(Nl, sim.n_split, Ng, 0, w["dx"], w["dt"], w["L"], w["p2c"],
                              (C.c_double * 2)(-E_CH, E_CH), (C.c_double * 2)(ME, MP))

            def ptrs(vals):
                return (C.c_void_p * len(vals))(*vals)

            def host_steps(n):
                it, res = (C.c_int * n)(), (C.c_double * n)()
                o = [outs[b % len(outs)] for b in range(n)]
                _lib.call("pic_host_dd_step_batches", C.byref(P), n, ptrs([hx0.data_ptr()] * n), ptrs([hu0.data_ptr()] * n),
                          ptrs([hE0.data_ptr()] * n), w["tol"], w["maxiter"], ptrs([d["x1"].data_ptr() for d in o]),
                          ptrs([d["u1"].data_ptr() for d in o]), ptrs([d["act"].data_ptr() for d in o]),
                          ptrs([d["E1"].data_ptr() for d in o]), ptrs([d["j1"].data_ptr() for d in o]), it, res)
                return list(it)
            api = ("pic_host_dd_step_batches (C ABI, pinned host buffers, %d batches pipelined through two device slots)" % nb)
        else:
            # N > 1 is ONE coupled system (every Picard iteration sums the ranks' accumulators), so the host-buffer
            # path goes through the resident driver, whose Picard loop holds the reduction: per step every rank
            # copies its shard in from pinned memory, all ranks take the coupled step, results are copied back out
            from pypic_b200.hostpipe import HostPipelinedSheath
            pipe = HostPipelinedSheath(sim)

            def host_steps(n):
                return pipe.run([dict(x0=hx0, u0=hu0, E0=hE0)] * n, outs)
            api = ("HostPipelinedSheath over SheathSim (pinned host buffers per rank; COUPLED: %s per Picard iteration; "
                   "copies of step b+1 / b-1 overlap the Picard loop of step b)"
                   % ("all-reduce of [jh|j1|counts]" if sim.p2p is None else "peer-memory sum inside the field kernel"))
        host_steps(min(nb, 2))           # warm-up (allocates the device slots)
        torch.cuda.synchronize()
        comm.barrier()
        t0 = time.perf_counter()
        its = host_steps(nb)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        t_e2e = comm.max_float(t_e2e, device=dev)
        # ---- what the host<->device link allows: the same bytes per step in both directions at once, all ranks
        # concurrently, nothing else running -- the ceiling of any host-buffer path on this box
        cs_up, cs_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        dbuf_in = torch.empty(2 * Nl, dtype=torch.float64, device=dev)
        dbuf_out = torch.empty(2 * Nl + Nl // 8 + 1, dtype=torch.float64, device=dev)
        hin = torch.empty(2 * Nl, dtype=torch.float64, pin_memory=True)
        hout = torch.empty(2 * Nl + Nl // 8 + 1, dtype=torch.float64, pin_memory=True)
        reps = 3
        torch.cuda.synchronize(); comm.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(cs_up):
                dbuf_in.copy_(hin, non_blocking=True)
            with torch.cuda.stream(cs_dn):
                hout.copy_(dbuf_out, non_blocking=True)
        torch.cuda.synchronize()
        t_copy = comm.max_float(time.perf_counter() - t0, device=dev) / reps
        del dbuf_in, dbuf_out, hin, hout
        e2e = {"value": w["N"] * nb / t_e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": nb, "picard_iterations": its[-1],
               "ms_per_step": 1e3 * t_e2e / nb, "coupled_over_ranks": bool(world > 1),
               "pcie_gbs": {"h2d": h2d * nb / t_e2e / 1e9, "d2h": d2h * nb / t_e2e / 1e9},
               "pcie_ceiling": {"ms_per_step": 1e3 * t_copy, "h2d_gbs": Nl * 16 / t_copy / 1e9, "d2h_gbs": Nl * 17 / t_copy / 1e9,
                                "particle_steps_per_s": w["N"] / t_copy,
                                "note": "this step's bytes copied both ways concurrently from/to pinned memory on every rank at once, "
                                        "no compute: the bound of any host-buffer path on this box (per rank)"},
               "api": api}
        if world == 1:
            _lib.call("pic_host_release")
        else:
            del pipe
        del hx0, hu0, outs

    enq = sim.enqueue_ahead
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_port_run(w, args.cpu_sample, 3, 1)
        cpu = {"value": res["value"], "unit": "particle-steps/s", "cores": res["cores"], "kind": "port",
               "sample": res["sample"]}

    # ---- the same workload the way the REFERENCE'S OWN API drives it (PIC_L_DD.main_i -> SheathSim(rng="host",
    # carry_vw=True, sort_every=8)): per step the host draws the re-injected particles from the legacy MT19937 stream
    # in the reference's index order (the thermostat's uniforms are jumped over), copies them in, and reads the
    # step's diagnostics (iteration count, residuals, EE, KE, jbias, kBTe) and the absorption log back
    api = None
    if not args.no_api_leg and world == 1:
        from pypic_b200.rng import LegacyDraws
        sim.close(); del sim
        torch.cuda.empty_cache()
        sm = SheathSim(w["N"], w["Ng"], w["dx"], w["dt"], w["p2c"], tol=w["tol"], maxiter=w["maxiter"],
                       kBT=(w["kBTe"], w["kBTi"]), carry_vw=True, rng="host", comm=comm, device=dev, sort_every=args.sort_every,
                       draws=LegacyDraws(np.random.RandomState(1)), vion_after=2000)
        gen = torch.Generator(device=dev); gen.manual_seed(4321)
        sm.x0.uniform_(0.0, 1.0, generator=gen).mul_(w["L"]).clamp_(1e-12, w["L"] * (1 - 1e-12))
        sm.u0.normal_(0.0, 1.0, generator=gen)
        sm.u0[:sm.n_split].mul_(float(np.sqrt(w["kBTe"] / ME))); sm.u0[sm.n_split:].mul_(float(np.sqrt(w["kBTi"] / MP)))
        torch.cuda.synchronize()
        na = max(1, args.api_steps)
        sm.fused_moments = True                  # as main_i: np.std(u0) / KE ride the first Picard iteration
        with sm.draws.hold():
            for _ in range(max(3, args.warmup)):
                sm.step()
            torch.cuda.synchronize()
            l0, dead0 = sm.kernel_launches, 0
            t0 = time.perf_counter()
            its_api = []
            for s_ in range(na):
                its_api.append(sm.step()[0])
                m1_, m2_ = sm.pre_step_moments()
                d_last = dict(sm.step_stats(), KE_previous_step=ME / 2. * m2_, kBTe=sm.kBTe_from(m1_, m2_))
            torch.cuda.synchronize()
            t_api = time.perf_counter() - t0
        sm.check()
        api = {"value": w["N"] * na / t_api, "unit": "particle-steps/s", "ms_per_step": 1e3 * t_api / na, "steps": na,
               "picard_iterations_per_step": float(np.mean(its_api)), "gpu_launches": int(sm.kernel_launches - l0),
               "mt_jumps": int(sm.draws.jumps), "mt_jumps_prefetched": int(sm.draws.prefetch_hits), "sorts": int(sm._sorts),
               "timing": "host wall clock around the loop (every step ends in a device->host read)",
               "api": "SheathSim as PIC_L_DD.main_i builds it (rng='host': legacy MT19937 draws in original-index order; "
                      "carry_vw=True; cell sort every %d steps with the original-index payload)" % args.sort_every,
               "last_step": {k_: float(v_) for k_, v_ in d_last.items()}}
        sim = sm

    # ---- N > 1: the same workload on SPATIAL SLABS (pypic_b200/spatial.py: every rank owns a cell range and the
    # particles inside it; halo exchange of the guard strips per Picard iteration, particle migration with the sort,
    # routed re-injection) as a sub-record, so that the driver's scaling sweep holds a number for that decomposition too
    slab = None
    if world > 1 and not args.no_slab_leg and args.decomposition == "particle":
        try:
            from pypic_b200.spatial import SlabSheathSim
            n_keep, p2p_keep, enq_keep = sim.N, sim.p2p is not None, sim.enqueue_ahead
            sim.close(); del sim
            torch.cuda.empty_cache()
            ss = SlabSheathSim(w["N"], w["Ng"], w["dx"], w["dt"], w["p2c"], kBT=(w["kBTe"], w["kBTi"]), tol=w["tol"],
                               maxiter=w["maxiter"], seed=1, comm=comm, device=dev, sort_every=min(args.sort_every, 8), guard=16,
                               field=args.slab_field)     # 8 steps between migrations: the setting the 16 guard cells were validated with
            ss.init_device(seed=1234)

            def agreed_check():
                """check() on every rank, then ONE collective so that all ranks leave the leg together when any of them
                failed (a rank that raised alone would meet the others in different collectives afterwards)."""
                msg = ""
                try:
                    ss.check()
                except Exception as ex_:
                    msg = repr(ex_)[:300]
                if comm.max_float(1.0 if msg else 0.0, device=dev) > 0.0:
                    raise RuntimeError(msg or "slab check failed on another rank")
            for _ in range(3):
                ss.step()
            agreed_check()
            torch.cuda.synchronize(); comm.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            its_sl = []
            e0.record()
            for _ in range(max(1, args.slab_steps)):
                its_sl.append(ss.step()[0])
            e1.record()
            torch.cuda.synchronize(); comm.barrier()
            ms_sl = comm.max_float(e0.elapsed_time(e1), device=dev)
            agreed_check()
            slab = {"value": w["N"] * len(its_sl) / (ms_sl * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms_sl / len(its_sl),
                    "steps": len(its_sl), "picard_iterations_per_step": float(np.mean(its_sl)), "n_gpus": world,
                    "migration": dict(ss.stat), "guard_cells": 16, "field_update": args.slab_field,
                    "sort_every": min(args.sort_every, 8),
                    "note": "spatial slab decomposition of the same workload; the particle decomposition above is the default"}
            del ss
            torch.cuda.empty_cache()
        except Exception as ex:                      # the sub-record must never take the main line down
            slab = {"error": repr(ex)[:300]}
        sim = None

    # ---- the north-star strong-scaling case (BASELINE config 4): a fixed TOTAL of particles over the N GPUs, same
    # code, same run -- the driver's 1/2/4/8 sweep then holds the whole curve (sub-record `strong_scaling`)
    strong = None
    if args.strong_total and not args.total_particles and args.strong_total != w["N"]:
        if sim is not None:
            n_main, p2p_main = sim.N, sim.p2p is not None
            sim.close()
            del sim
        else:
            n_main, p2p_main = n_keep, p2p_keep
        torch.cuda.empty_cache()
        a2 = argparse.Namespace(**vars(args)); a2.total_particles = args.strong_total
        w2 = workload(a2, world)
        sim = make_sim(w2)
        tr2 = timed_steps(sim, max(1, args.strong_steps), max(3, min(args.warmup, 3)))
        kms2 = [e[0].elapsed_time(e[1]) for e in tr2["events"]]
        kbar2 = float(np.mean(tr2["iters"]))
        ach2 = sim.N * (32.0 + 16.0 / kbar2) / (float(np.mean(kms2)) * 1e-3) / 1e9
        strong = {"particles_total": w2["N"], "particles_per_gpu": sim.N, "n_gpus": world, "scaling": "strong",
                  "value": w2["N"] * len(tr2["iters"]) / (tr2["ms"] * 1e-3), "unit": "particle-steps/s",
                  "ms_per_step": tr2["ms"] / len(tr2["iters"]), "steps": len(tr2["iters"]), "picard_iterations_per_step": kbar2,
                  "kernel_ms_mean": float(np.mean(kms2)), "kernel_share_of_step": float(np.sum(kms2) / tr2["ms"]),
                  "roofline_frac": ach2 / peak, "gpu_launches": int(tr2["launches"]),
                  "note": "speed-up at N GPUs = this value / the same sub-record of the --gpus 1 run (one code version)"}
        n_local = n_main
    elif sim is not None:
        n_local, p2p_main = sim.N, sim.p2p is not None
        sim.close()
    else:
        n_local, p2p_main = n_keep, p2p_keep

    if rank == 0:
        clocks = sampler.summary() if sampler else {}
        line = {
            "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "1D sheath (PIC_L_DD physics): %d particles (%d per GPU, e-/p+ halves), %d-node grid, "
                                   "implicit CN/Picard tol=1e-5" % (w["N"], n_local, w["Ng"]),
                       "parallelism": ("particle decomposition x%d, %s of [jh|j1|counts] per Picard iteration"
                                       % (world, "fp64 all-reduce" if not p2p_main else
                                          "sum over NVLink peer memory inside the field kernel (rank order)"))
                       if world > 1 else "single GPU",
                       "picard_iterations_per_step": kbar, "deposit": args.deposit, "sort_every": args.sort_every,
                       "sort": "electrons every %d steps, ions with them every %d steps" % (args.sort_every,
                                                                                             args.sort_every * max(1, args.heavy_sort_every)),
                       "reinjection": "device Philox4x32-10 (statistical parity)",
                       "picard_loop": ("enqueue-ahead: the iterations the previous step needed are queued behind a device flag, "
                                       "one host round trip per step") if enq else "one host round trip per iteration",
                       "parity_tolerance": "bit-exact cell indices, flags, counts, iteration counts; fields and particles after N steps "
                                           "within 1e-11 relative of the reference (fp64; north_star asks 1e-12 -- the order of the "
                                           "parallel deposit re-associates sums at 1e-13 per step and that round-off feeds back, "
                                           "DESIGN.md section 4)",
                       "l2_policy": "inputs (%.1f GB of particle arrays per GPU) are far larger than the 126 MB L2" % (n_local * 32 / 1e9)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "reference_api": api, "strong_scaling": strong,
            "slab_decomposition": slab,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


# ----------------------------------------------------------------------------- slab decomposition
def run_slab(args):
    """BASELINE config 5: the sheath on spatial slabs (pypic_b200/spatial.py), device-resident."""
    import torch
    import torch.distributed as dist
    from pypic_b200.dist import Comm
    from pypic_b200.spatial import SlabSheathSim
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    w = workload(args, world)
    sim = SlabSheathSim(w["N"], w["Ng"], w["dx"], w["dt"], w["p2c"], kBT=(w["kBTe"], w["kBTi"]), tol=w["tol"],
                        maxiter=w["maxiter"], seed=1, comm=comm, device=dev, sort_every=args.sort_every, guard=16,
                        field=args.slab_field)
    sim.init_device(seed=1234)
    for _ in range(args.warmup):
        sim.step()
    sim.check()
    torch.cuda.synchronize(); comm.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    sim.iter_events = []
    l0 = sim.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = []
    ev0.record()
    for _ in range(args.steps):
        k, r = sim.step(); iters.append(k)
    ev1.record()
    torch.cuda.synchronize(); comm.barrier()
    if sampler:
        sampler.stop_flag = True
    ms = comm.max_float(ev0.elapsed_time(ev1), device=dev)
    sim.check()
    kms = [a.elapsed_time(b) for a, b in sim.iter_events]
    kbar = float(np.mean(iters))
    nloc = sim.local_particles()
    by_rank = [float(np.mean(kms))]
    if world > 1:
        by_rank = [None] * world
        dist.all_gather_object(by_rank, (float(np.mean(kms)), nloc))
    peak, peak_src = measured_peak()
    per_launch = nloc * 40.0                       # x0,u0,x1 in; x1,u1 out (u1 is always stored on this path)
    achieved = per_launch / (float(np.mean(kms)) * 1e-3) / 1e9
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": w["N"] * args.steps / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "1D sheath (PIC_L_DD physics) on %d spatial slabs: %d particles, %d-node grid, implicit "
                                       "CN/Picard tol=1e-5" % (world, w["N"], w["Ng"]),
                           "parallelism": ("slab decomposition x%d: %s, migration with the sort every %d steps, routed re-injection"
                                           % (world, "field update distributed over the slabs (own nodes + 16 guard nodes); per "
                                              "Picard iteration one all-gather of the boundary bands and partial sums and one of "
                                              "the residual partials" if args.slab_field == "distributed" else
                                              "halo exchange of 16 guard nodes + all-gather of owned segments per Picard "
                                              "iteration, field update replicated", args.sort_every)),
                           "picard_iterations_per_step": kbar, "sort_every": args.sort_every,
                           "migration": sim.stat, "local_particles_rank0": nloc, "kernel_ms_and_particles_by_rank": by_rank,
                           "l2_policy": "particle arrays (%.1f GB per GPU) are far larger than the 126 MB L2" % (nloc * 32 / 1e9)},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "kernel": "dd_picard_iter_v6_k (one launch per species block)",
                             "peak_source": peak_src, "kernel_ms_mean": float(np.mean(kms)),
                             "kernel_share_of_step": float(np.sum(kms) / ms), "algorithmic_bytes_per_launch": per_launch},
                "cpu_baseline": None, "e2e": None, "gpu_launches": int(sim.kernel_launches - l0),
                "clocks": sampler.summary() if sampler else {}}
    if world > 1:
        dist.destroy_process_group()
    return line


# ----------------------------------------------------------------------------- other movers
def run_other(args):
    """The other hot-path rows at benchmark size, one GPU, device-resident state:
    explicit = PIC_L explicit leapfrog full step (push+gather+deposit fused, periodic Poisson by PCR);
    pypic    = pypic.py periodic implicit CN/Picard full step;
    boris    = pygcpic Boris 1D3V step (fused gather+push+walls+deposit, Newton-Boltzmann solve)."""
    import torch
    import torch.distributed as dist
    from pypic_b200.dist import Comm
    from pypic_b200.periodic import ExplicitSim, PeriodicImplicitSim
    from pypic_b200.gcstore import GridDev, ParticleStore
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    if world > 1:
        if args.workload != "boris":
            raise SystemExit("bench.py: multi-GPU runs of --workload %s are not wired into bench.py "
                             "(tools/mgpu_check.py covers their sharded parity)" % args.workload)
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    N = int(args.particles_per_gpu); N -= N % 2
    cells = args.cells
    dx, dt = 1e-5, 1e-12
    kT = KB * 116000.
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    kernel_events = []

    def timed_call(fn):
        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = fn(*a, **k); e1.record()
            kernel_events.append((e0, e1))
            return out
        return wrapped
    iters = []
    if args.workload == "explicit":
        L = dx * (cells - 1)
        sim = ExplicitSim(N, cells, dx, dt, (L + dx) * 1e19 / N, q=(-E_CH, E_CH), m=(ME, MP), n_split=N // 2, device=dev,
                          sort_every=args.sort_every, track_order=False)
        sim.x.uniform_(0., 1., generator=gen).mul_(L + dx).clamp_(1e-12, (L + dx) * (1 - 1e-12))
        sim.v.normal_(0., 1., generator=gen)
        sim.v[:N // 2].mul_(float(np.sqrt(kT / ME))); sim.v[N // 2:].mul_(float(np.sqrt(kT / MP)))
        sim.push = timed_call(sim.push)
        step, check, launches = sim.step, sim.check, (lambda: sim.kernel_launches)
        alg, kname = (lambda k: 32.0), "l_push_deposit_v2_k"
        desc = "PIC_L explicit leapfrog: %d particles (e-/p+ halves), %d-cell periodic grid, Poisson solve every step" % (N, cells)
    elif args.workload == "pypic":
        L = dx * cells
        sim = PeriodicImplicitSim(N, cells, dx, dt, L, L * 1e19 / N, tol=1e-3, maxiter=20, device=dev,
                                  sort_every=args.sort_every, track_order=False)
        sim.x0.uniform_(0., 1., generator=gen).mul_(L).clamp_(1e-12, L * (1 - 1e-12))
        sim.v0.normal_(0., 1., generator=gen).mul_(float(np.sqrt(kT / ME)))
        sim.iter_events = kernel_events

        def step():
            k, r = sim.push(); iters.append(k)
        check, launches = sim.check, (lambda: sim.kernel_launches)
        alg, kname = (lambda k: 32.0 + 16.0 / k), "pypic_picard_iter_v2_k"
        desc = "pypic periodic implicit CN/Picard (tol=1e-3): %d electrons, %d-cell grid" % (N, cells)
    else:
        ng = cells + 1
        Te, Ti = 60. * 11600., 50. * 11600.
        lamD = np.sqrt(8.854e-12 * KB * Te / (1e19 * E_CH ** 2))
        Lg = 100. * lamD * (ng - 1) / 149.       # the reference's resolution (150 nodes per 100 Debye lengths)
        alpha = 86. * np.pi / 180.
        grid = GridDev(ng, Lg, Te, device=dev, comm=comm if world > 1 else None)
        store = ParticleStore(N, B=(2. * np.cos(alpha), 2. * np.sin(alpha), 0.), device=dev)
        store.r[0].uniform_(0., 1., generator=gen).mul_(Lg).clamp_(Lg * 1e-9, Lg * (1 - 1e-9))
        vth = float(np.sqrt(KB * Ti / MP))
        for c in (3, 4, 5):
            store.r[c].normal_(0., vth, generator=gen)
        p2c = Lg * 1e19 / (N * world)
        store.charge_state.fill_(1.); store.m.fill_(MP); store.p2c.fill_(p2c); store.Z.fill_(1)
        store.carry_yzt = bool(args.boris_full_store)
        dtg = 1e-10
        store.push_6D = timed_call(store.push_6D)
        state = {"t": 0, "launches": 0}
        grid.weight_particles_to_grid_boltzmann(store, dtg)

        def step():
            if state["t"] % max(1, args.sort_every) == 0:
                store.sort_by_cell(grid); state["launches"] += 16
            if grid.have_fused_n:
                grid.finish_fused_deposit(1.0, dtg)
            grid.smooth_rho(); grid.reset_added_particles()
            grid.solve_for_phi_dirichlet_boltzmann(); grid.differentiate_phi_to_E_dirichlet()
            store.push_6D(dtg, grid, deposit=True)
            state["t"] += 1; state["launches"] += 6

        def check():
            store.check(); grid.check()
        launches = lambda: state["launches"]
        alg, kname = (lambda k: 112.0 if args.boris_full_store else 64.0), "gc_push_boris_v2_k"
        desc = ("pygcpic Boris 1D3V (B=2 T at 86 deg, H+, Ti=50 eV, Te=60 eV): %d particles, %d-node grid, fused "
                "gather+push+walls+deposit, Newton-Boltzmann field solve, store re-sorted every %d steps; %s"
                % (N, ng, max(1, args.sort_every),
                   "full store (x,y,z,v,t streamed: 112 B per particle-step)" if args.boris_full_store else
                   "lean store (x,vx,vy,vz streamed: 64 B per particle-step; y,z not tracked, clocks implicit)"))
    for _ in range(args.warmup):
        step()
    check()
    torch.cuda.synchronize(); comm.barrier()
    kernel_events.clear(); iters.clear()
    sampler = ClockSampler(local); sampler.start()
    l0 = launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize(); comm.barrier()
    sampler.stop_flag = True
    check()
    ms = comm.max_float(ev0.elapsed_time(ev1), device=dev)
    kms = [a.elapsed_time(b) for a, b in kernel_events]
    kbar = float(np.mean(iters)) if iters else 1.0
    peak, peak_src = measured_peak()
    per_launch = N * alg(kbar)
    achieved = per_launch / (float(np.mean(kms)) * 1e-3) / 1e9
    line = {"metric": METRIC, "value": N * world * args.steps / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "parallelism": "single GPU" if world == 1 else
                       "particle decomposition x%d (%d particles per GPU), all-reduce of the deposited density per step" % (world, N),
                       "sort_every": args.sort_every,
                       "picard_iterations_per_step": kbar if iters else None,
                       "l2_policy": "particle arrays (%.1f GB) are far larger than the 126 MB L2" % (N * alg(kbar) / 2 / 1e9)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": kname, "peak_source": peak_src, "kernel_ms_mean": float(np.mean(kms)),
                         "kernel_share_of_step": float(np.sum(kms) / ms), "algorithmic_bytes_per_launch": per_launch},
            "cpu_baseline": None, "e2e": None, "gpu_launches": int(launches() - l0), "clocks": sampler.summary()}
    try:        # DRAM bytes per launch of this workload's kernel from its ncu --set full capture (profiles/traffic.json)
        tk = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["other_kernels"][args.workload]
        line["roofline"]["traffic"] = (tk["dram_bytes_read"] + tk["dram_bytes_write"]) / tk["particles"] * N
    except Exception:
        pass
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


class _StdoutGuard:
    """Keeps the real stdout for the ONE JSON line: anything libraries print to fd 1 meanwhile
    (e.g. NCCL's version banner) goes to stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


def main():
    args = parse()
    with _StdoutGuard():
        if args.impl == "reference":
            line = run_reference(args)
        elif args.workload != "sheath":
            line = run_other(args)
        elif args.decomposition == "slab":
            line = run_slab(args)
        else:
            line = run_cuda(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
