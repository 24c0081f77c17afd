#!/usr/bin/env python
"""bench.py -- body-substeps/sec of the XPBD frame step on the north-star workload (BASELINE.json):
4096 independent worlds x (256 cubes + floor) per GPU, dt = 1/60, 20 substeps, 1 positional iteration, collisions on.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--worlds 4096]

One "step" = one 60 Hz frame (= 20 substeps) of every world. The workload's window is frames 0..59 from the initial
poses (nothing can fall asleep before 1.0 s of simulated quiet time, pbd.cpp:490-498, so it is not deflated
by sleeping). Work per frame GROWS through the window (the cubes land on each other one after the other), so for K < 60
the timed steps are the LAST K frames of the window (frames 60-K..59, the heavier end; the frames before them are
stepped untimed after the W warm-up steps and a reset) -- a short run never reports a lighter workload than the default
K = 60, which times the whole window; K > 60 runs on past it. `value` is timed on the device with CUDA events on the
stream the kernels are launched on, with state resident in HBM; `e2e` is the same metric through the host-buffer call
rp_batch_step_host (pinned host state in, pinned host state out, both copies inside the timed region).
`--impl reference` times the UNMODIFIED reference (oracle/_ref, else the CPU restatement) on all host cores.
Prints ONE JSON line on rank 0.
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DT = 1.0 / 60.0
SUBSTEPS = 20
ITERS = 1
WINDOW = 60  # frames of the workload's window


def lead_in(steps):
    """untimed frames stepped from the initial poses before the timed ones, so that K < WINDOW times frames WINDOW-K..WINDOW-1"""
    return max(0, WINDOW - steps)


def window_text(steps):
    return "frames %d..%d from the initial poses" % (lead_in(steps), lead_in(steps) + steps - 1)

# Algorithmic work per item, derived in DESIGN.md ("Kernels and rooflines"): bytes that must move and FP64 operations
# (mul/add/div/sqrt = 1 each, no FMA credit) for one body-substep / pair test / EPA+manifold run / solved contact.
BYTES = dict(integrate=288.0, gjk=400.0, epa=16.0 + 96.0 + 384.0 + 40.0, manifold=40.0 + 16.0 + 672.0 + 2 * 56.0, manifold_contact=64.0, solve_pos_pair=2 * (112.0 + 56.0), solve_contact=80.0,
             solve_vel_pair=2 * (128.0 + 48.0), solve_contact_vel=64.0)
FLOPS = dict(integrate=590.0, gjk=620.0, epa=800.0, manifold=1200.0, solve_pos_contact=1100.0, solve_vel_contact=650.0)
FLOP_PER_BODY_SUBSTEP = 4.0e3  # algorithmic FP64 flop per body-substep on the W256 world


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds", type=int, default=4096, help="worlds per GPU")
    ap.add_argument("--no-cull", action="store_true", help="run GJK on every broadphase pair (disables the exact-safe bounds cull)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / profile / cpu baseline (kernel timing only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (tuning runs)")
    ap.add_argument("--hetero", action="store_true",
                    help="also time the same window with every world started from a DIFFERENT pose set (random yaw and lateral "
                         "offset per cube): reported next to the headline as `heterogeneous`")
    ap.add_argument("--e2e-parts", type=int, default=2, help="sub-batches (host threads) of the pipelined end-to-end leg")
    ap.add_argument("--ncu-frame", type=int, default=-1,
                    help="profiling aid: run this many frames, then bracket ONE more frame with cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints nothing")
    return ap.parse_args()


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu --set full capture
    (profiles/ncu_traffic.json, written by profiles/summarise_ncu.py from the capture named inside it), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    for name, d in json.load(open(p)).get("kernels", {}).items():
        if name.split("<")[0] == kernel:  # template instances are listed as k_solve_pos<0>
            return d["dram_read"] + d["dram_write"]
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_single_thread_baseline(flavour, budget_s=12.0):
    """The reference's own single-threaded step on ONE host core: fresh W256 worlds, first 60 frames each."""
    import refdrv
    import scenes
    desc = scenes.w256()
    frames, worlds, spent = 60, 0, 0.0
    while spent < budget_s and worlds < 16:
        w = refdrv.RefWorld(flavour).load(desc)
        spent += w.run_timed(frames, DT, SUBSTEPS, ITERS, True)
        worlds += 1
    nb = len(desc.bodies)
    return {"value": nb * SUBSTEPS * frames * worlds / spent, "unit": "body-substeps/s", "cores": 1,
            "kind": "reference" if flavour == "strict" else "port",
            "sample": "%d fresh W256 worlds x first %d frames, one thread, %.1f s (%.2f ms/frame/world)" % (worlds, frames, spent, 1e3 * spent / (frames * worlds))}


def _ref_worker(flavour, frames_warm, frames, barrier, q):
    import refdrv
    import scenes
    w = refdrv.RefWorld(flavour).load(scenes.w256())
    barrier.wait()
    if frames_warm:
        w.run_timed(frames_warm, DT, SUBSTEPS, ITERS, True)
        w = refdrv.RefWorld(flavour).load(scenes.w256())  # same timed frames as the GPU arm
    if lead_in(frames):
        w.run_timed(lead_in(frames), DT, SUBSTEPS, ITERS, True)
    barrier.wait()
    t0 = time.perf_counter()
    w.run_timed(frames, DT, SUBSTEPS, ITERS, True)
    q.put(time.perf_counter() - t0)


def run_reference(args):
    """Reference arm: the unmodified reference (one world per process, it is single-threaded with global state) on every
    host core at once; a step = one frame of `cores` worlds."""
    import multiprocessing as mp
    import refdrv
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    flavour = "strict" if refdrv.available("strict") else "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores)
    q = ctx.Queue()
    procs = [ctx.Process(target=_ref_worker, args=(flavour, args.warmup, args.steps, barrier, q)) for _ in range(cores)]
    for p in procs:
        p.start()
    times = [q.get() for _ in procs]
    for p in procs:
        p.join()
    elapsed = max(times)
    nb = 257
    value = nb * cores * SUBSTEPS * args.steps / elapsed
    line = {"impl": "reference", "metric": "body-substeps/sec", "value": value, "unit": "body-substeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "w256 (256 cubes + floor), %d worlds (one per host core), %s" % (cores, window_text(args.steps)),
                       "bodies_per_world": nb, "substeps": SUBSTEPS, "pos_iters": ITERS, "dt": DT},
            "cpu_baseline": {"value": value, "unit": "body-substeps/s", "cores": cores, "kind": "reference" if flavour == "strict" else "port",
                             "sample": "%d processes x 1 W256 world x %d frames" % (cores, args.steps)},
            "e2e": {"value": value, "unit": "body-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def hetero_state(init, W, NB, stride, rank):
    """Every cube of every world gets its own yaw in +-0.3 rad and a lateral offset in +-0.2: contacts form at different
    times and with different manifolds in every world (no cross-world coherence for the lane = world work order)."""
    rng = np.random.default_rng(1234 + rank)
    st = np.ascontiguousarray(np.broadcast_to(init[None], (W, NB, stride))).copy()
    yaw = rng.uniform(-0.3, 0.3, size=(W, NB - 1))
    st[:, 1:, 0] += rng.uniform(-0.2, 0.2, size=(W, NB - 1))
    st[:, 1:, 2] += rng.uniform(-0.2, 0.2, size=(W, NB - 1))
    st[:, 1:, 3] = 0.0
    st[:, 1:, 4] = np.sin(0.5 * yaw)
    st[:, 1:, 5] = 0.0
    st[:, 1:, 6] = np.cos(0.5 * yaw)
    return st


def run_ours(args):
    import torch
    import __graft_entry__ as ge
    import scenes
    pkg = ge.load_package()
    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or pkg.lib().rp_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world_size > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from rawphys_b200 import multi
    job = multi.Job(dist, "cuda")  # worlds are sharded over ranks with no data-path collective; NCCL only aggregates

    desc = scenes.w256()
    scene = pkg.Scene(desc)
    W = args.worlds
    batch = pkg.Batch(scene, n_worlds=W, device=local_rank, disable_cull=args.no_cull)
    batch.set_scene_forces(desc)
    NB = batch.NB
    init = batch.state(0, 1)[0].copy()

    def barrier():
        torch.cuda.synchronize()
        batch.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    max_over_ranks = job.max_over_ranks

    if args.ncu_frame >= 0:
        if args.hetero:
            batch.upload(hetero_state(init, W, NB, pkg.STATE_STRIDE, rank))
        batch.run(max(args.ncu_frame, 1), DT, SUBSTEPS, ITERS, True)
        barrier()
        torch.cuda.profiler.start()
        batch.run(1, DT, SUBSTEPS, ITERS, True)
        barrier()
        torch.cuda.profiler.stop()
        batch.close()
        return

    # ---- device-resident throughput
    for _ in range(args.warmup):
        batch.step(DT, SUBSTEPS, ITERS, True)

    def rewind():
        """every world back to the initial poses, then the untimed lead-in frames"""
        batch.broadcast(init)
        for _ in range(lead_in(args.steps)):
            batch.step(DT, SUBSTEPS, ITERS, True)

    rewind()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms = batch.run(args.steps, DT, SUBSTEPS, ITERS, True)
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    units = NB * W * SUBSTEPS * args.steps * world_size
    value = units / (ms * 1e-3)
    status = batch.status()
    # kernels launched in the timed region: per frame 7 prologue kernels + per substep reset, integrate, cull, transform,
    # gjk, epa, manifold, ONE positional and ONE velocity sweep (cooperative grids that walk the dependency levels with grid
    # barriers), + the end-of-frame derive and frame counter
    launches_total = args.steps * (7 + SUBSTEPS * (7 + 2) + 2)

    line = {"metric": "body-substeps/sec", "value": value, "unit": "body-substeps/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "w256x%d per GPU (256 cubes + floor per world), %s" % (W, window_text(args.steps)),
                       "worlds_per_gpu": W, "bodies_per_world": NB, "substeps": SUBSTEPS, "pos_iters": ITERS, "dt": DT, "mode": "batched-worlds, reference Gauss-Seidel order (level schedule)",
                       "l2": "per-GPU state %.0f MB + transformed hulls %.0f MB > 126 MB L2: inputs larger than L2, no flush" % (
                           W * NB * 208 / 1e6, W * NB * 336 / 1e6)},
            "clocks": clocks, "gpu_launches": launches_total, "status_bits": int(np.bitwise_or.reduce(status))}

    if not args.no_extras:
        # ---- end to end through the host-buffer call (pinned host state in and out every step)
        nrec = W * NB * pkg.STATE_STRIDE
        h_in = torch.empty(nrec, dtype=torch.float64).pin_memory()
        h_out = torch.empty(nrec, dtype=torch.float64).pin_memory()
        h_in.copy_(torch.from_numpy(np.ascontiguousarray(np.broadcast_to(init[None], (W, NB, pkg.STATE_STRIDE))).reshape(-1)))
        for _ in range(min(args.warmup, 2)):
            batch.step_host(h_in.data_ptr(), h_out.data_ptr(), DT, SUBSTEPS, ITERS, True)
        rewind()
        h_in.copy_(torch.from_numpy(batch.state().reshape(-1)))  # the state the timed frames start from, in host memory
        barrier()
        t0 = time.perf_counter()
        a, b = h_in, h_out
        for _ in range(args.steps):
            batch.step_host(a.data_ptr(), b.data_ptr(), DT, SUBSTEPS, ITERS, True)
            a, b = b, a
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        line["e2e"] = {"value": units / e2e_s, "unit": "body-substeps/s", "h2d_bytes_per_step": nrec * 8, "d2h_bytes_per_step": nrec * 8,
                       "ms_per_step": 1e3 * e2e_s / args.steps, "api": "rp_batch_step_host (upload state, step, download state, sync)"}

        # ---- the same end-to-end loop with the worlds split over P sub-batches, each driven by its own host thread on its
        # own stream: one sub-batch's PCIe copies overlap the others' kernels. Same public call
        # (rp_batch_step_host), same bytes per step, same worlds; reported as `e2e` when it is the faster of the two.
        P = args.e2e_parts
        if P > 1 and W % P == 0 and W // P >= 32:
            rewind()
            start_state = torch.from_numpy(batch.state().reshape(-1))
            halves = [pkg.Batch(scene, n_worlds=W // P, device=local_rank, disable_cull=args.no_cull) for _ in range(P)]
            for hb in halves:
                hb.set_scene_forces(desc)
            part = nrec // P
            h_in.copy_(start_state)
            bufs = [(h_in[i * part:(i + 1) * part], h_out[i * part:(i + 1) * part]) for i in range(P)]
            for hb, (a, b) in zip(halves, bufs):
                hb.step_host(a.data_ptr(), b.data_ptr(), DT, SUBSTEPS, ITERS, True)  # warm-up (graph capture)
            h_in.copy_(start_state)
            gate = threading.Barrier(P + 1)

            def drive(hb, a, b):
                gate.wait()
                for _ in range(args.steps):
                    hb.step_host(a.data_ptr(), b.data_ptr(), DT, SUBSTEPS, ITERS, True)
                    a, b = b, a

            threads = [threading.Thread(target=drive, args=(hb, a, b)) for hb, (a, b) in zip(halves, bufs)]
            for t in threads:
                t.start()
            barrier()
            gate.wait()
            t0 = time.perf_counter()
            for t in threads:
                t.join()
            barrier()
            piped_s = max_over_ranks(time.perf_counter() - t0)
            bits = int(np.bitwise_or.reduce(np.concatenate([hb.status() for hb in halves])))
            for hb in halves:
                hb.close()
            single = dict(line["e2e"])
            piped = {"value": units / piped_s, "unit": "body-substeps/s", "h2d_bytes_per_step": nrec * 8, "d2h_bytes_per_step": nrec * 8,
                     "ms_per_step": 1e3 * piped_s / args.steps, "status_bits": bits,
                     "api": "rp_batch_step_host on %d sub-batches of %d worlds from %d host threads (the copies of one overlap the kernels of the others)" % (P, W // P, P)}
            if piped["value"] > single["value"] and bits == 0:
                line["e2e"] = piped
                line["e2e_single_batch"] = single
            else:
                line["e2e_two_half_batches"] = piped

        # ---- per-kernel device time over the same window + roofline of the dominant kernel
        rewind()
        barrier()
        c0 = batch.counters()
        fam = batch.profile(args.steps, DT, SUBSTEPS, ITERS, True)
        c1 = batch.counters()
        tests = c1["pair_tests"] - c0["pair_tests"]
        hits = c1["gjk_hits"] - c0["gjk_hits"]
        contacts = c1["contacts"] - c0["contacts"]
        bs = NB * W * SUBSTEPS * args.steps
        alg_bytes = {"integrate": BYTES["integrate"] * bs, "gjk": BYTES["gjk"] * tests,
                     "epa": BYTES["epa"] * hits,
                     "manifold": BYTES["manifold"] * hits + BYTES["manifold_contact"] * contacts,
                     "solve_pos": BYTES["solve_pos_pair"] * hits + BYTES["solve_contact"] * contacts,
                     "solve_vel": BYTES["solve_vel_pair"] * hits + BYTES["solve_contact_vel"] * contacts}
        alg_flops = {"integrate": FLOPS["integrate"] * bs, "gjk": FLOPS["gjk"] * tests, "epa": FLOPS["epa"] * hits, "manifold": FLOPS["manifold"] * hits,
                     "solve_pos": FLOPS["solve_pos_contact"] * contacts, "solve_vel": FLOPS["solve_vel_contact"] * contacts}
        total_ms = sum(fam.values())
        top = max(fam, key=fam.get)
        hbm_peak, hbm_src = measured_peaks()
        fp64_nofma, fp64_fma = pkg.measure_fp64_peak(local_rank)
        launches = SUBSTEPS * args.steps if top in alg_bytes else args.steps
        line["kernels"] = {k: {"ms": round(v, 3), "share": round(v / total_ms, 4)} for k, v in fam.items()}
        if top in alg_bytes:
            gbs = alg_bytes[top] / (fam[top] * 1e-3) / 1e9
            line["roofline"] = {"kernel": "k_" + top, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                "traffic": ncu_traffic("k_" + top), "peak_source": hbm_src, "avg_launch_ms": fam[top] / launches,
                                "note": "schema bound; the binding resource of this path is the FP64 CUDA-core pipe, see fp64"}
            tf = alg_flops[top] / (fam[top] * 1e-3) / 1e12
            line["fp64"] = {"kernel": "k_" + top, "achieved_tflops": tf, "peak_tflops_no_fma": fp64_nofma, "peak_tflops_fma": fp64_fma,
                            "frac_of_no_fma_peak": tf / fp64_nofma,
                            "whole_step_tflops": FLOP_PER_BODY_SUBSTEP * (NB * W * SUBSTEPS * args.steps) / (ms * 1e-3) / 1e12,
                            "whole_step_frac": FLOP_PER_BODY_SUBSTEP * (NB * W * SUBSTEPS * args.steps) / (ms * 1e-3) / 1e12 / fp64_nofma,
                            "peak_source": "rp_measure_fp64_peak (DMUL+DADD chains, this run)"}
        line["work_per_world_substep"] = {"pair_tests": tests / (W * SUBSTEPS * args.steps), "epa_runs": hits / (W * SUBSTEPS * args.steps),
                                          "contacts": contacts / (W * SUBSTEPS * args.steps)}
        if dist is not None:  # NCCL only gathers aggregate statistics
            t = job.sum_over_ranks([tests, hits, contacts])
            line["aggregate_work"] = {"pair_tests": float(t[0]), "epa_runs": float(t[1]), "contacts": float(t[2])}
        if args.hetero:
            # The headline workload is BASELINE.json's: 4096 COPIES of one scene, so the lanes of a warp (same pair index,
            # neighbouring worlds) follow the same control flow. This leg breaks that (hetero_state). Same kernels, same window.
            st = hetero_state(init, W, NB, pkg.STATE_STRIDE, rank)
            batch.upload(st)
            for _ in range(lead_in(args.steps)):
                batch.step(DT, SUBSTEPS, ITERS, True)
            barrier()
            c0 = batch.counters()
            hms = max_over_ranks(batch.run(args.steps, DT, SUBSTEPS, ITERS, True))
            barrier()
            c1 = batch.counters()
            batch.upload(st)
            for _ in range(lead_in(args.steps)):
                batch.step(DT, SUBSTEPS, ITERS, True)
            hfam = batch.profile(args.steps, DT, SUBSTEPS, ITERS, True)
            line["heterogeneous"] = {"value": units / (hms * 1e-3), "unit": "body-substeps/s", "ms_per_step": hms / args.steps,
                                     "status_bits": int(np.bitwise_or.reduce(batch.status())),
                                     "kernels_ms": {k: round(v, 1) for k, v in hfam.items()},
                                     "contacts_per_world_substep": (c1["contacts"] - c0["contacts"]) / (W * SUBSTEPS * args.steps),
                                     "note": "every world started from its own poses (per-cube yaw +-0.3 rad, offset +-0.2): no cross-world coherence"}
        if rank == 0 and world_size == 1 and not args.no_cpu:
            import refdrv
            line["cpu_baseline"] = cpu_single_thread_baseline("strict" if refdrv.available("strict") else "port")

    if rank == 0:
        print(json.dumps(line), flush=True)
    batch.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    # Exactly ONE line reaches stdout: libraries that write to file descriptor 1 on their own (NCCL prints its version
    # there) are sent to stderr for the length of the run; the JSON line is printed through the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
