#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): optimised trajectories/s (M2) and constraint+Jacobian
evaluations/s (M1) on the batched 2-D obstacle-avoidance workload (config C2, 65,536 problems per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (scipy SLSQP + its closures)

Prints ONE JSON line (rank 0).  A "step" is one pass of the hot path over one batch: every problem of the
batch is solved from its initial guess (M2); the same batch is then pushed through the evaluation kernel (M1).
Multi-GPU: one process per GPU (torchrun), independent problems sharded by rank (weak scaling: every rank owns a
full batch with its own seed), no collective inside the solve, one NCCL all-gather of the result rows at the end
of each step.
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

# dram__bytes_read.sum + dram__bytes_write.sum of the QP-stage kernel working through 65,536 problems in round 10 of a
# solve, from the committed `ncu --set full` captures (profiles/README.md).  C2 runs as two slices: the captured
# launch covers 32,768 problems and moves 417.2 MB; the figure below is per 65,536 problems like the others.
NCU_TRAFFIC = {"C2": 2 * 417.2e6, "C3": 1693.5e6, "C4": 4035.6e6}      # tg_sqp_qp_kernel (profiles/r01_prof_qp_*.raw.csv)
NCU_TRAFFIC_EVAL = {"C2": 150.9e6}                                     # tg_eval_kernel, 65,536 evaluations
NCU_TRAFFIC_SAMPLE_PER_SAMPLE = {"C2": 486.3e6 / (65536 * 512)}        # tg_sample_kernel, bytes per sample (d = 2)

METRIC = "optimized_trajectories_per_sec"
UNIT = "trajectories/s"
L2_FLUSH_BYTES = 512 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2", help="C2 (default, BASELINE configs[1]) | C3 | C4 | C5a | C5c")
    ap.add_argument("--batch", type=int, default=None, help="problems per GPU (default: the config's full batch)")
    ap.add_argument("--jacobian", default="fd", choices=["analytic", "fd"],
                    help="fd (default): scipy's forward differences emulated on the GPU -- the mode whose converged control points "
                         "match the reference within 1e-5; analytic: closed-form Jacobians (reported next to it as analytic_mode)")
    ap.add_argument("--cpu-sample", type=int, default=None, help="problems in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused", action="store_true", help="one persistent solve kernel instead of lock-step stage kernels")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# CPU path: the reference's scipy SLSQP call on its own closures (oracle/tg_oracle.py restates the Python closures,
# the native steps run in the reference's own C++ compiled unmodified when oracle/_ref exists)
# --------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    name, seed_batch, indices, native_kind = args[:4]
    perturb = len(args) > 4 and args[4]
    import warnings
    warnings.simplefilter("ignore")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import tg_oracle
    from trajectory_generator_b200 import synthetic
    batch = synthetic.make(name, seed_batch)
    out = []
    for i in indices:
        d, cc, kw = synthetic.container_for(batch, i)
        op = tg_oracle.OracleProblem(d, cc, kw.get("objective_function_type", synthetic.OBJECTIVE[name]),
                                     kw.get("num_intervals_free_space"), native_kind=native_kind)
        if perturb:
            # the same reference solve started one unit in the last place away from x0 (its own reproducibility)
            op.x0 = np.nextafter(op.x0, np.inf)
        t = time.perf_counter()
        res = op.solve()
        out.append((i, int(res.status), int(res.nit), time.perf_counter() - t, res.x.tolist()))
    return out


def cpu_path_available():
    ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libTrajectoryConstraints.so"))
    port = os.path.exists(os.path.join(ROOT, "oracle", "_build", "libtg_oracle.so"))
    return "ref" if ref else ("oracle" if port else None)


def run_cpu_sample(name, gen_batch, sample, procs, perturb=False):
    """Solves problems [0, sample) of the synthetic batch with a process pool.  Returns (seconds, results)."""
    import multiprocessing as mp
    kind = cpu_path_available()
    if kind is None:
        raise RuntimeError("oracle libraries are not built (run __graft_entry__.build())")
    chunks = [list(range(sample))[k::procs] for k in range(procs)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("spawn")
    with ctx.Pool(len(chunks)) as pool:
        pool.map(_noop, range(len(chunks)))            # start the workers before the clock
        t = time.perf_counter()
        parts = pool.map(_cpu_worker, [(name, gen_batch, c, kind, perturb) for c in chunks])
        dt = time.perf_counter() - t
    res = sorted(r for p in parts for r in p)
    return dt, res, kind


def _noop(_):
    import scipy.optimize  # noqa: F401  (import cost outside the timed region)
    return 0


def cpu_kind_label(kind):
    # Python closures are the oracle's restatement; with kind 'ref' the native geometry is the reference's own C++
    return "port"


# --------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _claim_stdout():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner on the first collective) goes to stderr;
    the returned file object is the real stdout, used for the ONE JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse()
    out_stream = _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from trajectory_generator_b200 import synthetic
    name = args.config
    B = args.batch or synthetic.FULL_BATCH[name]
    workload = {"C2": "batched 2D obstacle avoidance (test_obstacle_trajectory_2D shape, 8 circular obstacles) x %d problems per GPU",
                "C3": "batched 2D intermediate-waypoint trajectories with curvature + velocity bounds x %d problems per GPU",
                "C4": "batched 3D safe-flight-corridor trajectories (4 corridor boxes) x %d problems per GPU",
                "C5a": "bicycle/unicycle kinematic trajectories, angular-rate + acceleration bounds x %d problems per GPU",
                "C5c": "bicycle/unicycle kinematic trajectories, curvature + acceleration bounds x %d problems per GPU"}[name] % B
    cores = os.cpu_count() or 1

    # ---------------------------------------------------------------- reference arm: CPU path only
    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.cpu_sample or max(64, 8 * cores)
        times = []
        for it in range(args.warmup + args.steps):
            if it < args.warmup and it > 0:
                continue                    # one warm-up pass is enough for a process pool
            dt, res, kind = run_cpu_sample(name, B, sample, cores)
            if it >= args.warmup:
                times.append(dt)
        sec = float(np.mean(times))
        val = sample / sec
        st = [r[1] for r in res]
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "config": name, "sample_per_step": sample,
                           "solver": "scipy SLSQP (2-point finite differences) on the reference's closures"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": cpu_kind_label(kind),
                                 "native": "reference C++ (oracle/_ref)" if kind == "ref" else "plain-C oracle",
                                 "sample": "%d problems of the %s batch per step, process pool of %d" % (sample, name, cores),
                                 "status_histogram": {str(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), file=out_stream, flush=True)
        return

    # ---------------------------------------------------------------- CUDA path
    import torch
    import torch.distributed as dist
    from trajectory_generator_b200 import _native, batch as tgb
    if not torch.cuda.is_available():
        # a driver that is momentarily busy (seen once right after another process exited) answers "initialization
        # failed": wait and start over in a fresh process a few times before giving up
        tries = int(os.environ.get("TG_BENCH_CUDA_RETRY", "0"))
        if tries < 4:
            time.sleep(5 + 5 * tries)
            os.environ["TG_BENCH_CUDA_RETRY"] = str(tries + 1)
            os.dup2(out_stream.fileno(), 1)
            os.execv(sys.executable, [sys.executable] + sys.argv)
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.lib()

    # every rank owns a full batch generated with its own seed (weak scaling)
    gen = {"C2": synthetic.make_c2, "C3": synthetic.make_c3, "C4": synthetic.make_c4}.get(name)
    if gen is not None:
        bt = gen(B, seed=synthetic.SEED0 + int(name[1]) + 1000 * rank)
    else:
        bt = synthetic.make_c5(B, "angular_rate" if name == "C5a" else "curvature", seed=synthetic.SEED0 + 5 + 1000 * rank)
    L = bt.layout
    par_h = torch.from_numpy(bt.par).pin_memory()
    x0_h = torch.from_numpy(bt.x0).pin_memory()
    par = par_h.to(dev)
    x0 = x0_h.to(dev)
    x = torch.empty_like(x0)
    bufs = tgb.SolveBuffers(bt.spec, B, dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    rows = L.n + 4            # result row: x | status | nit | violation | f
    result = torch.empty((B, rows), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * B, rows), dtype=torch.float64, device=dev) if world > 1 else None

    def solve_step():
        out = tgb.solve(bt.spec, par, x, jacobian=args.jacobian, buffers=bufs, fused=args.fused)
        if world > 1:
            result[:, :L.n] = x
            result[:, L.n] = out["status"]; result[:, L.n + 1] = out["nit"]
            result[:, L.n + 2] = out["violation"]; result[:, L.n + 3] = out["f"]
            dist.all_gather_into_tensor(gathered, result)

    def timed(fn, steps, warmup, prepare=None):
        ev = []
        for it in range(warmup + steps):
            if prepare:
                prepare()
            flush.fill_(it & 1)                      # evict L2 between iterations
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            if it >= warmup:
                ev.append((s, e))
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in ev]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.tg_launch_count()
    barrier()
    ms = timed(solve_step, args.steps, args.warmup, prepare=lambda: x.copy_(x0))
    barrier()
    solve_launches = (lib.tg_launch_count() - launches0) * args.steps // (args.steps + args.warmup)
    tot = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_step = tot.item() / args.steps
    value = world * B / (ms_step * 1e-3)
    status = bufs.status.cpu().numpy(); nit = bufs.nit.cpu().numpy()
    x_gpu = x.cpu().numpy()

    # ---- per-kernel share of a step and the dominant kernel's launch duration: one extra solve with CUDA events
    #      around every stage launch on the launching stream (the C library records them; torch events would only
    #      see torch's stream)
    import ctypes
    stats = (ctypes.c_double * 6)()
    lib.tg_set_stage_timing(1)
    x.copy_(x0); flush.fill_(1)
    tgb.solve(bt.spec, par, x, jacobian=args.jacobian, buffers=bufs, fused=args.fused)
    torch.cuda.synchronize()
    lib.tg_set_stage_timing(0)
    lib.tg_last_solve_stats(stats, 6)
    ms_ls, ms_qp, flops_qp, n_ls, n_qp, rounds = [float(v) for v in stats]
    fp64_peak = ctypes.c_double(0.0)
    _native.check(lib.tg_measure_fp64_peak(ctypes.byref(fp64_peak)), "tg_measure_fp64_peak")

    # ---- the other Jacobian mode on the same batch (2 timed steps), reported next to the headline
    other = "analytic" if args.jacobian == "fd" else "fd"
    def other_step():
        tgb.solve(bt.spec, par, x, jacobian=other, buffers=bufs, fused=args.fused)
    barrier()
    ms_o = timed(other_step, 2, 1, prepare=lambda: x.copy_(x0))
    barrier()
    toto = torch.tensor([sum(ms_o)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(toto, op=dist.ReduceOp.MAX)
    other_ms = toto.item() / 2
    other_status = bufs.status.cpu().numpy(); other_nit = bufs.nit.cpu().numpy(); other_x = x.cpu().numpy()

    # ---- M1: evaluation kernel on the same batch
    xe = torch.from_numpy(synthetic.evaluation_points(bt)).to(dev)
    ev_out = {}
    barrier()
    ms_e = timed(lambda: tgb.evaluate(bt.spec, par, xe, out=ev_out), args.steps, args.warmup)
    barrier()
    tote = torch.tensor([sum(ms_e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tote, op=dist.ReduceOp.MAX)
    ms_eval = tote.item() / args.steps
    eval_bytes = 8 * (L.n + L.P + L.m + L.m_nl * L.n + 1 + L.n)

    # ---- f1: output sampling of the solved batch (positions, 2048 samples per trajectory) -- a pure HBM-write stream
    from trajectory_generator_b200 import matrix_evaluation as tgs
    SAMPLES = 2048
    samp = torch.empty((B, L.d, SAMPLES), dtype=torch.float64, device=dev)
    barrier()
    ms_s = timed(lambda: tgs.sample_batch((x, L.d, L.N), num_points=SAMPLES, out=samp), args.steps, args.warmup)
    barrier()
    tots = torch.tensor([sum(ms_s)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tots, op=dist.ReduceOp.MAX)
    ms_samp = tots.item() / args.steps
    samp_bytes = 8 * L.d * SAMPLES * B + 8 * (L.d * L.N) * B

    # ---- e2e: host buffers through the C-ABI (tg_solve_host / tg_eval_host), copies inside the timed call
    x_host = bt.x0.copy()
    e2e_times = []
    for it in range(1 + min(args.steps, 3)):
        x_host[:] = bt.x0
        barrier()
        t = time.perf_counter()
        oh = tgb.solve_host(bt.spec, bt.par, x_host, jacobian=args.jacobian, fused=args.fused)
        barrier()
        if it > 0:
            e2e_times.append(time.perf_counter() - t)
    e2e_t = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * B / e2e_t.item()
    h2d = 8 * B * (L.P + L.n)
    d2h = 8 * B * (L.n + 1) + 4 * B * 3
    te = []
    for it in range(3):
        t = time.perf_counter()
        tgb.evaluate_host(bt.spec, bt.par, synthetic.evaluation_points(bt) if it == 0 else xe_h)
        if it == 0:
            xe_h = xe.cpu().numpy()
        else:
            te.append(time.perf_counter() - t)
    eval_e2e = world * B / float(np.mean(te))
    # ---- f3: problems of different shapes in ONE call (tg_solve_mixed_host: the buckets' solves overlap on the
    #      device, each on its own stream) against bucket-by-bucket tg_solve_host calls; host buffers, end to end
    mixed = None
    if world == 1:
        MB = 4096
        gens = (("C2", synthetic.make_c2, 2), ("C3", synthetic.make_c3, 3), ("C4", synthetic.make_c4, 4))
        mbs = [g(MB, seed=synthetic.SEED0 + k + 100 * r) for _, g, k in gens for r in (1, 2)]
        mbs += [synthetic.make_c5(MB, kind, seed=synthetic.SEED0 + 5 + 100 * r) for kind in ("angular_rate", "curvature") for r in (1,)]
        buckets = [(b.spec, b.par, b.x0) for b in mbs]
        runs = {"bucket_by_bucket": lambda: [tgb.solve_host(sp_, p_, x_, jacobian=args.jacobian) for sp_, p_, x_ in buckets],
                "one_call": lambda: tgb.solve_mixed_host(buckets, jacobian=args.jacobian)}
        tm, res_m = {}, {}
        for label, fn in runs.items():
            fn()
            ts = []
            for _ in range(2):
                t = time.perf_counter(); res_m[label] = fn(); ts.append(time.perf_counter() - t)
            tm[label] = min(ts)
        same = all(np.array_equal(a["x"], b["x"]) and np.array_equal(a["status"], b["status"])
                   for a, b in zip(res_m["bucket_by_bucket"], res_m["one_call"]))
        mixed = {"workload": "%d buckets of %d problems: C2, C3, C4 (two seeds each), C5a, C5c" % (len(buckets), MB),
                 "problems": MB * len(buckets), "unit": UNIT, "api": "tg_solve_mixed_host (C-ABI, host buffers)",
                 "one_call": MB * len(buckets) / tm["one_call"], "bucket_by_bucket": MB * len(buckets) / tm["bucket_by_bucket"],
                 "identical_results": bool(same)}
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # per-GPU figures for the rooflines (one launch = one batch on one GPU)
    solve_bytes = 8 * (L.P + 2 * L.n + 4) * B
    mean_nit = float(nit.mean())
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload, "config": name, "problems_per_gpu": B, "n": L.n, "m": L.m, "meq": L.meq,
                       "maxiter": 100, "ftol": 1e-6, "jacobian": args.jacobian, "schedule": "fused persistent kernel" if args.fused else "lock-step stage kernels", "l2": "flushed between iterations (512 MiB fill)",
                       "multi_gpu": "independent problems sharded by rank, one NCCL all-gather of result rows per step"},
            "solve_stats": {"mean_nit": mean_nit, "max_nit": int(nit.max()),
                            "status_histogram": {str(k): int(v) for k, v in zip(*np.unique(status, return_counts=True))}},
            # dominant kernel of a step: the QP-stage kernel (share below).  It works out of shared memory; its bound is
            # the FP64 pipe, so `achieved` is algorithmic fp64 operations (model count accumulated by the kernel per
            # problem: 2 per multiply-add of the factor updates, products and scans it performs) / its launch time.
            "roofline": ({"kernel": "tg_sqp_qp_kernel", "bound": "fp64", "achieved": flops_qp / (ms_qp * 1e-3) / 1e12,
                          "peak": fp64_peak.value, "unit": "TFLOP/s", "frac": flops_qp / (ms_qp * 1e-3) / 1e12 / fp64_peak.value,
                          "traffic": NCU_TRAFFIC.get(name) if B == 65536 or name != "C2" else None,
                          "traffic_note": "dram bytes read + written by the kernel over 65,536 problems (round 10 of a solve) from the "
                                          "committed ncu --set full capture; the state lives in HBM between stage kernels",
                          "launches_per_step": int(n_qp), "avg_launch_ms": ms_qp / max(n_qp, 1.0),
                          "share_of_step": ms_qp / (ms_ls + ms_qp), "line_search_kernel_ms": ms_ls, "qp_kernel_ms": ms_qp,
                          "algorithmic_flops_per_trajectory": flops_qp / B,
                          "peak_source": "measured in this run: DFMA kernel, 8 chains per thread, every SM full "
                                         "(MEASURED_PEAKS.json has no fp64 figure; nominal 148 x 64 x 2 x 1.965 GHz = 37.2)",
                          "hbm": {"achieved": solve_bytes / (ms_step * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                  "frac": solve_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                                  "note": "algorithmic HBM bytes per trajectory are 8(P + 2n + 4): not the bound"}}
                         if not args.fused and ms_qp > 0 else
                         {"kernel": "tg_solve_kernel", "bound": "hbm", "achieved": solve_bytes / (ms_step * 1e-3) / 1e9,
                          "peak": hbm_peak, "unit": "GB/s", "frac": solve_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak,
                          "traffic": None, "peak_source": peak_src}),
            "evals": {"metric": "constraint_jacobian_evaluations_per_sec", "value": world * B / (ms_eval * 1e-3),
                      "unit": "evaluations/s", "ms_per_step": ms_eval, "bytes_per_eval": eval_bytes,
                      "e2e": {"value": eval_e2e, "unit": "evaluations/s", "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": 8 * B * (1 + L.n + L.m + L.m_nl * L.n)},
                      "roofline": {"kernel": "tg_eval_kernel", "bound": "hbm",
                                   "achieved": eval_bytes * B / (ms_eval * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                   "frac": eval_bytes * B / (ms_eval * 1e-3) / 1e9 / hbm_peak,
                                   "traffic": NCU_TRAFFIC_EVAL.get(name) if B == 65536 else None,
                                   "peak_source": peak_src}},
            "sampling": {"metric": "trajectory_samples_per_sec", "value": world * B * SAMPLES / (ms_samp * 1e-3),
                         "unit": "samples/s", "ms_per_step": ms_samp, "samples_per_trajectory": SAMPLES,
                         "roofline": {"kernel": "tg_sample_kernel", "bound": "hbm",
                                      "achieved": samp_bytes / (ms_samp * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": samp_bytes / (ms_samp * 1e-3) / 1e9 / hbm_peak,
                                      "traffic": (NCU_TRAFFIC_SAMPLE_PER_SAMPLE[name] * B * SAMPLES
                                                  if name in NCU_TRAFFIC_SAMPLE_PER_SAMPLE else None),
                                      "bytes_per_sample": 8 * L.d, "peak_source": peak_src}},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "tg_solve_host (C-ABI, host buffers; copies inside the call)"},
            other + "_mode": {"value": world * B / (other_ms * 1e-3), "unit": UNIT, "ms_per_step": other_ms, "steps": 2,
                              "mean_nit": float(other_nit.mean()),
                              "status_histogram": {str(k): int(v) for k, v in zip(*np.unique(other_status, return_counts=True))}},
            "gpu_launches": int(solve_launches),
            "clocks": sampler.summary()}
    if mixed is not None:
        line["mixed_shapes"] = mixed

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only): bounded sample of the same problems
    if world == 1 and not args.no_cpu_baseline and cpu_path_available():
        sample = args.cpu_sample or max(128, 16 * cores)
        try:
            dt, res, kind = run_cpu_sample(name, B, sample, cores)
            st_ref = np.array([r[1] for r in res]); x_ref = np.array([r[4] for r in res])
            k = L.ia + 1
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": cpu_kind_label(kind),
                                    "native": "reference C++ (oracle/_ref)" if kind == "ref" else "plain-C oracle",
                                    "sample": "first %d problems of the same batch, scipy SLSQP with 2-point finite differences, process pool of %d" % (sample, cores),
                                    "seconds": dt,
                                    "status_histogram": {str(a): int(b) for a, b in zip(*np.unique(st_ref, return_counts=True))}}
            # how reproducible the reference is against ITSELF: the same scipy solves started from x0 + 1 ulp.  Its
            # forward differences (h = 1.5e-8) amplify last-place differences of the closures by 1/h, so that long
            # solves separate; agreement of the CUDA path is therefore also reported on the subset of problems whose
            # reference solution is stable to 1e-5 under that perturbation.
            _, res2, _ = run_cpu_sample(name, B, sample, cores, perturb=True)
            st2 = np.array([r[1] for r in res2]); x2 = np.array([r[4] for r in res2])
            stable = (st_ref == 0) & (st2 == 0) & (np.abs(x2[:, :k] - x_ref[:, :k]).max(1) <= 1e-5)
            def agreement(xg, sg, mode):
                dcp = np.abs(xg[:, :k] - x_ref[:, :k]).max(1)
                both = (st_ref == 0) & (sg == 0)
                return {"jacobian": mode, "problems": int(sample), "reference_status0": int((st_ref == 0).sum()),
                        "both_status0": int(both.sum()), "same_success_flag": int(((st_ref == 0) == (sg == 0)).sum()),
                        "same_status": int((st_ref == sg).sum()),
                        "status0_within_1e-5": int((dcp[both] <= 1e-5).sum()),
                        "status0_within_1e-3": int((dcp[both] <= 1e-3).sum()),
                        "reference_stable_problems": int(stable.sum()),
                        "reference_stable_within_1e-5": int((dcp[stable & (sg == 0)] <= 1e-5).sum())}
            # the timed mode (first) and the other one
            line["parity_sample"] = [agreement(x_gpu[:sample], status[:sample], args.jacobian),
                                     agreement(other_x[:sample], other_status[:sample], other)]
            line["reference_self_consistency"] = {
                "perturbation": "x0 + 1 ulp", "problems": int(sample), "both_status0": int(((st_ref == 0) & (st2 == 0)).sum()),
                "same_status": int((st_ref == st2).sum()),
                "status0_within_1e-5": int(stable.sum()),
                "status0_within_1e-3": int(((st_ref == 0) & (st2 == 0) & (np.abs(x2[:, :k] - x_ref[:, :k]).max(1) <= 1e-3)).sum())}
        except Exception as exc:      # the baseline is reported, never required
            line["cpu_baseline"] = {"error": repr(exc)}
    print(json.dumps(line), file=out_stream, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
