#!/usr/bin/env python
"""bench.py -- constraint+Jacobian evaluations per second of the batched evaluator.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one batched evaluation (constraint values g and Jacobian values, fused in one kernel
launch) of BASELINE.json configs[1]: 65,536 synthetic 4-contact flat-ground instances per GPU.
Instances shard by index across ranks with no collective on the data path (weak scaling: every
rank evaluates its own 65,536-instance shard).  Prints ONE JSON line (rank 0).

  value     instances/s, inputs resident in HBM, K back-to-back launches bracketed by CUDA events
            on the launch stream; successive steps rotate through buffer sets whose total size
            is >> L2, so every step reads and writes HBM, not L2.
  e2e       the same metric through the host-buffer API: HOST (pinned) buffers in, host buffers out, every
            step's H2D and D2H copies inside the timed region; a queue of batches (cplb_eval_host_begin /
            _wait), with the one-batch-at-a-time cplb_eval_host figure beside it.
  roofline  algorithmic bytes per launch (8*(n+m+nnz) per instance) / average launch
            duration over the timed region, against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline  the CPU oracle (port of the reference's evaluation) timed on this box's host cores.

--impl reference times the reference's own CPU evaluation (oracle/_ref when it was built from the
reference sources, else the oracle port) on all host cores, same config/metric.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PER_GPU = 65536
NC = 4
WORKLOAD = "configs[1]: 65,536 synthetic 4-contact flat-ground instances per GPU, fused g+Jacobian eval"
L2_BYTES = 126 * 1024 * 1024


def algorithmic_bytes_per_instance(n, m, nnz):
    return 8 * (n + m + nnz)  # read x[n], write g[m] and jac[nnz], fp64


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed regions run (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(rank):
    from centroidalplanner_b200 import synthetic

    return synthetic.ground_batch(N_PER_GPU, NC, seed=1002 + 7919 * rank)


def configure(problem_like, env_like):
    from centroidalplanner_b200 import synthetic

    tb = synthetic.TESTBASIC
    env_like.SetGroundZ(tb["ground_z"])
    env_like.SetMu(tb["mu"])
    synthetic.configure_testbasic(problem_like, synthetic.NAMES4)


def oracle_problem():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import OracleProblem  # the oracle with the product's setter names

    op = OracleProblem(__import__("centroidalplanner_b200").synthetic.NAMES4, "ground", 100.0)
    configure(op, op)
    return op.o


def time_cpu(o, x, threads, budget_s=10.0, want=("g", "jac")):
    """Bounded CPU sample: repeat the 65,536-instance batch until ~budget_s of wall time is spent."""
    out = {"g": np.empty((x.shape[0], o.m)), "jac": np.empty((x.shape[0], o.nnz))}
    o.eval_batch(x[:4096], want=want, nthreads=threads)  # warm-up
    t0 = time.perf_counter()
    o.eval_batch(x, want=want, nthreads=threads, out=out)
    one = time.perf_counter() - t0
    reps = max(1, min(50, int(budget_s / max(one, 1e-6))))
    best = one
    for _ in range(reps):
        t0 = time.perf_counter()
        o.eval_batch(x, want=want, nthreads=threads, out=out)
        best = min(best, time.perf_counter() - t0)
    return x.shape[0] / best, reps + 1


def run_reference(args):
    """--impl reference: the reference's CPU evaluation on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    x = make_inputs(0)
    kind, sample = "port", ""
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libcpl_ref.so")
    per_step = []
    if os.path.exists(ref_lib):
        from oracle import cpl_ref_py

        r = cpl_ref_py.RefProblem(__import__("centroidalplanner_b200").synthetic.NAMES4, "ground", 100.0)
        configure(r, r)
        kind = "reference"
        n_s = min(N_PER_GPU, 8192)
        fn = lambda: r.eval_batch(x[:n_s], nthreads=cores)  # noqa: E731
        sample = (f"{n_s} of the 65,536 instances per step through the reference's own ifopt components "
                  f"(oracle/_ref: reference sources compiled against stand-in Eigen/ifopt headers), {cores} threads")
    else:
        o = oracle_problem()
        n_s = N_PER_GPU
        buf = {"g": np.empty((n_s, o.m)), "jac": np.empty((n_s, o.nnz))}
        fn = lambda: o.eval_batch(x, want=("g", "jac"), nthreads=cores, out=buf)  # noqa: E731
        sample = f"all 65,536 instances per step through the C oracle port, {cores} threads"
    for _ in range(max(1, min(args.warmup, 3))):
        fn()
    steps = max(1, min(args.steps, 20))
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        per_step.append(time.perf_counter() - t0)
    total = time.perf_counter() - t_all
    value = n_s * steps / total
    line = {"impl": "reference", "metric": "constraint+Jacobian evals/sec (instances/s)", "value": value, "unit": "instances/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * total / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "num_contacts": NC, "env": "Ground"},
            "cpu_baseline": {"value": value, "unit": "instances/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def pinned_array(lib, shape):
    nbytes = int(np.prod(shape)) * 8
    ptr = C.c_void_p()
    st = lib.cplb_host_alloc(nbytes, C.byref(ptr))
    if st != 0:
        raise RuntimeError(lib.cplb_last_error().decode())
    arr = np.ctypeslib.as_array((C.c_double * (nbytes // 8)).from_address(ptr.value)).reshape(shape)
    return arr, ptr


def run_ours(args):
    import torch
    import torch.distributed as dist

    import centroidalplanner_b200 as cpl
    from centroidalplanner_b200 import _cabi, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the evaluator has no CPU fallback (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    env = cpl.Ground()
    prob = cpl.BatchedCplProblem(synthetic.NAMES4, 100.0, env, device=local)
    configure(prob, env)
    n, m, nnz = prob.n, prob.m, prob.nnz
    N = N_PER_GPU
    bytes_per_launch = algorithmic_bytes_per_instance(n, m, nnz) * N
    layout = cpl.INSTANCE_MAJOR if args.layout == "instance" else cpl.COMPONENT_MAJOR

    x_host = make_inputs(rank)
    stream = torch.cuda.current_stream(dev)
    W, K = max(3, args.warmup), args.steps
    # buffer sets: total footprint >> L2 so no step finds its lines in L2
    sets = max(4, int(np.ceil(8 * L2_BYTES / bytes_per_launch)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def time_layout(lay):
        """W warm-up + K timed back-to-back evaluations in one buffer layout; returns (ms_per_step, launches, ...)."""
        shp = (lambda length: (N, length)) if lay == cpl.INSTANCE_MAJOR else (lambda length: (length, N))
        base = torch.from_numpy(x_host if lay == cpl.INSTANCE_MAJOR else np.ascontiguousarray(x_host.T)).to(dev)
        xs = [base.clone() for _ in range(sets)]
        gs = [torch.empty(shp(m), dtype=torch.float64, device=dev) for _ in range(sets)]
        js = [torch.empty(shp(nnz), dtype=torch.float64, device=dev) for _ in range(sets)]

        def step(i):
            s = i % sets
            # inputs_ready: the x buffers are resident and never written by a kernel (the metric's "inputs already in HBM")
            prob.eval(xs[s], g=True, jac=True, layout=lay, out={"g": gs[s], "jac": js[s]}, inputs_ready=not args.no_inputs_ready)

        for i in range(W):
            step(i)
        barrier()
        # The K timed steps are issued as replays of a CUDA graph holding one rotation over the buffer sets
        # (`sets` evaluation kernels, captured from the same cplb_eval_device calls) plus a plain-launch remainder:
        # at ~20 us per kernel the per-launch host path and inter-kernel launch gap would otherwise be >10% of the step.
        graph = None
        gsteps = min(K, sets)
        if not args.no_graph and gsteps >= 2:
            side = torch.cuda.Stream(dev)
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    for i in range(gsteps):
                        step(W + i)
            stream.wait_stream(side)
            graph.replay()  # one untimed replay (upload / first-run cost)
            barrier()
        reps, rem = (K // gsteps, K % gsteps) if graph is not None else (0, K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            graph.replay()
        for i in range(rem):
            step(W + i)
        e1.record(stream)
        barrier()
        ms_total = e0.elapsed_time(e1)
        n_launch = reps * gsteps + rem  # evaluation kernels executed inside the timed region
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # per-launch device time with an event pair around every launch (same stream), second pass
        prob.timing_begin()
        for i in range(min(K, 200)):
            step(W + K + i)
        torch.cuda.synchronize(dev)
        ev_ms, ev_k = prob.timing_end()
        del xs, gs, js
        return float(t.item()) / K, n_launch, ev_ms, ev_k

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    ms_per_step, launches, ev_ms, ev_k = time_layout(layout)
    other = cpl.COMPONENT_MAJOR if layout == cpl.INSTANCE_MAJOR else cpl.INSTANCE_MAJOR
    other_ms, _, other_ev_ms, _ = time_layout(other)
    value = world * N / (ms_per_step * 1e-3)
    lname = {cpl.INSTANCE_MAJOR: "instance-major", cpl.COMPONENT_MAJOR: "component-major"}
    kname = {cpl.INSTANCE_MAJOR: "eval_instance_major", cpl.COMPONENT_MAJOR: "eval_component_major_split"}
    shape = (lambda length: (N, length))  # the e2e leg below uses instance-major host buffers

    # ---- e2e: host buffers through cplb_eval_host -------------------------------------------
    lib = _cabi.load()
    hx, px = pinned_array(lib, shape(n))
    hg, pg = pinned_array(lib, shape(m))
    hj, pj = pinned_array(lib, shape(nnz))
    hx[...] = x_host
    e2e_steps = max(3, min(K, 20))
    e2e_layout = cpl.INSTANCE_MAJOR  # what an IPOPT thread consumes: its instance's x[n] -> g[m], values[nnz] slices
    for _ in range(2):
        prob.eval(hx, g=True, jac=True, layout=e2e_layout, out={"g": hg, "jac": hj})
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prob.eval(hx, g=True, jac=True, layout=e2e_layout, out={"g": hg, "jac": hj})  # synchronous: outputs landed
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_check = float(np.abs(hg).sum())  # touch the result on the host
    # the same call for a queue of batches (cplb_eval_host_begin / _wait, two pinned buffer sets: begin k+1, wait k) -- the
    # host-side counterpart of the device-side `value`, which also runs a queue of independent batches back to back
    hx2, px2 = pinned_array(lib, shape(n))
    hg2, pg2 = pinned_array(lib, shape(m))
    hj2, pj2 = pinned_array(lib, shape(nnz))
    hx2[...] = x_host
    bufsets = [(hx, {"g": hg, "jac": hj}), (hx2, {"g": hg2, "jac": hj2})]
    for b in range(2):
        prob.eval_host_wait(prob.eval_host_begin(bufsets[b][0], bufsets[b][1], g=True, jac=True, layout=e2e_layout)[0])
    barrier()
    t0 = time.perf_counter()
    pending = None
    for k in range(e2e_steps):
        ticket, _ = prob.eval_host_begin(bufsets[k % 2][0], bufsets[k % 2][1], g=True, jac=True, layout=e2e_layout)
        if pending is not None:
            prob.eval_host_wait(pending)     # batch k-1 has landed in its host buffers
        pending = ticket
    prob.eval_host_wait(pending)
    e2e_q_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_q_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_q_s = float(t.item())
    e2e_q_check = float(np.abs(hg2).sum())
    assert e2e_q_check == e2e_check, "queued host evaluation differs from the synchronous one"
    # component-major host buffers, x-independent Jacobian slots (whole rows there) pre-filled once and not re-transferred
    cmask, _ = prob.GetJacobianConstants()
    n_var = int((~cmask).sum())
    hxc, hgc, hjc = hx.reshape(n, N), hg.reshape(m, N), hj.reshape(nnz, N)
    hxc[...] = x_host.T
    prob.FillJacobianConstants(hjc, layout=cpl.COMPONENT_MAJOR)
    prob.eval(hxc, g=True, jac=True, layout=cpl.COMPONENT_MAJOR, out={"g": hgc, "jac": hjc}, jac_constants_present=True)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prob.eval(hxc, g=True, jac=True, layout=cpl.COMPONENT_MAJOR, out={"g": hgc, "jac": hjc}, jac_constants_present=True)
    torch.cuda.synchronize(dev)
    e2e_cm_s = (time.perf_counter() - t0) / e2e_steps

    # optional final gather of the per-rank output slices (NCCL all_gather over NVLink), timed apart from the evaluation
    gather = None
    if world > 1:
        from centroidalplanner_b200 import sharding

        gl = {"g": torch.empty((N, m), dtype=torch.float64, device=dev), "jac": torch.empty((N, nnz), dtype=torch.float64, device=dev)}
        sharding.gather_outputs(gl, world * N)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(5):
            full = sharding.gather_outputs(gl, world * N)
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / 5], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"ms": float(tg.item()), "bytes_per_rank_out": 8 * (m + nnz) * N * world,
                  "note": "all_gather of g and jac slices to every rank; NOT part of value/ms_per_step"}
        del full, gl

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = bytes_per_launch / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": "constraint+Jacobian evals/sec (instances/s)", "value": value, "unit": "instances/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "num_contacts": NC, "env": "Ground", "instances_per_gpu": N,
                       "layout": lname[layout],
                       "outputs": "g+jac", "params": "shared",
                       "inputs_ready": (not args.no_inputs_ready) and "x buffers are resident and not produced by the preceding kernel: "
                                       "CPLB_DEVICE_INPUTS_READY lets each kernel load x before waiting for the previous one (outputs still in stream order)",
                       "launch": "plain launches" if args.no_graph or min(K, sets) < 2 else
                       f"CUDA graph of {min(K, sets)} evaluation kernels replayed {K // min(K, sets)}x + {K % min(K, sets)} plain launches",
                       "l2": f"inputs larger than L2: steps rotate through {sets} buffer sets, {sets * bytes_per_launch / 2**20:.0f} MiB total vs 126 MiB L2"},
            "e2e": {"value": world * N / e2e_q_s, "unit": "instances/s", "h2d_bytes_per_step": world * 8 * n * N,
                    "d2h_bytes_per_step": world * 8 * (m + nnz) * N, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_q_s,
                    "api": "cplb_eval_host_begin / cplb_eval_host_wait: a queue of batches on two sets of instance-major pinned host "
                           "buffers (begin k+1, wait k); every step uploads its x and downloads its g + Jacobian; chunked "
                           "H2D/kernel/D2H on 3 streams",
                    "synchronous_call": {"value": world * N / e2e_s, "ms_per_step": 1e3 * e2e_s,
                                         "api": "cplb_eval_host: one batch at a time, returns when its outputs have landed"},
                    "component_major_constants_skipped": {
                        "value": N / e2e_cm_s, "ms_per_step": 1e3 * e2e_cm_s, "d2h_bytes_per_step": 8 * (m + n_var) * N,
                        "note": f"rank 0 only; component-major host buffers, the {nnz - n_var} x-independent of {nnz} Jacobian slot rows pre-filled once"},
                    "checksum": e2e_check},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "bytes_per_launch": bytes_per_launch,
                         "kernel": kname[layout],
                         "avg_launch_ms": ms_per_step, "avg_launch_ms_event_pairs": ev_ms, "event_pair_launches": ev_k},
            "other_layout": {"layout": lname[other], "kernel": kname[other], "ms_per_step": other_ms,
                             "value": world * N / (other_ms * 1e-3), "roofline_frac": bytes_per_launch / (other_ms * 1e-3) / 1e9 / peak,
                             "avg_launch_ms_event_pairs": other_ev_ms},
            "clocks": clocks,
        }
        if gather is not None:
            line["gather"] = gather
        traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(traffic_file):
            try:
                line["roofline"]["traffic"] = json.load(open(traffic_file)).get(line["roofline"]["kernel"])
            except Exception:
                pass
        if world == 1 and not args.no_cpu:
            from oracle import cpl_oracle_py, cpl_ref_py  # noqa: F401  (cpu_baseline leg: the CPU path as the timed baseline)

            cores = host_cores()
            o = oracle_problem()
            o.set_call_all_pairs(0)
            v_port, reps = time_cpu(o, x_host, cores, budget_s=args.cpu_seconds / 2)
            port = {"value": v_port, "unit": "instances/s", "cores": cores, "kind": "port",
                    "sample": f"{reps} passes over the same 65,536 instances, best pass; C oracle (-O2), {cores} threads, "
                              "(set, variable-set) pairs that produce no Jacobian entry skipped -- a lean port, faster than "
                              "the reference's own ifopt assembly"}
            if cpl_ref_py.available():
                from centroidalplanner_b200 import synthetic as syn

                r = cpl_ref_py.RefProblem(syn.NAMES4, "ground", 100.0)
                configure(r, r)
                n_s = 16384
                r.eval_batch(x_host[:2048], want=("g", "jac"), nthreads=cores)
                best, t_spent, passes = 1e9, 0.0, 0
                while t_spent < args.cpu_seconds / 2 and passes < 50:
                    t0 = time.perf_counter()
                    r.eval_batch(x_host[:n_s], want=("g", "jac"), nthreads=cores)
                    dt = time.perf_counter() - t0
                    best, t_spent, passes = min(best, dt), t_spent + dt, passes + 1
                line["cpu_baseline"] = {
                    "value": n_s / best, "unit": "instances/s", "cores": cores, "kind": "reference",
                    "sample": f"{passes} passes over the first {n_s} of the 65,536 instances, best pass; the reference's own sources "
                              "(CplProblem + its IFOPT components) compiled in place against stand-in Eigen/ifopt headers "
                              f"(oracle/_ref; std::map-based sparse blocks, so indicative), one CplProblem per thread, {cores} threads",
                    "lean_port": port}
            else:
                line["cpu_baseline"] = port
        print(json.dumps(line), flush=True)

    for p in (px, pg, pj):
        lib.cplb_host_free(p)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default="component", choices=["instance", "component"],
                    help="buffer layout of the headline number (the other one is timed too and reported under other_layout)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-inputs-ready", action="store_true",
                    help="do not tell the evaluator that x is independent of the preceding kernel (CPLB_DEVICE_INPUTS_READY)")
    ap.add_argument("--no-graph", action="store_true", help="issue every timed step as a separate launch instead of CUDA-graph replays")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
