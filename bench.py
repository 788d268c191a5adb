#!/usr/bin/env python
"""bench.py — TT-SVD GElements/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape 64,64,64,64,64] [--rank 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A "step" = one complete TT-SVD (tn.Tensor(X, ranks_tt=r)) of one dense fp32 tensor per GPU.
Workload: BASELINE.json configs[1] names 64^8 (2^48 elements = 1.1 PB) which cannot exist on any
machine; the stand-in is the largest 64^d that fits one GPU, randn(64,64,64,64,64) fp32 (4 GiB),
target TT-rank 32.  Multi-GPU: weak scaling, one tensor per
rank (the batch dimension shards), no data-path collective, one NCCL all-gather of the final cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="64,64,64,64,64")
    ap.add_argument("--rank", type=int, default=32)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="generic CUDA-core kernels only (A/B runs)")
    ap.add_argument("--cpu-shape", default="32,32,32,32,32", help="bounded sample timed on the host cores")
    ap.add_argument("--reserve-sms", type=int, default=-1, help="SMs left free by the persistent kernels (measured: no gain on B200, default 0)")
    ap.add_argument("--per-gpu-batch", type=int, default=6,
                    help="independent tensors decomposed concurrently per GPU (one CUDA stream + host thread each): the "
                         "latency-bound eigen phases of one overlap the bandwidth-bound Gram/projection phases of another")
    ap.add_argument("--step-barrier", action="store_true",
                    help="join all in-flight tensors after every step instead of once after the K steps")
    ap.add_argument("--no-concurrent-flag", action="store_true",
                    help="A/B: do not pass TNB_FLAG_CONCURRENT when several tensors are in flight")
    return ap.parse_args()


METRIC = "TT-SVD GElements/s"
_json_out = sys.stdout  # replaced in __main__ by a duplicate of the real fd 1 (everything else goes to stderr)


# ------------------------------------------------------------------------------------------------
# CPU arm: the REAL reference (rballester/tntorch, staged unmodified into oracle/_ref/ by __graft_entry__.build())
# timed on the box's host cores; the NumPy port (oracle/tt_oracle.py) only when the staged copy is missing.
# ------------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def workload_config(args):
    """The `config` both arms print (nothing run-dependent in it, so the two lines carry the same dict)."""
    shape = [int(s) for s in args.shape.split(",")]
    return {"workload": f"TT-SVD randn{shape} fp32 -> TT-rank {args.rank} (stand-in for the infeasible 64^8: 1.1 PB)",
            "per_gpu_batch": max(1, args.per_gpu_batch),
            "parallelism": "batch-sharded over the GPUs (independent tensors), all-gather of the final cores",
            "l2": "input 4 GiB >> 126 MB L2 (no flush needed)"}


class CpuArm:
    """tn.Tensor(X, ranks_tt=r, algorithm=...) of the reference on a bounded sample of the workload."""

    def __init__(self, shape, rank):
        import numpy as np

        self.shape, self.rank = tuple(shape), rank
        self.n = int(np.prod(shape))
        self.tn = None
        try:
            from oracle import stage_ref

            self.tn = stage_ref.load()
        except Exception:
            self.tn = None
        self.kind = "reference" if self.tn is not None else "port"
        self.algorithm = "eig"
        self.threads = cpu_threads()

    def make_input(self, seed=0):
        import numpy as np

        return np.random.default_rng(seed).standard_normal(self.shape, dtype=np.float32)

    def step(self, X, algorithm=None):
        """One decomposition; returns (seconds, cores as numpy arrays)."""
        alg = algorithm or self.algorithm
        if self.tn is not None:
            import torch

            Xt = torch.from_numpy(X)
            t0 = time.perf_counter()
            t = self.tn.Tensor(Xt, ranks_tt=self.rank, algorithm=alg)
            dt = time.perf_counter() - t0
            return dt, [c.numpy() for c in t.cores]
        from oracle import tt_oracle as orc

        t0 = time.perf_counter()
        cores = orc.tt_svd(X, ranks_tt=self.rank, algorithm=alg)
        return time.perf_counter() - t0, cores

    def tune(self):
        """LAPACK/BLAS on these shapes does not scale to every core of a 100+ core host (and torchrun exports
        OMP_NUM_THREADS=1): try a few thread counts and both reference algorithms on the actual sample, keep the
        fastest combination (the reference's default 'svd' computes and discards a full Vh; 'eig' is its Gram form)."""
        avail = cpu_threads()
        X = self.make_input(0)
        best = (float("inf"), avail, "eig")
        for n in sorted({min(avail, c) for c in (8, 16, 32, 64, avail)}):
            self.set_threads(n)
            for alg in ("eig", "svd"):
                dt, _ = self.step(X, alg)
                if dt < best[0]:
                    best = (dt, n, alg)
        _, self.threads, self.algorithm = best
        self.set_threads(self.threads)
        return best

    def set_threads(self, n):
        if self.tn is not None:
            import torch

            torch.set_num_threads(int(n))
        else:
            try:
                from threadpoolctl import threadpool_limits

                self._lim = threadpool_limits(limits=int(n))
            except Exception:
                pass

    def describe(self, steps):
        what = ("rballester/tntorch tn.Tensor(X, ranks_tt=%d, algorithm='%s') from oracle/_ref" % (self.rank, self.algorithm)
                if self.tn is not None else "oracle/tt_oracle.py::tt_svd (NumPy port; oracle/_ref was not staged)")
        return (f"randn{list(self.shape)} fp32 r={self.rank} (bounded sample of the 64^5 workload), {what}, {steps} step(s), "
                f"threads and algorithm picked among 8/16/32/64/all x eig/svd; CPU: {cpu_model()}")


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation on a bounded sample of the workload."""
    rank_env = int(os.environ.get("RANK", "0"))
    if rank_env != 0:
        return
    shape = tuple(int(s) for s in args.cpu_shape.split(","))
    arm = CpuArm(shape, args.rank)
    arm.tune()
    for i in range(min(args.warmup, 1)):
        arm.step(arm.make_input(100 + i))
    times = []
    for i in range(args.steps):
        X = arm.make_input(i)
        dt, _ = arm.step(X)
        times.append(dt)
    tot = sum(times)
    value = arm.n * args.steps / tot / 1e9
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GElements/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "GElements/s", "cores": arm.threads, "cores_available": cpu_threads(),
                         "kind": arm.kind, "cpu_model": cpu_model(), "algorithm": arm.algorithm,
                         "sample": arm.describe(args.steps)},
        "e2e": {"value": value, "unit": "GElements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), file=_json_out, flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(shape, rank, esz=4):
    """exact R->L TT-SVD reads the tensor twice and reads+writes every later carry once."""
    n = 1
    for s in shape:
        n *= s
    N = len(shape)
    total = 2 * n
    rows = n // shape[-1]
    r = 1
    carries = []
    for mu in range(N - 1, 0, -1):
        cols = shape[mu] * r
        r = min(rank, rows, cols)
        carries.append(rows * r)
        rows //= shape[mu - 1]
    total += 2 * sum(carries)
    return total * esz, carries


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from tntorch_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank_id = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shape = tuple(int(s) for s in args.shape.split(","))
    numel = int(np.prod(shape))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    bf16_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))

    PB = max(1, args.per_gpu_batch)
    concurrent = PB > 1 and not args.no_concurrent_flag
    # concurrent mode: gram_tc2 already leaves 4 SMs idle (72 CTA pairs); keep the same 4 free in every whole-GPU kernel
    reserve = args.reserve_sms if args.reserve_sms >= 0 else (4 if concurrent else 0)
    ops.set_reserved_sms(reserve)
    Xs, plans, streams = [], [], []
    for b in range(PB):
        g = torch.Generator(device=dev).manual_seed(1234 + rank_id * 16 + b)
        Xs.append(torch.randn(shape, generator=g, device=dev, dtype=torch.float32))  # 4 GiB each >> 126 MB L2
        plans.append(ops.TTSVDPlan(shape, torch.float32, rmax=args.rank, device=dev, use_tensorcore=not args.no_tc,
                                   concurrent=concurrent))
        streams.append(torch.cuda.Stream(device=dev))
    X, plan = Xs[0], plans[0]
    prof_plan = ops.TTSVDPlan(shape, torch.float32, rmax=args.rank, device=dev, use_tensorcore=not args.no_tc, profile=True)
    prof_plan.ws = plan.ws  # share the workspace
    prof_plan.cores_buf = plan.cores_buf
    pool = None
    if PB > 1:
        from concurrent.futures import ThreadPoolExecutor

        pool = ThreadPoolExecutor(PB)

    def gather_cores(cores_list):
        if world == 1:
            return
        flat = torch.cat([c.reshape(-1) for cores in cores_list for c in cores])
        out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=dev)
        dist.all_gather_into_tensor(out, flat)  # the final factor broadcast (north_star)

    def run_many(b, k):
        """Worker b: k decompositions back to back on its own stream (no barrier between steps)."""
        torch.cuda.set_device(local)
        with torch.cuda.stream(streams[b]):
            for _ in range(k):
                cores = plans[b].run(Xs[b])
        return cores

    def run_steps(k):
        """k steps = k * PB decompositions.  With several tensors in flight the workers stream through their k
        tensors independently (a step boundary is not a barrier: the eigen chain that ends one tensor overlaps the
        Gram of the next), joined once at the end; --step-barrier restores a join after every step."""
        if PB == 1:
            for _ in range(k):
                cores_list = [plan.run(X)]
                gather_cores(cores_list)
            return cores_list[0]
        cur = torch.cuda.current_stream()
        rounds = [1] * k if args.step_barrier else [k]
        for kk in rounds:
            for sb in streams:
                sb.wait_stream(cur)
            cores_list = list(pool.map(lambda b: run_many(b, kk), range(PB)))  # one host thread per in-flight tensor
            for sb in streams:
                cur.wait_stream(sb)
            for _ in range(kk):  # the final-factor all-gather of each step (the plan buffers hold the last step's cores)
                gather_cores(cores_list)
        return cores_list[0]

    def step():
        return run_steps(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # ---- timed region: K steps, CUDA events on the launching stream, barrier + sync on both sides ----
    sampler = ClockSampler(local)
    barrier()
    if rank_id == 0:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    cores = run_steps(args.steps)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ops.launch_count() - l0
    clocks = sampler.stop() if rank_id == 0 else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * PB * numel / (ms_step * 1e-3) / 1e9
    ranks = list(plan.ranks)

    # ---- per-phase device timings (same kernels, CUDA events inside the library, separate short run) ----
    phase = {"gram_ms": [], "eig_ms": [], "factor_ms": []}
    nprof = 3
    acc = np.zeros(32)
    prof_plan.run(X)  # untimed: the single-stream schedule uses kernels (resident filter) the concurrent steps did not load yet
    for _ in range(nprof):
        prof_plan.run(X)
        acc += np.array(list(prof_plan.info))
    acc /= nprof
    nsteps = int(round(acc[7]))
    for s in range(nsteps):
        phase["gram_ms"].append(round(float(acc[8 + 3 * s]), 4))
        phase["eig_ms"].append(round(float(acc[9 + 3 * s]), 4))
        phase["factor_ms"].append(round(float(acc[10 + 3 * s]), 4))
    # dominant kernel = the largest single phase
    cand = []
    B_alg, carries = algorithmic_bytes(shape, args.rank)
    rows0 = numel // shape[-1]
    # the Gram and projection phases are ONE kernel each (gram_tc_kernel / project_f32_kernel); the eigen phases are
    # chains of ~100 small launches and are reported as a share instead (phases_ms.eig_ms)
    for s in range(nsteps):
        cand.append((phase["gram_ms"][s], f"Gram of step {s}", s, "gram"))
        cand.append((phase["factor_ms"][s], f"project_tc_kernel (3xTF32 projection of step {s})", s, "factor"))
    cand.sort(reverse=True)
    top_ms, top_name, top_s, top_kind = cand[0]
    # algorithmic work of that phase
    rows = numel // shape[-1]
    r_prev = 1
    dims = []
    for mu in range(len(shape) - 1, 0, -1):
        cols = shape[mu] * r_prev
        r = ranks[mu]
        dims.append((rows, cols, r))
        r_prev = r
        rows //= shape[mu - 1]
    rws, cls, rr = dims[top_s]
    if top_kind == "gram":
        if cls <= 512:  # narrow Gram: HBM-bound, one read of the carry
            roof = {"kernel": "gram_tc_kernel (" + top_name + ")", "bound": "hbm", "achieved": rws * cls * 4 / top_ms / 1e6,
                    "peak": hbm_peak, "unit": "GB/s", "alg_bytes": rws * cls * 4}
        else:  # compute-bound symmetric Gram: rows*cols^2 MACs on the upper triangle -> rows*cols*(cols+1) flops
            fl = rws * cls * (cls + 1)
            roof = {"kernel": "gram_tc2_kernel, cta_group::2 (" + top_name + ")", "bound": "tensor",
                    "achieved": fl / top_ms / 1e9, "peak": bf16_sus / 2,
                    "unit": "TFLOP/s", "alg_flops": fl,
                    "peak_note": "TF32 dense peak taken as half the measured sustained bf16 cuBLAS rate (no TF32 entry in MEASURED_PEAKS.json)"}
    elif top_kind == "factor":
        by = (rws * cls + rws * rr) * 4
        roof = {"kernel": top_name, "bound": "hbm", "achieved": by / top_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                "alg_bytes": by}
    else:
        by = cls * cls * 4
        roof = {"kernel": top_name, "bound": "hbm", "achieved": by / top_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                "alg_bytes": by, "note": "latency-bound subspace eigensolver (dependent chain of small GEMMs)"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    # (profiles/r01_ncu_summaries.md) for the default 64^5 / r=32 workload; null for other shapes
    roof["traffic"] = None
    if list(shape) == [64] * 5 and args.rank == 32 and not args.no_tc:
        ncu_traffic = {("gram", 0): 4.295e9 + 0.004e9, ("factor", 0): 4.296e9 + 2.109e9, ("gram", 1): 4.185e9 + 0.008e9}
        roof["traffic"] = ncu_traffic.get((top_kind, top_s))
        roof["traffic_source"] = "profiles/r01_ncu_summaries.md (ncu --set full, per launch)"
    roof["ms"] = top_ms
    roof["peak_source"] = peak_src
    # the other single-kernel phases against their own bound (same CUDA-event timings), largest first
    others = []
    for ms_k, name_k, s_k, kind_k in cand[:4]:
        rws_k, cls_k, rr_k = dims[s_k]
        if kind_k == "gram" and cls_k > 512:
            others.append({"kernel": "gram_tc2_kernel " + name_k, "ms": ms_k, "bound": "tensor",
                           "frac": rws_k * cls_k * (cls_k + 1) / ms_k / 1e9 / (bf16_sus / 2)})
        else:
            by_k = rws_k * cls_k * 4 + (rws_k * rr_k * 4 if kind_k == "factor" else 0)
            others.append({"kernel": ("gram_tc_kernel " if kind_k == "gram" else "") + name_k, "ms": ms_k, "bound": "hbm",
                           "frac": by_k / ms_k / 1e6 / hbm_peak})
    roof["kernels"] = others
    sweep_roof = {"alg_bytes_per_tensor": B_alg, "achieved_GBps": PB * B_alg / ms_step / 1e6, "peak_GBps": hbm_peak,
                  "frac": PB * B_alg / ms_step / 1e6 / hbm_peak}

    # ---- parity of what was just timed (device-side fp64 error kernel; not in the timed region) ----
    relerr = ops.tt_relative_error(X, cores)

    # ---- e2e: host buffers through the public plan API, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        del prof_plan
        # two host-buffer pipelines per GPU: the H2D copy of one tensor overlaps the kernels of the other (the PCIe
        # link is the bound: 4 GiB per tensor); every step still moves its own input and its own result
        EB = 2 if PB > 1 else 1
        hplans, Xhs = [], []
        for b in range(EB):
            hp = ops.TTSVDPlan(shape, torch.float32, rmax=args.rank, device=dev, use_tensorcore=not args.no_tc, host_io=True)
            hp.ws = plans[b % PB].ws
            hp.cores_buf = plans[b % PB].cores_buf
            hplans.append(hp)
            xh = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            xh.copy_(Xs[b % PB])
            Xhs.append(xh)

        def run_host_one(b):
            torch.cuda.set_device(local)
            with torch.cuda.stream(streams[b]):
                hc = hplans[b].run_host(Xhs[b])
                return float(hc[0][0, 0, 0])  # the result is on the host

        def e2e_step():
            if EB == 1:
                return [run_host_one(0)]
            cur = torch.cuda.current_stream()
            for sb in streams[:EB]:
                sb.wait_stream(cur)
            r = list(pool.map(run_host_one, range(EB)))
            for sb in streams[:EB]:
                cur.wait_stream(sb)
            return r

        for _ in range(2):
            e2e_step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        ksteps = max(2, min(args.steps, 5))
        for _ in range(ksteps):
            e2e_step()
        f1.record()
        barrier()
        tt = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ems = float(tt.item()) / ksteps
        e2e = {"value": world * EB * numel / (ems * 1e-3) / 1e9, "unit": "GElements/s", "ms_per_step": ems,
               "h2d_bytes_per_step": EB * numel * 4, "d2h_bytes_per_step": EB * int(hplans[0].cap) * 4, "steps": ksteps,
               "tensors_per_step": EB}

    cpu = None
    same_sample = None
    if rank_id == 0 and not args.no_cpu_baseline:
        # free the e2e / batch buffers first: the CPU leg needs host RAM and cores, not HBM
        cshape = tuple(int(s) for s in args.cpu_shape.split(","))
        arm = CpuArm(cshape, args.rank)
        arm.tune()
        CX = arm.make_input(0)
        dt, ccores = arm.step(CX)
        v = arm.n / dt / 1e9
        cpu = {"value": v, "unit": "GElements/s", "cores": arm.threads, "cores_available": cpu_threads(), "kind": arm.kind,
               "cpu_model": cpu_model(), "algorithm": arm.algorithm, "sample": arm.describe(1) + f"; 1 pass = {dt:.2f} s"}
        # the SAME sample through our path: device-resident and end to end from host memory (same_config ratios)
        Xs_dev = torch.from_numpy(CX).to(dev)
        splan = ops.TTSVDPlan(cshape, torch.float32, rmax=args.rank, device=dev, use_tensorcore=not args.no_tc)
        for _ in range(3):
            scores = splan.run(Xs_dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        s0.record()
        for _ in range(reps):
            scores = splan.run(Xs_dev)
        s1.record()
        torch.cuda.synchronize()
        ms_dev = s0.elapsed_time(s1) / reps
        hplan = ops.TTSVDPlan(cshape, torch.float32, rmax=args.rank, device=dev, use_tensorcore=not args.no_tc, host_io=True)
        xh = torch.from_numpy(CX).pin_memory()
        hplan.run_host(xh)
        t0 = time.perf_counter()
        for _ in range(reps):
            hplan.run_host(xh)
        torch.cuda.synchronize()
        ms_host = (time.perf_counter() - t0) * 1e3 / reps
        import numpy as _np

        from oracle import tt_oracle as orc  # checker only: the error of both results against the same input

        e_ref = orc.relative_error(CX, ccores)
        e_ours = ops.tt_relative_error(Xs_dev, scores)
        same_sample = {"shape": list(cshape), "cpu_ms": dt * 1e3, "ours_device_ms": ms_dev, "ours_e2e_ms": ms_host,
                       "same_config_ratio_device": dt * 1e3 / ms_dev, "same_config_ratio_e2e": dt * 1e3 / ms_host,
                       "rel_error_reference": float(e_ref), "rel_error_ours": float(e_ours),
                       "note": "one tensor, one call in flight; the same bounded sample the CPU arm is timed on"}
        del Xs_dev, splan, hplan, xh

    if rank_id == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "GElements/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (Gram: tcgen05 kind::tf32, fp32 TMEM accumulation; projections: 3xTF32 on tcgen05 = fp32 accuracy; Gram matrices, eigenproblems and rank rule in fp64)"
            if not args.no_tc else "f32 (fp64-accumulated Gram, fp32 projections)",
            "data": "synthetic",
            "config": {"workload": f"TT-SVD randn{list(shape)} fp32 -> TT-rank {args.rank} (stand-in for the infeasible 64^8: 1.1 PB)",
                       "per_gpu_batch": PB, "in_flight_per_gpu": PB, "reserved_sms": reserve,
                       "step_join": "every step" if (args.step_barrier or PB == 1) else "once after the K steps (workers stream)",
                       "concurrent_flag": bool(concurrent),
                       "parallelism": f"batch-sharded x{world} ({PB} independent tensors in flight per GPU on {PB} streams), all-gather of final cores",
                       "l2": "input 4 GiB >> 126 MB L2 (no flush needed)", "ranks": ranks},
            "rel_error": relerr,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "sweep_roofline": sweep_roof,
            "phases_ms": phase,
            "e2e": e2e,
            "cpu_baseline": cpu,
            "same_sample": same_sample,
        }
        print(json.dumps(out), file=_json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner at communicator
    # creation, torchrun notices) is routed to stderr by pointing fd 1 at fd 2 for the whole run
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
