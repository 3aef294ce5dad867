#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): quadrature-point basis+Jacobian evaluations per
second on config C2 (3-D thick-walled cylinder patch, p=3, 128^3 elements per GPU, 4^3 Gauss points; outputs N, dN, x,
Jt, detJ for every point).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the evaluation over ALL 2,097,152 elements (134,217,728 points) owned by the rank, issued as
element-block launches that write into a ring of device output buffers (the full output, 289 GB, does not fit HBM).
Weak scaling: the N-GPU patch has 128 x 128 x 128N elements, rank r owns the k-slab [128r, 128r+128) (knot-span blocks,
SURVEY.md 8e); knots / control net are replicated; no data-path collective.

Prints ONE JSON line (rank 0).  torch is used only for device buffers, the stream handed to the C-ABI, CUDA events and
torch.distributed plumbing; the compute is dune-iga_b200/libigab200.so.
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

P_DEG, NQ1, LEVEL = 3, 4, 7           # C2: p=3, 4^3 Gauss points, 2^7 = 128 elements per direction
LAYERS_PER_LAUNCH = 8                 # k-layers of elements per launch (131,072 elements, 18 GB of output)
FIELDS = ("N", "dN", "x", "Jt", "detJ")
NB = (P_DEG + 1) ** 3
NQ = NQ1 ** 3
# algorithmic bytes per point, SURVEY.md 8(d): outputs written once + the element's control points read once / nq
ALG_BYTES_PER_PT = 8 * (NB * (1 + 3) + 3 + 9 + 1) + 8 * NB * 4 // NQ


def measured_peak_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_patch(n_gpus):
    import igab200  # noqa: F401
    from dune_iga_b200 import patchdata as PD
    extra = int(np.log2(n_gpus))
    assert 2 ** extra == n_gpus, "gpus must be a power of two"
    return PD.thickCylinder(P_DEG, (LEVEL, LEVEL, LEVEL + extra))


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = the reference headers
    compiled where they lie; falls back to the pinned C restatement), all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import igab200  # noqa: F401
    from dune_iga_b200 import patchdata as PD
    import refbind as RF
    import portbind as PT
    pd = build_patch(args.gpus)
    if RF.available():
        O, kind = RF.RefPatch.create(pd.degree, pd.knotSpans, pd.cp, 3), "reference"
    else:
        O, kind = PT.PortPatch(pd.degree, pd.knotSpans, pd.cp, 3), "port"
    xi = [PD.gaussLegendre01(NQ1)[0]] * 3
    layers = 2  # 32,768 elements = 2,097,152 points per step (4.5 GB of output, ~0.7 s on 16 cores)
    nel = 128 * 128 * layers
    ncpu = len(os.sched_getaffinity(0))  # torchrun exports OMP_NUM_THREADS=1: ask for every core explicitly
    # the output buffers are allocated and touched once, outside the timed region (the reference arm is timed on its
    # arithmetic, not on page faults of fresh numpy arrays)
    out = None
    for _ in range(max(1, args.warmup)):
        out = O.eval_tensor(xi, order=1, elem_begin=0, elem_end=nel, want=FIELDS, nthreads=ncpu, out=out)
    t0 = time.perf_counter()
    threads = 1
    for s in range(args.steps):
        eb = (s % 32) * nel
        threads = O.eval_tensor(xi, order=1, elem_begin=eb, elem_end=eb + nel, want=FIELDS, nthreads=ncpu, out=out)["threads"]
    dt = time.perf_counter() - t0
    val = args.steps * nel * NQ / dt
    sample = "%d k-layer(s) of the C2 patch per step (%d elements, %d points), outputs %s" % (
        layers, nel, nel * NQ, "+".join(FIELDS))
    print(json.dumps({
        "impl": "reference", "metric": "quad-pt basis+Jacobian evals/sec", "value": val, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def config_dict(n):
    return {"workload": "C2: 3-D thick-walled quarter cylinder NURBS patch p=3, 128x128x%d elements (128^3 per GPU), "
                        "4^3 Gauss points, order-1 outputs N+dN+x+Jt+detJ" % (128 * n),
            "elements_per_gpu": 128 ** 3, "points_per_gpu": 128 ** 3 * NQ, "bytes_per_point": ALG_BYTES_PER_PT,
            "elements_per_launch": 128 * 128 * LAYERS_PER_LAUNCH,
            "l2": "each launch writes 18 GB and reads a 72 MB control net: working set >> 126 MB L2, no flush needed",
            "parallelism": "element k-slabs, %d rank(s), no collective" % n}


def measured_fp64_peak():
    """DMMA / DFMA TFLOP/s of this GPU from tools/fp64_peak (built by __graft_entry__.build()); the driver's
    MEASURED_PEAKS.json holds no FP64 figure."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        dmma = max(float(l.split()[-2]) for l in out.splitlines() if l.startswith("DMMA"))
        dfma = max(float(l.split()[-2]) for l in out.splitlines() if l.startswith("DFMA"))
        return dmma, dfma, "measured live by tools/fp64_peak (mma.sync.m8n8k4.f64 / DFMA loops)"
    except Exception:
        return 37.0, 34.0, "fallback: tools/fp64_peak measured on this pool in round 1"


def bench_assembly(args, rank, world, local, barrier, dist):
    import torch
    import igab200  # noqa: F401
    from dune_iga_b200 import capi, patchdata as PD
    p, nq1, lev = 4, 5, 6
    extra = int(np.log2(world))
    pd = PD.thickCylinder(p, (lev, lev, lev + extra))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(nq1)
    xi, ww = [x] * 3, [w] * 3
    n = (p + 1) ** 3
    per_call = 64 * 64 * 4  # 16,384 elements = 2.05 GB of element matrices per call
    calls = 64 // 4
    my_begin = rank * 64 ** 3
    ring = [torch.empty((per_call, n, n), dtype=torch.float64, device="cuda") for _ in range(2)]
    stream = torch.cuda.current_stream()

    def one_pass():
        for j in range(calls):
            eb = my_begin + j * per_call
            P.assemble(capi.ASM_POISSON, xi, ww, elem_begin=eb, elem_end=eb + per_call, Ke=ring[j & 1], fe=None,
                       stream=stream.cuda_stream)

    one_pass()
    barrier()
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    npass = 3
    e0.record(stream)
    for _ in range(npass):
        one_pass()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = int(capi.launch_count() - l0)
    nel = 64 ** 3 * world * npass
    flops_el = 125 * (2 * 3 * n * n + 2 * 9 * n)  # nq (2 s n^2 + 2 s^2 n)
    exec_el = 3360 * 512 + 5 * 400 * 55  # DMMA instructions x 512 flop + stage-1 FMAs of gram_sf_kernel<4,5>
    dmma, dfma, src = measured_fp64_peak() if rank == 0 else (37.0, 34.0, "")
    achieved = flops_el * nel / world / (ms * 1e-3) / 1e12  # per GPU
    del ring
    # end to end through the reference-facing call: host patch (handle creation = control-net H2D), igab_global_assemble
    # (element matrices in device chunks, CSR gather), CSR values + load vector back to host memory
    asm_e2e = None
    if not args.no_e2e:
        k_layers = 16  # 64 x 64 x 16 elements per e2e step
        sub = PD.thickCylinder(p, (lev, lev, 4))
        Ps = capi.Patch.fromPatchData(sub)
        rp, _ = Ps.csrPattern()
        v, bb = capi.pinned_empty(int(rp[-1])), capi.pinned_empty(len(rp) - 1)  # page-locked host outputs
        Ps.globalAssemble(capi.ASM_POISSON, xi, ww, out=(v, bb))  # warm-up
        barrier()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            Ps.updateControlPoints(sub.cp)
            Ps.globalAssemble(capi.ASM_POISSON, xi, ww, out=(v, bb))
        dt = (time.perf_counter() - t0) / reps
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        asm_e2e = {"value": Ps.nelem * world / dt, "unit": "elements/s", "h2d_bytes_per_step": int(sub.cp.nbytes),
                   "d2h_bytes_per_step": int(v.nbytes + bb.nbytes),
                   "sample": "%d elements per rank per step (64 x 64 x %d), control net H2D, igab_global_assemble, CSR "
                             "values (%d non-zeros) + load vector D2H" % (Ps.nelem, k_layers, v.size)}
        capi.pinned_free(v)
        capi.pinned_free(bb)
        del Ps
    # CPU beside it (rank 0, N=1): the oracle's plain-loop element matrices (transliteration of poisson.py:37-81), all cores
    asm_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import portbind as PT
        small = PD.thickCylinder(p, (3, 3, 3))
        Op = PT.PortPatch(small.degree, small.knotSpans, small.cp, 3)
        ncpu = len(os.sched_getaffinity(0))
        Op.assemble(0, xi, ww, elem_end=16, nthreads=ncpu)
        t0 = time.perf_counter()
        Op.assemble(0, xi, ww, nthreads=ncpu)
        dt = time.perf_counter() - t0
        asm_cpu = {"value": Op.nelem / dt, "unit": "elements/s", "cores": ncpu, "kind": "port",
                   "sample": "8^3-element p=4 patch (512 element matrices, evaluation + plain-loop B^T B), %.1f s wall" % dt}
    return {"metric": "element matrices/sec", "value": nel / (ms * 1e-3), "unit": "elements/s",
            "config": {"workload": "C3: 3-D Poisson stiffness p=4 (n=125), 5^3 Gauss, 64x64x%d elements (64^3 per GPU)"
                                   % (64 * world)},
            "ms_per_pass": ms / npass, "gpu_launches": launches, "e2e": asm_e2e, "cpu_baseline": asm_cpu,
            "roofline": {"bound": "tensor", "kernel": "gram_sf_kernel<4,5> (evaluation + sum-factorised Gram contraction, "
                                                      "stages 2 and 3 on the FP64 tensor cores, one kernel)",
                         "achieved": achieved, "peak": dmma, "unit": "TFLOP/s", "frac": achieved / dmma,
                         "traffic": None, "peak_source": src, "dfma_peak": dfma,
                         "algorithmic_flops_per_element": flops_el,
                         "executed_flops_per_element": exec_el,
                         "executed": achieved * exec_el / flops_el,
                         "executed_frac": achieved * exec_el / flops_el / dmma,
                         "hbm_write_gbs": 8.0 * n * n * nel / world / (ms * 1e-3) / 1e9,
                         "note": "achieved = algorithmic flops of the dense B^T D B (2*3*125*125^2 + ...), the SURVEY "
                                 "8(d) convention; the kernel factorises the sum over the quadrature points direction "
                                 "by direction over index pairs (i_d, j_d): per element 3,360 DMMA instructions (stage "
                                 "2: 5 x 16 tiles x 22 k-steps, stage 3: 5 x 16 x 20) + 0.11 MFLOP of scalar FMA "
                                 "(stage 1) = 1.83 MFLOP executed instead of 12.0 MFLOP, so a fraction above 1 is the "
                                 "algorithmic saving and `executed_frac` is the share of the measured DMMA rate the "
                                 "kernel sustains; `hbm_write_gbs` is the K_e stream it writes (125 kB per element); "
                                 "evaluation is fused into the same kernel (IGAB_GRAM_SF=0 selects the dense-Gram "
                                 "kernel gram_fused_kernel it replaced: 3.3e6 element matrices/s)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-assembly", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import igab200  # noqa: F401
    from dune_iga_b200 import capi, patchdata as PD

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pd = build_patch(world)
    P = capi.Patch.fromPatchData(pd)
    assert P.nelem == 128 * 128 * 128 * world
    P.setKernel(capi.KERNEL_TENSOR)  # fail loudly if the tuned kernel is not the one running
    gx = PD.gaussLegendre01(NQ1)[0]
    xi = [gx, gx, gx]
    el_per_launch = 128 * 128 * LAYERS_PER_LAUNCH
    pts_per_launch = el_per_launch * NQ
    launches_per_step = 128 // LAYERS_PER_LAUNCH
    my_begin = rank * 128 ** 3
    sh = P.shapes(pts_per_launch)
    ring = [{f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in FIELDS} for _ in range(2)]
    stream = torch.cuda.current_stream()

    def step(evs=None):
        for j in range(launches_per_step):
            eb = my_begin + j * el_per_launch
            if evs is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                evs.append(e)
            P.evalTensor(xi, order=1, elem_begin=eb, elem_end=eb + el_per_launch, want=FIELDS, out=ring[j & 1],
                         stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    l0 = capi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = []  # one event before every launch of the timed region -> per-launch durations of the dominant kernel
    barrier()
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step(evs)
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = capi.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    evs.append(ev1)
    per_launch = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)])
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pts_per_step = 128 ** 3 * NQ * world
    value = pts_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (eval_sf3d_kernel<3,4>, > 99 % of the step, profiles/launches_r01.csv):
    # average launch duration over the timed region, CUDA events on the launch stream (each interval = 1 main kernel
    # + 3 table kernels of ~6 us)
    k_ms = float(per_launch.mean())
    peak, peak_src = measured_peak_hbm()
    achieved = ALG_BYTES_PER_PT * pts_per_launch / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["eval_sf3d_p3q4_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "eval_sf3d_kernel<3,4>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, copy read+write)",
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_PT * pts_per_launch, "ms_per_launch": k_ms,
                "ms_per_launch_min_median_max": [float(per_launch.min()), float(np.median(per_launch)),
                                                 float(per_launch.max())],
                "launches_timed": int(len(per_launch))}

    # ---- e2e through the C-ABI with HOST buffers (rank-local; control net H2D + all outputs D2H inside the region)
    e2e = None
    if not args.no_e2e:
        e_layers = 2
        e_el = 128 * 128 * e_layers
        e_pts = e_el * NQ
        hs = P.shapes(e_pts)
        host = {f: capi.pinned_empty(hs[f]) for f in FIELDS}
        cp_host = capi.pinned_empty(pd.cp.shape)
        cp_host[...] = pd.cp
        d2h = sum(int(np.prod(hs[f])) * 8 for f in FIELDS)
        h2d = cp_host.nbytes

        def e2e_step(s):
            P.updateControlPoints(cp_host)
            eb = my_begin + (s % 32) * e_el
            P.evalTensor(xi, order=1, elem_begin=eb, elem_end=eb + e_el, want=FIELDS, out=host)
            return float(host["detJ"][0])

        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 5))
        for s in range(n_e2e):
            e2e_step(s)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": e_pts * n_e2e * world / dt, "unit": "points/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "sample": "%d k-layers per rank per e2e step (%d points), control net H2D + all outputs D2H to pinned "
                         "host memory through igab_eval_tensor(IGAB_HOST); PCIe-bound" % (e_layers, e_pts)}
        for f in FIELDS:
            capi.pinned_free(host[f])
        capi.pinned_free(cp_host)

    # ---- the same end-to-end call chain when the consumer lives on the GPU (the assembly kernels): control net H2D,
    # the full evaluation pass with outputs left in HBM, one patch-level reduction (volume) read back to the host
    e2e_dev = None
    if not args.no_e2e:
        gw = PD.gaussLegendre01(NQ1)[1]
        cp_host = capi.pinned_empty(pd.cp.shape)
        cp_host[...] = pd.cp

        def e2e_dev_step():
            P.updateControlPoints(cp_host)
            step()
            return P.volume(xi, [gw, gw, gw], my_begin, my_begin + 128 ** 3, stream=stream.cuda_stream)

        vol = e2e_dev_step()
        barrier()
        t0 = time.perf_counter()
        n_dev = 3
        for _ in range(n_dev):
            vol = e2e_dev_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e_dev = {"value": 128 ** 3 * NQ * world * n_dev / dt, "unit": "points/s",
                   "h2d_bytes_per_step": int(cp_host.nbytes), "d2h_bytes_per_step": 8,
                   "result": "slab volume %.12f (exact pi/4*(r_o^2-r_i^2)*L/%d = %.12f)" % (
                       vol, 1, np.pi / 4 * 3.0 * 3.0 / world),
                   "note": "all 134,217,728 points per rank evaluated per step with N+dN+x+Jt+detJ written to HBM (ring "
                           "buffers) for a device-side consumer; only the reduction crosses PCIe"}
        capi.pinned_free(cp_host)

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's own headers (oracle/_ref), all host threads
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refbind as RF
        import portbind as PT
        if RF.available():
            O, kind = RF.RefPatch.create(pd.degree, pd.knotSpans, pd.cp, 3), "reference"
        else:
            O, kind = PT.PortPatch(pd.degree, pd.knotSpans, pd.cp, 3), "port"
        c_layers = 4
        c_el = 128 * 128 * c_layers
        ncpu = len(os.sched_getaffinity(0))
        # first call allocates and touches the 9 GB of output (untimed), the timed call reuses the buffers
        r = O.eval_tensor(xi, order=1, elem_begin=0, elem_end=c_el, want=FIELDS, nthreads=ncpu)
        t0 = time.perf_counter()
        r = O.eval_tensor(xi, order=1, elem_begin=c_el, elem_end=2 * c_el, want=FIELDS, nthreads=ncpu, out=r)
        dt = time.perf_counter() - t0
        cpu = {"value": c_el * NQ / dt, "unit": "points/s", "cores": r["threads"], "kind": kind,
               "sample": "%d k-layers of the same patch (%d elements, %d points), same outputs, %.1f s wall" % (
                   c_layers, c_el, c_el * NQ, dt)}
        del r
        # the path AS SHIPPED: per-point calls that copy the whole patch's weights / coordinates
        # (nurbsbasis.hh:637-663, nurbspatchgeometry.hh:76-82,163-185), on a 16^3 patch only, labelled as such
        if kind == "reference":
            small = PD.thickCylinder(3, (4, 4, 4))
            Os = RF.RefPatch.create(small.degree, small.knotSpans, small.cp, 3)
            rng = np.random.default_rng(0)
            u = rng.random((1 << 18, 3))  # ~2 s on 16 cores
            Os.as_shipped_points(u[:256], nthreads=ncpu)
            t0 = time.perf_counter()
            used, _ = Os.as_shipped_points(u, nthreads=ncpu)
            dt = time.perf_counter() - t0
            cpu["as_shipped"] = {"value": len(u) / dt, "unit": "points/s", "cores": used,
                                 "sample": "16^3-element p=3 patch (19^3 control points), %d random points, per point "
                                           "evaluateFunction + evaluateJacobian + global + jacobianTransposed + "
                                           "integrationElement exactly as the shipped classes call them (whole-patch "
                                           "weight / coordinate copies per call); %.2f s wall" % (len(u), dt)}

    # ---- second half of BASELINE.json's metric: element matrices/sec (config C3: p=4 Poisson stiffness, 5^3 Gauss,
    # FP64 tensor-core contraction), each rank assembling its own slab of a 64 x 64 x 64N patch
    assembly = None
    if not args.no_assembly:
        assembly = bench_assembly(args, rank, world, local, barrier, dist if world > 1 else None)

    if rank == 0:
        print(json.dumps({
            "metric": "quad-pt basis+Jacobian evals/sec", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(world),
            "clocks": clocks, "e2e": e2e, "e2e_device_consumer": e2e_dev, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "assembly": assembly}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
