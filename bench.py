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


def build_patch(n_gpus, scaling="weak"):
    import igab200  # noqa: F401
    from dune_iga_b200 import patchdata as PD
    extra = int(np.log2(n_gpus)) if scaling == "weak" else 0
    assert 2 ** int(np.log2(n_gpus)) == n_gpus, "gpus must be a power of two"
    return PD.thickCylinder(P_DEG, (LEVEL, LEVEL, LEVEL + extra))


def run_reference(args, rank, world, emit):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref = the reference headers
    compiled where they lie; falls back to the pinned C restatement), all host threads, bounded sample per step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import igab200  # noqa: F401
    from dune_iga_b200 import patchdata as PD
    import refbind as RF
    import portbind as PT
    pd = build_patch(args.gpus, args.scaling)
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
    emit(json.dumps({
        "impl": "reference", "metric": "quad-pt basis+Jacobian evals/sec", "value": val, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.gpus, args.scaling),
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def config_dict(n, scaling="weak"):
    if scaling == "strong":
        lay = min(LAYERS_PER_LAUNCH, 128 // n)
        return {"workload": "C2 (strong scaling): the ONE 3-D thick-walled quarter cylinder NURBS patch p=3 with 128^3 = "
                            "2,097,152 elements split into %d k-slabs of %d element layers, 4^3 Gauss points, order-1 "
                            "outputs N+dN+x+Jt+detJ" % (n, 128 // n),
                "elements_per_gpu": 128 ** 3 // n, "points_per_gpu": 128 ** 3 * NQ // n, "bytes_per_point": ALG_BYTES_PER_PT,
                "elements_per_launch": 128 * 128 * lay,
                "l2": "each launch writes %.1f GB and reads the control net: working set >> 126 MB L2, no flush needed"
                      % (128 * 128 * lay * NQ * ALG_BYTES_PER_PT / 1e9),
                "parallelism": "element k-slabs of one patch, %d rank(s), no collective in the evaluation pass" % n}
    return {"workload": "C2: 3-D thick-walled quarter cylinder NURBS patch p=3, 128x128x%d elements (128^3 per GPU), "
                        "4^3 Gauss points, order-1 outputs N+dN+x+Jt+detJ" % (128 * n),
            "elements_per_gpu": 128 ** 3, "points_per_gpu": 128 ** 3 * NQ, "bytes_per_point": ALG_BYTES_PER_PT,
            "elements_per_launch": 128 * 128 * LAYERS_PER_LAUNCH,
            "l2": "each launch writes 18 GB and reads a 72 MB control net: working set >> 126 MB L2, no flush needed",
            "parallelism": "element k-slabs, %d rank(s), no collective in the evaluation pass (multi_gpu times the "
                           "NCCL reductions and the interface-row exchange)" % n}


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

    npass = 3

    def timed_passes():
        one_pass()
        barrier()
        l0 = capi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(npass):
            one_pass()
        e1.record(stream)
        barrier()
        t_ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, int(capi.launch_count() - l0)

    # the default kernel: gram_sym_kernel<4,5,0> (upper block triangle + mirror, assemble_sym.cu)
    ms, launches = timed_passes()
    # beside it, the kernel that contracts every block (gram_sf_kernel<4,5,0>, round 1 / 2's default)
    prev_sym = os.environ.get("IGAB_GRAM_SYM")
    os.environ["IGAB_GRAM_SYM"] = "0"
    ms_all, _ = timed_passes()
    if prev_sym is None:
        del os.environ["IGAB_GRAM_SYM"]
    else:
        os.environ["IGAB_GRAM_SYM"] = prev_sym
    nel = 64 ** 3 * world * npass
    flops_el = 125 * (2 * 3 * n * n + 2 * 9 * n)  # nq (2 s n^2 + 2 s^2 n)
    # executed per element: DMMA instructions x 512 flop + the scalar FP64 of stage 1
    #   gram_sym_kernel<4,5,0>: stage 1 384 + stage 2 880 + stage 3 960 = 2,224 DMMA; 6,144 flop of operand products
    #   gram_sf_kernel<4,5,0>:  stage 2 1,760 + stage 3 1,600 = 3,360 DMMA; stage 1 in FMAs, 0.11 MFLOP
    exec_el = 2224 * 512 + 6144
    exec_el_all = 3360 * 512 + 5 * 400 * 55
    dmma, dfma, src = measured_fp64_peak() if rank == 0 else (37.0, 34.0, "")
    achieved = flops_el * nel / world / (ms * 1e-3) / 1e12  # per GPU, dense-equivalent
    executed = achieved * exec_el / flops_el
    executed_all = flops_el * nel / world / (ms_all * 1e-3) / 1e12 * exec_el_all / flops_el
    tr = _traffic().get("gram_sym_poisson_p4q5_bytes_per_element")
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
        # median of five separately timed steps: single calls are hit by host-side jitter on the shared boxes (one call in ten
        # took 20-45 ms instead of 11 in tools/e2e_asm_probe.py), which a mean over three steps turns into the headline
        reps, tsteps = 5, []
        for _ in range(reps):
            t0 = time.perf_counter()
            Ps.updateControlPoints(sub.cp)
            Ps.globalAssemble(capi.ASM_POISSON, xi, ww, out=(v, bb))
            tsteps.append(time.perf_counter() - t0)
        dt = float(np.median(tsteps))
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        asm_e2e = {"value": Ps.nelem * world / dt, "unit": "elements/s", "h2d_bytes_per_step": int(sub.cp.nbytes),
                   "d2h_bytes_per_step": int(v.nbytes + bb.nbytes),
                   "sample": "%d elements per rank per step (64 x 64 x %d), control net H2D, igab_global_assemble, CSR "
                             "values (%d non-zeros) + load vector D2H; median of %d steps" % (Ps.nelem, k_layers, v.size, reps),
                   "ms_per_step_min_median_max": [1e3 * min(tsteps), 1e3 * dt, 1e3 * max(tsteps)]}
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
            "all_blocks_kernel": {"kernel": "gram_sf_kernel<4,5,0> (IGAB_GRAM_SYM=0: every (i2, j2) block contracted; the "
                                            "default before gram_sym_kernel)",
                                  "value": nel / (ms_all * 1e-3), "unit": "elements/s",
                                  "executed_flops_per_element": exec_el_all, "achieved": executed_all,
                                  "frac": executed_all / dmma},
            "roofline": {"bound": "tensor", "kernel": "gram_sym_kernel<4,5,0> (evaluation + sum-factorised Gram contraction of "
                                                      "the upper block triangle, all three stages on the FP64 tensor cores, "
                                                      "mirror from shared memory; one kernel)",
                         "achieved": executed, "peak": dmma, "unit": "TFLOP/s", "frac": executed / dmma,
                         "traffic": (tr * per_call if tr else None), "traffic_source": _traffic().get("source_gram_sym"),
                         "algorithmic_bytes_per_launch": 8.0 * n * n * per_call,
                         "peak_source": src, "dfma_peak": dfma,
                         "algorithmic_flops_per_element": flops_el,
                         "executed_flops_per_element": exec_el,
                         "dense_equivalent": achieved, "dense_equivalent_frac": achieved / dmma,
                         "hbm_write_gbs": 8.0 * n * n * nel / world / (ms * 1e-3) / 1e9,
                         "equivalent_all_blocks": executed * exec_el_all / exec_el,
                         "equivalent_all_blocks_frac": executed * exec_el_all / exec_el / dmma,
                         "note": "achieved = the flops the kernel EXECUTES per element (2,224 DMMA instructions x 512 + "
                                 "6 kFLOP of scalar FP64 = 1.14 MFLOP; DMMA and DFMA share one FP64 datapath) x elements "
                                 "/ time, against the DMMA rate measured live by tools/fp64_peak: `frac` is the share of "
                                 "the FP64 tensor pipe the kernel sustains.  The kernel contracts only the blocks j2 >= i2 "
                                 "of the symmetric matrix: it executes 0.62x the flops of gram_sf_kernel (all_blocks_kernel, "
                                 "1.83 MFLOP per element, timed beside it), so its frac is lower at a higher throughput; "
                                 "equivalent_all_blocks rates its throughput in that kernel's flops.  dense_equivalent = the SURVEY 8(d) count of "
                                 "the dense B^T D B contraction it replaces (12.0 MFLOP per element): a value above the "
                                 "peak is the algorithmic saving of sum factorisation, not a roofline fraction.  "
                                 "hbm_write_gbs is the K_e stream (125 kB per element); traffic = ncu dram bytes read + "
                                 "written per launch"}}


def median_ms(fn, stream, reps=5, warm=2, target_ms=2.0):
    """Median device time of one call of fn in ms (CUDA events on `stream`).  Short calls are queued back to back
    between the two events -- as they are in a real element loop -- so that a sample lasts >= target_ms."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn()
    e1.record(stream)
    torch.cuda.synchronize()
    inner = max(1, min(200, int(target_ms / max(e0.elapsed_time(e1), 1e-3))))
    ts = []
    for _ in range(reps):
        e0.record(stream)
        for _ in range(inner):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    return float(np.median(ts)), inner


def bench_configs(peak_hbm, dmma_peak):
    """The other configs of BASELINE.json on one GPU, device-resident, each with its own roofline from CUDA events:
    C1 (2-D p=2 256^2), C3 elasticity (p=4, n=375), C4 (shell, second derivatives), C5 (tensor part + trimmed point lists).
    C2 is the headline line, C3 Poisson the `assembly` object."""
    import torch
    import igab200  # noqa: F401
    from dune_iga_b200 import capi, patchdata as PD
    stream = torch.cuda.current_stream()
    out = []

    def eval_cfg(name, pd, nq1, order, fields, kernel_note):
        P = capi.Patch.fromPatchData(pd)
        P.setKernel(capi.KERNEL_TENSOR)  # a tuned kernel or nothing
        xi = [PD.gaussLegendre01(n)[0] for n in nq1]
        nq = int(np.prod(nq1))
        npts = P.nelem * nq
        sh = P.shapes(npts)
        o = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in fields}
        l0 = capi.launch_count()
        ms, inner = median_ms(lambda: P.evalTensor(xi, order=order, want=fields, out=o, stream=stream.cuda_stream), stream)
        nb, dim, dw, nh = P.nb, P.dim, P.dimworld, P.nh
        per = dict(N=nb, dN=nb * dim, d2N=nb * nh, x=dw, Jt=dim * dw, H=nh * dw, detJ=1, JinvT=dim * dw)
        bpp = 8 * sum(per[f] for f in fields) + 8 * nb * (dw + 1) / nq
        gbs = bpp * npts / ms / 1e6
        out.append({"config": name, "metric": "quad-pt basis+Jacobian evals/sec", "value": npts / ms * 1e3,
                    "unit": "points/s", "elements": P.nelem, "points": npts, "ms_per_launch": ms,
                    "launches_queued_per_sample": inner, "outputs": "+".join(fields), "gpu_launches": int(capi.launch_count() - l0),
                    "roofline": {"bound": "hbm", "kernel": kernel_note, "achieved": gbs, "peak": peak_hbm, "unit": "GB/s",
                                 "frac": gbs / peak_hbm, "traffic": None, "algorithmic_bytes_per_point": bpp}})
        del o
        P.close()

    F1 = ("N", "dN", "x", "Jt", "detJ")
    eval_cfg("C1: 2-D quarter plate with hole p=2, 256x256 elements, 3x3 Gauss, order 1", PD.plateWithHole(7, 8), [3, 3], 1,
             F1, "eval_2d_kernel<2,2,3,1,false> (175 MB per launch: launch-latency share matters)")
    eval_cfg("C4: torus shell p=2 in R^3, 1024x1024 elements, 4x4 Gauss, order 2 (second derivatives, H)",
             PD.torusSurface(8), [4, 4], 2, ("N", "dN", "d2N", "x", "Jt", "H", "detJ"), "eval_2d_kernel<3,2,4,2,false>")
    eval_cfg("C5 tensor part: 2-D square plate p=3, 1024x1024 elements, 4x4 Gauss, order 1", PD.squarePlate(3, 10), [4, 4],
             1, F1, "eval_2d_kernel<2,3,4,1,false>")
    # C5 trimmed batches: 5 % of 512^2 elements with 3..40 sub-triangles x 6-point rule each (fillquadraturerule.hh:13-19)
    pd = PD.squarePlate(3, 9)
    P = capi.Patch.fromPatchData(pd)
    rng = np.random.default_rng(42)
    nel_t = P.nelem // 20
    elems = np.sort(rng.choice(P.nelem, nel_t, replace=False)).astype(np.int32)
    per = rng.integers(3, 41, nel_t) * 6
    elem = np.repeat(elems, per)
    d_elem, d_xi = torch.from_numpy(elem).cuda(), torch.from_numpy(rng.random((len(elem), 2))).cuda()
    sh = P.shapes(len(elem))
    o = {f: torch.empty(sh[f], dtype=torch.float64, device="cuda") for f in F1}
    l0 = capi.launch_count()
    ms, inner = median_ms(lambda: P.evalPoints(d_elem, d_xi, order=1, want=F1, out=o, stream=stream.cuda_stream), stream)
    bpp = 8 * (16 * 3 + 2 + 4 + 1) + 8 * 3  # outputs + the point's (elem, xi) read
    gbs = bpp * len(elem) / ms / 1e6
    out.append({"config": "C5 trimmed batches: 5 % of 512x512 elements (13,107), 18..240 own points each (point lists)",
                "metric": "quad-pt basis+Jacobian evals/sec", "value": len(elem) / ms * 1e3, "unit": "points/s",
                "elements": int(nel_t), "points": int(len(elem)), "ms_per_launch": ms, "launches_queued_per_sample": inner,
                "outputs": "+".join(F1), "gpu_launches": int(capi.launch_count() - l0),
                "roofline": {"bound": "hbm", "kernel": "eval_2d_kernel<2,3,4,1,true> (A2.3 in registers per point)",
                             "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm, "traffic": None,
                             "algorithmic_bytes_per_point": bpp}})
    del o, d_elem, d_xi
    P.close()
    # C3 elasticity: p=4, n=375, isotropic D (E=1, nu=0.3)
    pd = PD.thickCylinder(4, (5, 5, 5))
    P = capi.Patch.fromPatchData(pd)
    x, w = PD.gaussLegendre01(5)
    lam, mu = 0.3 / (1.3 * 0.4), 1 / 2.6
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] = lam + 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    n, nel = 375, 2368  # 16 x 148 elements = 2.66 GB of element matrices
    Ke = torch.empty((nel, n, n), dtype=torch.float64, device="cuda")
    l0 = capi.launch_count()
    ms, inner = median_ms(lambda: P.assemble(capi.ASM_ELASTICITY, [x] * 3, [w] * 3, D=D, elem_end=nel, Ke=Ke, fe=None,
                                             stream=stream.cuda_stream), stream, reps=3, warm=1)
    flops = 125 * (2 * 6 * n * n + 2 * 36 * n)
    t = _traffic()
    tr = t.get("gram_sf_elasticity_p4q5_bytes_per_element")
    ex = t.get("gram_sf_elasticity_p4q5_executed_flops_per_element", 9 * (3360 * 512) + 9 * 5 * 400 * 55)
    out.append({"config": "C3 elasticity: 3-D linear-elasticity element stiffness p=4 (n=375), 5^3 Gauss, isotropic D",
                "metric": "element matrices/sec", "value": nel / ms * 1e3, "unit": "elements/s", "elements": nel,
                "ms_per_launch": ms, "launches_queued_per_sample": inner, "gpu_launches": int(capi.launch_count() - l0),
                "roofline": {"bound": "tensor", "kernel": "gram_ws_kernel<4,5,2> (warp-specialised: DMMA warps + helper warps, complete rows stored)",
                             "achieved": ex * nel / ms / 1e9, "peak": dmma_peak, "unit": "TFLOP/s",
                             "frac": ex * nel / ms / 1e9 / dmma_peak,
                             "traffic": tr * nel if tr else None, "algorithmic_bytes": 8.0 * n * n * nel,
                             "algorithmic_flops_per_element": flops, "executed_flops_per_element": ex,
                             "dense_equivalent_tflops": flops * nel / ms / 1e9,
                             "hbm_write_gbs": 8.0 * n * n * nel / ms / 1e6,
                             "note": "achieved = flops the kernel EXECUTES (DMMA + FMA, shared FP64 datapath) over the "
                                     "measured DMMA peak; dense_equivalent_tflops is the SURVEY 8(d) count of the dense "
                                     "B^T D B it replaces"}})
    del Ke
    P.close()
    return out


def _traffic():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def host_bandwidth_probe(world, barrier, dist, nbytes=1 << 30, reps=4):
    """Plain cudaMemcpyAsync D2H into page-locked memory (igab_host_alloc: NUMA-local to the GPU), all ranks at once:
    the ceiling of any end-to-end path whose outputs go to host memory.  Returns (aggregate GB/s, this rank's NUMA node)."""
    import torch
    from dune_iga_b200 import capi
    h = capi.pinned_empty(nbytes // 8)
    d = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    ht = torch.from_numpy(h)
    ht.copy_(d, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        ht.copy_(d, non_blocking=True)
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    capi.pinned_free(h)
    return nbytes * reps * world / dt / 1e9, int(capi.lib().igab_device_numa_node())


def bench_multi_gpu(args, rank, world, P, pd, xi, barrier, dist):
    """The collectives of the path, timed and checked on all ranks through the PRODUCT's communicator (igab_comm_create):
    (a) patch volume and L2 norm / energy of a field over the rank's C2 slab, all-reduced inside igab_reduce_volume /
    igab_reduce_norms (ncclAllReduce on the call's stream), against closed forms; (b) the C3 patch (64^3 elements, p=4)
    sharded into k-slabs: igab_global_assemble per slab + igab_exchange_interface_rows (ncclSend / ncclRecv + ordered
    add), against this rank's own single-GPU assembly of the same rows."""
    import torch
    import igab200  # noqa: F401
    from dune_iga_b200 import capi, partition, patchdata as PD
    comm = capi.Communicator.fromTorch(dist, device="cuda")
    crank, cworld = comm.info()
    gw = PD.gaussLegendre01(NQ1)[1]
    ww = [gw, gw, gw]
    eb, ee = partition.slab_range(pd.elementsPerDirection, world, rank)
    stream = torch.cuda.current_stream()
    res = {"comm_nranks": cworld, "communicator": "ncclComm_t from igab_comm_unique_id / igab_comm_create (C-ABI)"}

    def timed(fn, reps=3):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return r, dt

    # (a) reductions.  Field u_i = x-coordinate of control point i is reproduced exactly (isoparametric): u_h = x, so
    # ||u_h||^2 = int x^2 dV = L pi/4 (r_o^4 - r_i^4)/4 and |grad u_h|^2 integrates to the volume.
    vol, t_vol = timed(lambda: P.volume(xi, ww, eb, ee, nccl_comm=comm, stream=stream.cuda_stream))
    exact_vol = np.pi / 4 * (2.0 ** 2 - 1.0 ** 2) * 3.0
    u = torch.from_numpy(np.ascontiguousarray(pd.cp[:, 0])).cuda()
    nrm, t_nrm = timed(lambda: P.norms(xi, ww, u, elem_begin=eb, elem_end=ee, nccl_comm=comm, stream=stream.cuda_stream))
    exact_l2 = 3.0 * np.pi / 4 * (2.0 ** 4 - 1.0 ** 4) / 4
    vols = [None] * world
    if world > 1:
        dist.all_gather_object(vols, (vol, float(nrm[0]), float(nrm[1])))
    else:
        vols = [(vol, float(nrm[0]), float(nrm[1]))]
    npts = (ee - eb) * NQ * world
    res["reductions"] = {
        "workload": "volume and L2 norm / energy over every rank's slab of the bench patch (%d points in total), one "
                    "ncclAllReduce of 1 / 2 doubles inside the call" % npts,
        "volume": vol, "volume_exact": exact_vol, "volume_rel_err": abs(vol - exact_vol) / exact_vol,
        "l2_sq": float(nrm[0]), "l2_sq_exact": exact_l2, "l2_rel_err": abs(float(nrm[0]) - exact_l2) / exact_l2,
        "energy": float(nrm[1]), "energy_rel_err": abs(float(nrm[1]) - exact_vol) / exact_vol,
        "identical_on_all_ranks": len(set(vols)) == 1,
        "ms_volume": 1e3 * t_vol, "ms_norms": 1e3 * t_nrm, "points_per_s_volume": npts / t_vol,
        "points_per_s_norms": npts / t_nrm}
    ok = res["reductions"]["volume_rel_err"] < 1e-9 and res["reductions"]["l2_rel_err"] < 1e-9 and \
        res["reductions"]["energy_rel_err"] < 1e-9 and res["reductions"]["identical_on_all_ranks"]

    # (b) sharded global assembly of C3 (strong: ONE 64^3 patch over all ranks) + interface rows
    p, nq1, lev = 4, 5, 6
    pd3 = PD.thickCylinder(p, (lev, lev, lev))
    P3 = capi.Patch.fromPatchData(pd3)
    x5, w5 = PD.gaussLegendre01(nq1)
    xi5, ww5 = [x5] * 3, [w5] * 3
    rowptr, _ = P3.csrPattern(1)
    nnz, nrows = int(rowptr[-1]), len(rowptr) - 1
    e0, e1 = partition.slab_range(pd3.elementsPerDirection, world, rank)
    vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    b = torch.empty(nrows, dtype=torch.float64, device="cuda")

    def sharded():
        P3.globalAssemble(capi.ASM_POISSON, xi5, ww5, out=(vals, b), elem_begin=e0, elem_end=e1, stream=stream.cuda_stream)
        P3.exchangeInterfaceRows(comm, values=vals, b=b, stream=stream.cuda_stream)

    sharded()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 3
    ms_asm = ms_x = 0.0
    for _ in range(reps):
        ev[0].record(stream)
        P3.globalAssemble(capi.ASM_POISSON, xi5, ww5, out=(vals, b), elem_begin=e0, elem_end=e1, stream=stream.cuda_stream)
        ev[1].record(stream)
        P3.exchangeInterfaceRows(comm, values=vals, b=b, stream=stream.cuda_stream)
        ev[2].record(stream)
        barrier()
        ms_asm += ev[0].elapsed_time(ev[1]) / reps
        ms_x += ev[1].elapsed_time(ev[2]) / reps
    tt = torch.tensor([ms_asm, ms_x, ms_asm + ms_x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_asm, ms_x, ms_tot = [float(v) for v in tt.tolist()]
    # check: this rank's own single-GPU assembly of the whole patch, compared on the rows its slab touches
    full = torch.empty(nnz, dtype=torch.float64, device="cuda")
    bfull = torch.empty(nrows, dtype=torch.float64, device="cuda")
    P3.globalAssemble(capi.ASM_POISSON, xi5, ww5, out=(full, bfull), stream=stream.cuda_stream)
    rb, re = partition.touched_row_ranges(pd3, world)
    z0, z1 = int(rowptr[rb[rank]]), int(rowptr[re[rank]])
    err = float((vals[z0:z1] - full[z0:z1]).abs().max() / full.abs().max())
    errb = float((b[rb[rank]:re[rank]] - bfull[rb[rank]:re[rank]]).abs().max() / bfull.abs().max())
    te = torch.tensor([err, errb], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    err, errb = [float(v) for v in te.tolist()]
    shared_rows = sum(max(0, min(re[rank], re[s]) - max(rb[rank], rb[s])) for s in range(world) if s != rank)
    res["sharded_assembly"] = {
        "workload": "C3 Poisson p=4: ONE 64^3-element patch (262,144 elements, %d rows, %d non-zeros) in %d k-slabs; per rank "
                    "igab_global_assemble of its slab into the shared CSR pattern + load vector, then "
                    "igab_exchange_interface_rows" % (nrows, nnz, world),
        "ms_assemble_slab": ms_asm, "ms_exchange": ms_x, "ms_total": ms_tot,
        "elements_per_s": 64 ** 3 / (ms_tot * 1e-3),
        "interface_bytes_sent_rank0": int(8 * sum(
            max(0, int(rowptr[min(re[0], re[s])]) - int(rowptr[max(rb[0], rb[s])])) for s in range(1, world))),
        "shared_rows_this_rank": int(shared_rows),
        "max_rel_err_values_vs_single_gpu": err, "max_rel_err_b_vs_single_gpu": errb}
    ok = ok and err < 1e-12 and errb < 1e-12
    res["checked"] = bool(ok)
    del vals, b, full, bfull
    P3.close()
    comm.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-assembly", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 128^3 elements per GPU (default); strong: the one 128^3 patch split over the ranks")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config lines (C1, C3 elasticity, C4, C5)")
    ap.add_argument("--no-multi-gpu", action="store_true", help="skip the timed NCCL reductions / interface-row exchange")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries write banners to file descriptor 1 (NCCL prints its
    # version there when the first communicator is created), so everything before the final print goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (line + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    import igab200  # noqa: F401
    from dune_iga_b200 import capi, partition, patchdata as PD

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    strong = args.scaling == "strong"
    pd = build_patch(world, args.scaling)
    P = capi.Patch.fromPatchData(pd)
    assert P.nelem == 128 * 128 * 128 * (1 if strong else world)
    P.setKernel(capi.KERNEL_TENSOR)  # fail loudly if the tuned kernel is not the one running
    gx = PD.gaussLegendre01(NQ1)[0]
    xi = [gx, gx, gx]
    my_begin, my_end = partition.slab_range(pd.elementsPerDirection, world, rank)
    my_layers = (my_end - my_begin) // (128 * 128)
    layers_per_launch = min(LAYERS_PER_LAUNCH, my_layers)
    el_per_launch = 128 * 128 * layers_per_launch
    pts_per_launch = el_per_launch * NQ
    launches_per_step = my_layers // layers_per_launch
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
    pts_per_rank = (my_end - my_begin) * NQ
    pts_per_step = pts_per_rank * world
    value = pts_per_step * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (eval_sf3d_kernel<3,4>, > 99 % of the step, profiles/launches_r02.csv):
    # average launch duration over the timed region, CUDA events on the launch stream (each interval = 1 main kernel;
    # the table kernels run once per rule and are cached)
    k_ms = float(per_launch.mean())
    peak, peak_src = measured_peak_hbm()
    achieved = ALG_BYTES_PER_PT * pts_per_launch / (k_ms * 1e-3) / 1e9
    tj = _traffic()
    traffic = tj.get("eval_sf3d_p3q4_bytes_per_launch")
    if traffic is not None and layers_per_launch != LAYERS_PER_LAUNCH:
        traffic = traffic * layers_per_launch / LAYERS_PER_LAUNCH
    roofline = {"bound": "hbm", "kernel": "eval_sf3d_kernel<3,4>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": tj.get("source"),
                "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, copy read+write)",
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_PT * pts_per_launch, "ms_per_launch": k_ms,
                "ms_per_launch_min_median_max": [float(per_launch.min()), float(np.median(per_launch)),
                                                 float(per_launch.max())],
                "launches_timed": int(len(per_launch))}

    # ---- e2e through the C-ABI with HOST buffers (rank-local; control net H2D + all outputs D2H inside the region).
    # The host buffers come from igab_host_alloc (page-locked, on the NUMA node of this rank's GPU); `ceiling` is what a
    # plain cudaMemcpyAsync D2H of the same bytes reaches with all ranks copying at once, measured just before.
    e2e = None
    if not args.no_e2e:
        probe_gbs, numa = host_bandwidth_probe(world, barrier, dist if world > 1 else None)
        e_layers = min(2, my_layers)
        e_el = 128 * 128 * e_layers
        e_pts = e_el * NQ
        hs = P.shapes(e_pts)
        host = {f: capi.pinned_empty(hs[f]) for f in FIELDS}
        cp_host = capi.pinned_empty(pd.cp.shape)
        cp_host[...] = pd.cp
        d2h = sum(int(np.prod(hs[f])) * 8 for f in FIELDS)
        h2d = cp_host.nbytes
        n_chunks = max(1, my_layers // e_layers)

        def e2e_step(s):
            P.updateControlPoints(cp_host)
            eb = my_begin + (s % n_chunks) * e_el
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
        e2e_val = e_pts * n_e2e * world / dt
        numas = [numa]
        if world > 1:
            numas = [None] * world
            dist.all_gather_object(numas, numa)
        ceiling_pts = probe_gbs * 1e9 / ((d2h + h2d) / e_pts)
        e2e = {"value": e2e_val, "unit": "points/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "ceiling": {"value": ceiling_pts, "unit": "points/s", "d2h_gbs_all_ranks": probe_gbs,
                           "frac_of_ceiling": e2e_val / ceiling_pts, "numa_node_of_each_rank_gpu": numas,
                           "how": "plain cudaMemcpyAsync D2H of 1 GiB x 4 into igab_host_alloc memory on every rank at "
                                  "once (aggregate GB/s), divided by the bytes per point this path moves over PCIe"},
               "sample": "%d k-layers per rank per e2e step (%d points), control net H2D + all outputs D2H to pinned "
                         "host memory through igab_eval_tensor(IGAB_HOST), 512 MB chunks, copy of chunk k overlapping "
                         "the kernel of chunk k+1; PCIe-bound" % (e_layers, e_pts)}
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
            return P.volume(xi, [gw, gw, gw], my_begin, my_end, stream=stream.cuda_stream)

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
        e2e_dev = {"value": pts_per_step * n_dev / dt, "unit": "points/s",
                   "h2d_bytes_per_step": int(cp_host.nbytes), "d2h_bytes_per_step": 8,
                   "result": "slab volume %.12f (exact pi/4*(r_o^2-r_i^2)*L/%d = %.12f)" % (
                       vol, world, np.pi / 4 * 3.0 * 3.0 / world),
                   "note": "all points of the rank evaluated per step with N+dN+x+Jt+detJ written to HBM (ring "
                           "buffers) for a device-side consumer; only the reduction crosses PCIe"}
        capi.pinned_free(cp_host)

    # ---- the collectives of the path on all ranks (also with one rank: a single-rank communicator)
    multi = None
    if not args.no_multi_gpu:
        if world == 1:
            try:  # one rank: a side measurement must not take the headline line down (N > 1: fail loudly, torchrun ends all ranks)
                multi = bench_multi_gpu(args, rank, world, P, pd, xi, barrier, dist)
            except Exception as ex:  # noqa: BLE001
                multi = {"error": "%s: %s" % (type(ex).__name__, ex)}
        else:
            multi = bench_multi_gpu(args, rank, world, P, pd, xi, barrier, dist)

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

    del ring
    # ---- second half of BASELINE.json's metric: element matrices/sec (config C3: p=4 Poisson stiffness, 5^3 Gauss,
    # FP64 tensor-core contraction), each rank assembling its own slab of a 64 x 64 x 64N patch
    assembly = None
    if not args.no_assembly:
        assembly = bench_assembly(args, rank, world, local, barrier, dist if world > 1 else None)

    # ---- the remaining configs of BASELINE.json, one GPU each (N = 1 line only)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        try:
            dm = assembly["roofline"]["peak"] if assembly else measured_fp64_peak()[0]
            configs = bench_configs(peak, dm)
        except Exception as ex:  # noqa: BLE001
            configs = [{"error": "%s: %s" % (type(ex).__name__, ex)}]

    if rank == 0:
        emit(json.dumps({
            "metric": "quad-pt basis+Jacobian evals/sec", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(world, args.scaling), "clocks": clocks, "e2e": e2e, "e2e_device_consumer": e2e_dev,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "multi_gpu": multi,
            "assembly": assembly, "configs": configs}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
