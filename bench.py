#!/usr/bin/env python
"""bench.py -- PVConv forward+backward points/sec (BASELINE.json metric).

Workload: one modules.PVConv(64, 64, kernel_size=3, resolution=32) in train mode,
features ~ N(0,1) [16,64,4096], coords ~ U(0,1.5)xU(0,1.5)xU(0,3.0) [16,3,4096], fixed grad_out ~ N(0,1),
seed 1588147245.  A "step" = zero grads, forward, backward (+ one NCCL all-reduce of the flat parameter
gradients when N > 1).  Weak scaling: B=16 per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|tf32] [--scaling weak|strong]

`--impl reference`: the reference has NO CPU implementation of this path (modules/functional/src/utils.hpp:7),
so the reference arm is the CPU oracle (oracle/: C restatement of the reference kernels + the same torch dense
ops the reference calls, on the host cores), on a bounded sample of the same workload.
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

B, N, C, R = 16, 4096, 64, 32
SEED = 1588147245
CONV_FLOPS = 2.0 * B * R ** 3 * C * C * 27  # one 3x3x3 conv pass at the metric shape (115.96 GFLOP)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: one long-running
    `nvidia-smi -lms 20` process (started before the region, killed after it); only the rows that arrived
    between mark_start() and stop() are used."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.t0 = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        if not self.proc:
            return
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.strip().split(",")]))

    def start(self):
        deadline = time.perf_counter() + 3.0
        while not self.rows and time.perf_counter() < deadline:  # first sample can take a second to appear
            time.sleep(0.02)
        self.t0 = time.perf_counter()

    def stop(self):
        t1 = time.perf_counter()
        time.sleep(0.05)
        if self.proc:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if self.t0 is not None and self.t0 <= t <= t1 + 0.03 and len(r) >= 7]
        if not rows:
            rows = [r for (_, r) in self.rows[-3:] if len(r) >= 7]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = sorted(v for v in (num(r[0]) for r in rows) if v is not None)
        mx = [v for v in (num(r[1]) for r in rows) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [v for v in (num(r[2]) for r in rows) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


def make_inputs(torch, device, batch=B):
    g = torch.Generator(device="cpu").manual_seed(SEED)
    feats = torch.randn(batch, C, N, generator=g)
    coords = torch.rand(batch, 3, N, generator=g) * torch.tensor([1.5, 1.5, 3.0]).view(1, 3, 1)
    gout = torch.randn(batch, C, N, generator=g)
    return feats, coords, gout


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    os.environ["PVCNN_B200_PRECISION"] = args.precision
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # bind this process (and therefore its first-touch pinned host buffers) to the GPU's NUMA node: measured on a 2-socket
    # box at N=1, the end-to-end step (host buffers copied every step) takes 3.7 ms unpinned vs 2.5 ms pinned
    from pvcnn_b200.parallel import pin_process_to_gpu_numa_node
    numa = pin_process_to_gpu_numa_node(local)
    import modules
    from pvcnn_b200 import _lib
    torch.manual_seed(SEED)
    m = modules.PVConv(C, C, 3, R).to(dev).train()
    params = [p for p in m.parameters()]
    if world > 1:
        for p in params:
            dist.broadcast(p.data, 0)
    strong = args.scaling == "strong"
    if strong and B % world:
        raise SystemExit("--scaling strong needs the global batch (%d) divisible by the number of GPUs" % B)
    bl = B // world if strong else B     # clouds per GPU: strong = global B=16 sharded, weak = B=16 on every GPU
    full = make_inputs(torch, dev)
    feats_h, coords_h, gout_h = [(t[rank * bl:(rank + 1) * bl] if strong else t).contiguous().pin_memory() for t in full]
    feats = feats_h.to(dev).requires_grad_(True)
    coords, gout = coords_h.to(dev), gout_h.to(dev)
    from pvcnn_b200.parallel import GradBucket
    # p.grad of every parameter is a view of ONE flat fp32 buffer; the fused backward writes into it directly
    bucket = GradBucket(params, dev).attach(m)

    def step(f, c, go):
        bucket.zero()   # replaces `p.grad = None` (the fused block overwrites its gradients, nothing to clear)
        f.grad = None
        out, _ = m((f, c))
        out.backward(go)
        # ONE collective per step: NCCL all-reduce (AVG) of the flat bucket, launched inside the backward right after the
        # last parameter gradient and overlapped with the input-gradient kernels; finish() orders the stream behind it
        bucket.finish()
        return out

    # L2 hygiene: the step streams > 2 GB of activations through a 126 MB L2, so consecutive steps
    # cannot serve each other's data from cache; no explicit flush is needed ("inputs larger than L2").
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None  # nvidia-smi -lms starts sampling during the warm-up
    for _ in range(args.warmup):
        step(feats, coords, gout)
    if sampler:
        sampler.start()
    barrier()  # every rank enters the timed region together (the sampler's settle time is before the barrier)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(feats, coords, gout)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (_lib.launch_count() - launches0) // args.steps
    clocks = sampler.stop() if sampler else None

    minimal = os.environ.get("PVCNN_BENCH_MINIMAL") == "1"  # profiling runs: timed loop only
    # ---- end-to-end: host buffers in, host buffers out, copies inside the timed region.
    # Every step copies ITS inputs from pinned host memory and returns ITS outputs (fused features + input gradient) to
    # pinned host memory; the copies run on two side streams (one per direction) with two buffer sets, so step i+1's upload and step i's
    # download overlap step i+1's compute (ordering by CUDA events; nothing is skipped or reused across steps).
    nbuf = 2
    out_h = [torch.empty(bl, C, N).pin_memory() for _ in range(nbuf)]
    gfe_h = [torch.empty(bl, C, N).pin_memory() for _ in range(nbuf)]
    f_d = [torch.empty(bl, C, N, device=dev, requires_grad=True) for _ in range(nbuf)]
    c_d = [torch.empty(bl, 3, N, device=dev) for _ in range(nbuf)]
    g_d = [torch.empty(bl, C, N, device=dev) for _ in range(nbuf)]
    res_o = [torch.empty(bl, C, N, device=dev) for _ in range(nbuf)]
    res_g = [torch.empty(bl, C, N, device=dev) for _ in range(nbuf)]
    # one stream per PCIe direction: uploads and downloads use different copy engines and the link is full duplex, so
    # step i+1's inputs and step i's results travel at the same time (a single copy stream serialises them: 68 MB per
    # step at ~29 GB/s took longer than the 2.07 ms of compute it was meant to hide behind)
    up_stream = torch.cuda.Stream(device=dev)
    down_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    up_done = [torch.cuda.Event() for _ in range(nbuf)]
    comp_done = [torch.cuda.Event() for _ in range(nbuf)]
    down_done = [torch.cuda.Event() for _ in range(nbuf)]

    def upload(k):
        with torch.cuda.stream(up_stream), torch.no_grad():
            up_stream.wait_event(comp_done[k])        # the previous user of this buffer set has finished computing
            f_d[k].copy_(feats_h, non_blocking=True)
            c_d[k].copy_(coords_h, non_blocking=True)
            g_d[k].copy_(gout_h, non_blocking=True)
            up_done[k].record(up_stream)

    def e2e_loop(nsteps):
        upload(0)
        for i in range(nsteps):
            k = i % nbuf
            if i + 1 < nsteps:
                upload((i + 1) % nbuf)                # next step's inputs travel while this step computes
            main_stream.wait_event(up_done[k])
            main_stream.wait_event(down_done[k])      # results of the step that used this set have left the device
            out = step(f_d[k], c_d[k], g_d[k])
            with torch.no_grad():
                res_o[k].copy_(out.detach())
                res_g[k].copy_(f_d[k].grad)
            comp_done[k].record(main_stream)
            with torch.cuda.stream(down_stream):
                down_stream.wait_event(comp_done[k])
                out_h[k].copy_(res_o[k], non_blocking=True)
                gfe_h[k].copy_(res_g[k], non_blocking=True)
                down_done[k].record(down_stream)
        up_stream.synchronize()
        down_stream.synchronize()

    for ev in comp_done + down_done:
        ev.record(main_stream)
    ms_e2e = float("nan")
    if not minimal:
        e2e_loop(max(3, args.warmup // 2))
        barrier()
        e0.record()
        e2e_loop(args.steps)
        main_stream.wait_stream(up_stream)
        main_stream.wait_stream(down_stream)   # the timed region ends after the last result has reached host memory
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1) / args.steps

    t = torch.tensor([ms, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- the same step in the other precision / with activity skipping off (both must be driver-visible).
    # PVCNN_B200_PRECISION and PVCNN_B200_SPARSE are read per call, so the modes run in this process on the same inputs.
    modes = {}
    if not minimal:
        def timed(nsteps, nwarm):
            for _ in range(nwarm):
                step(feats, coords, gout)
            barrier()
            e0.record()
            for _ in range(nsteps):
                step(feats, coords, gout)
            e1.record()
            barrier()
            tt = torch.tensor([e0.elapsed_time(e1) / nsteps], device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt[0])
        other = "tf32" if args.precision == "fp32" else "fp32"
        for name, prec, sparse in ((other, other, "1"), (args.precision + "_dense", args.precision, "0"),
                                   (other + "_dense", other, "0")):
            os.environ["PVCNN_B200_PRECISION"], os.environ["PVCNN_B200_SPARSE"] = prec, sparse
            t_ms = timed(max(5, args.steps // 2), 3)
            modes[name] = {"ms_per_step": t_ms, "value": world * bl * N / t_ms * 1e3, "unit": "points/s",
                           "precision": prec, "activity_skipping": sparse == "1"}
        os.environ["PVCNN_B200_PRECISION"] = args.precision
        os.environ.pop("PVCNN_B200_SPARSE", None)
        if world > 1 and not strong and B % world == 0:
            # strong scaling next to the (default) weak number: the SAME global batch of 16 clouds sharded over the ranks
            bs = B // world
            fs, cs, gs_ = [t[rank * bs:(rank + 1) * bs].contiguous().to(dev) for t in full]
            fs.requires_grad_(True)
            f_keep, c_keep, g_keep = feats, coords, gout
            feats, coords, gout = fs, cs, gs_
            t_ms = timed(max(5, args.steps // 2), 3)
            feats, coords, gout = f_keep, c_keep, g_keep
            modes["strong_scaling"] = {"global_batch": B, "per_gpu_batch": bs, "ms_per_step": t_ms,
                                       "value": B * N / t_ms * 1e3, "unit": "points/s"}

    extra = {}
    if rank == 0 and world == 1 and not minimal:
        extra = single_gpu_extras(torch, dev, m, args)
    if rank == 0:
        pk = peaks()
        line = {
            "metric": "PVConv fwd+bwd points/sec (B=16,N=4096,C=64,R=32)", "value": world * bl * N / ms * 1e3,
            "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32 (3xTF32 error-compensated tensor-core passes)" if args.precision == "fp32" else "tf32",
            "data": "synthetic",
            "config": {"workload": "single PVConv(64,64,k=3,R=32) block, train mode, fwd+bwd, B=%d/GPU N=4096%s" % (
                           bl, " (global B=16 sharded)" if strong else ""),
                       "parallelism": "dp%d" % world, "precision": args.precision, "numa_node": numa,
                       "l2": "activations (>2 GB/step) exceed the 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": world * bl * N / ms_e2e * 1e3, "unit": "points/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 4 * (2 * bl * C * N + bl * 3 * N), "d2h_bytes_per_step": 4 * 2 * bl * C * N},
            "gpu_launches": int(launches),
            "modes": modes,
            "peaks": pk,
        }
        ops_table = extra.pop("ops_vs_reference", None)
        line.update(extra)
        print(json.dumps(line))
        if ops_table is not None:   # bulky and informational: kept off the ONE stdout line (stderr)
            sys.stderr.write("[bench ops_vs_reference] " + json.dumps(ops_table) + "\n")
    if world > 1:
        dist.destroy_process_group()


def single_gpu_extras(torch, dev, m, args):
    """roofline of the dominant kernel (timed alone with CUDA events), GPU-reference arm, cpu_baseline."""
    from pvcnn_b200 import dense
    pk = peaks()
    extra = {}
    npass = 3 if args.precision == "fp32" else 1
    x = torch.randn(B, R, R, R, C, device=dev)
    w_hi, w_lo = dense.prep_weight(m.voxel_layers[3].weight)
    x_hi, x_lo = dense.split_tf32(x, want_hi=False)
    for _ in range(3):
        dense.igemm_conv(x_hi, x_lo, w_hi, w_lo, None, npass=npass)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        dense.igemm_conv(x_hi, x_lo, w_hi, w_lo, None, npass=npass)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / reps
    achieved = CONV_FLOPS / k_ms / 1e9
    tf32_peak = pk["bf16_tflops"] / 2.0   # kernel timed alone -> burst figure; kind::tf32 is the half-rate kind
    traffic, traffic_src = None, None
    try:  # dram bytes per launch of THIS tree's kernel, from the committed ncu --set full capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_conv_halo.json")))
        ent = prof["npass%d" % npass]
        traffic, traffic_src = ent["dram_bytes_read"] + ent["dram_bytes_write"], prof["source"]
    except Exception:  # noqa: BLE001
        pass
    extra["roofline"] = {
        "kernel": "conv_halo_kernel (3x3x3 conv fwd / dgrad, tcgen05 kind::tf32)", "bound": "tensor",
        "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
        "traffic": traffic, "kernel_ms": k_ms, "traffic_source": traffic_src,
        "note": "achieved = algorithmic FLOPs (%.2f GFLOP) / CUDA-event time of a dense launch run alone; executed tensor "
                "FLOPs are %dx that; peak = %s bf16 burst / 2 (no tf32 peak measured)" % (CONV_FLOPS / 1e9, npass, pk["source"]),
    }
    # the same kernel in the other precision (sub-result)
    npass2 = 1 if npass == 3 else 3
    for _ in range(3):
        dense.igemm_conv(x_hi, x_lo, w_hi, w_lo, None, npass=npass2)
    e0.record()
    for _ in range(reps):
        dense.igemm_conv(x_hi, x_lo, w_hi, w_lo, None, npass=npass2)
    e1.record()
    torch.cuda.synchronize()
    k2 = e0.elapsed_time(e1) / reps
    extra["roofline"]["other_precision"] = {"npass": npass2, "kernel_ms": k2, "achieved": CONV_FLOPS / k2 / 1e9,
                                            "frac": CONV_FLOPS / k2 / 1e9 / tf32_peak}
    # GPU reference arm (reference CUDA ops + cuDNN), informational
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))  # checker-side tool: runs oracle/_ref (reference .so)
        import ref_gpu_time
        extra["reference_gpu"] = [ref_gpu_time.run(True, steps=10, warmup=5), ref_gpu_time.run(False, steps=5, warmup=2)]
        extra["reference_gpu_note"] = "reference CUDA ops + cuDNN, same GPU and inputs: [0] its default (TF32), [1] fp32-strict"
    except Exception as e:  # noqa: BLE001
        extra["reference_gpu"] = {"unavailable": repr(e)[:200]}
    # test-time voting around the network (pvcnn_b200/evaluate.py): device arm with CUDA events next to
    # the reference's own host steps (numpy tiling / shuffle / gather + merge) on this box's cores; informational
    try:
        import voting_bench   # tests/tools: its host arm is a CPU baseline leg (reference numba function or oracle restatement)
        dv, hs = voting_bench.device_arm(20), voting_bench.host_arm(3)
        extra["voting"] = {"workload": "one S3DIS evaluation batch around the network (eval.py:149-179): 10 windows x 8192 "
                                       "points x 9 ch, 81920 voted points, 13 classes",
                           "device_ms": round(dv["ms_total"], 4), "device_GBps_algorithmic": round(dv["achieved_gbs"], 1),
                           "host_ms": round(hs["ms_total"], 2), "host_merge": hs["merge"],
                           "host_merge_counted": hs["merge_counted_in_total"]}
    except Exception as e:  # noqa: BLE001
        extra["voting"] = {"unavailable": repr(e)[:200]}
    if os.environ.get("PVCNN_BENCH_CONFIGS", "1") != "0":
        # whole networks and the classic-op table run in child processes under ONE time budget (the driver's scaling run
        # gives each N 870 s; these are informational and must never cost the line)
        deadline = time.perf_counter() + float(os.environ.get("PVCNN_BENCH_EXTRAS_S", "300"))
        extra["cuda_graph"] = graph_subresult(args, deadline)
        extra["configs"] = configs_subresults(args, deadline)
        extra["ops_vs_reference"] = ops_subresults(deadline)
    extra["cpu_baseline"] = cpu_baseline(sample_batch=2, iters=2)
    return extra


def _child(cmd, timeout):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    return p.returncode, [ln for ln in p.stdout.splitlines() if ln.startswith("{")], (p.stderr or "")[-120:]


def graph_subresult(args, deadline):
    """the metric step replayed as one CUDA graph (run_graph_probe in a child process): device-timed and end to end"""
    left = deadline - time.perf_counter()
    if left < 30:
        return {"unavailable": "sub-result time budget spent"}
    try:
        rc, lines, err = _child([sys.executable, os.path.abspath(__file__), "--graph-probe", "--steps", str(args.steps),
                                 "--warmup", str(args.warmup), "--precision", args.precision], min(120, left))
        if rc != 0 or not lines:
            return {"unavailable": "rc=%d %s" % (rc, err)}
        d = json.loads(lines[-1])
        return {"ms_per_step": round(d["ms_per_step"], 4), "e2e_ms_per_step": round(d["e2e_ms_per_step"], 4),
                "matches_eager": bool(d["matches_eager"] and d["e2e_host_output_matches"]),
                "max_rel_diff": float("%.2e" % max(d["rel_diff_vs_eager"].values())),
                "what": "same step as one CUDA-graph replay (graphs.GraphedTrainStep, child process); headline stays eager"}
    except subprocess.TimeoutExpired:
        return {"unavailable": "timeout"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:120]}


def ops_subresults(deadline):
    """The stand-alone ("classic") ops through the C ABI next to the reference's own CUDA kernels (oracle/_ref, built from
    the unmodified reference sources) on this GPU: tests/tools/ops_bench.py in a child process; median of 20 launches each,
    algorithmic GB/s where SURVEY.md 8d defines the bytes (profiles/r02_ops_vs_reference.md is the builder-run copy)."""
    left = deadline - time.perf_counter()
    if left < 30:
        return {"unavailable": "sub-result time budget spent"}
    try:
        rc, lines, err = _child([sys.executable, os.path.join(ROOT, "tests", "tools", "ops_bench.py")], min(150, left))
        rows = [json.loads(ln) for ln in lines]
        if not rows:
            return {"unavailable": "rc=%d %s" % (rc, err)}
        return {"columns": ["op", "ours_us", "reference_us", "ours_GBps_algorithmic"],
                "rows": [[r["op"], r["ours_us"], r["reference_us"], r.get("ours_GBps_algorithmic")] for r in rows]}
    except subprocess.TimeoutExpired:
        return {"unavailable": "timeout"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:120]}


def configs_subresults(args, deadline):
    """BASELINE.json configs 2-5 (whole networks, pvcnn_b200/zoo.py) as sub-results of the default line, so that the
    driver's own run records them: each is `bench.py --config <name>` in a child process (a failure or a hang there
    cannot touch this line), reduced to the numbers profiles/r02_configs.md tabulates: ms per step eager / as one CUDA
    graph / comparison arm (same network and weights on the stand-alone sm_100a point ops + torch cuDNN/cuBLAS dense layers
    with TF32 allowed).  10 timed steps after 3 warm-up, same precision mode as the line."""
    out = {"columns": ["ms_per_step", "cuda_graph_ms", "comparison_arm_ms", "gpu_launches"]}
    for name in ("s3dis_pvcnn", "shapenet_c0p25_train", "pvcnn2", "frustum_pvcnne"):
        left = deadline - time.perf_counter()
        if left < 30:
            out[name] = "time budget spent"
            continue
        try:
            rc, lines, err = _child([sys.executable, os.path.abspath(__file__), "--config", name, "--steps", "10",
                                     "--warmup", "3", "--precision", args.precision], min(150, left))
            if rc != 0 or not lines:
                out[name] = "unavailable: rc=%d %s" % (rc, err)
                continue
            d = json.loads(lines[-1])
            g, c = d.get("cuda_graph") or {}, d.get("comparison_arm") or {}
            rnd = lambda v: None if v is None else round(v, 4)   # noqa: E731
            out[name] = [rnd(d["ms_per_step"]), rnd(g.get("ms_per_step")), rnd(c.get("ms_per_step")), d.get("gpu_launches")]
        except subprocess.TimeoutExpired:
            out[name] = "unavailable: timeout"
        except Exception as e:  # noqa: BLE001
            out[name] = "unavailable: " + repr(e)[:120]
    return out


def _host_thread_candidates():
    """Thread counts worth trying for the CPU arm: every hardware thread, the scheduler affinity, the cgroup CPU quota,
    and a few smaller powers of two (oversubscription beyond the quota is much slower than fewer threads)."""
    total = os.cpu_count() or 1
    cand = {total}
    try:
        cand.add(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            cand.add(max(1, int(int(quota) / int(period) + 0.5)))
    except (OSError, ValueError):
        pass
    cand.update(t for t in (8, 16, 32, 64) if t < total)
    return sorted(c for c in cand if 1 <= c <= total)


def cpu_baseline(sample_batch=2, iters=2):
    """The oracle (CPU port of the reference kernels + torch CPU dense ops) timed on the host cores."""
    import numpy as np
    import torch
    import oracle
    torch.manual_seed(SEED)
    feats, coords, gout = make_inputs(torch, None, batch=sample_batch)
    ref = torch.nn.Sequential()  # parameters with the reference's shapes / default init
    conv1, conv2 = torch.nn.Conv3d(C, C, 3, padding=1), torch.nn.Conv3d(C, C, 3, padding=1)
    bn1, bn2 = torch.nn.BatchNorm3d(C, eps=1e-4), torch.nn.BatchNorm3d(C, eps=1e-4)
    cp, bnp = torch.nn.Conv1d(C, C, 1), torch.nn.BatchNorm1d(C)
    params = {"voxel_layers.0.weight": conv1.weight, "voxel_layers.0.bias": conv1.bias,
              "voxel_layers.1.weight": bn1.weight, "voxel_layers.1.bias": bn1.bias,
              "voxel_layers.3.weight": conv2.weight, "voxel_layers.3.bias": conv2.bias,
              "voxel_layers.4.weight": bn2.weight, "voxel_layers.4.bias": bn2.bias,
              "point_features.layers.0.weight": cp.weight, "point_features.layers.0.bias": cp.bias,
              "point_features.layers.1.weight": bnp.weight, "point_features.layers.1.bias": bnp.bias}
    params = {k: v.detach().numpy() for k, v in params.items()}
    f, c, g = feats.numpy(), coords.numpy(), gout.numpy()
    # "all the host threads it can use": the box reports every hardware thread of the node (os.cpu_count()), but the
    # container's CPU quota / affinity can be much smaller and torch's CPU conv3d collapses when oversubscribed (measured:
    # 4.2 s/step with 128 threads vs 0.16 s/step with 8 threads for the same sample).  So the thread count is tuned: one
    # timed pass per candidate, the fastest one is the baseline's configuration.
    oracle.pvconv_forward_backward(params, f, c, g, R, threads=min(_host_thread_candidates()))  # warm-up
    warm = 1
    best = None
    for th in _host_thread_candidates():
        t0 = time.perf_counter()
        oracle.pvconv_forward_backward(params, f, c, g, R, threads=th)
        dt1 = time.perf_counter() - t0
        warm += 1
        if best is None or dt1 < best[0]:
            best = (dt1, th)
    threads = best[1]
    t0 = time.perf_counter()
    for _ in range(iters):
        oracle.pvconv_forward_backward(params, f, c, g, R, threads=threads)
    dt = (time.perf_counter() - t0) / iters
    return {"value": sample_batch * N / dt, "unit": "points/s", "cores": threads, "kind": "port",
            "ms_per_step": dt * 1e3, "iterations": iters, "warmup_passes": warm,
            "sample": "B=%d of the 16 clouds, fwd+bwd, %d timed iterations after %d untimed passes (warm-up + thread-count "
                      "tuning); oracle port: the reference has no CPU path" % (sample_batch, iters, warm)}


def run_reference(args):
    """CPU arm: the oracle port on a bounded sample.  It runs (and REPORTS) its own iteration counts: at most 20 timed and
    3 warm-up passes of a B=2 sample, whatever --steps / --warmup ask for (requested values are echoed in `config`)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sample = 2
    iters = max(1, min(args.steps, 20))
    cb = cpu_baseline(sample_batch=sample, iters=iters)
    ms = cb["ms_per_step"]
    line = {"impl": "reference", "metric": "PVConv fwd+bwd points/sec (B=16,N=4096,C=64,R=32)", "value": cb["value"],
            "unit": "points/s", "n_gpus": args.gpus, "steps": iters, "warmup": cb["warmup_passes"], "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "single PVConv(64,64,k=3,R=32) block, train mode, fwd+bwd, bounded sample B=%d "
                                   "(NOT the B=16 metric batch: same_config=false)" % sample,
                       "parallelism": "host threads", "same_config": False, "sample_batch": sample,
                       "requested_steps": args.steps, "requested_warmup": args.warmup,
                       "what": "CPU oracle port (the reference has no CPU path); the reference's own CUDA+cuDNN path on "
                               "the same GPU is the `reference_gpu` key of the --impl ours line"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_graph_probe(args):
    """Child process of the default run (`modes.cuda_graph`): the SAME fwd+bwd step replayed as one CUDA graph
    (pvcnn_b200/graphs.py::GraphedTrainStep), checked against an eager step on the same inputs, timed on the device, and
    timed end to end with host buffers (two captured instances, double-buffered like the eager e2e loop).  Informational:
    the headline numbers of the line stay the eager ones."""
    import torch
    import modules
    from pvcnn_b200.graphs import GraphedTrainStep
    from pvcnn_b200.parallel import GradBucket, pin_process_to_gpu_numa_node
    os.environ["PVCNN_B200_PRECISION"] = args.precision
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    pin_process_to_gpu_numa_node(0)
    torch.manual_seed(SEED)
    m = modules.PVConv(C, C, 3, R).to(dev).train()
    feats_h, coords_h, gout_h = [t.contiguous().pin_memory() for t in make_inputs(torch, dev)]
    feats = feats_h.to(dev).requires_grad_(True)
    coords, gout = coords_h.to(dev), gout_h.to(dev)
    bucket = GradBucket(list(m.parameters()), dev).attach(m)
    for _ in range(3):   # eager reference step (also the warm-up of every lazily created buffer)
        bucket.zero()
        feats.grad = None
        out, _ = m((feats, coords))
        out.backward(gout)
    torch.cuda.synchronize()
    o0, g0, p0 = out.detach().clone(), feats.grad.detach().clone(), bucket.flat.clone()
    gts = [GraphedTrainStep(m, feats, coords, gout, bucket=bucket) for _ in range(2)]
    o1, g1 = gts[0]()
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    diffs = {"out": rel(o1, o0), "grad_features": rel(g1, g0), "param_grads": rel(bucket.flat, p0)}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(max(3, args.warmup)):
        gts[0].graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        gts[0].graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # end to end: pinned host buffers in and out every step, copies on two side streams (one per direction), two captured instances
    out_h = [torch.empty(B, C, N).pin_memory() for _ in range(2)]
    gfe_h = [torch.empty(B, C, N).pin_memory() for _ in range(2)]
    up_stream, down_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    up_done, comp_done, down_done = ([torch.cuda.Event() for _ in range(2)] for _ in range(3))

    def upload(k):
        with torch.cuda.stream(up_stream), torch.no_grad():
            up_stream.wait_event(comp_done[k])
            gts[k].f.copy_(feats_h, non_blocking=True)
            gts[k].c.copy_(coords_h, non_blocking=True)
            gts[k].go.copy_(gout_h, non_blocking=True)
            up_done[k].record(up_stream)

    def e2e_loop(nsteps):
        upload(0)
        for i in range(nsteps):
            k = i % 2
            if i + 1 < nsteps:
                upload((i + 1) % 2)
            main_stream.wait_event(up_done[k])
            main_stream.wait_event(down_done[k])
            gts[k].graph.replay()
            comp_done[k].record(main_stream)
            with torch.cuda.stream(down_stream):
                down_stream.wait_event(comp_done[k])
                out_h[k].copy_(gts[k].out, non_blocking=True)
                gfe_h[k].copy_(gts[k].f.grad, non_blocking=True)
                down_done[k].record(down_stream)
        up_stream.synchronize()
        down_stream.synchronize()

    for ev in comp_done + down_done:
        ev.record(main_stream)
    e2e_loop(4)
    torch.cuda.synchronize()
    e0.record()
    e2e_loop(args.steps)
    main_stream.wait_stream(up_stream)
    main_stream.wait_stream(down_stream)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    ok = max(diffs.values()) < 1e-4
    print(json.dumps({"ms_per_step": ms, "value": B * N / ms * 1e3, "e2e_ms_per_step": ms_e2e,
                      "e2e_value": B * N / ms_e2e * 1e3, "unit": "points/s", "steps": args.steps,
                      "precision": args.precision, "rel_diff_vs_eager": diffs, "matches_eager": ok,
                      "e2e_host_output_matches": rel(out_h[(args.steps - 1) % 2].to(dev), o0) < 1e-4}))


def run_config(args):
    """BASELINE.json configs 2-5: whole networks rebuilt from our modules (pvcnn_b200/zoo.py), synthetic inputs of the
    reference's shapes; one line per run.  `comparison_arm` = the same network and weights with the stand-alone sm_100a
    point ops around torch's cuDNN/cuBLAS dense layers (TF32 allowed, the reference's default precision)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pvcnn_b200 import zoo, _lib
    from pvcnn_b200.parallel import GradBucket, broadcast_parameters, pin_process_to_gpu_numa_node
    os.environ["PVCNN_B200_PRECISION"] = args.precision
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        # weak scaling, one process per GPU: the train config (3) is data parallel -- every rank a replica on its own
        # batch of the configured size, ONE NCCL all-reduce of the flat gradient bucket per step; the inference configs
        # (2, 4, 5) are replicas only, no collective
        dist.init_process_group("nccl", device_id=dev)
        pin_process_to_gpu_numa_node(dev.index)
    torch.manual_seed(SEED)
    model, spec = zoo.build(args.config)
    train = spec["mode"] == "train"
    model = model.to(dev).train(train)
    if world > 1:
        broadcast_parameters(model)
    g = torch.Generator().manual_seed(SEED + rank)
    x = zoo.synthetic_input(spec, g)
    x = {k: v.to(dev) for k, v in x.items()} if isinstance(x, dict) else x.to(dev)
    bsz, npts = spec["batch"], spec["points"]
    target = torch.randint(0, 50, (bsz, npts), generator=g).to(dev) if train else None
    opt = torch.optim.Adam(model.parameters(), lr=1e-3) if train else None   # configs/shapenet/__init__.py:43-44
    # data parallel: p.grad of every parameter is a view of one flat buffer (fused PVConv blocks write there directly,
    # the other layers accumulate through autograd), averaged by a single all-reduce in bucket.finish()
    bucket = GradBucket(list(model.parameters()), dev).attach(model) if (train and world > 1) else None

    def step():
        if train:
            if bucket is not None:
                bucket.zero()
            else:
                opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(x), target)
            loss.backward()
            if bucket is not None:
                bucket.finish()
            opt.step()
            return loss
        with torch.no_grad():
            return model(x)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, nwarm):
        np.random.seed(0)
        for _ in range(nwarm):
            step()
        barrier()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / nsteps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the slowest rank defines the step
        return float(t[0]), (_lib.launch_count() - l0) // nsteps

    sampler = ClockSampler(dev.index or 0) if rank == 0 else None
    if sampler:
        sampler.start()
    ms, launches = timed(args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    what = {"s3dis_pvcnn": "S3DIS PVCNN (1xC) forward", "shapenet_c0p25_train": "ShapeNet PVCNN (0.25xC) train step (Adam)",
            "pvcnn2": "S3DIS PVCNN++ forward", "frustum_pvcnne": "KITTI Frustum-PVCNN(E) end-to-end inference"}[args.config]
    if world > 1:
        if rank == 0:
            print(json.dumps({
                "metric": "%s points/sec (B=%d/GPU,N=%d)" % (what, bsz, npts), "value": world * bsz * npts / ms * 1e3,
                "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
                "dtype": "f32 (3xTF32)" if args.precision == "fp32" else "tf32",
                "config": {"workload": "%s, random-init weights, B=%d per GPU N=%d, pvcnn_b200/zoo.py:%s" % (
                               what, bsz, npts, args.config),
                           "precision": args.precision,
                           "parallelism": ("dp%d: one NCCL all-reduce (AVG) of the flat gradient bucket per step" % world)
                                          if train else ("%d independent replicas, no collective" % world)},
                "clocks": clocks, "gpu_launches": int(launches)}))
        dist.destroy_process_group()
        return
    graphed = None
    if not train:
        try:   # the same forward as one CUDA-graph replay (eager runs are bound by the host's launch rate)
            from pvcnn_b200.graphs import GraphedInference
            np.random.seed(0)
            gi = GraphedInference(model, x)
            for _ in range(3):
                gi(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                gi(x)
            e1.record()
            torch.cuda.synchronize()
            ms_g = e0.elapsed_time(e1) / args.steps
            graphed = {"ms_per_step": ms_g, "value": bsz * npts / ms_g * 1e3,
                       "what": "same forward captured once into a CUDA graph (pvcnn_b200/graphs.py), input copy + one replay per step"}
        except Exception as e:  # noqa: BLE001
            graphed = {"unavailable": repr(e)[:200]}
    os.environ["PVCNN_B200_PVCONV"], os.environ["PVCNN_B200_MLP"] = "composed", "torch"
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    ms_cmp, _ = timed(max(3, args.steps // 2), args.warmup)
    os.environ.pop("PVCNN_B200_PVCONV"); os.environ.pop("PVCNN_B200_MLP")
    print(json.dumps({
        "metric": "%s points/sec (B=%d,N=%d)" % (what, bsz, npts), "value": bsz * npts / ms * 1e3, "unit": "points/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "data": "synthetic",
        "dtype": "f32 (3xTF32)" if args.precision == "fp32" else "tf32",
        "config": {"workload": "%s, random-init weights, B=%d N=%d, pvcnn_b200/zoo.py:%s" % (what, bsz, npts, args.config),
                   "precision": args.precision},
        "clocks": clocks, "gpu_launches": int(launches), "cuda_graph": graphed,
        "comparison_arm": {"ms_per_step": ms_cmp, "value": bsz * npts / ms_cmp * 1e3,
                           "what": "same network/weights: stand-alone sm_100a point ops + torch cuDNN/cuBLAS dense layers, TF32 allowed"},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: B=16 clouds on every GPU (default, what the driver's scaling run uses); "
                         "strong: the global B=16 batch sharded over the GPUs (SURVEY.md 8d reports both)")
    ap.add_argument("--precision", default=os.environ.get("PVCNN_B200_PRECISION", "fp32"), choices=["fp32", "tf32"])
    ap.add_argument("--config", default="metric", choices=["metric", "s3dis_pvcnn", "shapenet_c0p25_train", "pvcnn2",
                                                            "frustum_pvcnne"],
                    help="metric (default): BASELINE.json's single-PVConv metric; the others: BASELINE configs 2-5 "
                         "(whole networks, 1 GPU)")
    ap.add_argument("--graph-probe", action="store_true", help=argparse.SUPPRESS)   # child of the default run
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.graph_probe:
        run_graph_probe(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.config != "metric":
        run_config(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
