#!/usr/bin/env python3
"""Headline benchmark: AlexNet-224 INT8 images/s on 1/2/4/8 B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path
                                                           # (oracle/_ref: its src/*.cc + stand-in GEMM)
  python bench.py --config simple_conv | fc_mnist          # BASELINE configs 2 / 1 (same line format)

A step = one pass of the INT8 hot path (quantise -> conv/fc stack -> dequantise, i.e.
i8ie.Module.__call__) over one synthetic batch. N=1: batch 100 (BASELINE config 3); N>1: global
batch 1000 sharded over the ranks (config 4), weights replicated, and the result exchange (logits +
top-1 agreement count) over NVLink peer memory — forward + exchange are ONE CUDA graph per step.
The model runs with the (scale, zero_point) the compiled reference calibrated, so the logits of the
timed path are compared BIT FOR BIT with the reference's own output (`parity`).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from int8inferenceengine_b200 import workloads as W  # noqa: E402

UNIT = "images/s"
SPEC_INT8_TOPS = 4500.0
WORKLOAD_TAG = {"alexnet": "alexnet224", "simple_conv": "simpleconv_cifar32", "fc_mnist": "fc_mnist784"}


def metric_name(topo):
    return f"{WORKLOAD_TAG[topo]}_int8_images_per_s"


def default_batch(topo, world):
    return 1000 if (topo == "alexnet" and world > 1) else 100


def workload_config(topo, gbatch, world):
    """The workload description — computed identically by both arms (the driver compares them)."""
    c, h, w = W.TOPOLOGIES[topo]["input"]
    lo, hi = W.TOPOLOGIES[topo]["range"]
    return {
        "workload": f"{WORKLOAD_TAG[topo]}_int8_b{gbatch}",
        "topology": topo,
        "global_batch": gbatch,
        "input": f"{c}x{h}x{w} fp32 U({lo},{hi}), seed 2 + rank; weights He-uniform seed 0",
        "calibration": "one batch of 100 (seed 1) through the fp32 path, ranges from the compiled reference's "
                       "own calibrator (calibrator.cc) — the B200 arm injects the same (scale, zero_point)",
        "l2": "inputs rotate over a ring of distinct batches totalling > 126 MB (L2) between timed steps",
        "sharding": f"batch split contiguously over {world} rank(s), weights replicated",
    }


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    if os.path.exists(p):
        d = json.load(open(p))
        out = {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
               "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    # kind::i8 tensor peak measured on this pool with a UTCIMMA-only kernel (tools/ubench/i8_peak.cu)
    p = os.path.join(ROOT, "profiles", "i8_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        out["i8_tops_burst"] = d.get("i8_tops_burst")
        out["i8_tops_sustained"] = d.get("i8_tops_sustained")
    return out


# ------------------------------------------------------------------------------------------
# clocks: poll NVML during the timed region
# ------------------------------------------------------------------------------------------
SAMPLE_PERIOD_S = float(os.environ.get("I8IE_BENCH_SAMPLE_MS", "10")) * 1e-3


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.active = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sw_power_cap": 0x4, "sync_boost": 0x10}
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self.active:
                    self.samples.append(mhz)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            # 10 ms: ~50 samples over the 0.5 s loaded window; a faster poll only competes with the launch
            # loop for the GIL and, at N ranks, N pollers queue on the driver's NVML lock
            time.sleep(SAMPLE_PERIOD_S)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# the reference on the host cores (oracle/_ref = its own src/*.cc, else the C port)
# ------------------------------------------------------------------------------------------
def host_threads():
    threads = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every worker; the reference is ONE process that is
    # meant to use all host cores (its conv loop is `omp parallel for` over images, conv2d.cc:125)
    if "TORCHELASTIC_RUN_ID" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    threads = int(os.environ["OMP_NUM_THREADS"])
    try:   # an OpenMP runtime that was initialised before the variable changed
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(threads)
    except Exception:  # noqa: BLE001
        pass
    return threads


def set_omp_threads(n):
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except Exception:  # noqa: BLE001
        pass


def pick_threads(model, probe, threads_all):
    """The reference gets the thread count that serves it best: all host cores for the conv nets
    (its conv loop is an OpenMP loop over images), one thread where the OpenMP fork/join costs more
    than the work (the 784->10 layer). Returns the chosen count; leaves the runtime set to it."""
    best_n, best_t = threads_all, None
    for n in ([threads_all, 1] if threads_all > 1 else [1]):
        set_omp_threads(n)
        model.forward_int8(probe)
        t0 = time.perf_counter()
        for _ in range(2):
            model.forward_int8(probe)
        t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best_n, best_t = n, t
    set_omp_threads(best_n)
    return best_n


def ref_model_and_qparams(topology):
    """Builds + calibrates the compiled reference (or the C port when oracle/_ref is absent).
    Returns (model, kind, qparams {layer: (scale, zp)}, calibrate+convert wall ms)."""
    from oracle import models, ref
    sd = W.make_weights(topology, 0)
    x_cal = W.make_images(topology, 100, 1)
    t0 = time.perf_counter()
    if ref.available():
        r = models.RefModel(topology, sd)
        r.calibrate(x_cal)
        ms = (time.perf_counter() - t0) * 1e3
        return r, "reference", r.qparams(W.make_images(topology, 2, 3)), ms
    p = models.PortModel(topology, sd)
    qp = p.calibrate_minmax(x_cal)
    p.convert(qp)
    ms = (time.perf_counter() - t0) * 1e3
    return p, "port", {k: (np.float32(v[0]), int(v[1])) for k, v in qp.items()}, ms


def time_cpu_forward(model, x, steps, warmup):
    for _ in range(warmup):
        model.forward_int8(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.forward_int8(x)
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    topo = args.config
    batch = args.batch or default_batch(topo, args.gpus)
    model, kind, _, cal_ms = ref_model_and_qparams(topo)
    threads = pick_threads(model, W.make_images(topo, min(batch, 16), 2), threads)
    # bounded sample: size the per-step sample so (steps + warmup) steps end within ~2.5 minutes
    probe = W.make_images(topo, 4, 2)
    model.forward_int8(probe)
    t0 = time.perf_counter()
    model.forward_int8(probe)
    per_img = (time.perf_counter() - t0) / 4
    budget = 150.0 / max(1, args.steps + args.warmup)
    sample = int(max(1, min(batch, budget / max(per_img, 1e-6))))
    x = W.make_images(topo, sample, 2)
    dt = time_cpu_forward(model, x, args.steps, args.warmup)
    val = sample / dt
    vnni = None
    try:
        import ctypes
        from oracle import ref
        vnni = bool(ctypes.CDLL(ref.so_path()).i8ie_shim_uses_vnni()) if kind == "reference" else None
    except Exception:  # noqa: BLE001
        pass
    line = {
        "impl": "reference", "metric": metric_name(topo), "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "u8*s8->s32",
        "data": "synthetic",
        "config": workload_config(topo, batch, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{sample} images/step of the batch-{batch} workload, {args.steps} steps; "
                                   f"reference C++ + stand-in GEMM (MKL unavailable offline), vnni={vnni}",
                         "calibrate_convert_ms": round(cal_ms, 2)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "sample_images_per_step": sample,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# our arm (CUDA)
# ------------------------------------------------------------------------------------------
def conv_fc_macs(topology, batch):
    """Per-layer algorithmic MACs with the reference's un-padded dims."""
    t = W.TOPOLOGIES[topology]
    c, h, w = t["input"]
    out = {}
    for op in t["ops"]:
        if op[0] == "conv":
            _, name, cin, cout, k, s, p = op
            oh, ow = W.conv_out_hw(h, w, k, s, p)
            out[name] = batch * oh * ow * cout * cin * k * k
            c, h, w = cout, oh, ow
        elif op[0] == "pool":
            h, w = (h - op[1]) // op[2] + 1, (w - op[1]) // op[2] + 1
        elif op[0] == "fc":
            out[op[1]] = batch * op[2] * op[3]
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: anything libraries print to fd 1 meanwhile (e.g. NCCL's
    # version banner when the box sets NCCL_DEBUG) is sent to stderr; print_line() restores fd 1
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def print_line(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(obj), flush=True)

    verbose = os.environ.get("I8IE_BENCH_VERBOSE") is not None
    t_start = time.time()

    def stage(msg):
        if verbose:
            print(f"[bench rank {rank} +{time.time() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    if verbose:   # a hung rank prints its Python stack to stderr
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("I8IE_BENCH_VERBOSE") or 90), exit=False, file=sys.stderr)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        stage("process group up")

    import i8ie
    from int8inferenceengine_b200 import _lib, backend as B, sharding
    from int8inferenceengine_b200.runner import build_module

    _lib.check(_lib.load().i8ie_device_check(), "device_check")
    topo = args.config
    gbatch = args.batch or default_batch(topo, world)
    assert gbatch % world == 0
    lbatch = gbatch // world
    in_shape = W.TOPOLOGIES[topo]["input"]
    ncls = [op for op in W.TOPOLOGIES[topo]["ops"] if op[0] == "fc"][-1][3]
    dev = torch.device("cuda", local)

    # ---- the oracle (rank 0): the compiled reference calibrates; its (scale, zp) go to every rank
    sd = W.make_weights(topo, 0)
    ref_model = ref_kind = None
    ref_cal_ms = None
    box = [None]
    if rank == 0:
        threads = host_threads()
        ref_model, ref_kind, qp, ref_cal_ms = ref_model_and_qparams(topo)
        box[0] = {k: (float(v[0]), int(v[1])) for k, v in qp.items()}
        stage(f"oracle ({ref_kind}) calibrated in {ref_cal_ms:.0f} ms")
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    qparams = {k: (np.float32(v[0]), int(v[1])) for k, v in box[0].items()}

    # ---- our own calibrate + convert, timed (BASELINE configs 1/2 include it; notebooks: %%time of
    # prepare(); model(x_cal); convert()). Rank 0 only; this model is then dropped.
    cal_ms = None
    own_qp = None
    if rank == 0:
        x_cal = W.make_images(topo, 100, 1)
        for rep in range(2):   # the second pass has warm plans / allocator
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m_cal = build_module(topo, sd, calib=x_cal)
            torch.cuda.synchronize()
            cal_ms = (time.perf_counter() - t0) * 1e3
        own_qp = {n: (float(L.layer._scale), int(L.layer._zp)) for n, L in m_cal.layers().items()}
        del m_cal
    model = build_module(topo, sd, qparams=qparams)
    stage("model built (reference qparams injected)")

    # ---- inputs: a ring of distinct device-resident batches larger than L2 (126 MB) in total
    bytes_per_batch = lbatch * int(np.prod(in_shape)) * 4
    ring = max(2, min(1024 if world == 1 else 64, int(np.ceil(260e6 / bytes_per_batch))))
    lo, hi = W.TOPOLOGIES[topo]["range"]

    def gen_batches(r, count):
        rng = np.random.default_rng(2 + r)
        return [rng.uniform(lo, hi, size=(lbatch,) + tuple(in_shape)).astype(np.float32) for _ in range(count)]

    dev_inputs, host_inputs = [], []
    for a in gen_batches(rank, ring):
        ht = torch.from_numpy(a).pin_memory()
        host_inputs.append(ht)
        dev_inputs.append(i8ie.Tensor(B.tensor_from_torch(ht)))

    # ---- oracle outputs: rank 0 runs the reference on ring slot 0 of EVERY rank (all slots at N=1):
    # expected logits for the bit-for-bit parity check and the reference argmax of the agreement count
    exp_logits = None          # rank 0: [gbatch, ncls] of slot 0 in rank order
    ref_arg_np = [None] * ring
    if rank == 0:
        t0 = time.perf_counter()
        slots = ring if world == 1 else 1
        per_rank = [[ref_model.forward_int8(x) for x in gen_batches(r, slots)] for r in range(world)]
        exp_logits = np.concatenate([per_rank[r][0] for r in range(world)], 0)
        exp_all_slots = per_rank[0] if world == 1 else None
        arg0 = [per_rank[r][0].argmax(1).astype(np.int64) for r in range(world)]
        if world == 1:
            ref_arg_np = [p.argmax(1).astype(np.int64) for p in per_rank[0]]
        stage(f"oracle forward of {world * slots * lbatch} images: {time.perf_counter() - t0:.1f} s")
    if world > 1:
        scat = [arg0 if rank == 0 else None]
        dist.broadcast_object_list(scat, src=0)
        ref_arg_np[0] = scat[0][rank]
    # N > 1, slots >= 1: the engine's own argmax of a warm-up run stands in (parity on slot 0 has
    # established that it equals the oracle's); the count's arithmetic and exchange are what is timed
    for j in range(ring):
        if ref_arg_np[j] is None:
            ref_arg_np[j] = model(dev_inputs[j]).numpy().argmax(1).astype(np.int64)
    ref_argmax = [torch.from_numpy(a).to(dev) for a in ref_arg_np]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        sync_all()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- the step. N = 1: the public call model(x). N > 1: forward + result exchange as ONE CUDA
    # graph per ring buffer (sharding.ShardedStep) — a single host enqueue per step.
    exchange = sharded = None
    alone_ms = None
    if world > 1:
        alone_ms = timed(lambda i: model(dev_inputs[i % ring]), args.steps, max(args.warmup, 3)) / args.steps
        exchange = sharding.make_exchange(gbatch, ncls, dev)
        sharded = sharding.ShardedStep(model, dev_inputs, ref_argmax, exchange)
        sync_all()
        stage(f"sharded step captured ({type(exchange).__name__}, {sharded.kernels_per_step} kernels per step)")

        def step(i):
            return sharded(i)
    else:
        def step(i):
            return model(dev_inputs[i % ring])

    # set-up, not warm-up: the first calls of a model run eagerly (they build plans) and every ring
    # buffer gets its CUDA graph captured on first use — all of that happens here, before the W warm-up steps
    if world == 1:
        for i in range(ring + 3):
            step(i)
    # the NVML poller is created and started BEFORE the warm-up (nvmlInit takes a rank-dependent number of
    # milliseconds: done after the barrier it made the ranks enter the timed loop that far apart, and with a
    # lock-step exchange in every step the early ranks' 50-step window absorbed the whole skew)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    stage("warm-up done")

    launches0 = _lib.launch_count() + model.graph_launches() + (sharded.launches() if sharded else 0)
    sampler.active = True
    if args.profiler_range:      # ncu --profile-from-start off: capture exactly the timed steps
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync_all()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    sampler.active = False
    ms = e0.elapsed_time(e1)
    stage("timed region done")
    launches = _lib.launch_count() + model.graph_launches() + (sharded.launches() if sharded else 0) - launches0
    _lib.check_tc_error()
    clock_window = "timed region"
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # the timed region is only a few milliseconds (NVML answers in ~ms): keep sampling over ~0.5 s of
    # the very same steps so that the clocks line describes the loaded state. Every step holds an
    # exchange at N > 1, so whether and how many extra steps run is decided identically on all
    # ranks (from the max-reduced time and a MAX-reduced flag), never per rank.
    need = 1 if (len(sampler.samples) < 5 and not args.profiler_range) else 0
    if world > 1:
        t = torch.tensor([need], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        need = int(t.item())
    if need:
        extra = max(args.steps, int(0.5 / max(ms / args.steps * 1e-3, 1e-6)))
        sampler.active = True
        for i in range(extra):
            step(i)
        sync_all()
        sampler.active = False
        clock_window = f"timed region + {extra} further identical steps (untimed, ~0.5 s)"
    ms_per_step = ms / args.steps
    value = gbatch / (ms_per_step * 1e-3)

    # ---- parity: the output of the TIMED path on ring slot 0 (gathered over all ranks at N > 1) vs
    # the compiled reference's logits for the same images, bit for bit; agreement count vs its argmax
    if world > 1:
        logits_all, agree = sharded(0)
        sync_all()
        got = logits_all.cpu().numpy()
        agree_n = int(agree.item())
    else:
        got = model(dev_inputs[0]).numpy()
        agree_n = int((got.argmax(1) == ref_arg_np[0]).sum())
    _lib.check_tc_error()
    parity = None
    if rank == 0:
        same = bool(np.array_equal(got, exp_logits))
        rows_checked = gbatch
        if world == 1:            # every ring slot at N = 1
            for j in range(1, ring):
                same = same and bool(np.array_equal(model(dev_inputs[j]).numpy(), exp_all_slots[j]))
                rows_checked += gbatch
        parity = {"rows": rows_checked, "bit_identical": same, "oracle": ref_kind,
                  "max_abs_diff": float(np.max(np.abs(got - exp_logits))) if got.shape == exp_logits.shape else None,
                  "argmax_equal": bool(np.array_equal(got.argmax(1), exp_logits.argmax(1))),
                  "top1_agreement_count": agree_n, "of": gbatch,
                  "what": "logits of the timed path (gathered over all ranks) vs the reference's own INT8 forward "
                          "(oracle/_ref: src/*.cc compiled unmodified) on the same images and (scale, zero_point)"}

    # ---- e2e: through the public API with HOST buffers, H2D + D2H inside the timed region
    def e2e_fn(inputs):
        def fn(i):
            out = model(i8ie.tensor(inputs[i % ring]))
            if world > 1:
                la, ag = exchange(out.data.buf.view(lbatch, ncls), ref_argmax[i % ring])
                return la.cpu() if rank == 0 else ag.cpu()
            return out.numpy()
        return fn

    # (8 warm-up calls: every chunk shape runs twice eagerly, then its graphs are captured per device address)
    ms_e2e = timed(e2e_fn(host_inputs), args.steps, 8)
    e2e_val = gbatch / (ms_e2e / args.steps * 1e-3)
    # the same with what a drop-in script passes: a pageable numpy array (blocking staged copy)
    np_inputs = [h.numpy().copy() for h in host_inputs[:2]] + [None] * (ring - 2)
    np_inputs = [np_inputs[i % 2] for i in range(ring)]
    ms_pg = timed(e2e_fn(np_inputs), max(3, args.steps // 4), 4)
    e2e_pageable = gbatch / (ms_pg / max(3, args.steps // 4) * 1e-3)
    sampler.stop()
    _lib.check_tc_error()

    # ---- per-layer timing + roofline of the dominant kernel (CUDA events on the launch stream)
    pk = peaks()
    layer_rows, roof = [], None
    if rank == 0:
        model.record = []
        model(dev_inputs[0])
        rec = model.record
        model.record = None
        macs = conv_fc_macs(topo, lbatch)
        ins = {}
        prev = i8ie.Tensor(B.quantize(dev_inputs[0].data, W.INPUT_SCALE, W.INPUT_ZP))
        for op, (tag, t) in zip([o for o in W.TOPOLOGIES[topo]["ops"] if o[0] != "flatten"], rec):
            if op[0] in ("conv", "fc"):
                ins[tag] = prev
            prev = t
        for op in W.TOPOLOGIES[topo]["ops"]:   # fc inputs need the flattened view
            if op[0] == "fc":
                ins[op[1]] = ins[op[1]].reshape(-1, op[2])
        reps = 10
        relu_after = {}
        ops = W.TOPOLOGIES[topo]["ops"]
        for i, op in enumerate(ops):
            if op[0] in ("conv", "fc"):
                relu_after[op[1]] = i + 1 < len(ops) and ops[i + 1][0] == "relu"
        for name, layer in model.layers().items():
            x_in = ins[name].data
            x_in.buf  # materialise the layer input outside the timed graph
            fwd = lambda: layer.layer._forward_u8(x_in, relu=relu_after[name])  # noqa: E731
            for _ in range(3):
                fwd()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):     # `reps` back-to-back launches of this layer's kernel(s)
                for _ in range(reps):
                    fwd()
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(5):
                a.record()
                g.replay()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b) / reps * 1e3)
            us = best
            tops = 2 * macs[name] / (us * 1e-6) / 1e12
            layer_rows.append({"layer": name, "us": round(us, 2), "tops": round(tops, 1)})

        # dominant kernel = the kernel FUNCTION with the largest share of the step (the ncu launch list
        # under profiles/ ranks the same way): group the layers by the kernel that serves them and
        # take achieved = algorithmic ops of its launches / their summed duration.
        def kernel_of(name):
            impl = getattr(model.layers()[name].layer, "_last_impl", None)
            if name.startswith("fc"):
                return "tc_igemm_kernel / fc_head_kernel (fully_connected)"
            return {3: "tc_stem2_kernel (fused input quantise + stem conv)",
                    2: "tc_igemm2_kernel (CTA-pair tcgen05 implicit GEMM + fused requant epilogue)",
                    4: "tc_igemm2_kernel (CTA-pair tcgen05 implicit GEMM + fused requant epilogue)",
                    1: "simt_igemm_kernel"}.get(impl, "tc_igemm_kernel")
        groups = {}
        for r in layer_rows:
            r["kernel"] = kernel_of(r["layer"])
            g = groups.setdefault(r["kernel"], {"us": 0.0, "ops": 0.0, "layers": []})
            g["us"] += r["us"]
            g["ops"] += 2 * macs[r["layer"]]
            g["layers"].append(r["layer"])
        kname, kg = max(groups.items(), key=lambda kv: kv[1]["us"])
        achieved = kg["ops"] / (kg["us"] * 1e-6) / 1e12
        traffic = traffic_src = None
        tpath = os.path.join(ROOT, "profiles", "ncu_top_kernel.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("layers") == kg["layers"] and tj.get("batch") == lbatch:
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "profiles/ncu_top_kernel.json (ncu capture of this kernel at this batch; not re-measured by this run)"
        bf16x2 = 2 * pk["bf16_tflops"]
        int8_peak = pk.get("i8_tops_burst") or bf16x2
        best_layer = max((r for r in layer_rows if r["layer"] in kg["layers"]), key=lambda r: r["tops"])
        roof = {"bound": "tensor", "kernel": kname, "launches_per_step": len(kg["layers"]), "layers": kg["layers"],
                "achieved": round(achieved, 1), "peak": int8_peak, "unit": "TOP/s",
                "frac": achieved / int8_peak, "traffic": traffic, "traffic_source": traffic_src,
                "share_of_step": round(kg["us"] / (ms_per_step * 1e3), 3),
                "best_launch": {"layer": best_layer["layer"], "tops": best_layer["tops"],
                                "frac": round(best_layer["tops"] / int8_peak, 4)},
                "frac_of_2x_bf16": round(achieved / bf16x2, 4), "frac_of_spec": round(achieved / SPEC_INT8_TOPS, 4),
                "peak_source": (f"kind::i8 tensor peak measured on this pool with a UTCIMMA-only kernel "
                                f"(profiles/i8_peak.json, burst clocks): {int8_peak} TOP/s" if pk.get("i8_tops_burst") else
                                f"2 x {pk['source']} cuBLAS bf16 burst ({pk['bf16_tflops']} TFLOP/s)") +
                               f"; 2 x bf16 = {bf16x2:.0f} TOP/s, spec {SPEC_INT8_TOPS:.0f} TOP/s",
                "note": "achieved = algorithmic ops (2*M*N*K, un-padded reference dims) of the kernel's launches in one "
                        "step / their summed duration, each launch replayed back-to-back from a CUDA graph and timed "
                        "with CUDA events on the launch stream; operands are L2-resident as in the real forward (the "
                        "producer just wrote them)"}

    # ---- HBM-bound kernels standalone on tensors far larger than L2 (north_star: >= 80 % of HBM)
    hbm_rows = []
    if rank == 0 and world == 1 and not args.no_hbm_kernels:
        L = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        n = 1 << 28
        xf = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-3, 3)
        xq = torch.empty(n, dtype=torch.uint8, device="cuda")
        xi = torch.randint(-200000, 200000, (n,), dtype=torch.int32, device="cuda")
        ws = torch.zeros(int(L.i8ie_minmax_workspace_bytes()), dtype=torch.uint8, device="cuda")
        mm = torch.empty(2, dtype=torch.float32, device="cuda")
        cases = [
            ("quantize_f32_u8", 5, lambda: L.i8ie_quantize_f32_u8(xf.data_ptr(), xq.data_ptr(), n, 0.025, 127, st)),
            ("dequantize_u8_f32", 5, lambda: L.i8ie_dequantize_u8_f32(xq.data_ptr(), xf.data_ptr(), n, 0.025, 127, st)),
            ("downscale_s32_u8", 5, lambda: L.i8ie_downscale_s32_u8(xi.data_ptr(), xq.data_ptr(), n, 0.025, 0.003, 0.05, 116, st)),
            ("minmax_f32", 4, lambda: L.i8ie_minmax_f32(xf.data_ptr(), n, mm.data_ptr(), ws.data_ptr(), st)),
            ("relu_u8", 2, lambda: L.i8ie_relu_u8(xq.data_ptr(), xq.data_ptr(), n, 127, st)),
        ]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for name, bpe, fn in cases:
            for _ in range(3):
                fn()
            best = 1e9
            for _ in range(5):
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            gbs = bpe * n / (best * 1e-3) / 1e9
            hbm_rows.append({"kernel": name, "elements": n, "bytes_per_element": bpe, "ms": round(best, 4),
                             "gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / pk["hbm_gbs"], 3)})
        del xf, xq, xi
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N=1 only): the compiled reference on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = pick_threads(ref_model, W.make_images(topo, 16, 2), threads)
        xs = W.make_images(topo, 4, 2)
        ref_model.forward_int8(xs)
        t0 = time.perf_counter()
        ref_model.forward_int8(xs)
        per_img = (time.perf_counter() - t0) / 4
        sample = int(max(4, min(gbatch, 6.0 / max(per_img, 1e-6))))
        dt = time_cpu_forward(ref_model, W.make_images(topo, sample, 2), 3, 1)
        cpu = {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": ref_kind,
               "sample": f"{sample}-image batch x 3 forwards (1 warm-up) of the {topo} INT8 forward; "
                         "reference C++ + stand-in GEMM (MKL unavailable offline)",
               "calibrate_convert_ms": round(ref_cal_ms, 2)}

    if rank == 0:
        line = {
            "metric": metric_name(topo), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u8*s8->s32",
            "data": "synthetic",
            "config": workload_config(topo, gbatch, world),
            "run": {"per_gpu_batch": lbatch, "ring_batches": ring, "ring_mb": round(ring * bytes_per_batch / 1e6),
                    "step": (("forward + result exchange captured as ONE CUDA graph per ring buffer (one host enqueue per "
                              "step); exchange = PeerExchange: pack kernel pushes [agreement count | logits] into every "
                              "rank's buffer with NVLink peer stores, unpack kernel waits for all ranks' step flags"
                              if type(exchange).__name__ == "PeerExchange" else
                              "forward captured as a CUDA graph per ring buffer, then ResultExchange: pack kernel, one NCCL "
                              "all-gather, unpack kernel (the peer-memory exchange could not be set up on this box)")
                             if world > 1 else "i8ie.Module.__call__ on a device-resident batch (CUDA-graph replay)"),
                    "collectives": "none on the data path; result exchange only" if world > 1 else "none",
                    "scaling_note": "N=1 runs BASELINE config 3 (batch 100), N>1 config 4 (batch 1000 sharded): compare "
                                    "N>1 lines with each other (strong scaling), not with the N=1 line"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": bytes_per_batch,
                    "d2h_bytes_per_step": (gbatch if world > 1 else lbatch) * ncls * 4},
            "e2e_pageable": {"value": e2e_pageable, "unit": UNIT,
                             "what": "same call with a pageable numpy batch (what an unmodified reference script passes): "
                                     "tensor() makes one blocking staged copy"},
            "gpu_launches": int(launches),
            "parity": parity,
            "calibrate_convert_ms": round(cal_ms, 2) if cal_ms is not None else None,
            "calibrated_qparams": {"reference": box[0], "b200_minmax": own_qp},
            "clocks": dict(sampler.summary(), window=clock_window),
            "roofline": roof,
            "cpu_baseline": cpu,
            "layers": layer_rows,
            "hbm_kernels": hbm_rows,
            "hbm_peak_gbs": pk["hbm_gbs"],
        }
        if alone_ms is not None:
            line["per_gpu_alone_ms"] = alone_ms
            line["exchange_ms"] = ms_per_step - alone_ms
        print_line(line)
    if world > 1:
        sync_all()
        if hasattr(exchange, "close"):
            exchange.close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="alexnet", choices=sorted(WORKLOAD_TAG),
                    help="BASELINE config: alexnet (3/4, the headline), simple_conv (2), fc_mnist (1)")
    ap.add_argument("--batch", type=int, default=0, help="global batch (default 100; alexnet at N>1: 1000)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm-kernels", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
