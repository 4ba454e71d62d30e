#!/usr/bin/env python
"""bench.py - images/s of one FrostNet-Large QAT training step (fwd + bwd + GradBoost QSGD step),
bs=256 per GPU, 224x224x3 synthetic fp32 batches (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with the batch resident in HBM;
`e2e` = the same step through the public nn.Module/Optimizer API with the fp32 batch copied from
pinned host memory and the loss read back every step.  `roofline` describes the dominant C-ABI
entry point (per-call CUDA-event timing in a second pass over the same steps); `cpu_baseline` is the
CPU oracle (a port of the reference path: oracle/frost_oracle.py) on a bounded sample.
`--impl reference` times that CPU path alone (the reference itself is Python/torch wiring that
cannot travel to the GPU box; the oracle is bit-identical to it - tests/golden/make_golden.py).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec FrostNet-L QAT fwd+bwd+GradBoost step, bs=256/GPU, 224x224"
UNIT = "images/s"
FLOP_PER_IMG = 2.563e9           # fwd+bwd conv FLOPs per image (FrostNet-L 1.0 @224)
ALG_BYTES_PER_IMG = 26.6e6       # ideal bottleneck-fused HBM traffic per image, fwd+bwd (6.81 GB at bs=256)


class Args:                       # Classification/setting/train.json
    learning_rate, weight_decay, nesterov, clip_by, toss_coin, noise_decay, amsgrad = 5e-3, 1e-5, True, 1e-3, True, 1e-2, False


def param_groups(model, weight_decay):
    """Classification/train.py:121-137 - one group per tensor, weight decay by shape."""
    groups = []
    for _, p in model.named_parameters():
        if p.dim() == 4:
            wd = 0.0 if p.shape[1] == 1 else weight_decay
        else:
            wd = weight_decay * 0.01
        groups.append({"params": [p], "weight_decay": wd})
    return groups


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--id=%d" % index, "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        top = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(top) if top else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference path) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_qat_images_per_s(batch, steps, warmup, threads=None):
    import numpy as np
    from oracle import frost_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1882)
    np.random.seed(1882)
    spec = O.net_spec("large", 1.0, 1000)
    net = O.OracleNet(spec, O.fresh_state_dict(spec, seed=1882))
    params = net.parameters()
    wds = [(0.0 if p.shape[1] == 1 else Args.weight_decay) if p.dim() == 4 else Args.weight_decay * 0.01 for p in params]
    opt = O.GradBoost("QSGD", params, Args.learning_rate, momentum=0.9, nesterov=True, clip_by=Args.clip_by,
                      toss_coin=True, noise_decay=Args.noise_decay)
    opt.is_warmup = False
    crit = torch.nn.CrossEntropyLoss()
    x = torch.randn(batch, 3, 224, 224)
    y = torch.randint(0, 1000, (batch,))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in params:
            p.grad = None
        loss = crit(net.forward(x, training=True, drop_rate=0.2), y)
        loss.backward()
        with torch.no_grad():
            opt.step([p.grad for p in params], wds=wds)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), threads, sum(times) / len(times)


CPU_BYTES_PER_IMAGE = 0.23e9      # measured peak RSS of the reference's QAT step: 3.3 GB at bs=16 (fp32 autograd graph)


def reference_batch_that_fits(batch):
    """The reference's CPU step keeps ~0.2 GB of fp32 autograd state per image: bs=256 needs ~60 GB of host RAM."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        return batch
    bs = batch
    while bs > 16 and bs * CPU_BYTES_PER_IMAGE * 1.3 > avail:
        bs //= 2
    return bs


def run_reference(args):
    """The reference arm: the UNMODIFIED reference files (baseline/_ref, see oracle/ref_harness.py) running their own
    QAT training step on the host cores, same workload (FrostNet-L, bs=256, 224x224); the oracle port only if the
    reference copy did not travel.  Each step is one full bs=256 batch: 1 warm-up + up to 3 timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import ref_harness as R
    bs = reference_batch_that_fits(args.batch)
    steps = max(1, min(args.steps, 3))
    warm = 1
    if R.available():
        kind = "reference"
        ips, threads, spstep = R.qat_images_per_s(bs, steps, warm)
        c1_ms, _ = R.c1_small_fp32_ms_per_image()
    else:
        kind = "port"
        ips, threads, spstep = cpu_qat_images_per_s(bs, steps, warm)
        c1_ms = None
    sample = "%d timed + %d warm-up QAT steps (fwd+bwd+QSGD GradBoost) of FrostNet-L at bs=%d on %d host threads, %.1f s/step" % (
        steps, warm, bs, threads, spstep)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "FrostNet-Large 1.0 QAT (StatAssist+GradBoost QSGD) bs=%d/GPU 224x224 synthetic ImageNet" % bs,
                       "global_batch": bs * max(1, args.gpus), "parallelism": "dp%d" % max(1, args.gpus),
                       "l2": "CPU arm"},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "c1_cpu_reference": None if c1_ms is None else {
                "workload": "BASELINE configs[0]: FrostNet-Small 1.0 fp32 forward, bs=1, 224x224, eval, CPU",
                "ms_per_image": c1_ms, "images_per_s": 1e3 / c1_ms, "cores": threads, "kind": kind},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if bs != args.batch:
        line["config"]["note"] = "host RAM too small for bs=%d on the CPU (needs ~%d GB); ran bs=%d" % (
            args.batch, int(args.batch * CPU_BYTES_PER_IMAGE * 1.3 / 1e9), bs)
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(batch, steps):
    """cpu_baseline leg of the B200 line: the reference arm on a bounded sample (bs=`batch`), in a process of its own
    (the reference needs torch-global shims, oracle/ref_harness.py).  Returns (cpu_baseline dict, c1 dict)."""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--batch", str(batch),
                        "--steps", str(steps), "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                       text=True, env=env, timeout=1500)
    for ln in reversed(r.stdout.strip().splitlines()):
        if ln.startswith("{"):
            d = json.loads(ln)
            return d["cpu_baseline"], d.get("c1_cpu_reference")
    raise RuntimeError("the reference arm printed no JSON line (exit code %d)" % r.returncode)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def algorithmic_cost(name, a):
    """(bytes, flops) of one C-ABI call from its arguments: what the op must move / compute."""
    if name in ("frost_pw_conv_forward", "frost_pw_conv_forward_simt"):
        M, K, co = a[5], a[6], a[7]
        return M * K + co * K + 4 * M * co, 2 * M * K * co
    if name in ("frost_pw_fused_forward", "frost_pw_fused_bwd_reduce", "frost_pw_fused_bwd_apply"):
        args = getattr(a[0], "_obj", None)          # ctypes.byref(struct) keeps the struct in _obj
        if args is not None:
            M, K, co = args.op.M, args.op.K, args.op.cout
            if name == "frost_pw_fused_forward":    # x twice (statistics pass + quantise pass), q once
                return 2 * M * K + co * K + M * co, 4 * M * K * co
            if name == "frost_pw_fused_bwd_reduce":  # x + dy
                return M * K + co * K + 4 * M * co, 2 * M * K * co
            return M * K + co * K + 8 * M * co, 2 * M * K * co      # x + dy + dz planes
    if name == "frost_pw_chain_backward":
        args = getattr(a[0], "_obj", None)
        if args is not None:                        # x + dy read, dx written (read too when accumulating); I recompute + dgrad + wgrad
            M, K, co = args.op.M, args.op.K, args.op.cout
            return M * K + 4 * M * co + 4 * M * K * (2 if args.accumulate else 1) + 2 * co * K, 6 * M * K * co
    if name == "frost_bnq_apply":
        M, Cc = a[2], a[3]
        return 5 * M * Cc, 0
    if name in ("frost_pw_dgrad", "frost_pw_dgrad_tc"):
        M, K, co = a[4], a[5], a[6]
        return 4 * M * co + co * K + 4 * M * K, 2 * M * K * co
    if name == "frost_pw_wgrad":
        M, K, co = a[4], a[5], a[6]
        return 4 * M * co + M * K + 4 * co * K, 2 * M * K * co
    if name == "frost_pw_wgrad_tc":
        M, K, co = a[6], a[7], a[8]
        return 4 * M * co + M * K + 4 * co * K, 2 * M * K * co
    if name in ("frost_bn_backward", "frost_bn_backward_reduce", "frost_bn_backward_apply"):
        args = getattr(a[0], "_obj", None)
        if args is not None:                        # reduce: dy+I (8 B); apply: dy+I+dz (12 B)
            per = {"frost_bn_backward": 20, "frost_bn_backward_reduce": 8, "frost_bn_backward_apply": 12}[name]
            return per * args.M * args.C, 0
    if name == "frost_dw_conv_forward":
        N, H, W, Cc, k, s = a[5], a[6], a[7], a[8], a[9], a[10]
        Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
        return N * H * W * Cc + 4 * N * Ho * Wo * Cc, 2 * N * Ho * Wo * Cc * k * k
    if name == "frost_dw_dgrad":
        N, H, W, Cc, k, s = a[4], a[5], a[6], a[7], a[8], a[9]
        Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
        return 4 * N * H * W * Cc + 4 * N * Ho * Wo * Cc, 2 * N * Ho * Wo * Cc * k * k
    if name == "frost_dw_wgrad":
        N, H, W, Cc, k, s = a[5], a[6], a[7], a[8], a[9], a[10]
        Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
        return N * H * W * Cc + 4 * N * Ho * Wo * Cc, 2 * N * Ho * Wo * Cc * k * k
    return 0, 0


def run_b200(args):
    import torch.distributed as dist
    import frostnet_b200 as F
    from frostnet_b200 import _lib as L
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (b200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner (and anything NCCL_DEBUG asks for) to stdout
        # unless told otherwise
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)
    bs = args.batch
    torch.manual_seed(1882)
    model = F.frostnet_quant_large_1_0().to(dev)
    model.train()
    if world > 1:
        F.parallel.broadcast_parameters(model)
    opt = F.get_optimizer("QSGD", param_groups(model, Args.weight_decay), Args)
    crit = torch.nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(1882 + rank)
    x_host = torch.randn(bs, 3, 224, 224, generator=g).pin_memory()
    y_host = torch.randint(0, 1000, (bs,), generator=g).pin_memory()
    x = x_host.to(dev)
    y = y_host.to(dev)

    def step(xb, yb):
        opt.zero_grad()
        loss = crit(model(xb), yb)
        loss.backward()
        opt.step()
        return loss

    # StatAssist (Classification/train.py:149-173): FP warm-up steps collect GradBoost sensitivity,
    # then fuse + prepare_qat; the timed region is steady-state QAT.
    for _ in range(2):
        step(x[:64], y[:64])
    opt.is_warmup = False
    model.fuse_model()
    F.prepare_qat(model)
    if world > 1:
        F.parallel.distribute(model)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
    host_ms = [0.0]

    W, K = max(3, args.warmup), args.steps
    for _ in range(W):
        step(x, y)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = L.launch_count()
    # ncu --nvtx --nvtx-include "frost_step" profiles exactly these steps (a start/end range: backward runs on the
    # autograd thread, which a push/pop range of this thread would not cover)
    nvtx_id = torch.cuda.nvtx.range_start("frost_step")
    ms = timed(lambda: step(x, y), K)
    torch.cuda.nvtx.range_end(nvtx_id)
    launches = L.launch_count() - l0
    value = world * bs * K / (ms * 1e-3)

    # host cost of ENQUEUEING one step: a single step after a full sync, so the launch queue never fills
    barrier()
    h0 = time.perf_counter()
    step(x, y)
    host_ms[0] = (time.perf_counter() - h0) * 1e3
    barrier()

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms / K, "steps": K,
                              "gpu_launches_per_step": launches / K, "host_enqueue_ms_per_step": host_ms[0], "quick": True}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # end to end through the public API: every step's fp32 batch comes from pinned host memory (copy issued by
    # frostnet_b200.DevicePrefetcher on a side stream, inside the timed region, one copy per step) and every step's
    # loss is read back to the host (the read of step i is waited for after step i+1 has been queued)
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_losses = []

    def e2e_epoch(k):
        def run():
            for i, (xb, yb) in enumerate(F.DevicePrefetcher(((x_host, y_host) for _ in range(k)), dev)):
                loss = step(xb, yb)
                loss_host[i & 1:(i & 1) + 1].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_ev[i & 1].record()
                if i > 0:
                    loss_ev[(i - 1) & 1].synchronize()
                    e2e_losses.append(float(loss_host[(i - 1) & 1]))
            loss_ev[(k - 1) & 1].synchronize()
            e2e_losses.append(float(loss_host[(k - 1) & 1]))
        return run
    e2e_epoch(2)()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_epoch(K)()
    e1.record()
    barrier()
    ms_e2e_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e_t, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms_e2e_t)
    assert len(e2e_losses) == K + 2 and all(v == v for v in e2e_losses), "e2e losses were not read back"
    e2e_value = world * bs * K / (ms_e2e * 1e-3)
    clocks = sampler.stop() if sampler else None

    # per-entry-point CUDA-event pass: roofline of the dominant kernel.  Every rank runs the same steps (the
    # gradient all-reduce inside backward is a collective); only rank 0 records events.
    roofline, breakdown = None, None
    pk = min(K, 5)
    prof = {}
    call_log = []
    import frostnet_b200.engine as E
    orig_call = L.call

    def prof_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_call(name, *a)
        e1.record()
        prof.setdefault(name, []).append((e0, e1, algorithmic_cost(name, a)))
        if args.dump_calls:
            if name.startswith("frost_bn_backward"):
                sig = [a[0]._obj.M, a[0]._obj.C]
            elif name.startswith("frost_pw_fused"):
                sig = [a[0]._obj.op.M, a[0]._obj.op.K, a[0]._obj.op.cout]
            else:
                sig = [v for v in a[:-1] if isinstance(v, int) and 0 <= v < (1 << 40)]
            call_log.append((name, sig, e0, e1, algorithmic_cost(name, a)[0]))
    if rank == 0:
        E.L.call = prof_call
    try:
        torch.cuda.synchronize()
        for _ in range(pk):
            step(x, y)
        torch.cuda.synchronize()
    finally:
        E.L.call = orig_call
    if rank == 0 and args.dump_calls:
        per = len(call_log) // pk
        rows = [{"name": n, "sig": sig, "us": round(a.elapsed_time(b) * 1e3, 1), "bytes": byt}
                for n, sig, a, b, byt in call_log[(pk - 1) * per:]]
        with open(args.dump_calls, "w") as fh:
            json.dump(rows, fh)
    if rank == 0:
        agg = {}
        for name, evs in prof.items():
            t = sum(a.elapsed_time(b) for a, b, _ in evs) / pk
            agg[name] = (t, sum(c[0] for _, _, c in evs) / pk, sum(c[1] for _, _, c in evs) / pk, len(evs) // pk)
        tot = sum(v[0] for v in agg.values())
        breakdown = {k: {"ms_per_step": round(v[0], 4), "share": round(v[0] / tot, 4), "calls_per_step": v[3],
                         "algorithmic_GBps": round(v[1] / (v[0] * 1e-3) / 1e9, 1) if v[1] else None,
                         "algorithmic_TFLOPps": round(v[2] / (v[0] * 1e-3) / 1e12, 2) if v[2] else None}
                     for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:12]}
        breakdown["all_entry_points_ms"] = round(tot, 3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        top = next((k for k, _ in sorted(agg.items(), key=lambda kv: -kv[1][0]) if agg[k][1] > 0), None)
        if top:
            t_ms, byt, flop, calls = agg[top]
            ai = flop / max(byt, 1)
            if ai > tf_peak * 1e12 / (hbm_peak * 1e9):
                ach = flop / (t_ms * 1e-3) / 1e12
                roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                            "frac": ach / tf_peak, "traffic": None}
            else:
                ach = byt / (t_ms * 1e-3) / 1e9
                roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "traffic": None}
            kernel_of = {"frost_bn_backward_apply": "bn_bwd_apply_kernel", "frost_bn_backward_reduce": "bn_bwd_reduce_kernel",
                         "frost_pw_conv_forward": "pw_conv_fwd_tc_kernel", "frost_pw_wgrad_tc": "pw_wgrad_tc_kernel",
                         "frost_pw_fused_forward": "pw_fused_kernel<0>", "frost_pw_fused_bwd_reduce": "pw_fused_kernel<1>",
                         "frost_pw_fused_bwd_apply": "pw_fused_kernel<2>", "frost_pw_chain_backward": "pw_chain_bwd_kernel",
                         "frost_pw_dgrad_tc": "pw_dgrad_tc_kernel", "frost_bnq_apply": "bnq_apply_kernel",
                         "frost_dw_conv_forward": "dw_conv_fwd_kernel", "frost_dw_wgrad": "dw_wgrad_kernel",
                         "frost_dw_dgrad": "dw_dgrad_kernel"}
            roofline["cuda_kernel"] = kernel_of.get(top)
            # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture, if there is one
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                ent = tr.get(roofline["cuda_kernel"])
                if ent:
                    roofline["traffic"] = ent["dram_bytes_per_launch"]
                    roofline["traffic_source"] = ent["source"]
            except Exception:
                pass
            roofline["algorithmic_bytes_per_launch"] = byt / max(calls, 1)
            roofline["peak_source"] = src
            roofline["launches_per_step"] = calls
            roofline["avg_launch_ms"] = t_ms / max(calls, 1)
            roofline["algorithmic_bytes_per_step"] = byt
            roofline["algorithmic_flops_per_step"] = flop

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu, c1 = None, None
    if world == 1 and not args.no_cpu_baseline:
        cpu, c1 = cpu_baseline_subprocess(32, 2)
    # step-level roofline: the whole step against the ideal bottleneck-fused traffic and the conv FLOPs
    step_roofline = None
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        peaks = {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    alg_bytes = ALG_BYTES_PER_IMG * bs
    alg_flops = FLOP_PER_IMG * bs
    t_step = ms / K * 1e-3
    step_roofline = {
        "algorithmic_bytes_per_step": alg_bytes, "algorithmic_flops_per_step": alg_flops,
        "hbm_frac": alg_bytes / t_step / 1e9 / hbm_peak, "tensor_frac": alg_flops / t_step / 1e12 / tf_peak,
        "hbm_peak_GBps": hbm_peak, "tensor_peak_TFLOPps": tf_peak,
        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
        "definition": "SURVEY 8d: 26.6 MB/img = block-boundary tensors only, fp32, fwd+bwd (ideal bottleneck fusion); "
                      "2.563 GFLOP/img conv fwd+bwd",
        "measured_dram_bytes_per_step": None}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if "step_total" in tr:
            step_roofline["measured_dram_bytes_per_step"] = tr["step_total"]["dram_bytes"]
            step_roofline["measured_dram_source"] = tr["step_total"]["source"]
            step_roofline["dram_over_algorithmic"] = tr["step_total"]["dram_bytes"] / alg_bytes
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 activations x s8 weights -> s32 (exact), fp32 BN/grad/optimizer",
        "data": "synthetic",
        "config": {"workload": "FrostNet-Large 1.0 QAT (StatAssist+GradBoost QSGD) bs=%d/GPU 224x224 synthetic ImageNet" % bs,
                   "global_batch": bs * world, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (~8 GB of activations) >> 126 MB L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": 4,
                "how": "DevicePrefetcher: H2D of step i+1 on a side stream overlaps step i; one copy and one loss read per step, all inside the timed region"},
        "gpu_launches": int(launches),
        "gpu_launches_per_step": launches / K,
        "clocks": clocks,
        "roofline": roofline,
        "conv_roofline": {"algorithmic_tflops": FLOP_PER_IMG * bs * K / (ms * 1e-3) / 1e12,
                          "note": "2.563 GFLOP/img conv fwd+bwd (SURVEY 8d) / step time; tensor peak %s" % "1407 TF/s sustained bf16"},
        "step_roofline": step_roofline,
        "breakdown": breakdown,
        "cpu_baseline": cpu,
        "c1_cpu_reference": c1,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident timing only (for ncu launch lists)")
    ap.add_argument("--dump-calls", default="", help="write the per-call CUDA-event times of one profiled step to this JSON file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
