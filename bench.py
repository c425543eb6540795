#!/usr/bin/env python
"""Benchmark of the LAUD hot path: images/sec of LAUD-ResNet101 channel-2222
target-0.5 at batch 256 per GPU (BASELINE.json metric, configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]

One JSON line on stdout (rank 0).  A "step" is one forward pass of the hot path
over one synthetic batch of 256 images per GPU:
  * value : whole-job images/s, inputs already resident in HBM, CUDA-graphed forward
            (+ the logits all-gather when N > 1), device-timed with CUDA events;
  * e2e   : the same through the public module API from PINNED HOST memory:
            H2D copy of the batch and D2H read of the logits inside the timed region;
  * roofline : the mask-conditioned conv kernel (all launches of one step), algorithmic
            bytes / summed per-launch CUDA-event time, against MEASURED_PEAKS.json;
  * cpu_baseline : the CPU oracle (a port of the reference's PyTorch path) on a bounded sample.
`--impl reference` times that CPU oracle alone (the reference is pure Python and cannot travel
to the GPU box; oracle/laud_oracle.py is its restatement, pinned to it by tests/golden/).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

METRIC = "images/sec LAUD-ResNet101 ch-2222 t0.5 bs256"
UNIT = "images/s"
BATCH = 256
SEED = 1
CALIB_IMAGES = 32


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    tflops_burst=p["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        return False

    def summary(self):
        if not self.path or not os.path.exists(self.path):
            return None
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [c.strip() for c in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "samples": len(sm), "reasons": sorted(reasons)}


def build_model(device):
    import laudnet_b200 as L
    from laudnet_b200 import synth
    model = L.uni_resnet101(**synth.HEADLINE_KWARGS)
    calib = synth.synth_images(CALIB_IMAGES, 224, SEED + 100).to(device)
    sd = synth.synth_calibrated_state_dict(model, SEED, calib, channel_rate=0.6)
    model.load_state_dict(sd)
    return model, sd


# ----------------------------------------------------------------------------- CPU oracle arm
def cpu_oracle_rate(sd, n_images: int, repeats: int):
    """img/s of the CPU restatement of the reference forward (masked-dense fp32 torch)."""
    from laudnet_b200 import synth
    from oracle import laud_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.ResNetCfg()                                           # ResNet-101 channel-2222 defaults
    sd_cpu = {k: v.cpu() for k, v in sd.items()}
    x = synth.synth_images(n_images, 224, SEED)
    with torch.no_grad():
        O.resnet_forward(sd_cpu, cfg, x)   # warm-up at the timed shape (oneDNN primitives are created per shape)
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            out = O.resnet_forward(sd_cpu, cfg, x)
            times.append(time.perf_counter() - t0)
    flops_ratio = float(out[6]) / float(O.dense_flops(cfg))
    return n_images / statistics.median(times), cores, flops_ratio


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import laudnet_b200 as L
    from laudnet_b200 import synth
    model = L.uni_resnet101(**synth.HEADLINE_KWARGS)
    calib = synth.synth_images(8, 224, SEED + 100)
    sd = synth.synth_calibrated_state_dict(model, SEED, calib, channel_rate=0.6)
    from oracle import laud_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.ResNetCfg()
    n = args.cpu_sample
    x = synth.synth_images(n, 224, SEED)
    with torch.no_grad():
        for _ in range(args.warmup):
            O.resnet_forward(sd, cfg, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.resnet_forward(sd, cfg, x)
        dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} of the 256 images per step (fp32, torch {torch.__version__} CPU, {cores} threads)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "LAUD-ResNet101 channel-2222 target-0.5, 3x224x224 synthetic, CPU oracle port of the "
                               "reference PyTorch masked-dense path", "batch_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- GPU arm
def run_graft(args):
    from laudnet_b200 import _engine, _lib, dist as ldist, roofline, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL prints its version banner on stdout: keep stdout clean for the one JSON line
    out_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    _lib.lib()                                               # fail loudly when the extension is missing
    B = args.batch
    model, sd = build_model(dev)
    model = model.to(dev).eval()
    lo = rank * B                                            # weak scaling: every rank owns B images
    x_host = synth.synth_images(B, 224, SEED, start=lo).to(torch.float16).pin_memory()
    x_dev = x_host.to(dev)
    gathered = torch.empty((world * B, 1000), dtype=torch.float32, device=dev) if world > 1 else None

    with torch.no_grad():
        # one eager forward: statistics (measured densities) + launch census
        n0 = _lib.launch_count()
        out = model(x_dev, 1.0)
        torch.cuda.synchronize()
        launches_per_step = _lib.launch_count() - n0
        rho_c = torch.cat(out[4]).tolist()
        r3, r2, r1 = (torch.cat(out[i]).tolist() for i in (1, 2, 3))
        flops_ref_counter = float(out[6])
        plans = model._engine.plans
        work = roofline.network_work(plans, rho_c, r3, r2, r1, 224, 64, 1000, B)

        graphed = model.capture(x_dev)

        def step_resident():
            logits, _ = graphed.replay()
            if world > 1:
                dist.all_gather_into_tensor(gathered, logits)
            return logits

        def timed(fn, steps, warmup):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dist.barrier()
                ms = t.item()
            return ms

        clk = ClockSampler(local)          # sampled across BOTH timed regions (resident and end to end)
        clk.__enter__()
        ms_total = timed(step_resident, args.steps, args.warmup)
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)

        # ---- end to end from pinned host memory, double-buffered H2D on a copy stream
        copy_stream = torch.cuda.Stream()
        logits_host = torch.empty((B if world == 1 else world * B, 1000), dtype=torch.float32).pin_memory()
        stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"i": 0}

        def issue_copy(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i % 2])
                stage[i % 2].copy_(x_host, non_blocking=True)
                ready[i % 2].record(copy_stream)

        def step_e2e():
            i = state["i"]
            if i == 0:
                issue_copy(0)
            issue_copy(i + 1)                               # prefetch the next step's batch
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[i % 2])
            graphed.static_x.copy_(stage[i % 2], non_blocking=True)
            consumed[i % 2].record(cur)
            logits, _ = graphed.replay()
            src = logits
            if world > 1:
                dist.all_gather_into_tensor(gathered, logits)
                src = gathered
            logits_host.copy_(src, non_blocking=True)       # D2H of the step's result
            state["i"] = i + 1

        for ev in consumed:
            ev.record()
        ms_e2e = timed(step_e2e, args.steps, args.warmup)
        copy_stream.synchronize()
        clk.__exit__(None, None, None)
        clocks = clk.summary()
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
        h2d = x_host.numel() * x_host.element_size()
        d2h = logits_host.numel() * logits_host.element_size()

        # ---- per-kernel timing of the conv launches of one step: the library brackets each kernel (only the
        #      kernel, not the host-side descriptor encoding) with CUDA events on the launching stream
        conv_ms, by_tag, n_conv = [], {}, 0
        L = _lib.lib()
        for _ in range(3):
            L.laud_conv_profile(1)
            model.forward_logits(x_dev)
            torch.cuda.synchronize()
            tot = ctypes.c_float(0.0)
            n_conv = L.laud_conv_profile_read(ctypes.byref(tot))
            conv_ms.append(tot.value)
            L.laud_conv_profile(0)
        with _engine.conv_profile() as prof:            # per-layer split (includes host launch gaps: shares only)
            model.forward_logits(x_dev)
            torch.cuda.synchronize()
        by_tag = prof.by_tag()
        conv_ms_step = statistics.median(conv_ms)

    peaks = _peaks()
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01E_conv_traffic.json")
    if os.path.exists(tpath):                  # dram__bytes_read+write of the conv launches from the committed ncu capture
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_per_launch_avg"]
        traffic_src = "profiles/r01E_conv_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, average per conv launch)"
    conv_bytes = work.conv_bytes_per_image * B
    conv_flops = work.conv_flops_per_image * B
    ach_gbs = conv_bytes / (conv_ms_step * 1e-3) / 1e9
    roof = {
        "kernel": "laud::conv_tma_kernel (mask-conditioned tcgen05 conv, TMA-staged; all %d launches of one step)" % n_conv,
        "bound": "hbm", "achieved": ach_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
        "frac": ach_gbs / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peaks["source"],
        "launches_per_step": n_conv, "avg_launch_us": 1e3 * conv_ms_step / max(n_conv, 1),
        "algorithmic_bytes_per_step": conv_bytes, "kernel_ms_per_step": conv_ms_step,
        "share_of_step": conv_ms_step / ms_step,
        "share_note": "summed device time of the conv launches (eager, one stream) over the CUDA-graphed step: the graph runs two "
                      "chains whose kernels overlap, so this exceeds the serialised share of the ncu launch list "
                      "(profiles/r01E_launch_summary.txt: conv 89 %, masker 5 %, stem 4 %, head 1 %)",
        "tensor": {"achieved": conv_flops / (conv_ms_step * 1e-3) / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s",
                   "frac": conv_flops / (conv_ms_step * 1e-3) / 1e12 / peaks["tflops"],
                   "note": "sparsity-adjusted algorithmic FLOPs (2 x reference counter, conv terms)"},
        "by_layer_ms": {k: round(t, 4) for k, (n, t) in sorted(by_tag.items())},
    }
    net = {
        "flops_per_image": work.flops_per_image, "dense_flops_per_image": work.dense_flops_per_image,
        "flops_ratio": work.flops_per_image / work.dense_flops_per_image,
        "reference_counter_flops": 2.0 * flops_ref_counter, "bytes_per_image": work.bytes_per_image,
        "hbm_frac_whole_net": work.bytes_per_image * value / world / 1e9 / peaks["hbm_gbs"],
        "tensor_frac_whole_net": work.flops_per_image * value / world / 1e12 / peaks["tflops"],
        "mean_channel_density": sum(rho_c) / len(rho_c),
    }

    dynet = None
    dpath = os.path.join(ROOT, "profiles", "dynet_prediction_b200.json")
    if os.path.exists(dpath):        # the reference's own analytic latency model, evaluated in the build container
        dj = json.load(open(dpath))
        key = "channel_2222_density_0.587_measured_mean"
        dynet = {"predicted_images_per_s_per_gpu": dj[key]["images_per_s"],
                 "predicted_static_dense_images_per_s_per_gpu": dj["static_dense"]["images_per_s"],
                 "measured_images_per_s_per_gpu": value / world, "hardware_parameters": dj["hardware_parameters"],
                 "scope": dj["network"], "caveat": dj["caveat"], "source": "profiles/dynet_prediction_b200.json "
                 "(scripts/make_dynet_prediction.py: DyNetSimulator imported unchanged from the reference)"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, ratio = cpu_oracle_rate(sd, args.cpu_sample, 2)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_sample} of the 256 images, 2 timed forwards after a warm-up (fp32 torch CPU oracle, "
                         f"{cores} threads; flops ratio {ratio:.3f})"}
    if rank == 0:
        sys.stdout.flush()
        os.write(out_fd, (json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "LAUD-ResNet101 channel-2222 target-0.5, batch 256 x 3x224x224 per GPU (configs[1])",
                       "batch_per_gpu": B, "global_batch": world * B, "parallelism": f"batch-sharded x{world}, "
                       "replicated weights, one NCCL all-gather of logits per step" if world > 1 else "single GPU",
                       "l2": "inputs + activations per step (>2 GB) exceed the 126 MB L2; no explicit flush",
                       "weights": "seeded synthetic, BN stats + gate biases calibrated to channel density 0.6",
                       "cuda_graph": True, "channel_exec": model._engine.channel_exec,
                       "graph_chains": graphed.splits},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": graphed.launches * args.steps, "launches_per_step": graphed.launches,
            "eager_launches_per_step": launches_per_step, "conv_paths": _lib.conv_path_counts(),
            "clocks": clocks, "roofline": roof, "net": net, "dynet_simulator": dynet, "cpu_baseline": cpu,
        }) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU per step (metric is quoted at 256)")
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_graft(args)


if __name__ == "__main__":
    main()
