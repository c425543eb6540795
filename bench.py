#!/usr/bin/env python
"""Benchmark of the LAUD hot path (BASELINE.json metric): images/sec of LAUD-ResNet101 channel-2222 target-0.5 at
batch 256 per GPU (configs[1]) - or, with --config, one of the other BASELINE configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference] [--config 0|1|2|3|4]

One JSON line on stdout (rank 0).  A "step" is one forward pass of the hot path over one synthetic batch:
  * value : whole-job images/s, inputs already resident in HBM, CUDA-graphed forward (two parallel chains at batch
            >= 128; + the logits all-gather, captured in the same graph, when N > 1), device-timed with CUDA events;
  * e2e   : the same through the public module API from PINNED HOST memory: H2D copy of the batch and D2H read of
            this rank's logits inside the timed region;
  * parity: the graphed logits and every gating decision of the CPU-sample images against the CPU oracle;
  * roofline : the mask-conditioned conv kernel (all launches of one step) timed INSIDE a graph execution (event-record
            nodes around every conv kernel of a single-chain graph): algorithmic bytes / summed kernel time against
            MEASURED_PEAKS.json, and per layer class (conv1 / conv2 / conv3 / downsample per stage) achieved GB/s,
            executed and credited (sparsity-adjusted) TFLOP/s;
  * gpu_baseline : stock PyTorch/cuDNN running the reference's masked-dense scheme (fp16, channels_last, CUDA graph);
  * cpu_baseline : the CPU oracle (a port of the reference's PyTorch path) on a bounded sample.
`--impl reference` times that CPU oracle alone (the reference is pure Python and does not exist on the GPU box;
oracle/laud_oracle.py is its restatement, pinned to the reference's own outputs by tests/golden/).
"""
from __future__ import annotations

import argparse
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

UNIT = "images/s"
SEED = 1
CALIB_IMAGES = 32

# BASELINE.json configs (index = position in `configs`); configs[3] (AdaViT) has no reference code in the tree: it runs
# against the declared self-oracle oracle/adavit_oracle.py (run_adavit below).
CONFIGS = {
    3: dict(metric="images/sec AdaViT-DeiT-S token+head+layer skipping bs512", arch="adavit_deit_s", batch=512,
            rates=dict(token_rate=0.65, head_rate=0.7, layer_rate=0.85),
            workload="AdaViT on DeiT-S (D=384, 6 heads, 12 blocks) token + head + layer skipping, batch 512 x 3x224x224 per GPU "
                     "(configs[3]; self-oracle: no AdaViT code in the reference tree)"),
    0: dict(metric="images/sec LAUD-ResNet50 spatial t0.5 bs8", arch="resnet50", kw="SPATIAL", batch=8,
            rates=dict(spatial_rate=0.4),
            workload="LAUD-ResNet50 spatial-skip 4-4-2-1 target-0.5, batch 8 x 3x224x224 (configs[0])"),
    1: dict(metric="images/sec LAUD-ResNet101 ch-2222 t0.5 bs256", arch="resnet101", kw="HEADLINE", batch=256,
            rates=dict(channel_rate=0.6),
            workload="LAUD-ResNet101 channel-2222 target-0.5, batch 256 x 3x224x224 per GPU (configs[1])"),
    2: dict(metric="images/sec LAUD-ResNet101 layer-skip t0.5 bs256", arch="resnet101", kw="LAYER", batch=256,
            rates=dict(layer_rate=0.47),
            workload="LAUD-ResNet101 layer-skip target-0.5, batch 256 x 3x224x224 per GPU (configs[2])"),
    4: dict(metric="images/sec LAUD-RegNetY-800MF spatial t0.3 bs2048/8", arch="regnety800", kw="SPATIAL", batch=256,
            rates=dict(spatial_rate=0.22),
            workload="LAUD-RegNetY-800MF spatial-skip 4-4-2-1 target-0.3, batch 256 x 3x224x224 per GPU = 2048 over 8 GPUs (configs[4])"),
}


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


# ----------------------------------------------------------------------------- workload
def build_model(conf, device):
    """Drop-in model of this configuration with seeded synthetic weights; BatchNorm statistics and gate biases are
    calibrated on a seeded batch on `device` (weight synthesis - stands in for training)."""
    import laudnet_b200 as L
    from laudnet_b200 import synth
    kw = dict({"HEADLINE": synth.HEADLINE_KWARGS, "SPATIAL": synth.SPATIAL_KWARGS, "LAYER": synth.LAYER_KWARGS}[conf["kw"]])
    if conf["arch"] == "regnety800":
        kw.pop("lr_mult", None)
        model = L.lad_regnet_y_800mf(**kw)
    else:
        model = (L.uni_resnet50 if conf["arch"] == "resnet50" else L.uni_resnet101)(**kw)
    calib = synth.synth_images(CALIB_IMAGES if device.type == "cuda" else 8, 224, SEED + 100).to(device)
    sd = synth.synth_calibrated_state_dict(model, SEED, calib, **conf["rates"])
    model.load_state_dict(sd)
    return model, sd, kw


def oracle_cfg(conf, kw):
    from oracle import laud_oracle as O
    t = {k: tuple(v) if isinstance(v, list) else v for k, v in kw.items() if k != "lr_mult"}
    if conf["arch"] == "regnety800":
        return O.RegNetCfg(**t), O.regnet_forward
    layers = (3, 4, 6, 3) if conf["arch"] == "resnet50" else (3, 4, 23, 3)
    return O.ResNetCfg(layers=layers, **t), O.resnet_forward


def cpu_oracle(conf, kw, sd, n_images: int, repeats: int):
    """The CPU restatement of the reference forward (masked-dense fp32 torch) on the first `n_images` images:
    -> (img/s, cores, outputs of the last forward, per-block traces)."""
    from laudnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, fwd = oracle_cfg(conf, kw)
    sd_cpu = {k: v.cpu() for k, v in sd.items()}
    x = synth.synth_images(n_images, 224, SEED)
    with torch.no_grad():
        fwd(sd_cpu, cfg, x)        # warm-up at the timed shape (oneDNN primitives are created per shape)
        times = []
        for _ in range(repeats):
            traces = []
            t0 = time.perf_counter()
            out = fwd(sd_cpu, cfg, x, traces)
            times.append(time.perf_counter() - t0)
    return n_images / statistics.median(times), cores, out, traces


def run_reference(args, conf):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, sd, kw = build_model(conf, torch.device("cpu"))
    from laudnet_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, fwd = oracle_cfg(conf, kw)
    n = min(args.cpu_sample, conf["batch"])
    x = synth.synth_images(n, 224, SEED)
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd(sd, cfg, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd(sd, cfg, x)
        dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} of the {conf['batch']} images per step (fp32, torch {torch.__version__} CPU, {cores} threads)"
    print(json.dumps({
        "impl": "reference", "metric": conf["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": conf["workload"] + ", CPU oracle port of the reference PyTorch masked-dense path",
                   "batch_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- parity of the benched path
def parity_report(model, graphed_logits, x_dev, oracle_out, traces, n):
    """The timed path against the CPU oracle on the first n images: gating decisions of an eager forward in execution
    order per sample (laudnet_b200/parity.py - bit-exact except where the oracle's own keep/drop margin lies inside the
    fp16 activation budget; a sample is compared up to its first differing block), our gate logits, and the GRAPHED
    logits of the samples whose gates all agree."""
    from laudnet_b200.parity import compare_traces
    TOL = 5e-3
    keep = []
    with torch.no_grad():
        model(x_dev[:n], 1.0, keep=keep)
        torch.cuda.synchronize()
    gp = compare_traces(keep, traces, n, TOL)
    rep = gp.summary()
    err = gp.logits_error(graphed_logits[:n], oracle_out[0])
    ref = oracle_out[0].double()
    ours = graphed_logits[:n].double().cpu()
    rep.update({
        "images": n, "max_rel_err": err, "logits_tolerance": TOL,
        "n_gates": rep["decisions_compared"] + rep["decisions_skipped_after_divergence"],
        "gate_flips": rep["first_flips_within_margin"] + rep["unexplained_flips"],
        "top1_agree_all_samples": float((ours.argmax(1) == ref.argmax(1)).float().mean()),
        "ok": bool(rep["unexplained_flips"] == 0 and rep["max_gate_logit_err_rel"] <= TOL and (err != err or err <= TOL)),
        "how": "graphed forward (the timed path): logits of the samples whose gates all agree vs the fp32 CPU oracle "
               "(max_rel_err, normalised by max|logits|); gates of an eager forward over the same images, compared in execution "
               "order per sample up to the sample's first differing block: a differing decision is explained only if the oracle's "
               "|keep-drop| <= margin_tol_rel x max|logits| (the fp16 activation budget), everything else counts as unexplained"})
    return rep


# ----------------------------------------------------------------------------- GPU arm
def run_graft(args, conf):
    from laudnet_b200 import _engine, _lib, roofline
    from laudnet_b200 import synth
    from laudnet_b200.build import source_hash
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
    B = args.batch or conf["batch"]
    is_resnet = conf["arch"] != "regnety800"
    model, sd, kw = build_model(conf, dev)
    model = model.to(dev).eval()
    lo = rank * B                                            # weak scaling: every rank owns B images
    x_host = synth.synth_images(B, 224, SEED, start=lo).to(torch.float16).pin_memory()
    x_dev = x_host.to(dev)
    ncls = 1000
    gathered = torch.empty((world * B, ncls), dtype=torch.float16, device=dev) if world > 1 else None

    with torch.no_grad():
        # one eager forward: statistics (measured densities) + launch census
        n0 = _lib.launch_count()
        out = model(x_dev, 1.0)
        torch.cuda.synchronize()
        launches_per_step = _lib.launch_count() - n0
        rho_c = torch.cat(out[4]).tolist()
        r3, r2, r1 = (torch.cat(out[i]).tolist() for i in (1, 2, 3))
        flops_ref_counter = float(out[6])
        eng = model._engine
        work = roofline.network_work(eng.plans, rho_c, r3, r2, r1, 224, 64, ncls, B) if is_resnet else None

        # ---- the timed graph: forward (+ fp16 logits all-gather over NVLink inside the same graph when sharded)
        gather_in_graph = False
        post = None
        if world > 1:
            warm = torch.zeros((B, ncls), dtype=torch.float16, device=dev)
            dist.all_gather_into_tensor(gathered, warm)       # communicator set-up outside any capture
            torch.cuda.synchronize()
            if not os.environ.get("LAUD_EAGER_ALLGATHER"):
                def post(logits):
                    lh = logits.to(torch.float16)
                    dist.all_gather_into_tensor(gathered, lh)
                    return lh
        try:
            graphed = _engine.GraphedForward(eng, x_dev, post=post)
            gather_in_graph = post is not None
        except Exception as e:                                # NCCL refused the capture: gather after the replay
            if post is None:
                raise
            sys.stderr.write(f"[bench] all-gather capture failed ({type(e).__name__}: {e}); gathering eagerly\n")
            torch.cuda.synchronize()
            graphed = _engine.GraphedForward(eng, x_dev)

        def step_resident():
            logits, _ = graphed.replay()
            if world > 1 and not gather_in_graph:
                dist.all_gather_into_tensor(gathered, logits.to(torch.float16))
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
        graphed_logits = graphed.logits.float().clone()

        # ---- end to end from pinned host memory, double-buffered H2D on a copy stream; D2H of this rank's logits
        copy_stream = torch.cuda.Stream()
        logits_host = torch.empty((B, ncls), dtype=torch.float32).pin_memory()
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
            if world > 1 and not gather_in_graph:
                dist.all_gather_into_tensor(gathered, logits.to(torch.float16))
            logits_host.copy_(logits, non_blocking=True)    # D2H of the step's result (this rank's rows)
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

        # ---- the conv kernels timed INSIDE a graph execution: a single-chain graph whose conv launches are bracketed
        #      by event-record nodes; kernel_ms_per_step is therefore <= that graph's step time by construction
        roof = None
        if rank == 0:
            pg = _engine.GraphedForward(eng, x_dev, splits=1, profile_convs=True)
            per_launch, pg_ms = [], []
            for it in range(2 + 5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pg.replay()
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    per_launch.append(pg.conv_times())
                    pg_ms.append(e0.elapsed_time(e1))
            tags = pg.conv_tags
            med = [statistics.median(col) for col in zip(*per_launch)]
            _lib.lib().laud_conv_profile(0)
            conv_ms_step = sum(med)
            pg_step = statistics.median(pg_ms)
            by_tag = {}
            for t, ms in zip(tags, med):
                n_, s_ = by_tag.get(t, (0, 0.0))
                by_tag[t] = (n_ + 1, s_ + ms)
            peaks = _peaks()
            roof = {
                "kernel": "laud::conv_tma_kernel (mask-conditioned tcgen05 conv, TMA-staged; all %d launches of one step)" % len(tags),
                "bound": "hbm", "peak": peaks["hbm_gbs"], "unit": "GB/s", "peak_source": peaks["source"],
                "launches_per_step": len(tags), "avg_launch_us": 1e3 * conv_ms_step / max(len(tags), 1),
                "kernel_ms_per_step": conv_ms_step, "graph_ms_per_step_single_chain": pg_step,
                "share_of_step": conv_ms_step / pg_step,
                "timing": "event-record nodes around every conv kernel inside a single-chain CUDA graph, median of 5 replays; "
                          "`value` times the production graph (chains: %d)" % graphed.splits,
            }
            if work is not None:
                conv_bytes = work.conv_bytes_per_image * B
                conv_flops = work.conv_flops_per_image * B
                ach = conv_bytes / (conv_ms_step * 1e-3) / 1e9
                roof.update({"achieved": ach, "frac": ach / peaks["hbm_gbs"], "algorithmic_bytes_per_step": conv_bytes,
                             "tensor": {"achieved": conv_flops / (conv_ms_step * 1e-3) / 1e12, "peak": peaks["tflops"],
                                        "unit": "TFLOP/s", "frac": conv_flops / (conv_ms_step * 1e-3) / 1e12 / peaks["tflops"],
                                        "note": "sparsity-adjusted algorithmic FLOPs (2 x reference counter, conv terms) - the CREDITED rate"}})
                classes = {}
                # MACs the kernels really execute: all of them when a gate is executed masked-dense, the gated share when it
                # is executed as a skip (layer skip over sample lists, gathered channel execution)
                skipping = ((conf["kw"] == "LAYER" and eng.layer_exec == "skip") or
                            (conf["kw"] == "HEADLINE" and eng.channel_exec == "sparse") or
                            (conf["kw"] == "SPATIAL" and getattr(eng, "spatial_exec", "mask") != "mask"))
                # "nskip": only the 3x3 layers of the blocks engine.uses_nskip() names skip - and only their gated OUTPUT
                # channels (N), not the input channels: executed MACs = dense x rho_c there, dense elsewhere
                nskip_tags = {f"s{p.stage + 1}.conv2": sum(rho_c[q.index] for q in eng.plans if q.stage == p.stage and eng.uses_nskip(q)) /
                              max(1, sum(1 for q in eng.plans if q.stage == p.stage and eng.uses_nskip(q)))
                              for p in eng.plans if hasattr(eng, "uses_nskip") and eng.uses_nskip(p)}
                roof["executes"] = ("gated work skipped" if skipping else
                                    ("masked-dense, except the gated output channels of " + ", ".join(sorted(nskip_tags)) + " (TMA-free "
                                     "weight-row gather + expanding epilogue)" if nskip_tags else
                                     "masked-dense (every MAC executed, gates applied in the epilogue)"))
                for t, (n_, ms) in sorted(by_tag.items()):
                    w = work.by_class.get(t)
                    if w is None:
                        classes[t] = {"launches": n_, "ms": round(ms, 4)}
                        continue
                    sec = ms * 1e-3
                    gbs = w["bytes"] * B / sec / 1e9
                    classes[t] = {"launches": n_, "ms": round(ms, 4), "achieved_gbs": round(gbs, 1),
                                  "hbm_frac": round(gbs / peaks["hbm_gbs"], 4),
                                  "executed_tflops": round(2 * (w["macs"] if skipping else w["macs_dense"] * nskip_tags.get(t, 1.0)) * B / sec / 1e12, 1),
                                  "credited_tflops": round(2 * w["macs"] * B / sec / 1e12, 1)}
                roof["by_layer_class"] = classes
            else:
                roof.update({"achieved": None, "frac": None,
                             "by_layer_ms": {t: round(ms, 4) for t, (n_, ms) in sorted(by_tag.items())}})
            if work is not None:
                # per block: measured densities + the summed conv-kernel time of the block's launches (a block's launches
                # are consecutive and start with its conv1) - the walk DyNetSimulator's per-block prediction is set against
                blocks, cur = [], None
                for t, ms in zip(tags, med):
                    if t.endswith(".conv1") or cur is None:
                        cur = {"block": len(blocks), "stage": int(t[1]), "conv_ms": 0.0}
                        blocks.append(cur)
                    cur["conv_ms"] += ms
                for bi, brow in enumerate(blocks):
                    if bi < len(rho_c):
                        brow.update({"rho_c": rho_c[bi], "rho3": r3[bi], "rho2": r2[bi], "rho1": r1[bi], "conv_ms": round(brow["conv_ms"], 5)})
                dpb = os.path.join(ROOT, "profiles", "dynet_per_block_b200.json")
                if os.path.exists(dpb):        # predicted by the reference's simulator for THESE densities (build container)
                    pj = json.load(open(dpb)).get(str(args.config))
                    if pj and len(pj["blocks"]) == len(blocks) and all(
                            abs(a_["rho_c"] - b_["rho_c"]) < 1e-6 and abs(a_["rho3"] - b_["rho3"]) < 1e-6
                            for a_, b_ in zip(pj["blocks"], blocks)):
                        for a_, b_ in zip(pj["blocks"], blocks):
                            b_["dynet_predicted_ms"] = round(a_["predicted_ms"] * B / pj["batch"], 5)
                        roof["per_block_note"] = ("dynet_predicted_ms: DyNetSimulator (reference code, FP32 CUDA-core model, B200 "
                                                  "parameter set) walked over these blocks with these measured densities - "
                                                  "scripts/make_dynet_per_block.py, profiles/dynet_per_block_b200.json")
                    else:
                        roof["per_block_note"] = "profiles/dynet_per_block_b200.json is for other densities (stale): regenerate it"
                roof["per_block"] = blocks
            # DRAM traffic of the same launches: from the committed ncu capture of THIS build (hash-checked), else null
            traffic, traffic_src = None, "no ncu capture of this build (profiles/conv_traffic.json absent or of another build)"
            tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")
            if os.path.exists(tpath) and args.config == 1:
                tj = json.load(open(tpath))
                if tj.get("build_source_hash") == source_hash():
                    traffic = tj["dram_bytes_per_launch_avg"]
                    traffic_src = ("profiles/conv_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, average per conv "
                                   "launch; capture of build %s)" % tj["build_source_hash"])
            roof["traffic"], roof["traffic_source"] = traffic, traffic_src

    net = None
    if work is not None:
        peaks = _peaks()
        net = {
            "flops_per_image": work.flops_per_image, "dense_flops_per_image": work.dense_flops_per_image,
            "flops_ratio": work.flops_per_image / work.dense_flops_per_image,
            "reference_counter_flops": 2.0 * flops_ref_counter, "bytes_per_image": work.bytes_per_image,
            "hbm_frac_whole_net": work.bytes_per_image * value / world / 1e9 / peaks["hbm_gbs"],
            "tensor_frac_whole_net": work.flops_per_image * value / world / 1e12 / peaks["tflops"],
            "mean_channel_density": sum(rho_c) / len(rho_c), "mean_conv3_density": sum(r3) / len(r3),
        }
    else:
        net = {"reference_counter_flops": 2.0 * flops_ref_counter, "mean_conv3_density": sum(r3) / len(r3)}

    dynet = None
    dpath = os.path.join(ROOT, "profiles", "dynet_prediction_b200.json")
    if os.path.exists(dpath) and args.config == 1:        # the reference's own analytic latency model, evaluated in the build container
        dj = json.load(open(dpath))
        key = "channel_2222_density_0.587_measured_mean"
        dynet = {"predicted_images_per_s_per_gpu": dj[key]["images_per_s"],
                 "predicted_static_dense_images_per_s_per_gpu": dj["static_dense"]["images_per_s"],
                 "measured_images_per_s_per_gpu": value / world, "hardware_parameters": dj["hardware_parameters"],
                 "scope": dj["network"], "caveat": dj["caveat"], "source": "profiles/dynet_prediction_b200.json "
                 "(scripts/make_dynet_prediction.py: DyNetSimulator imported unchanged from the reference)"}
        if "per_block" in dj:
            dynet["per_block_source"] = "profiles/dynet_per_block_b200.json"

    cpu = parity = gpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n = min(args.cpu_sample, B)
        v, cores, o_out, traces = cpu_oracle(conf, kw, sd, n, 2)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} of the {B} images, 2 timed forwards after a warm-up (fp32 torch CPU oracle, {cores} threads)"}
        parity = parity_report(model, graphed_logits, x_dev, o_out, traces, n)
    if rank == 0 and world == 1 and is_resnet and not args.no_gpu_baseline:
        try:
            from laudnet_b200.torch_baseline import TorchMaskedDenseResNet
            tb = TorchMaskedDenseResNet(model, dev)
            rate, ms_b, lg = tb.measure(x_dev, steps=10, warmup=3)
            gpu_base = {"value": rate, "unit": UNIT, "ms_per_step": ms_b, "kind": "stock PyTorch/cuDNN masked-dense (the reference's "
                        "execution scheme): fp16 weights + activations, channels_last, cudnn.benchmark, CUDA graph, batch %d" % B,
                        "torch": torch.__version__, "finite": bool(torch.isfinite(lg).all()),
                        "note": "a speed baseline, not a parity instrument (fp16 BatchNorm / pooling change gating decisions)"}
            del tb
        except Exception as e:
            gpu_base = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        sys.stdout.flush()
        os.write(out_fd, (json.dumps({
            "metric": conf["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": conf["workload"], "baseline_config_index": args.config,
                       "batch_per_gpu": B, "global_batch": world * B,
                       "parallelism": (f"batch-sharded x{world}, replicated weights, one NCCL all-gather of fp16 logits per step "
                                       f"({'captured in the CUDA graph' if gather_in_graph else 'after the graph replay'})")
                       if world > 1 else "single GPU",
                       "l2": "inputs + activations per step exceed the 126 MB L2; no explicit flush",
                       "weights": "seeded synthetic, BN stats + gate biases calibrated (%s)" % ", ".join(f"{k}={v}" for k, v in conf["rates"].items()),
                       "cuda_graph": True, "channel_exec": getattr(eng, "channel_exec", None),
                       "layer_exec": getattr(eng, "layer_exec", None), "spatial_exec": getattr(eng, "spatial_exec", None),
                       "graph_chains": graphed.splits},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": graphed.launches * args.steps, "launches_per_step": graphed.launches,
            "eager_launches_per_step": launches_per_step, "conv_paths": _lib.conv_path_counts(),
            "clocks": clocks, "parity": parity, "roofline": roof, "net": net, "gpu_baseline": gpu_base,
            "dynet_simulator": dynet, "cpu_baseline": cpu,
        }) + "\n").encode())
    if world > 1:
        # The communicator's all-gather lives inside a captured CUDA graph: tearing NCCL down while that graph exists
        # blocks forever (measured: the 2-GPU run printed its line and then sat in destroy_process_group).  Release the
        # graphs first, meet at a barrier, and leave without the collective teardown.
        del graphed
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    if parity is not None and not parity["ok"]:
        sys.stderr.write("[bench] PARITY FAILURE: %s\n" % json.dumps(parity))
        sys.exit(3)


# ----------------------------------------------------------------------------- configs[3]: AdaViT (self-oracle)
def build_adavit(conf, device):
    from laudnet_b200 import synth
    from laudnet_b200.adavit import ada_deit_small_patch16_224
    model = ada_deit_small_patch16_224()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    kw = dict(img_size=224, patch_size=16, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4.0, num_classes=1000, keep_layers=1,
              ada_token=True, ada_head=True, ada_layer=True)
    calib = synth.synth_images(CALIB_IMAGES if device.type == "cuda" else 8, 224, SEED + 100).to(device)
    sd = synth.calibrate_adavit(synth.synth_adavit_state_dict(shapes, SEED), kw, calib, **conf["rates"])
    model.load_state_dict(sd, strict=True)
    return model, sd, kw


def run_adavit_reference(args, conf):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from laudnet_b200 import synth
    from oracle import adavit_oracle as A
    model, sd, kw = build_adavit(conf, torch.device("cpu"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = A.AdaViTCfg(**kw)
    n = min(args.cpu_sample, conf["batch"])
    x = synth.synth_images(n, 224, SEED)
    with torch.no_grad():
        for _ in range(args.warmup):
            A.forward(sd, cfg, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            A.forward(sd, cfg, x)
        dt = time.perf_counter() - t0
    v = n * args.steps / dt
    sample = f"{n} of the {conf['batch']} images per step (fp32 masked-dense self-oracle, torch {torch.__version__} CPU, {cores} threads)"
    print(json.dumps({
        "impl": "reference", "metric": conf["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": conf["workload"] + ", CPU self-oracle (masked-dense torch)", "batch_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_adavit(args, conf):
    from laudnet_b200 import _lib, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    out_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    _lib.lib()
    B = args.batch or conf["batch"]
    model, sd, kw = build_adavit(conf, dev)
    model = model.to(dev).eval()
    x_host = synth.synth_images(B, 224, SEED, start=rank * B).to(torch.float16).pin_memory()
    x_dev = x_host.to(dev)
    ncls = model.num_classes
    gathered = torch.empty((world * B, ncls), dtype=torch.float32, device=dev) if world > 1 else None
    peaks = _peaks()

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
            ms = t.item()
        return ms

    with torch.no_grad():
        # eager forward: decisions (measured keep rates), launch census, per-kernel-class times (events between launches)
        n0, g0 = _lib.launch_count(), int(_lib.lib().laud_tok_gemm_launch_count())
        model.forward_logits(x_dev)
        torch.cuda.synchronize()
        launches_per_step = _lib.launch_count() - n0
        gemm_launches = int(_lib.lib().laud_tok_gemm_launch_count()) - g0
        tok, head, layer = (t.bool().transpose(0, 1).cpu() for t in model.decisions(B))
        model.profile = []
        model.forward_logits(x_dev)
        torch.cuda.synchronize()
        marks, model.profile = model.profile, None
        by_class = {}
        for (tag, e), (_, e_next) in zip(marks[:-1], marks[1:]):
            by_class[tag] = by_class.get(tag, 0.0) + e.elapsed_time(e_next)

        graphed = model.capture(B)
        graphed.x.copy_(x_dev)                            # the graph's static input buffer: the batch stays resident in HBM

        def step_resident():
            lg = graphed.replay()
            if world > 1:
                dist.all_gather_into_tensor(gathered, lg)
            return lg

        clk = ClockSampler(local)
        clk.__enter__()
        ms_total = timed(step_resident, args.steps, args.warmup)
        ms_step = ms_total / args.steps
        value = world * B * args.steps / (ms_total * 1e-3)
        graphed_logits = graphed.logits.clone()

        logits_host = torch.empty((B, ncls), dtype=torch.float32).pin_memory()

        # end to end from pinned host memory: the H2D copy of step i+1 runs on a copy stream behind step i's compute
        copy_stream = torch.cuda.Stream()
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
            issue_copy(i + 1)
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[i % 2])
            graphed.x.copy_(stage[i % 2], non_blocking=True)
            consumed[i % 2].record(cur)
            lg = graphed.replay()
            if world > 1:
                dist.all_gather_into_tensor(gathered, lg)
            logits_host.copy_(lg, non_blocking=True)      # D2H of this rank's logits
            state["i"] = i + 1

        for ev in consumed:
            ev.record()
        ms_e2e = timed(step_e2e, args.steps, args.warmup)
        copy_stream.synchronize()
        clk.__exit__(None, None, None)
        clocks = clk.summary()
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- algorithmic work at the measured decisions (oracle-free arithmetic: the formula of simulate_adavit.py:83-182)
    D, H, L, Hd, dh = model.embed_dim, model.num_heads, model.seq_len, model.hidden, 64
    nt, nh = tok.sum(-1).double(), head.sum(-1).double()                       # [B, depth]
    la, lm = layer[..., 0].double(), layer[..., 1].double()
    macs_qkv = (la * nt * D * 3 * dh * nh).sum().item()
    macs_attn = (la * 2 * nt * nt * dh * nh).sum().item()
    macs_proj = (la * nt * dh * nh * D).sum().item()
    macs_mlp = (lm * 2 * nt * D * Hd).sum().item()
    macs_fixed = B * (model.num_patches * D * 3 * model.patch_size ** 2 + D * ncls)
    flops_step = 2.0 * (macs_qkv + macs_attn + macs_proj + macs_mlp + macs_fixed)
    dense_flops_step = 2.0 * B * (model.num_patches * D * 3 * model.patch_size ** 2 + D * ncls +
                                  model.depth * (L * D * 3 * D + 2 * L * L * D + L * D * D + 2 * L * D * Hd))
    # executed GEMM FLOPs: QKV computes every head of a kept m-tile unless the whole tile drops it; proj reads all D inputs
    rows_a, rows_m = (la * nt).sum().item(), (lm * nt).sum().item()
    gemm_flops_exec = 2.0 * (rows_a * D * 3 * D + rows_a * D * D + rows_m * 2 * D * Hd) + 2.0 * macs_fixed
    gemm_ms = sum(v for k, v in by_class.items() if k.startswith("gemm")) + by_class.get("embed", 0.0) + by_class.get("head", 0.0)
    eager_ms = sum(by_class.values())
    roof = {
        "kernel": "laud::tok_gemm_kernel (tcgen05 token GEMM over compact rows: patch projection, QKV, proj, fc1+GELU, fc2, classifier; "
                  "all %d launches of one step)" % gemm_launches,
        "bound": "tensor", "peak": peaks["tflops"], "unit": "TFLOP/s", "peak_source": peaks["source"],
        "launches_per_step": gemm_launches, "kernel_ms_per_step": gemm_ms, "avg_launch_us": 1e3 * gemm_ms / max(1, gemm_launches),
        "eager_ms_per_step": eager_ms, "share_of_step": gemm_ms / eager_ms,
        "achieved": gemm_flops_exec / (gemm_ms * 1e-3) / 1e12, "frac": gemm_flops_exec / (gemm_ms * 1e-3) / 1e12 / peaks["tflops"],
        "timing": "CUDA events between consecutive launches of one eager forward (single stream; includes the patchify / init / "
                  "LayerNorm launches of the embed and head groups); `value` times the CUDA-graph replay",
        "algorithmic_flops_per_step": flops_step, "executed_gemm_flops_per_step": gemm_flops_exec,
        "credited_whole_step": {"achieved": flops_step / (ms_step * 1e-3) / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s",
                                "frac": flops_step / (ms_step * 1e-3) / 1e12 / peaks["tflops"],
                                "note": "sparsity-adjusted algorithmic FLOPs of the whole network / graphed step time"},
        "ms_by_kernel_class": {k: round(v, 4) for k, v in sorted(by_class.items())}, "traffic": None,
    }
    net = {"flops_per_image": flops_step / B, "dense_flops_per_image": dense_flops_step / B, "flops_ratio": flops_step / dense_flops_step,
           "token_keep_rate_dynamic_blocks": tok[:, model.keep_layers:, 1:].float().mean().item(),
           "head_keep_rate_dynamic_blocks": head[:, model.keep_layers:].float().mean().item(),
           "attn_sublayer_rate": layer[:, model.keep_layers:, 0].float().mean().item(),
           "mlp_sublayer_rate": layer[:, model.keep_layers:, 1].float().mean().item()}

    dynet = None
    dpath = os.path.join(ROOT, "profiles", "dynet_adavit_b200.json")
    if os.path.exists(dpath):       # the reference's analytic latency model on the same geometry and keep rates (build container)
        dj = json.load(open(dpath))
        dynet = {"predicted_images_per_s_per_gpu": dj["token_head_layer_skipping"]["images_per_s"],
                 "predicted_static_dense_images_per_s_per_gpu": dj["static_dense"]["images_per_s"],
                 "measured_images_per_s_per_gpu": value / world, "keep_rates_of_the_prediction": dj["keep_rates"],
                 "hardware_parameters": dj["hardware_parameters"], "caveat": dj["caveat"], "source": "profiles/dynet_adavit_b200.json "
                 "(scripts/make_dynet_adavit.py: DyNetSimulator/adavit/simulate_adavit.py imported unchanged from the reference)"}

    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import adavit_oracle as A
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cfg = A.AdaViTCfg(**kw)
        n = min(args.cpu_sample, B)
        sd_cpu = {k: v.cpu() for k, v in sd.items()}
        xs = synth.synth_images(n, 224, SEED)
        with torch.no_grad():
            A.forward(sd_cpu, cfg, xs)
            times = []
            for _ in range(2):
                traces = []
                t0 = time.perf_counter()
                want, wt, wh, wl = A.forward(sd_cpu, cfg, xs, traces)
                times.append(time.perf_counter() - t0)
        cpu = {"value": n / statistics.median(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} of the {B} images, 2 timed forwards after a warm-up (fp32 masked-dense self-oracle, {cores} threads)"}
        # parity of the timed path: decisions per sample in execution order up to the first differing block (a difference
        # is explained only inside the fp16 operand margin), graphed logits of the samples whose decisions all agree, and
        # the logits of ALL samples with the oracle's decisions installed
        TOL = 5e-3
        agree = torch.ones(n, dtype=torch.bool)
        flips = unexplained = compared = 0
        for i, t in enumerate(traces):
            pol = t.policy
            for ours, want_d, lg in ((tok[:n, i, 1:], pol.token[:, 1:], pol.token_logits), (head[:n, i], pol.head, pol.head_logits),
                                     (layer[:n, i], pol.layer, pol.layer_logits)):
                if lg is None:
                    continue
                diff = (ours != want_d).reshape(n, -1)
                lgr = lg.reshape(n, -1)
                compared += int(agree.sum()) * lgr.shape[1]
                for b in torch.nonzero(agree & diff.any(1)).flatten().tolist():
                    inside = lgr[b][diff[b]].abs() <= TOL * lgr.abs().max()
                    flips += int(inside.sum())
                    unexplained += int((~inside).sum())
            same = (tok[:n, i] == pol.token).all(1) & (head[:n, i] == pol.head).all(1) & (layer[:n, i] == pol.layer).all(1)
            agree &= same
        rel = lambda a, b: ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
        err = rel(graphed_logits[:n].cpu()[agree], want[agree]) if agree.any() else float("nan")
        with torch.no_grad():
            forced = [(t.policy.token.to(dev), t.policy.head.to(dev), t.policy.layer.to(dev)) for t in traces]
            lf = model(x_dev[:n], forced=forced)[0]
        err_forced = rel(lf.cpu(), want)
        parity = {"kind": "self-oracle (oracle/adavit_oracle.py; the reference tree holds no AdaViT code: parity unpinned)",
                  "images": n, "decisions_compared": compared, "first_flips_within_margin": flips, "unexplained_flips": unexplained,
                  "margin_tol_rel": TOL, "samples_all_decisions_equal": int(agree.sum()), "max_rel_err": err,
                  "max_rel_err_teacher_forced_all_samples": err_forced, "logits_tolerance": TOL,
                  "ok": bool(unexplained == 0 and err_forced <= TOL and (err != err or err <= TOL))}

    gpu_base = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            from laudnet_b200.torch_baseline import TorchMaskedDenseAdaViT
            tb = TorchMaskedDenseAdaViT(model, dev)
            rate, ms_b, lg = tb.measure(x_dev, steps=10, warmup=3)
            gpu_base = {"value": rate, "unit": UNIT, "ms_per_step": ms_b, "kind": "stock PyTorch masked-dense AdaViT (every token / head / "
                        "sub-layer computed, decisions multiplied in): fp16 weights + activations, F.scaled_dot_product_attention with "
                        "the key mask, CUDA graph, batch %d" % B, "torch": torch.__version__, "finite": bool(torch.isfinite(lg).all()),
                        "note": "a speed baseline, not a parity instrument (fp16 LayerNorm / policies change decisions)"}
            del tb
        except Exception as e:
            gpu_base = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        sys.stdout.flush()
        os.write(out_fd, (json.dumps({
            "metric": conf["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": conf["workload"], "baseline_config_index": args.config, "batch_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"batch-sharded x{world}, replicated weights, one NCCL all-gather of logits per step" if world > 1
                       else "single GPU", "l2": "inputs + token stream per step exceed the 126 MB L2; no explicit flush",
                       "weights": "seeded synthetic DeiT-S, policy biases calibrated (%s)" % ", ".join(f"{k}={v}" for k, v in conf["rates"].items()),
                       "cuda_graph": True, "execution": "token / head / layer skipping executed (compact row lists, device-side row counts)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 2, "d2h_bytes_per_step": B * ncls * 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step, "clocks": clocks,
            "parity": parity, "roofline": roof, "net": net, "gpu_baseline": gpu_base, "dynet_simulator": dynet, "cpu_baseline": cpu,
        }) + "\n").encode())
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("[bench] PARITY FAILURE: %s\n" % json.dumps(parity))
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS),
                    help="index into BASELINE.json configs (default 1: the configuration the metric is quoted on)")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the configuration's)")
    ap.add_argument("--cpu-sample", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "graft" else args.warmup
    conf = CONFIGS[args.config]
    if args.config == 3:
        (run_adavit_reference if args.impl == "reference" else run_adavit)(args, conf)
    elif args.impl == "reference":
        run_reference(args, conf)
    else:
        run_graft(args, conf)


if __name__ == "__main__":
    main()
