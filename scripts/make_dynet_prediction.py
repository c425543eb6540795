"""Predicted latency of LAUD-ResNet101 on a B200 parameter set from the reference's own analytic model
(DyNetSimulator, imported UNCHANGED from /root/reference - which exists only in the build container, so the result is
committed as profiles/dynet_prediction_b200.json and bench.py reports it beside the measured numbers).

    python scripts/make_dynet_prediction.py            # writes profiles/dynet_prediction_b200.json

Block compositions are the reference's (DyNetSimulator/eval_example.py:12-122); the network walk mirrors its
__main__ (:203-330) for resnet101.  Caveat printed with the numbers: the model is an FP32 CUDA-core model (no tensor
cores, hardware_models/static_predictor.py:144-150), so it over-predicts on B200; it is reported for continuity with
the paper, not as a bound."""
import contextlib
import io
import json
import os
import sys
import types

REF = "/root/reference/DyNetSimulator"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()):
    import eval_example as E                                  # noqa: E402  (functions only; __main__ is guarded)
    from hardware_models.multi_cores import GPGPUDynamicPredictor   # noqa: E402

BATCH = 256
HW = dict(n_pes=148, pe_fp32s=128, frequency=1.9e9, mem_bandwidth=6.55e12)   # MEASURED_PEAKS.json copy bandwidth
widths = [56, 28, 14, 7]
last_channels = [256, 512, 1024, 2048]
first_channels = [64, 256, 512, 1024]
first_strides = [1, 2, 2, 2]
n_block = [3, 4, 23, 3]


def walk(fn, **kw):
    total = 0.0
    per_stage = []
    for s in range(4):
        first = fn(c_in=first_channels[s], c_out=last_channels[s], b=4, n_groups=1,
                   h=widths[s] * first_strides[s], w=widths[s] * first_strides[s], stride=first_strides[s],
                   down=first_strides[s], is_se=False, **kw)
        other = fn(c_in=last_channels[s], c_out=last_channels[s], b=4, n_groups=1, h=widths[s], w=widths[s], stride=1,
                   down=1, is_se=False, **kw)
        per_stage.append(float(first + other * (n_block[s] - 1)))
        total += per_stage[-1]
    return total, per_stage


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        pred = GPGPUDynamicPredictor(HW["n_pes"], HW["pe_fp32s"], HW["frequency"], HW["mem_bandwidth"], verbose=False,
                                     latency_mode="add", batch_size=BATCH)
        static, static_st = walk(lambda **k: E.get_static_block_latency(pred, **k))
        out = {}
        for name, dens in (("channel_2222_density_0.60", 0.60), ("channel_2222_density_0.587_measured_mean", 0.587)):
            t, st = walk(lambda **k: E.get_dynamic_block_latency_channel(
                pred, granul_size=1, c_granul_size=2, density_conv1=1.0, density_conv2=1.0, density_conv3=1.0,
                c_density=dens, layer=2, **k))
            out[name] = (t, st)
        t, st = walk(lambda **k: E.get_skipping_block_latency(
            pred, granul_size=1, c_granul_size=2, density_conv1=0.477, density_conv2=0.477, density_conv3=0.477,
            c_density=1.0, layer=2, **k))
        out["layer_skip_rate_0.477_measured_mean"] = (float(t), [float(v) for v in st])

    def row(t, st):
        return {"seconds_per_batch_blocks_only": t, "ms_per_image": 1e3 * t / BATCH, "images_per_s": BATCH / t,
                "seconds_per_stage": st}
    doc = {
        "model": "DyNetSimulator GPGPUDynamicPredictor (reference code, unchanged), latency_mode='add'",
        "generated_by": "scripts/make_dynet_prediction.py (build container; the GPU box has no /root/reference)",
        "hardware_parameters": dict(HW, batch_size=BATCH),
        "network": "ResNet-101 bottleneck trunk (33 blocks; stem and head are not modelled by the reference's walk)",
        "caveat": "FP32 CUDA-core model without tensor cores: over-predicts on B200; reported for continuity with the "
                  "paper, not as a bound",
        "static_dense": row(static, static_st),
    }
    for k, (t, st) in out.items():
        doc[k] = row(float(t), [float(v) for v in st])
    path = os.path.join(ROOT, "profiles", "dynet_prediction_b200.json")
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps({k: (v["images_per_s"] if isinstance(v, dict) and "images_per_s" in v else None) for k, v in doc.items()},
                     indent=1))


if __name__ == "__main__":
    main()
