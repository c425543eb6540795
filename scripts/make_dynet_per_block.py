"""Per-block latency predicted by the reference's own analytic model (DyNetSimulator, imported UNCHANGED from
/root/reference - build container only) for the blocks of our module tree at the MEASURED per-block densities, next to
the measured conv-kernel time of each block (reference walk: DyNetSimulator/eval_example.py:63-122).

    python scripts/make_dynet_per_block.py profiles/r02l_bench.json [profiles/r02l_bench_config2_layer.json ...]
        -> profiles/dynet_per_block_b200.json   (bench.py attaches `dynet_predicted_ms` to roofline.per_block from it when
                                                 the densities of the run equal the ones predicted for)

Block compositions are the reference's: channel mode -> get_dynamic_block_latency_channel (conv1 with oc_density, the channel
masker predictor, conv2 with ic x oc density, conv3 with ic density, scatter-add); layer mode -> get_skipping_block_latency;
spatial mode -> get_dynamic_block_latency_spatial (masker-fused conv1, gather conv2, conv3, scatter-add) with the measured
densities of mask_conv1 / mask_conv2 / mask_conv3.  Caveat: an FP32 CUDA-core model (no tensor cores)."""
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
    import eval_example as E                                  # noqa: E402
    from hardware_models.multi_cores import GPGPUDynamicPredictor   # noqa: E402

HW = dict(n_pes=148, pe_fp32s=128, frequency=1.9e9, mem_bandwidth=6.55e12)
LAYERS = {"resnet50": (3, 4, 6, 3), "resnet101": (3, 4, 23, 3)}
SPATIAL_GRAN = (4, 4, 2, 1)


def geometry(layers):
    """(c_in, c_out, H_in, stride) per block of a torchvision-style bottleneck ResNet at 224x224 (laud_resnet.py:208-250)."""
    out, c_prev = [], 64
    for s, n in enumerate(layers):
        c_out, h = 256 << s, 56 >> s
        for i in range(n):
            stride = 2 if (i == 0 and s > 0) else 1
            out.append(dict(stage=s, c_in=c_prev if i == 0 else c_out, c_out=c_out, h=h * stride, stride=stride,
                            has_down=i == 0))
        c_prev = c_out
    return out


def predict(bench):
    cfg = bench["config"]["baseline_config_index"]
    arch = "resnet50" if cfg == 0 else "resnet101"
    B = bench["config"]["batch_per_gpu"]
    blocks = bench["roofline"]["per_block"]
    geo = geometry(LAYERS[arch])
    assert len(geo) == len(blocks)
    with contextlib.redirect_stdout(io.StringIO()):
        pred = GPGPUDynamicPredictor(HW["n_pes"], HW["pe_fp32s"], HW["frequency"], HW["mem_bandwidth"], verbose=False,
                                     latency_mode="add", batch_size=B)
        rows = []
        for g, m in zip(geo, blocks):
            common = dict(c_in=g["c_in"], c_out=g["c_out"], b=4, n_groups=1, h=g["h"], w=g["h"], stride=g["stride"],
                          down=g["stride"], is_se=False)
            if cfg == 1:
                t = E.get_dynamic_block_latency_channel(pred, granul_size=1, c_granul_size=2, density_conv1=1.0, density_conv2=1.0,
                                                        density_conv3=1.0, c_density=m["rho_c"], layer=2, **common)
            elif cfg == 2:
                t = E.get_skipping_block_latency(pred, granul_size=1, c_granul_size=2, density_conv1=m["rho3"],
                                                 density_conv2=m["rho3"], density_conv3=m["rho3"], c_density=1.0, layer=2, **common)
            else:
                t = E.get_dynamic_block_latency_spatial(pred, granul_size=SPATIAL_GRAN[g["stage"]], c_granul_size=1,
                                                        density_conv1=m["rho1"], density_conv2=m["rho2"], density_conv3=m["rho3"],
                                                        c_density=1.0, **common)
            static = E.get_static_block_latency(pred, **common)
            rows.append({"block": m["block"], "stage": g["stage"] + 1, "rho_c": m["rho_c"], "rho3": m["rho3"], "rho2": m["rho2"],
                         "rho1": m["rho1"], "predicted_ms": 1e3 * float(t), "predicted_static_ms": 1e3 * float(static),
                         "measured_conv_ms": m["conv_ms"]})
    tot_p, tot_m = sum(r["predicted_ms"] for r in rows), sum(r["measured_conv_ms"] for r in rows)
    return {"batch": B, "arch": arch, "workload": bench["config"]["workload"], "blocks": rows,
            "predicted_ms_all_blocks": tot_p, "predicted_static_ms_all_blocks": sum(r["predicted_static_ms"] for r in rows),
            "measured_conv_ms_all_blocks": tot_m, "measured_over_predicted": tot_m / tot_p}


def main():
    out_path = os.path.join(ROOT, "profiles", "dynet_per_block_b200.json")
    doc = json.load(open(out_path)) if os.path.exists(out_path) else {}
    doc["_about"] = {"model": "DyNetSimulator GPGPUDynamicPredictor (reference code, unchanged), latency_mode='add'",
                     "hardware_parameters": HW, "generated_by": "scripts/make_dynet_per_block.py (build container)",
                     "caveat": "FP32 CUDA-core model without tensor cores: the B200 kernels run on tcgen05, so measured << predicted; "
                               "the per-block RATIOS (how latency follows density and stage) are the comparable part"}
    for path in sys.argv[1:]:
        bench = json.loads(open(path).read().strip().splitlines()[-1])
        res = predict(bench)
        doc[str(bench["config"]["baseline_config_index"])] = res
        print(f"{path}: {res['arch']} config {bench['config']['baseline_config_index']}: predicted {res['predicted_ms_all_blocks']:.2f} ms "
              f"(static {res['predicted_static_ms_all_blocks']:.2f} ms), measured conv kernels {res['measured_conv_ms_all_blocks']:.2f} ms")
        for r in res["blocks"][::max(1, len(res["blocks"]) // 8)]:
            print(f"   block {r['block']:2d} stage {r['stage']} rho_c {r['rho_c']:.3f} rho3 {r['rho3']:.3f}: predicted {r['predicted_ms']:.3f} ms, "
                  f"measured {r['measured_conv_ms']:.3f} ms")
    json.dump(doc, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
