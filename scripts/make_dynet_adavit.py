"""Predicted latency of AdaViT on DeiT-S (BASELINE configs[3]) on a B200 parameter set from the reference's own analytic
model: DyNetSimulator/adavit/simulate_adavit.py and hardware_models/predictor_transformer.py, imported UNCHANGED from
/root/reference (build container only), evaluated on the DeiT-S geometry with the keep rates the bench measures.  The
script in the reference uses an undefined global `predictor` and has no __main__; the global is injected
here.  Result: profiles/dynet_adavit_b200.json, reported by `bench.py --config 3` beside the measured numbers.

    python scripts/make_dynet_adavit.py

Caveat printed with the numbers: an FP32 CUDA-core model (no tensor cores) - continuity with the paper, not a bound."""
import contextlib
import importlib.util
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
    from hardware_models.predictor_transformer import PredictorTransformer      # noqa: E402
    spec = importlib.util.spec_from_file_location("simulate_adavit", os.path.join(REF, "adavit", "simulate_adavit.py"))
    S = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(S)

BATCH = 512
HW = dict(n_pes=148, pe_fp32s=128, frequency=1.9e9, mem_bandwidth=6.55e12)
DIM, HEADS, DEPTH, MLP, L, KEEP_LAYERS = 384, 6, 12, 4, 197, 1
RATES = dict(token=0.647, head=0.689, attn=0.825, mlp=0.833)      # measured by bench.py --config 3 (profiles/r03*_bench_config3*.json)


def network(pred, dynamic):
    S.predictor = pred
    t = pred.simulate_linear(x_shape=(BATCH, L - 1, 768), w_shape=(DIM, 768), out_shape=(BATCH, L - 1, DIM)).latency
    t += S.simulate_add_pos_embed(BATCH, L, DIM)
    per_block = []
    for i in range(DEPTH):
        dyn = dynamic and i >= KEEP_LAYERS
        per_block.append(float(S.simulate_ada_block(
            B=BATCH, L=L, in_dim=DIM, mlp_ratio=MLP, token_skip=dyn, token_density=RATES["token"] if dyn else 1.0, head_skip=dyn,
            head_num=HEADS, head_density=RATES["head"] if dyn else 1.0, layer_skip=dyn,
            layer_density_attn=RATES["attn"] if dyn else 1.0, layer_density_mlp=RATES["mlp"] if dyn else 1.0)))
    t += sum(per_block) + S.simulate_tail(BATCH, DIM)
    return float(t), per_block


def main():
    with contextlib.redirect_stdout(io.StringIO()):
        pred = PredictorTransformer(HW["n_pes"], HW["pe_fp32s"], HW["frequency"], HW["mem_bandwidth"], verbose=False,
                                    latency_mode="add", batch_size=1)      # the script puts the batch into its shapes
        t_static, pb_static = network(pred, False)
        t_dyn, pb_dyn = network(pred, True)
    out = {"network": "AdaViT on DeiT-S (D=384, 6 heads, 12 blocks, L=197), batch %d, patch projection + blocks + classifier" % BATCH,
           "hardware_parameters": dict(HW, batch_size=BATCH),
           "keep_rates": RATES, "keep_layers": KEEP_LAYERS,
           "static_dense": {"seconds_per_batch": t_static, "images_per_s": BATCH / t_static, "per_block_s": pb_static},
           "token_head_layer_skipping": {"seconds_per_batch": t_dyn, "images_per_s": BATCH / t_dyn, "per_block_s": pb_dyn},
           "caveat": "DyNetSimulator is an FP32 CUDA-core latency model without tensor cores (hardware_models/static_predictor.py:144-150): "
                     "reported for continuity with the paper, not as a bound",
           "source": "DyNetSimulator/adavit/simulate_adavit.py:83-191 and hardware_models/predictor_transformer.py imported unchanged; "
                     "the script's undefined global `predictor` is injected (scripts/make_dynet_adavit.py)"}
    path = os.path.join(ROOT, "profiles", "dynet_adavit_b200.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path, "static %.0f img/s, skipping %.0f img/s" % (out["static_dense"]["images_per_s"], out["token_head_layer_skipping"]["images_per_s"]))


if __name__ == "__main__":
    main()
