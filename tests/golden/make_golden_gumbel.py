"""Known answers for the TRAINING-mode gate with supplied Gumbel noise, produced by the reference's own
operator classes.  Build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_gumbel.py

`F.gumbel_softmax` draws its sample as `-torch.empty_like(logits).exponential_().log()` - the first consumer of torch's
generator inside the masker's forward.  Seeding the generator, drawing a tensor of the same shape ourselves, re-seeding and
calling the REFERENCE masker in train mode therefore gives us the exact noise the reference used.  Stored per case:
input, parameters, temperature, the noise and the reference's hard mask (tests/golden/kat_gumbel.npz)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/imagenet_classification")
sys.dont_write_bytecode = True
with contextlib.redirect_stdout(io.StringIO()):
    from models import utils as RU                  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    save = {}
    torch.manual_seed(7)
    cases = [("spatial", lambda: RU.Masker_spatial(16, 2, 4), (3, 16, 8, 8), 0.7),
             ("layer", lambda: RU.Masker_spatial(32, 1, 1), (5, 32, 6, 6), 0.3),
             ("mlp2", lambda: RU.Masker_channel_MLP(32, 16, layers=2, reduction=16), (4, 32, 5, 5), 1.0),
             ("mlp1", lambda: RU.Masker_channel_MLP(16, 8, layers=1), (4, 16, 7, 7), 0.05),
             ("convlin", lambda: RU.Masker_channel_conv_linear(32, 8, reduction=4), (3, 32, 6, 6), 0.5)]
    for tag, ctor, shape, tau in cases:
        with contextlib.redirect_stdout(io.StringIO()):
            mod = ctor()
        for p in mod.parameters():                   # data-dependent, undecided gates: the noise matters
            p.data = torch.randn_like(p) * 0.3
        mod.train()
        if tag == "convlin":                         # frozen BN, as the mmdet backbones run it (norm_eval)
            mod.conv[1].running_mean.normal_()
            mod.conv[1].running_var.uniform_(0.5, 2.0)
            mod.conv[1].eval()
        x = torch.randn(*shape)
        with torch.no_grad():
            mod.eval()
            eval_mask = mod(x, tau)[0]
            mod.train()
            if tag == "convlin":
                mod.conv[1].eval()
            b = shape[0]
            lshape = (b, 2) + tuple(eval_mask.shape[1:])
            torch.manual_seed(1000 + len(save))
            noise = -torch.empty(lshape).exponential_().log()
            torch.manual_seed(1000 + len(save))
            mask, sparsity, flops = mod(x, tau)
        hard = (mask > 0.5).float()
        assert float((mask - hard).abs().max()) < 1e-6            # forward value of the straight-through estimator
        assert not torch.equal(hard, eval_mask), "noise did not change any decision: case is not informative"
        save[f"{tag}.x"] = x.numpy()
        save[f"{tag}.tau"] = np.float32(tau)
        save[f"{tag}.noise"] = noise.reshape(b, -1, *eval_mask.shape[2:]).numpy()      # [B, 2G(, S, S)]: keep half first
        save[f"{tag}.mask"] = hard.numpy().astype(np.uint8)
        save[f"{tag}.eval_mask"] = eval_mask.numpy().astype(np.uint8)
        save[f"{tag}.sparsity"] = np.float32(sparsity.item())
        for k, v in mod.state_dict().items():
            save[f"{tag}.sd.{k}"] = v.numpy()
        print(tag, "mask density", float(hard.mean()), "decisions changed by the noise:", int((hard != eval_mask).sum()))
    np.savez_compressed(os.path.join(HERE, "kat_gumbel.npz"), **save)


if __name__ == "__main__":
    main()
