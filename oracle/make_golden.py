"""Generates tests/golden/*.npz.  Run in the build container (needs /root/reference for the
ranking vectors):  python oracle/make_golden.py

ranking_*.npz  : outputs of the reference's own extract_triplets_more_partitions + BTLoss
                 (oracle/ref_ranking.py) -- these PIN the oracle's a14/a15 restatement.
unetr_tiny.npz : outputs of the oracle itself on a reduced network (regression fixture for
                 the CUDA path on the GPU box; not a pin of the oracle).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_ranking, unetr_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def ranking_vectors():
    cases = {"feat": (8, 12, 0.1, 11), "recon": (3, 16, 0.1, 12), "feat_T05": (5, 8, 0.5, 13)}
    for name, (c, s, temp, seed) in cases.items():
        g = torch.Generator().manual_seed(seed)
        feat = torch.randn(4, c, s, s, s, generator=g)
        rec = {"feat": feat.numpy(), "temperature": np.float32(temp)}
        for sd in (2, 3, 4):
            loss, grad, idx, ntrip = ref_ranking.run_reference(feat, sd, temp, np_seed=seed + sd)
            assert ntrip == 576
            rec[f"loss_sd{sd}"] = np.float64(loss)
            rec[f"grad_sd{sd}"] = grad.numpy()
            rec[f"idx_sd{sd}"] = np.array(idx, dtype=np.int64)
            rec[f"npseed_sd{sd}"] = np.int64(seed + sd)
        np.savez_compressed(os.path.join(OUT, f"ranking_{name}.npz"), **rec)
        print("wrote ranking", name, {k: float(v) for k, v in rec.items() if k.startswith("loss")})
    # identical slices => 576 ln 2
    feat = torch.ones(4, 4, 8, 8, 8)
    loss, _, _, _ = ref_ranking.run_reference(feat, 2, 0.1, np_seed=0)
    np.savez(os.path.join(OUT, "ranking_const.npz"), loss=np.float64(loss))
    print("const", loss)


def tiny_unetr():
    torch.manual_seed(0)
    m = O.UNETR(1, 5, (32, 32, 32), 8, 64, 128, 4, "perceptron", "instance", res_block=True)
    with torch.no_grad():
        m.out.conv.conv.weight.mul_(4.0)
        m.out.conv.conv.bias.copy_(torch.linspace(-1, 1, 5))
    x, y = O.make_inputs(batch=2, img=32, n_classes=5, seed=3)
    inter = m(x, return_intermediates=True)
    loss, dice, ce = O.dice_ce_loss(inter["logits"], y, return_terms=True)
    loss.backward()
    rec = {"loss": loss.item(), "dice": dice.item(), "ce": ce.item(),
           "logits_sum": inter["logits"].double().sum().item(),
           "logits_absmax": inter["logits"].abs().max().item(),
           "enc4": inter["enc4"].detach().numpy(),
           "logits_corner": inter["logits"][:, :, :4, :4, :4].detach().numpy(),
           "argmax_hist": np.bincount(inter["logits"].argmax(1).flatten().numpy(), minlength=5)}
    for k, p in m.named_parameters():
        if p.grad is not None:
            rec["gradnorm/" + k] = p.grad.double().norm().item()
    np.savez_compressed(os.path.join(OUT, "unetr_tiny.npz"), **rec)
    print("wrote unetr_tiny: loss", rec["loss"])


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ranking_vectors()
    tiny_unetr()
