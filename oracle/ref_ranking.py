"""Loads the reference's own ranking functions, UNMODIFIED, for golden-vector generation.

TEST INFRASTRUCTURE.  Works only where /root/reference exists (this container); the GPU box
never imports it.  The two scripts cannot be imported (they import MONAI at module top,
rank:9-42), so the `FunctionDef` nodes of

    extract_triplets_more_partitions   unetr_ranking_pretraining_3d.py:59-133
    BTLoss                             unetr_ranking_pretraining_3d.py:202-217

are compiled from the file's AST and executed with the module globals they expect
(`num_partitions` rank:330, `temperature` rank:327, `cos` rank:467).  No source text is copied
into this repository.
"""
import ast
import contextlib
import io
import itertools
import os

import numpy as np
import torch

REF_FILE = "/root/reference/unetr_ranking_pretraining_3d.py"
WANTED = ("extract_triplets_more_partitions", "BTLoss")


def available() -> bool:
    return os.path.exists(REF_FILE)


def load(temperature: float, num_partitions: int = 4):
    tree = ast.parse(open(REF_FILE).read(), REF_FILE)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert len(keep) == len(WANTED)
    mod = ast.Module(body=keep, type_ignores=[])
    env = {
        "np": np, "torch": torch, "product": itertools.product, "permutations": itertools.permutations,
        "num_partitions": num_partitions, "temperature": temperature,
        "cos": torch.nn.CosineSimilarity(dim=-1, eps=1e-6),
    }
    exec(compile(mod, REF_FILE, "exec"), env)
    return env


class NullOptimizer:
    """BTLoss (rank:213-215) calls backward/step/zero_grad itself; keep the leaf grads readable."""

    def step(self):
        pass

    def zero_grad(self):
        pass


def run_reference(feat: torch.Tensor, slice_dimension: int, temperature: float, np_seed: int):
    """feat: [4,C,X,Y,Z] leaf tensor.  Returns (loss float, grad wrt feat, chosen slice indices)."""
    env = load(temperature)
    feat = feat.detach().clone().requires_grad_(True)
    f1, f2 = torch.split(feat, [2, 2], dim=0)  # rank:264
    np.random.seed(np_seed)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        ref, sim, dis = env["extract_triplets_more_partitions"](f1, f2, slice_dimension)
        loss = env["BTLoss"](ref, sim, dis, NullOptimizer())
    line = [l for l in sink.getvalue().splitlines() if l.startswith("Slice indices")][0]
    import re
    body = line.split("[", 1)[1]
    idx = [int(v) for v in re.findall(r"(?<![\w.])(\d+)(?=\)|,|\])", body)]
    assert len(idx) == 4, line
    return loss, feat.grad.detach().clone(), idx, len(ref)
