"""UNETR module backed by the sm_100a library -- the drop-in for the reference's `unetr.UNETR`
(/root/reference/unetr.py:21-208) and for `monai.networks.nets.UNETR` as used at
unetr_segmentation_3d.py:36,501-513.

Same constructor, same `forward(x_in, freeze_encoder=False)`, same 165-tensor state-dict, same
errors (unetr.py:60-67).  The sub-modules below only *hold parameters* under the reference's names; all arithmetic
runs in `csrc/` through two C-ABI calls (`b200_unetr_forward` / `b200_unetr_backward`).
"""
from __future__ import annotations

import ctypes
import os
from typing import Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import _lib

__all__ = ["UNETR", "MonaiUNETR"]


# ------------------------------------------------------------------------------------------------
# parameter containers (names / shapes / default initialisation of MONAI 0.6.0)
# ------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; call the parent UNETR")


def _conv(cin, cout, k, stride, transposed=False, bias=False) -> nn.Module:
    """`get_conv_layer(conv_only=True)` -> a module whose only child is `conv` (Appendix B.1)."""
    wrap = _Holder()
    cls = nn.ConvTranspose3d if transposed else nn.Conv3d
    wrap.conv = cls(cin, cout, k, stride, bias=bias)
    return wrap


def _res_block(cin, cout) -> nn.Module:  # UnetResBlock: conv1 conv2 conv3 (instance norm has no parameters)
    m = _Holder()
    m.conv1, m.conv2, m.conv3 = _conv(cin, cout, 3, 1), _conv(cout, cout, 3, 1), _conv(cin, cout, 1, 1)
    return m


def _pr_up(cin, cout, layers) -> nn.Module:  # UnetrPrUpBlock(conv_block=False)
    m = _Holder()
    m.transp_conv_init = _conv(cin, cout, 2, 2, transposed=True)
    m.blocks = nn.ModuleList([_conv(cout, cout, 2, 2, transposed=True) for _ in range(layers)])
    return m


def _up(cin, cout) -> nn.Module:  # UnetrUpBlock(res_block=True)
    m = _Holder()
    m.transp_conv = _conv(cin, cout, 2, 2, transposed=True)
    m.conv_block = _res_block(2 * cout, cout)
    return m


def _vit(in_channels, img_size, hidden, mlp_dim, heads, pos_embed) -> nn.Module:
    n_patches = (img_size[0] // 16) * (img_size[1] // 16) * (img_size[2] // 16)
    pe = _Holder()
    if pos_embed == "conv":
        pe.patch_embeddings = nn.Conv3d(in_channels, hidden, kernel_size=16, stride=16)
    else:
        pe.patch_embeddings = nn.Sequential(nn.Identity(), nn.Linear(4096 * in_channels, hidden))
    pe.position_embeddings = nn.Parameter(torch.zeros(1, n_patches, hidden))
    pe.cls_token = nn.Parameter(torch.zeros(1, 1, hidden))  # in the state-dict, never used (classification=False)
    nn.init.trunc_normal_(pe.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
    for mod in pe.modules():
        if isinstance(mod, nn.Linear):
            nn.init.trunc_normal_(mod.weight, mean=0.0, std=0.02, a=-2.0, b=2.0)
            nn.init.zeros_(mod.bias)
    vit = _Holder()
    vit.patch_embedding = pe
    blocks = []
    for _ in range(12):  # unetr.py:69
        blk = _Holder()
        blk.mlp = _Holder()
        blk.mlp.linear1, blk.mlp.linear2 = nn.Linear(hidden, mlp_dim), nn.Linear(mlp_dim, hidden)
        blk.norm1 = nn.LayerNorm(hidden)
        blk.attn = _Holder()
        blk.attn.out_proj, blk.attn.qkv = nn.Linear(hidden, hidden), nn.Linear(hidden, 3 * hidden, bias=False)
        blk.norm2 = nn.LayerNorm(hidden)
        blocks.append(blk)
    vit.blocks = nn.ModuleList(blocks)
    vit.norm = nn.LayerNorm(hidden)
    return vit


# ------------------------------------------------------------------------------------------------
# autograd bridge
# ------------------------------------------------------------------------------------------------
class _UnetrFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module: "UNETR", x: torch.Tensor, freeze_encoder: bool, needs_grad: bool, *params: torch.Tensor):
        lib = _lib.load()
        _lib.require_device(x)
        x = x.contiguous().float()
        batch = x.shape[0]
        handle = module._handle(batch, x.device)
        # the packed bf16 weight copies live in one persistent buffer per device (bf16 mode); they are current when every parameter
        # still has the storage and version counter they were packed from (FusedAdamW keeps them current itself)
        vkey = module._version_key(params)
        bf16 = module.compute_mode == "bf16"
        packed = _lib.FLAG_WEIGHTS_PACKED if bf16 and module._packed_key.get(x.device) == vkey else 0
        if needs_grad:
            ws = torch.empty(lib.b200_unetr_workspace_bytes(handle, 1), dtype=torch.uint8, device=x.device)
        else:
            # inference (sliding window: hundreds of calls on fixed weights): keep one workspace per batch size
            key = (batch, module.compute_mode, x.device)
            ws = module._infer_ws.get(key)
            if ws is None:
                ws = torch.empty(lib.b200_unetr_workspace_bytes(handle, 0), dtype=torch.uint8, device=x.device)
                if len(module._infer_ws) >= 3:     # a few batch sizes at most (sliding window: full chunks + one ragged tail)
                    module._infer_ws.clear()
                    module._graphs.clear()         # captured graphs point into the workspaces just dropped
                module._infer_ws[key] = ws
        fs, s = module.feature_size, module.img_size
        enc4 = torch.empty((batch, 8 * fs, s[0] // 8, s[1] // 8, s[2] // 8), dtype=torch.float32, device=x.device)
        logits = torch.empty((batch, module.out_channels, *s), dtype=torch.float32, device=x.device)
        table = module._param_table(params)
        flags = (0 if freeze_encoder else _lib.FLAG_NEED_ENCODER_GRAD) | packed | (0 if needs_grad else _lib.FLAG_NO_BACKWARD)
        _lib.check(lib.b200_unetr_forward(handle, table, _lib.ptr(x), _lib.ptr(ws), _lib.ptr(enc4), _lib.ptr(logits),
                                          flags, _lib.stream_ptr()), "b200_unetr_forward")
        if bf16:
            module._packed_key[x.device] = vkey
        ctx.module, ctx.freeze, ctx.handle = module, bool(freeze_encoder), handle
        ctx.save_for_backward(x, ws, *params)
        ctx.set_materialize_grads(False)
        if freeze_encoder:
            ctx.mark_non_differentiable(enc4)  # computed under no_grad in the reference (unetr.py:183-192)
        return enc4, logits

    @staticmethod
    def backward(ctx, d_enc4, d_logits):
        lib = _lib.load()
        x, ws, *params = ctx.saved_tensors
        module = ctx.module
        enc = not ctx.freeze
        if d_enc4 is not None and not enc:
            d_enc4 = None
        has_dl, has_de = d_logits is not None, d_enc4 is not None
        n = len(params)
        if not (has_dl or has_de):
            return (None, None, None, None) + (None,) * n
        flags = (_lib.FLAG_NEED_ENCODER_GRAD if enc else 0) | (_lib.FLAG_HAS_DLOGITS if has_dl else 0) | \
                (_lib.FLAG_HAS_DENC4 if has_de else 0)
        reach = module._grad_reach(has_dl, enc, has_dl or has_de)
        wanted = [reach[i] and params[i].requires_grad for i in range(n)]
        total = sum(p.numel() for p, wnt in zip(params, wanted) if wnt)
        # All parameter gradients are views of ONE flat buffer (a single all-reduce / one AdamW table).  The buffer is kept across
        # steps -- stable gradient addresses mean the optimizer's pointer tables and a captured CUDA graph stay valid -- unless a
        # parameter still holds a gradient (accumulation: autograd adds the new views to it, so they must not share storage).
        fkey = (x.device, tuple(wanted))
        flat = module._flat_grads.get(fkey)
        if flat is None or flat.numel() != total or any(w and p.grad is not None for p, w in zip(params, wanted)):
            flat = torch.empty(total, dtype=torch.float32, device=x.device)
            if all(p.grad is None for p, w in zip(params, wanted) if w):
                module._flat_grads = {fkey: flat}
        grads, off = [], 0
        gtab = (ctypes.c_void_p * n)()
        for i, (p, wnt) in enumerate(zip(params, wanted)):
            if wnt:
                g = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
                gtab[i] = g.data_ptr()
                grads.append(g)
            else:
                gtab[i] = None
                grads.append(None)
        if has_dl:
            d_logits = d_logits.contiguous().float()
        if has_de:
            d_enc4 = d_enc4.contiguous().float()
        # gradient-ready events (data-parallel overlap, parallel.GradientAllReduce): backward passes through the encoder whose
        # reachable parameters are all trainable -- the segmentation step and the ranking stages alike -- record one event per
        # gradient group; the groups cover contiguous ranges of `flat` because it is laid out in parameter order
        full = enc and x.is_cuda and wanted == list(reach)
        module._grad_ready = None
        if full and module.overlap_grad_reduce:
            ng = int(module.grad_groups)
            if ng not in (4, 7, 13):
                raise ValueError("grad_groups must be 4, 7 or 13 (transformer blocks per gradient group: 4, 2, 1)")
            evs = module._grad_events
            if evs is None or len(evs) != ng:
                evs = [torch.cuda.Event() for _ in range(ng)]
                for e in evs:
                    e.record()               # forces creation of the underlying cudaEvent_t
                module._grad_events = evs
            arr = (ctypes.c_void_p * ng)(*[e.cuda_event for e in evs])
            lib.b200_unetr_set_grad_events(ctx.handle, arr, ng)
            first = lambda i: sum(p.numel() for p, wnt in zip(params[:i], wanted) if wnt)
            gs = 12 // (ng - 1)
            conv0 = first(3 + 12 * 11 + 2)
            ranges = [(evs[0], conv0, total)]
            hi = conv0                        # vit.norm rides with the top group
            for k in range(1, ng):
                lo = first(3 + (12 - k * gs) * 11) if k < ng - 1 else 0
                ranges.append((evs[k], lo, hi))
                hi = lo
            # defer_conv_wgrads: the engine launches the conv-stack weight gradients AFTER the ViT backward and records event 0 last
            # (exec.cuh: defer_wg): reduce the conv range last
            if module.defer_conv_wgrads:
                ranges = ranges[1:] + ranges[:1]
            else:
                flags |= _lib.FLAG_INPLACE_WGRADS
            module._grad_ready = (flat, ranges, [(p, g.data_ptr()) for p, g in zip(params, grads) if g is not None])
        else:
            lib.b200_unetr_set_grad_events(ctx.handle, None, 0)
        _lib.check(lib.b200_unetr_backward(ctx.handle, module._param_table(params), gtab, _lib.ptr(x), _lib.ptr(ws),
                                           _lib.ptr(d_enc4), _lib.ptr(d_logits), flags, _lib.stream_ptr()),
                   "b200_unetr_backward")
        return (None, None, None, None) + tuple(grads)


# ------------------------------------------------------------------------------------------------
# the module
# ------------------------------------------------------------------------------------------------
class UNETR(nn.Module):
    """B200-native UNETR.  `forward` returns `(enc4, logits)` like the reference's local class (unetr.py:208)."""

    tuple_output = True

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        img_size: Tuple[int, int, int],
        feature_size: int,
        hidden_size: int,
        mlp_dim: int,
        num_heads: int,
        pos_embed: str,
        norm_name: Union[Tuple, str],
        conv_block: bool = False,
        res_block: bool = False,
        dropout_rate: float = 0.0,
    ) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise AssertionError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise AssertionError("hidden size should be divisible by num_heads.")
        if pos_embed not in ["conv", "perceptron"]:
            raise KeyError(f"Position embedding layer of type {pos_embed} is not supported.")
        # the configuration both reference scripts use (seg:501-513, rank:450-462); anything else must raise
        # rather than silently run something different (there is no fallback path)
        norm = norm_name[0] if isinstance(norm_name, (tuple, list)) else norm_name
        if str(norm).lower() != "instance" or not res_block or conv_block or dropout_rate != 0.0:
            raise NotImplementedError(
                "b200 UNETR implements norm_name='instance', res_block=True, conv_block=False, dropout_rate=0.0")
        img_size = tuple(int(v) for v in (img_size if isinstance(img_size, Sequence) else (img_size,) * 3))
        if len(img_size) != 3 or any(v < 16 or v % 16 for v in img_size):
            raise ValueError("img_size must be three multiples of the 16^3 patch size")
        if feature_size < 8 or feature_size & (feature_size - 1):
            raise NotImplementedError("feature_size must be a power of two >= 8")
        self.in_channels, self.out_channels, self.img_size = int(in_channels), int(out_channels), img_size
        self.feature_size, self.hidden_size, self.mlp_dim, self.num_heads = feature_size, hidden_size, mlp_dim, num_heads
        self.pos_embed = pos_embed
        self.num_layers, self.patch_size = 12, (16, 16, 16)
        self.feat_size = tuple(v // 16 for v in img_size)
        fs = feature_size
        self.vit = _vit(in_channels, img_size, hidden_size, mlp_dim, num_heads, pos_embed)
        self.encoder1 = _Holder()
        self.encoder1.layer = _res_block(in_channels, fs)
        self.encoder2 = _pr_up(hidden_size, fs * 2, 2)
        self.encoder3 = _pr_up(hidden_size, fs * 4, 1)
        self.encoder4 = _pr_up(hidden_size, fs * 8, 0)
        self.decoder5 = _up(hidden_size, fs * 8)
        self.decoder4 = _up(fs * 8, fs * 4)
        self.decoder3 = _up(fs * 4, fs * 2)
        self.decoder2 = _up(fs * 2, fs)
        self.out = _Holder()
        self.out.conv = _conv(fs, out_channels, 1, 1, bias=True)
        # "bf16" (throughput, default) or "fp32" (parity: logits within 1e-4 of the fp32 reference)
        self.compute_mode = os.environ.get("B200_UNETR_MODE", "bf16")
        self.inference_graph = False       # see _graph_forward; switched on by sliding_window_inference for its loop
        self.overlap_grad_reduce = False   # set by parallel.GradientAllReduce
        # with overlap_grad_reduce: launch the conv-stack weight gradients after the ViT backward and reduce the conv range last, so
        # that the 340 MB of ViT gradients cross NVLink behind them.  Measured equal to the in-place order within run-to-run noise
        # (N = 8: 5.83 / 5.94 ms deferred vs 5.86 / 5.90 in place; N = 2: 5.70 vs 5.66, profiles/r02_dp_sweep_n*.json) -- the
        # all-reduce tail is not what limits data parallel -- so the in-place order stays the default.  B200_DEFER_WGRAD=1 turns it on.
        self.defer_conv_wgrads = bool(os.environ.get("B200_DEFER_WGRAD"))
        self.grad_groups = 7               # gradient-ready events per backward when overlapping: conv stack + 6 groups of 2 blocks
        self._init_runtime()

    _RUNTIME = ("_handles", "_infer_ws", "_grad_events", "_graphs", "_grad_ready", "_ordered", "_packed", "_packed_key", "_flat_grads")

    def _init_runtime(self):
        """Per-process state that must never be copied: C handles, workspaces, CUDA graphs / events, the packed bf16 weights."""
        self._handles = {}
        self._infer_ws = {}
        self._grad_events = None
        self._graphs = {}
        self._grad_ready = None
        self._ordered = None
        self._packed = {}                  # device -> uint8 buffer of packed bf16 weight copies
        self._packed_key = {}              # device -> parameter (storage, version) key the buffer was packed from
        self._flat_grads = {}

    def __getstate__(self):
        # copy.deepcopy / pickle / torch.save(model): the raw C handles (freed in __del__), CUDA events and graphs belong to THIS
        # object; the copy rebuilds its own lazily
        state = dict(super().__getstate__()) if hasattr(super(), "__getstate__") else dict(self.__dict__)
        for k in self._RUNTIME:
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._init_runtime()

    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float(): parameters get new storage; cached tables, packed copies and graphs are rebuilt lazily
        out = super()._apply(fn, *a, **k)
        self._ordered = None
        self._packed_key = {}
        self._graphs = {}
        self._flat_grads = {}
        return out

    # ---- plumbing -------------------------------------------------------------------------------
    def set_mode(self, mode: str) -> "UNETR":
        if mode not in ("bf16", "fp32"):
            raise ValueError("mode must be 'bf16' or 'fp32'")
        self.compute_mode = mode
        return self

    def _ordered_params(self):
        """Parameters in the C-ABI table order (enum ParamIdx in csrc/exec.cuh)."""
        if self._ordered is None:
            pe = self.vit.patch_embedding
            lin = pe.patch_embeddings if self.pos_embed == "conv" else pe.patch_embeddings[1]
            out = [pe.position_embeddings, lin.weight, lin.bias]
            for b in self.vit.blocks:
                out += [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.out_proj.weight, b.attn.out_proj.bias,
                        b.norm2.weight, b.norm2.bias, b.mlp.linear1.weight, b.mlp.linear1.bias, b.mlp.linear2.weight,
                        b.mlp.linear2.bias]
            out += [self.vit.norm.weight, self.vit.norm.bias]
            e1 = self.encoder1.layer
            out += [e1.conv1.conv.weight, e1.conv2.conv.weight, e1.conv3.conv.weight]
            out += [self.encoder2.transp_conv_init.conv.weight, self.encoder2.blocks[0].conv.weight, self.encoder2.blocks[1].conv.weight]
            out += [self.encoder3.transp_conv_init.conv.weight, self.encoder3.blocks[0].conv.weight]
            out += [self.encoder4.transp_conv_init.conv.weight]
            for d in (self.decoder5, self.decoder4, self.decoder3, self.decoder2):
                out += [d.transp_conv.conv.weight, d.conv_block.conv1.conv.weight, d.conv_block.conv2.conv.weight,
                        d.conv_block.conv3.conv.weight]
            out += [self.out.conv.conv.weight, self.out.conv.conv.bias]
            assert len(out) == _lib.PARAM_COUNT
            self._ordered = out
        return self._ordered

    @staticmethod
    def _grad_reach(has_dlogits: bool, encoder: bool, any_grad: bool):
        """Which table entries a backward pass reaches (unreached parameters keep grad=None)."""
        reach = [False] * _lib.PARAM_COUNT
        vit_end = 3 + 12 * 11
        if encoder and any_grad:
            top = 12 if has_dlogits else 10          # enc4 only sees blocks 0..9
            for i in range(0, 3 + top * 11):
                reach[i] = True
            if has_dlogits:
                reach[vit_end] = reach[vit_end + 1] = True        # vit.norm
                for i in range(vit_end + 2, vit_end + 2 + 8):      # encoder1..3
                    reach[i] = True
            reach[vit_end + 2 + 8] = True                          # encoder4.transp_conv_init
        if has_dlogits:
            for i in range(vit_end + 2 + 9, _lib.PARAM_COUNT):     # decoder5..2, out
                reach[i] = True
        return reach

    def _param_table(self, params):
        tab = (ctypes.c_void_p * len(params))()
        for i, p in enumerate(params):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("b200 UNETR expects contiguous fp32 parameters")
            tab[i] = p.data_ptr()
        return tab

    @staticmethod
    def _version_key(params):
        return tuple((p.data_ptr(), p._version) for p in params)

    def _norm_device(self, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return device

    def packed_mirrors(self, device):
        """Per parameter of `_ordered_params()`: device address of its plain bf16 mirror in the packed-weight buffer, or 0 (C ABI:
        b200_unetr_packed_cast_offset).  FusedAdamW writes these while it updates the parameter.  None in fp32 mode."""
        if self.compute_mode != "bf16":
            return None
        lib = _lib.load()
        device = self._norm_device(device)
        h = self._handle(1, device)
        base = self._packed[device].data_ptr()
        offs = [lib.b200_unetr_packed_cast_offset(h, i) for i in range(_lib.PARAM_COUNT)]
        return [base + o if o >= 0 else 0 for o in offs]

    def repack_convs(self, device):
        """Refresh the re-laid-out conv / transposed-conv copies of the packed-weight buffer from the fp32 parameters (one launch)."""
        lib = _lib.load()
        device = self._norm_device(device)
        h = self._handle(1, device)
        _lib.check(lib.b200_unetr_pack_convs(h, self._param_table(self._ordered_params()), _lib.ptr(self._packed[device]), _lib.stream_ptr()),
                   "b200_unetr_pack_convs")

    def _handle(self, batch: int, device=None):
        device = self._norm_device(self._ordered_params()[0].device if device is None else device)
        key = (batch, self.compute_mode, device)
        h = self._handles.get(key)
        if h is None:
            lib = _lib.load()
            cfg = _lib.UnetrConfig(batch, self.in_channels, self.out_channels, *self.img_size, self.feature_size,
                                   self.hidden_size, self.mlp_dim, self.num_heads, 1 if self.pos_embed == "conv" else 0,
                                   0 if self.compute_mode == "fp32" else 1)
            h = lib.b200_unetr_create(ctypes.byref(cfg))
            if not h:
                raise RuntimeError("b200_unetr_create: " + _lib.last_error())
            if self.compute_mode == "bf16":
                buf = self._packed.get(device)
                if buf is None:
                    buf = torch.empty(lib.b200_unetr_packed_bytes(h), dtype=torch.uint8, device=device)
                    self._packed[device] = buf
                    self._packed_key.pop(device, None)      # a fresh buffer holds nothing yet
                lib.b200_unetr_set_packed_weights(h, _lib.ptr(buf))
            self._handles[key] = h
        return h

    def __del__(self):
        try:
            lib = _lib.load()
            handles, self._handles = list(self._handles.values()), {}
            for h in handles:
                lib.b200_unetr_destroy(h)
        except Exception:
            pass

    # ---- the reference's forward ---------------------------------------------------------------------
    def forward(self, x_in, freeze_encoder=False):
        if x_in.dim() != 5 or tuple(x_in.shape[1:]) != (self.in_channels, *self.img_size):
            raise ValueError(f"expected input [B,{self.in_channels},{self.img_size}], got {tuple(x_in.shape)}")
        params = self._ordered_params()
        # grad mode is off inside Function.forward, so decide here whether the backward workspace is needed
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if self.inference_graph and not needs_grad and x_in.is_cuda:
            out = self._graph_forward(x_in, params)
            if out is not None:
                return out if self.tuple_output else out[1]
        enc4, logits = _UnetrFunction.apply(self, x_in, bool(freeze_encoder), needs_grad, *params)
        return (enc4, logits) if self.tuple_output else logits

    # ---- inference through a captured CUDA graph ---------------------------------------------------------
    def _graph_forward(self, x_in, params):
        """Replays the whole forward (about 200 launches, ~2 ms of host work) as ONE graph launch.  Opt-in (`inference_graph`):
        the returned tensors are the graph's static output buffers and are overwritten by the next call, which is what a
        sliding-window loop wants (it accumulates each prediction at once) but not what arbitrary callers expect.  The graph is
        rebuilt when the batch size or any parameter (storage, version) changes."""
        key = (x_in.shape[0], self.compute_mode, x_in.device)
        vkey = tuple((p.data_ptr(), p._version) for p in params)
        ent = self._graphs.get(key)
        try:
            if ent is None or ent[0] != vkey:
                static_x = x_in.detach().contiguous().float().clone()
                with torch.no_grad():
                    for _ in range(2):      # eager: allocates the cached workspace, packs the weights, then reuses them
                        _UnetrFunction.apply(self, static_x, False, False, *params)
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        enc4, logits = _UnetrFunction.apply(self, static_x, False, False, *params)
                ent = (vkey, graph, static_x, enc4, logits, self._infer_ws[key])   # keeps the captured workspace alive
                self._graphs[key] = ent
            ent[2].copy_(x_in)
            ent[1].replay()
            return ent[3], ent[4]
        except Exception as exc:      # capture not possible on this driver: fall back to eager launches for good
            import warnings
            warnings.warn(f"b200 UNETR: CUDA-graph inference disabled ({exc})")
            self.inference_graph = False
            self._graphs = {}
            return None


class MonaiUNETR(UNETR):
    """`monai.networks.nets.UNETR` flavour (unetr_segmentation_3d.py:36): `forward(x) -> logits`."""

    tuple_output = False

    def forward(self, x_in):  # noqa: D102
        return super().forward(x_in, freeze_encoder=False)
