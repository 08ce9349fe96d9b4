"""FreqFusion x4 (3 frozen experts + fusion head) on the ffb200 kernels, with the reference's checkpoint contract.

Mirrors `CompleteEnhancedFusionSR.forward` in eval mode (reference src/models/enhanced_fusion.py:694-754) with
`ExpertEnsemble.forward_all` (src/models/expert_loader.py:768-777, experts in the order hat, dat, nafnet).
Checkpoint ingestion follows io._build_and_load (models/team29_FreqFusion/io.py:127-182), load_checkpoint_flexible
(expert_loader.py:99-169) and NAFNetSR.load_nafnet_weights (nafnet/__init__.py:84-115): same accepted container keys,
`module.` / `model.` prefix stripping, name+shape filter, missing files tolerated with a warning.
"""
import os
from collections import OrderedDict

import torch

from . import lib as L
from . import ops, weights
from .dat import DATRunner
from .hat import HATRunner, Workspace
from .head import HeadRunner
from .nafnet import NAFNetRunner

EXPERT_FILES = {
    "hat": os.path.join("pretrained", "hat", "HAT-L_SRx4_ImageNet-pretrain.pth"),
    "dat": os.path.join("pretrained", "dat", "DAT_x4.pth"),
    "nafnet": os.path.join("pretrained", "nafnet", "NAFNet-SIDD-width64.pth"),
}


def _extract_state(ckpt, keys):
    for k in keys:
        if isinstance(ckpt, dict) and k in ckpt:
            return ckpt[k]
    return ckpt


def _filtered_update(base, incoming, prefixes=("module.",), replace_anywhere=False):
    """name + shape filter of the reference loaders; returns the number of tensors taken."""
    n = 0
    for key, val in incoming.items():
        ck = key
        if replace_anywhere:
            ck = ck.replace("module.", "")
        else:
            for p in prefixes:
                if ck.startswith(p):
                    ck = ck[len(p):]
        if ck in base and torch.is_tensor(val) and tuple(val.shape) == tuple(base[ck].shape):
            base[ck] = val.detach().to("cpu", torch.float32)
            n += 1
    return n


# forwards up to this many LR pixels run as replayed CUDA graphs (launch bound otherwise)
GRAPH_MAX_LR_PIXELS = int(os.environ.get("FFB200_GRAPH_MAX_LR_PIXELS", str(4 * 128 * 128)))
MIN_SIDE = 9                             # smallest LR side the reference's padding paths accept
GRAPH_CACHE_MAX = 8                      # captured shapes kept (least recently used dropped first)
WORKSPACE_LIMIT_BYTES = int(float(os.environ.get("FFB200_WS_LIMIT_GB", "64")) * 2 ** 30)      # cached workspaces over all shapes


class FreqFusionB200:
    """Inference-only model object.  `state` holds four fp32 CPU state dicts (reference key names)."""

    def __init__(self, device="cuda", init_seed=0, verbose=True):
        device = torch.device(device)
        if device.type != "cuda":
            raise L.FFError("FreqFusionB200 runs on CUDA devices only: there is no CPU fallback")
        L.load()
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.verbose = verbose
        # random-init stand-ins (the reference keeps its random init when a checkpoint is missing)
        self.state = {m: weights.make_state_dict(m, init_seed) for m in ("hat", "dat", "nafnet", "fusion")}
        self._runners = None

    # ---- checkpoint contract ------------------------------------------------------------------
    def load_fusion_checkpoint(self, path):
        """io._build_and_load, reference io.py:164-176: the checkpoint is applied to the WHOLE module tree with the name+shape
        filter, so a checkpoint written with live experts (checkpoint_manager saves model.state_dict(), which then holds
        `expert_ensemble.{hat,dat,nafnet.nafnet}.*`) also overrides the expert weights -- not only the fusion head."""
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        sd = ckpt.get("model_state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
        n = self.load_state_dict(sd)
        if self.verbose:
            print(f"[team29_FreqFusion/b200] Loaded {n} weight tensors from the fusion checkpoint")
        return n

    def load_expert_checkpoint(self, name, path):
        if not os.path.exists(path):
            if self.verbose:
                print(f"[team29_FreqFusion/b200] WARNING {name} checkpoint not found: {path} (random init kept)")
            return 0
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        sd = _extract_state(ckpt, ("params_ema", "params", "state_dict", "model"))
        n = _filtered_update(self.state[name], sd, replace_anywhere=True)
        if self.verbose:
            print(f"[team29_FreqFusion/b200] {name} loaded: {n}/{len(self.state[name])} tensors")
        self._runners = None
        return n

    def state_dict(self):
        """Reference-compatible flat view: fusion keys + expert_ensemble.<expert>.<key> (NAFNet under .nafnet.)."""
        out = OrderedDict(self.state["fusion"])
        for k, v in self.state["hat"].items():
            out["expert_ensemble.hat." + k] = v
        for k, v in self.state["dat"].items():
            out["expert_ensemble.dat." + k] = v
        for k, v in self.state["nafnet"].items():
            out["expert_ensemble.nafnet.nafnet." + k] = v
        return out

    def load_state_dict(self, sd, strict=False):
        """name + shape filtered update of all four state dicts from one flat dict in the reference's module-tree naming
        (fusion keys bare, experts under `expert_ensemble.`; `module.` / `model.` prefixes stripped first, io.py:166-170).
        NAFNet tensors are registered twice by the reference (nafnet/__init__.py:76-82: `expert_ensemble.nafnet.nafnet.X`
        and the alias `expert_ensemble.nafnet.X`); both spellings are accepted and counted like the reference counts them."""
        clean = {}
        for k, v in sd.items():
            ck = k
            for p in ("module.", "model."):
                if ck.startswith(p):
                    ck = ck[len(p):]
            clean[ck] = v
        n = _filtered_update(self.state["fusion"], clean, prefixes=())
        for name, pres in (("hat", ("expert_ensemble.hat.",)), ("dat", ("expert_ensemble.dat.",)),
                           ("nafnet", ("expert_ensemble.nafnet.", "expert_ensemble.nafnet.nafnet."))):
            for pre in pres:      # the canonical spelling last, so it wins when both are present
                sub = {k[len(pre):]: v for k, v in clean.items() if k.startswith(pre)}
                if name == "nafnet":      # `body` is a second alias of `middle_blks` (nafnet/__init__.py:82)
                    sub = {("middle_blks." + k[5:] if k.startswith("body.") else k): v for k, v in sub.items()}
                n += _filtered_update(self.state[name], sub, prefixes=())
        self._runners = None
        return n

    # ---- execution ------------------------------------------------------------------------------
    def runners(self):
        if self._runners is None:
            dev = self.device
            with torch.cuda.device(dev):
                self._runners = dict(hat=HATRunner(self.state["hat"], dev), dat=DATRunner(self.state["dat"], dev),
                                     nafnet=NAFNetRunner(self.state["nafnet"], dev), head=HeadRunner(self.state["fusion"], dev))
            self._stacks = Workspace(dev)
            self._graphs = OrderedDict()
            self._epoch = 0
        return self._runners

    def _workspaces(self):
        return [r.ws for r in self._runners.values()] + [self._stacks]

    def workspace_bytes(self):
        return sum(w.nbytes() for w in self._workspaces()) if self._runners is not None else 0

    def _begin_forward(self):
        """Bounded buffer cache: every forward is an epoch; when the cached workspaces of all shapes seen so far exceed
        WORKSPACE_LIMIT_BYTES, the buffers of the least recently used shapes are released (oldest first) and the CUDA graphs,
        which hold raw pointers into them, are dropped and re-captured on demand."""
        self.runners()
        self._epoch += 1
        for w in self._workspaces():
            w.epoch = self._epoch
        if self.workspace_bytes() <= WORKSPACE_LIMIT_BYTES:
            return
        torch.cuda.synchronize(self.device)
        for age in (8, 4, 2, 1, 0):
            freed = sum(w.evict_unused_since(self._epoch - age) for w in self._workspaces())
            if freed:
                self._graphs.clear()
            if self.workspace_bytes() <= WORKSPACE_LIMIT_BYTES // 2:
                break

    def _stack(self, B, S0, S1):
        return self._stacks.get("expert_stack", B * 16 * S0 * S1, 12, torch.float32)

    def _check_input(self, lr, what):
        if not torch.is_tensor(lr) or not lr.is_cuda:
            raise L.FFError(f"{what} needs a CUDA tensor (no CPU fallback)")
        if lr.dim() != 4 or lr.shape[1] != 3:
            raise L.FFError(f"{what}: expected an NCHW batch with 3 channels, got {tuple(lr.shape)}")
        if lr.device != self.device:
            raise L.FFError(f"{what}: input lives on {lr.device}, the model on {self.device}")
        _, _, h, w = lr.shape
        if not self.supports_whole_image(h, w):
            raise L.FFError(f"{what}: sides must be at least {MIN_SIDE} px (got {h}x{w}): below that the reference's own reflect padding "
                            "(pad_to_window_size, expert_loader.py:83-91) fails")

    @torch.no_grad()
    def forward_experts(self, lr):
        """[B,3,S,S] fp32 on the device -> fp32 expert stack [B*4S*4S][12] (hat 0-2, dat 3-5, nafnet 6-8), each clamp(.,0,1)."""
        r = self.runners()
        B, _, h, w = lr.shape
        stack = self._stack(B, h, w)
        if os.environ.get("FFB200_EXPERT_STREAMS", "1") == "0":
            r["hat"].forward(lr, stack, 0)
            r["dat"].forward(lr, stack, 3)
            r["nafnet"].forward(lr, stack, 6)
            return stack
        # the experts are independent (disjoint workspaces, disjoint channels of `stack`): run them on three streams so the
        # small / low-occupancy kernels of one expert fill the gaps of the others.  Invariant: the side streams start after
        # an event on the caller's stream (so `lr` and `stack` are ready) and the caller's stream waits for both before this
        # function returns -- `lr` / `stack` are therefore never used by a side stream outside this call, which is why no
        # record_stream() is needed on them.
        if not hasattr(self, "_streams"):
            # FFB200_STREAM_PRIO="<dat>,<nafnet>" (0 = default, -1 = high): block-scheduling priority of the side streams
            prio = [int(v) for v in os.environ.get("FFB200_STREAM_PRIO", "0,0").split(",")]
            self._streams = [torch.cuda.Stream(device=self.device, priority=prio[i]) for i in range(2)]
        cur = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(cur)
        done = []
        for s, (name, off) in zip(self._streams, (("dat", 3), ("nafnet", 6))):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                r[name].forward(lr, stack, off)
                e = torch.cuda.Event()
                e.record(s)
                done.append(e)
        r["hat"].forward(lr, stack, 0)
        for e in done:
            cur.wait_event(e)
        return stack

    @torch.no_grad()
    def forward(self, lr, out=None, intermediates=None):
        """lr: fp32 NCHW [B,3,h,w] in [0,1] on the model's device, ANY h x w >= 9 -> fp32 NCHW [B,3,4h,4w]: the reference's
        whole-image forward (CompleteEnhancedFusionSR.forward, enhanced_fusion.py:694-754) with its padding paths -- reflect
        padding to the HAT / DAT window (expert_loader.py:63-91), DAT's zero padding to 32 with run-time masks
        (dat_arch.py:505-528), NAFNet's zero padding to 16 (nafnet_arch.py:219-225), the DCT's reflect padding to 8, DFTs of any
        length, floor-sized pyramids.
        `intermediates` (a dict to fill) forces the eager path: the CUDA-graph replay keeps no per-stage tensors."""
        self._check_input(lr, "FreqFusionB200.forward")
        lr = lr.contiguous().float()
        B, _, h, w = lr.shape
        with torch.cuda.device(self.device):
            self._begin_forward()
            if intermediates is None and B * h * w <= GRAPH_MAX_LR_PIXELS and os.environ.get("FFB200_GRAPHS", "1") != "0":
                return self._forward_graphed(lr, out)
            stack = self.forward_experts(lr)
            return self._runners["head"].forward(lr, stack, out=out, intermediates=intermediates)

    __call__ = forward

    @staticmethod
    def supports_whole_image(h, w):
        """Sizes the forward accepts: everything the reference's forward accepts.  Its first reflect padding (right / bottom, up
        to the next multiple of 16) needs pad < size, i.e. sides of at least 9 px."""
        return min(h, w) >= MIN_SIDE

    def forward_any(self, lr, out=None):
        """Batch of equal-size LR images / tiles (kept for the plugin pipeline; forward() takes any supported size)."""
        return self.forward(lr, out=out)

    def _forward_graphed(self, lr, out):
        """Small batches are bound by the ~2 000 host-side launches of a forward (~10 us each), not by the GPU: the forward
        of each (B, h, w) shape is captured once into a CUDA graph (static input / output buffers, the cached workspaces keep
        every address stable, the three expert streams fork and join inside the capture) and replayed.  At most
        GRAPH_CACHE_MAX shapes stay captured (least recently used first out)."""
        key = tuple(lr.shape)
        ent = self._graphs.get(key)
        if ent is None:
            B, _, h, w = lr.shape
            x_s = lr.clone()
            o_s = torch.empty(B, 3, 4 * h, 4 * w, dtype=torch.float32, device=self.device)
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):      # allocates the workspaces and configures the kernels outside the capture
                    self._runners["head"].forward(x_s, self.forward_experts(x_s), out=o_s)
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    self._runners["head"].forward(x_s, self.forward_experts(x_s), out=o_s)
            cur.wait_stream(side)
            ent = self._graphs[key] = (graph, x_s, o_s)
            while len(self._graphs) > GRAPH_CACHE_MAX:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)      # (an eviction of its workspaces drops the graph too: _begin_forward)
        graph, x_s, o_s = ent
        x_s.copy_(lr)
        graph.replay()
        if out is None:
            return o_s.clone()
        out.copy_(o_s)
        return out

    @torch.no_grad()
    def forward_with_precomputed(self, lr, expert_outputs, expert_features=None, out=None, intermediates=None):
        """Fusion head only, on pre-computed expert SR outputs (reference CompleteEnhancedFusionSR.forward_with_precomputed,
        src/models/enhanced_fusion.py:756-812, eval path; BASELINE.json configs[0]).
        expert_outputs: dict with keys 'hat', 'dat', 'nafnet' (the cached-dataset aliases 'drct' -> hat and 'grl' / 'mambair' -> dat
        of src/data/cached_dataset.py are accepted), each fp32 NCHW [B,3,4h,4w] on the device.
        expert_features: optional dict of the experts' intermediate features ({'hat': [B,180,h,w], 'dat': [B,180,h,w],
        'nafnet': [B,64,h,w]}); when given, the collaborative branch (EnhancedCollaborativeWithLKA) modulates the expert outputs first,
        exactly as the reference does in cached mode -- including its handling of other channel counts (truncated / zero padded), of
        larger feature maps (resized to the smallest, which must have the LR size) and of experts without features (zeros)."""
        alias = {"drct": "hat", "grl": "dat", "mambair": "dat"}
        ex = {alias.get(k, k): v for k, v in expert_outputs.items()}
        missing = [k for k in ("hat", "dat", "nafnet") if k not in ex]
        if missing:
            raise KeyError(f"forward_with_precomputed: missing expert outputs {missing}")
        self._check_input(lr, "forward_with_precomputed")
        lr = lr.contiguous().float()
        B, _, h, w = lr.shape
        with torch.cuda.device(self.device):
            self._begin_forward()
            stack = self._stack(B, h, w)
            for i, k in enumerate(("hat", "dat", "nafnet")):
                t = ex[k]
                if not t.is_cuda or tuple(t.shape) != (B, 3, 4 * h, 4 * w):
                    raise L.FFError(f"forward_with_precomputed: expert output '{k}' must be a CUDA tensor of shape {(B, 3, 4 * h, 4 * w)}")
                # NCHW -> channels 3i..3i+2 of the NHWC expert stack (layout plumbing; no arithmetic)
                stack.view(B, 4 * h, 4 * w, 12)[..., 3 * i:3 * i + 3].copy_(t.float().permute(0, 2, 3, 1))
            if expert_features is not None:
                # phase 4 of the reference: collaborative feature learning runs whenever features are passed
                # (apply_collaborative_learning, enhanced_fusion.py:466-496; MODEL_CONFIG enables it)
                # (an expert without features contributes zeros to the cross-expert attention, large_kernel_attention.py:374-377)
                fe = {alias.get(k, k): v for k, v in expert_features.items()}
                self._runners["head"].collaborative(fe, stack, B, h, w, intermediates=intermediates)
            return self._runners["head"].forward(lr, stack, out=out, intermediates=intermediates)

    def expert_outputs_nchw(self, lr):
        """Testing helper: dict of NCHW expert outputs like ExpertEnsemble.forward_all(return_dict=True)."""
        self._check_input(lr, "expert_outputs_nchw")
        B, _, h, w = lr.shape
        out = {}
        with torch.cuda.device(self.device):
            self._begin_forward()
            stack = self.forward_experts(lr.contiguous().float())
            for i, name in enumerate(("hat", "dat", "nafnet")):
                t = torch.empty(B, 3, 4 * h, 4 * w, dtype=torch.float32, device=self.device)
                ops.nhwc_to_nchw(stack, 3 * i, 3, t)
                out[name] = t
        return out
