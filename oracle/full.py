"""ORACLE (test infrastructure only).  Full FreqFusion eval forward from the four oracle pieces, mirroring
CompleteEnhancedFusionSR.forward (/root/reference/src/models/enhanced_fusion.py:694-754) with
ExpertEnsemble.forward_all (src/models/expert_loader.py:768-777)."""
import torch

from . import dat as odat
from . import hat as ohat
from . import head as ohead
from . import nafnet as onaf


@torch.no_grad()
def forward_experts(state, lr):
    return [ohat.forward_hat(state["hat"], lr), odat.forward_dat(state["dat"], lr), onaf.forward_nafnet(state["nafnet"], lr)]


@torch.no_grad()
def forward(state, lr, return_intermediates=False):
    ex = forward_experts(state, lr)
    out = ohead.head_forward(state["fusion"], lr, ex, return_intermediates)
    if return_intermediates:
        out[1]["expert_outputs"] = ex
    return out
