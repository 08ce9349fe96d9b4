"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the UNMODIFIED reference tree (/root/reference, read-only, exists only in the
build container) so golden vectors can be generated and the oracle restatement pinned.
Recipe follows SURVEY.md Appendix C: stub `timm.models.layers` (three helper symbols,
used by src/models/hat/__init__.py:18 and src/models/expert_loader.py:284), put the
reference root on sys.path, never write bytecode into the read-only tree.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("FF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "models"))


def install():
    """Make `import src.models...` / `import models.team29_FreqFusion.io` resolve to the reference."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    import torch
    sys.dont_write_bytecode = True
    if "timm.models.layers" not in sys.modules:
        tl = types.ModuleType("timm.models.layers")
        tl.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        tl.trunc_normal_ = torch.nn.init.trunc_normal_
        tl.DropPath = torch.nn.Identity
        timm = types.ModuleType("timm")
        tm = types.ModuleType("timm.models")
        timm.models = tm
        tm.layers = tl
        sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})
    # the reference's top-level package names (`src`, `models`, `utils`) collide with ours
    for name in [n for n in list(sys.modules) if n == "models" or n.startswith("models.") or n == "src" or n.startswith("src.")]:
        del sys.modules[name]
    if REF_ROOT in sys.path:
        sys.path.remove(REF_ROOT)
    sys.path.insert(0, REF_ROOT)


def build_reference(device="cpu", quiet=True):
    """Return (ensemble, fusion_model, io_module) of the reference, random-init, eval mode."""
    install()
    import contextlib
    import io as _io
    import torch
    ctx = contextlib.redirect_stdout(_io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        from src.models import expert_loader
        from src.models.enhanced_fusion import CompleteEnhancedFusionSR
        from models.team29_FreqFusion import io as ffio
        ens = expert_loader.ExpertEnsemble(upscale=4, device=device)
        ens.load_all_experts({"hat": "/nonexistent", "dat": "/nonexistent", "nafnet": "/nonexistent"})
        cfg = {k: v for k, v in ffio.MODEL_CONFIG.items() if k != "scale"}
        model = CompleteEnhancedFusionSR(expert_ensemble=ens, upscale=4, **cfg).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return ens, model, ffio
