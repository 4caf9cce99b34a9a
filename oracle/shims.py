"""Register the oracle restatement as ``basicsr`` / ``realesrgan`` so the UNMODIFIED reference
``/root/reference/nesr/nesr.py`` runs end to end on CPU.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference checks
``importlib.util.find_spec("basicsr")`` / ``("realesrgan")`` (``nesr/nesr.py:153,157``) before
importing ``basicsr.archs.rrdbnet_arch.RRDBNet`` and ``realesrgan.RealESRGANer``
(``nesr/nesr.py:161-162``), so the fake modules need real ``ModuleSpec`` objects.

``write_checkpoint`` writes a ``{'params_ema': state_dict}`` file in the published format at one
of the locations the reference searches (``nesr/nesr.py:181-188``).
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types


def _module(name: str, is_pkg: bool) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=is_pkg)
    if is_pkg:
        mod.__path__ = []
    sys.modules[name] = mod
    return mod


def install_shims(rrdbnet_cls=None, upsampler_cls=None) -> None:
    """Install ``basicsr`` and ``realesrgan`` stand-ins (oracle classes unless others are given)."""
    if rrdbnet_cls is None:
        from .rrdbnet import RRDBNet as rrdbnet_cls
    if upsampler_cls is None:
        from .realesrganer import RealESRGANer as upsampler_cls
    basicsr = _module("basicsr", True)
    archs = _module("basicsr.archs", True)
    arch = _module("basicsr.archs.rrdbnet_arch", False)
    arch.RRDBNet = rrdbnet_cls
    basicsr.archs = archs
    archs.rrdbnet_arch = arch
    rg = _module("realesrgan", True)
    rg.RealESRGANer = upsampler_cls


def remove_shims() -> None:
    for name in ("basicsr", "basicsr.archs", "basicsr.archs.rrdbnet_arch", "realesrgan"):
        sys.modules.pop(name, None)


def write_checkpoint(state_dict, root: str, key: str = "params_ema") -> str:
    """Save ``{key: state_dict}`` to ``<root>/models/weights/RealESRGAN_x2plus.pth`` (a path the
    reference finds when its working directory is ``root``)."""
    import torch
    path = os.path.join(root, "models", "weights", "RealESRGAN_x2plus.pth")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({key: state_dict}, path)
    return path


def import_reference(reference_root: str = "/root/reference"):
    """Import the reference's pipeline class without writing bytecode into the read-only tree."""
    sys.dont_write_bytecode = True
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from nesr.nesr import SuperResolutionPipeline
    return SuperResolutionPipeline
