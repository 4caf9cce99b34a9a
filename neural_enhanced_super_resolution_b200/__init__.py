"""nesr-b200: the Real-ESRGAN x2plus upscaling stage of NESR as hand-written sm_100a CUDA behind the
reference's own Python API.

    from neural_enhanced_super_resolution_b200 import RRDBNet, RealESRGANer, SuperResolutionPipeline

``RRDBNet`` / ``RealESRGANer`` mirror ``basicsr.archs.rrdbnet_arch.RRDBNet`` / ``realesrgan.RealESRGANer``;
``SuperResolutionPipeline`` mirrors ``nesr.SuperResolutionPipeline`` for the ESRGAN path.  All compute
goes through ``libnesr_b200.so`` (C ABI in ``include/nesr_b200.h``); importing this package never
imports the CPU oracle and never falls back to torch convolutions.
"""
from . import _ffi
from ._ffi import Engine
from .pipeline import SuperResolutionPipeline, install, install_shims
from .realesrganer import RealESRGANer
from .rrdbnet import RRDBNet

__all__ = ["Engine", "RRDBNet", "RealESRGANer", "SuperResolutionPipeline", "install", "install_shims", "_ffi"]
__version__ = "0.1.0"
