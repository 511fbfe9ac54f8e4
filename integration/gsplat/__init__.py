"""`import gsplat` shim: the reference's package name on top of gscodec_studio_b200.

Put this directory's parent (`integration/`) on PYTHONPATH *instead of* the reference's `gsplat` package and the
reference's trainers (`examples/simple_trainer*.py`, `profiling/main.py`, the compression benchmarks) import the
B200-native hot path without an edit:

    gsplat.rasterization, gsplat.fully_fused_projection, gsplat.isect_tiles, ...    -> gscodec_studio_b200
    gsplat.rendering / gsplat.cuda._wrapper / gsplat.distributed / gsplat.relocation
    gsplat.compression_simulation (+ .ops / .simulation / .entropy_model / .ada_mask)
    gsplat.optimizers (SelectiveAdam)

Everything else of the reference package (densification strategies, test-time codecs) is OUTSIDE the
hot path and is not re-implemented; when GSPLAT_REFERENCE_PATH points at a checkout of the reference repository those
sub-packages are served from there (by path, unmodified), otherwise asking for them raises an ImportError that says so.
"""
import importlib
import os
import sys

import gscodec_studio_b200 as _impl
from gscodec_studio_b200 import *  # noqa: F401,F403
from gscodec_studio_b200 import __version__  # noqa: F401

_ALIASES = {
    "gsplat.rendering": "gscodec_studio_b200.rendering",
    "gsplat.distributed": "gscodec_studio_b200.distributed",
    "gsplat.relocation": "gscodec_studio_b200.relocation",
    "gsplat.profile": "gscodec_studio_b200.profile",
    "gsplat._helper": "gscodec_studio_b200._helper",
    "gsplat.utils": "gscodec_studio_b200.utils",
    "gsplat.cuda": "gscodec_studio_b200.cuda",
    "gsplat.cuda._wrapper": "gscodec_studio_b200.cuda._wrapper",
    "gsplat.optimizers": "gscodec_studio_b200.optimizers",
    "gsplat.compression_simulation": "gscodec_studio_b200.compression_simulation",
    "gsplat.compression_simulation.ops": "gscodec_studio_b200.compression_simulation.ops",
    "gsplat.compression_simulation.simulation": "gscodec_studio_b200.compression_simulation.simulation",
    "gsplat.compression_simulation.entropy_model": "gscodec_studio_b200.compression_simulation.entropy_model",
    "gsplat.compression_simulation.ada_mask": "gscodec_studio_b200.compression_simulation.ada_mask",
}
for _alias, _target in _ALIASES.items():
    try:
        sys.modules[_alias] = importlib.import_module(_target)
    except ImportError:  # pragma: no cover - an optional piece that is not built
        pass
for _name in ("rendering", "distributed", "relocation", "cuda", "optimizers", "compression_simulation"):
    globals()[_name] = sys.modules.get("gsplat." + _name)

# sub-packages we do not provide resolve against the reference checkout, if one is named
_REF = os.environ.get("GSPLAT_REFERENCE_PATH")
if _REF and os.path.isdir(os.path.join(_REF, "gsplat")):
    __path__.append(os.path.join(_REF, "gsplat"))

_REFERENCE_ONLY = {
    "DefaultStrategy": "strategy", "MCMCStrategy": "strategy", "Strategy": "strategy", "STG_Strategy": "strategy",
    "PngCompression": "compression", "rasterization_inria_wrapper": None, "accumulate": None,
}
if hasattr(_impl.optimizers, "SelectiveAdam"):
    SelectiveAdam = _impl.optimizers.SelectiveAdam


def __getattr__(name):
    sub = _REFERENCE_ONLY.get(name, "?")
    if sub not in (None, "?") and _REF:
        return getattr(importlib.import_module("gsplat." + sub), name)
    if sub != "?":
        raise ImportError(
            f"gsplat.{name} is outside the rasterization hot path gscodec_studio_b200 re-implements"
            + ("" if sub is None else f"; set GSPLAT_REFERENCE_PATH to a checkout of the reference to have gsplat.{sub} served from it"))
    raise AttributeError(f"module 'gsplat' has no attribute {name!r}")
