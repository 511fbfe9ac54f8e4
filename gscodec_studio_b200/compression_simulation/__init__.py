"""In-loop compression simulation (fake quantisation + entropy models) on the B200 kernels.

Same import surface as the reference's ``gsplat.compression_simulation`` (__init__.py:1):
``CompressionSimulation``, ``STGCompressionSimulation``; ops in ``.ops``; models in ``.entropy_model``.
"""
from .simulation import CompressionSimulation, STGCompressionSimulation  # noqa: F401
from .ops import fake_quantize, fake_quantize_ste  # noqa: F401
