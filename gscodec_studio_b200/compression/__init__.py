"""On-disk side of the quantisation path: `PngCompression`'s quantise-to-grid, which shares the
per-attribute bounds-free min / max normalisation and bit widths with the in-loop simulation.  The entropy-coding / video
codecs of the reference (`gsplat/compression/{entropy_coding,hevc,hybrid,seq_hevc,stg}_compression.py`) are out of scope."""
from .png_compression import PngCompression, dequantize_from_grid, quantize_to_grid  # noqa: F401
