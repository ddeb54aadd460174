"""cellranger_b200 — B200-native barcode / UMI correction and counting (one hot path of Cell Ranger).

CUDA kernels and the C ABI live in csrc/ (built into libcrgpu.so); api.py is the host side that mirrors
the reference's interfaces for this path. There is no CPU implementation in this package.
"""
from .api import (BarcodeCorrector, ChemistryDef, CountMatrix, FeatureReference, GemWell, Posterior,  # noqa: F401
                  SegmentedBarcodeCorrector, Whitelist, WhitelistHistogram, ascii_matrix, check_barcodes_compatibility, comm_unique_id, tethered_min_read_length, tethered_offset, unpack_2bit)
from ._lib import CrgpuError, NO_FEATURE, NO_RANK  # noqa: F401

__all__ = ["BarcodeCorrector", "SegmentedBarcodeCorrector", "ChemistryDef", "CountMatrix", "FeatureReference", "GemWell", "Posterior",
           "Whitelist", "WhitelistHistogram", "check_barcodes_compatibility", "comm_unique_id", "CrgpuError", "NO_FEATURE", "NO_RANK", "ascii_matrix", "tethered_offset", "tethered_min_read_length", "unpack_2bit"]
