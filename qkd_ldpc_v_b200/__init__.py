"""qkd_ldpc_v_b200 -- B200-native batched LDPC syndrome decoder for QKD information reconciliation.

The product is libqkdldpc_cuda.so (hand-written sm_100a kernels behind the C ABI of include/qkdldpc.h) plus
host code that mirrors the reference's surface for the hot path. Importing this package does not touch the GPU;
creating an :class:`LdpcCode` does, and fails loudly when the CUDA library or a device is missing.
"""
from .decoder import (ALG_NAMES, DEC_ANMSA, DEC_AOMSA, DEC_NMSA, DEC_OMSA, DEC_SPA, DEC_SPA_APPROX,  # noqa: F401
                      FLAG_KEYS_MATCH, FLAG_SYNDROMES_MATCH, BatchResult, DecoderConfig, LdpcCode, pack_bits,
                      stats_from_tally, tally_len, unpack_bits)

__version__ = "0.1.0"
