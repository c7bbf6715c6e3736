"""promptttspp_b200 -- B200-native (sm_100a) inference hot path of PromptTTS++.

The package mirrors the reference's module surface (constructor kwargs,
``state_dict`` keys, ``infer`` / ``infer_batch`` / ``forward`` signatures) so
that the reference's Hydra ``_target_`` paths can be re-pointed by replacing
the ``promptttspp.`` prefix with ``promptttspp_b200.``.  All arithmetic of the
hot path runs in the hand-written CUDA kernels of ``csrc/`` behind the C ABI
declared in ``include/pttspp_b200.h``; there is no CPU fallback.
"""

__version__ = "0.1.0"
