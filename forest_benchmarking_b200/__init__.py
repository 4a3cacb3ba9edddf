"""forest_benchmarking_b200 -- B200-native batched quantum tomography and superoperator algebra.

Drop-in for the data-analysis hot path of rigetti/forest-benchmarking: same Python call signatures
(``tomography``, ``operator_tools``, ``distance_measures``), with ``*_batch`` twins that take and
return device tensors.  All arithmetic runs in hand-written sm_100a CUDA kernels (``csrc/``) reached
through a ctypes C ABI (``include/qtomo.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
