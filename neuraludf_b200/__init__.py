"""neuraludf_b200 -- B200-native (sm_100a) implementation of NeuralUDF's volume-rendering hot path.

Host side: the reference's module / renderer API (`neuraludf_b200.models.*` mirror `models.fields`,
`models.embedder`, `models.udf_renderer_blending` of xxlong0/NeuralUDF).  Device side: libnudf.so (C-ABI in
include/nudf.h), hand-written CUDA kernels.  No CPU fallback.
"""
__version__ = "0.1.0"
