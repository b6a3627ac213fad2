"""dffinthewild_b200 — B200-native (sm_100a) depth-from-focus network behind the reference's nn.Module API.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C-ABI library), `runtime.py` (ctypes
binding, weight packing, autograd glue) and the drop-in modules `Depth_Estimation_Network` / `End_to_End`.
"""
__all__ = ["Depth_Estimation_Network", "runtime"]
