"""rend3_b200 — rend3's GPU-driven per-object hot path (cull + uniform bake + PBR opaque forward)
as hand-written sm_100a CUDA behind a C ABI (include/rend3_b200.h).

Package contents: `csrc/` (CUDA kernels + the C ABI, built into librend3_b200.so), `backend.py`
(ctypes binding), `routines.py` (host mirror of rend3-routine's interface for the path),
`world.py` / `glam.py` / `scenes.py` (stand-ins for rend3's Rust managers that build the std430
input buffers), `runner.py` (rend3-test's TestRunner mirror).  There is no CPU fallback.
"""
from .backend import CAMERA_VIEWPORT, Backend, R3Error, load_cuda_backend, load_cuda_library  # noqa: F401
