"""stable-renderer_b200 — B200-native correspondence-map latent overlap step and UV-texture bake.

Python surface mirrors the reference's (`OverlapCorresponder.step_finished`, `CorrespondMap.update`,
`Overlap.__call__`, `Scheduler`, node classes); arithmetic runs in hand-written sm_100a CUDA kernels reached
through the C-ABI library `csrc/libsrx.so` (declared in `include/srx.h`).  There is no CPU fallback."""
__version__ = "0.1.0"
