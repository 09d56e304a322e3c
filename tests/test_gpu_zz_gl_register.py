"""-m gpu, runs last: the GL registration entry point must fail cleanly on a box without a GL context."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_gl_register_without_context_fails_cleanly():
    from stable_renderer_b200 import _lib
    lib = _lib.load()
    res = C.c_void_p()
    rc = lib.srx_gl_register_image(C.byref(res), 1, 0x0DE1, 0)     # GL_TEXTURE_2D; no GL context on the box
    assert rc == _lib.SRX_ERR_CUDA and not res.value
    assert b"cudaGraphicsGLRegisterImage" in lib.srx_last_error()
    # the failure must not poison later launches
    x = torch.ones(8, device="cuda")
    assert float((x * 2).sum()) == 16.0
