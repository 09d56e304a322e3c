"""Is there any way to get a GL context on this box?  Tries EGL (device platform, surfaceless), then reports which GL
libraries exist.  If a context comes up: creates an RGBA16F texture, registers it with cudaGraphicsGLRegisterImage through
srx_gl_register_image, writes it with Texture.set_data and reads it back.  Output: one line per step (profiles/r2_egl_probe.txt)."""
import ctypes as C
import ctypes.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

for name in ("EGL", "GL", "OpenGL", "GLESv2", "OSMesa", "GLX", "nvidia-eglcore", "EGL_nvidia"):
    print(f"find_library({name}) = {ctypes.util.find_library(name)}")
for path in ("/usr/lib/x86_64-linux-gnu/libEGL.so.1", "/usr/lib/x86_64-linux-gnu/libEGL_nvidia.so.0", "/usr/lib/x86_64-linux-gnu/libGL.so.1",
             "/usr/lib/x86_64-linux-gnu/libOSMesa.so.8", "/usr/lib/x86_64-linux-gnu/libnvidia-eglcore.so"):
    print(f"{path}: {'present' if os.path.exists(path) else 'absent'}")
os.system("ls /usr/lib/x86_64-linux-gnu | grep -i -E 'egl|libgl|mesa|glvnd' | head -20; ls /usr/share/glvnd/egl_vendor.d 2>/dev/null")

egl = None
for cand in ("libEGL.so.1", "libEGL_nvidia.so.0"):      # the glvnd dispatcher, else the driver's vendor library directly
    try:
        egl = C.CDLL(cand)
        print(f"loaded {cand}")
        break
    except OSError as e:
        print(f"{cand} cannot be loaded: {e}")
if egl is None or not hasattr(egl, "eglGetProcAddress"):
    print("RESULT: no GL context possible on this box; cudaGraphicsGLRegisterImage cannot be exercised")
    sys.exit(0)

EGL_PLATFORM_DEVICE_EXT, EGL_NONE, EGL_OPENGL_API = 0x313F, 0x3038, 0x30A2
egl.eglGetProcAddress.restype = C.c_void_p
egl.eglGetProcAddress.argtypes = [C.c_char_p]
q = egl.eglGetProcAddress(b"eglQueryDevicesEXT")
g = egl.eglGetProcAddress(b"eglGetPlatformDisplayEXT")
print(f"eglQueryDevicesEXT={q} eglGetPlatformDisplayEXT={g}")
if not q or not g:
    print("RESULT: EGL loads but has no device platform; no headless context")
    sys.exit(0)
QueryDevices = C.CFUNCTYPE(C.c_uint, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int))(q)
GetPlatformDisplay = C.CFUNCTYPE(C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_int))(g)
devs = (C.c_void_p * 16)()
n = C.c_int(0)
ok = QueryDevices(16, devs, C.byref(n))
print(f"eglQueryDevicesEXT ok={ok} devices={n.value}")
ctx_ok = False
for i in range(n.value):
    dpy = GetPlatformDisplay(EGL_PLATFORM_DEVICE_EXT, devs[i], None)
    major, minor = C.c_int(0), C.c_int(0)
    egl.eglInitialize.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    if not dpy or not egl.eglInitialize(dpy, C.byref(major), C.byref(minor)):
        print(f"device {i}: eglInitialize failed")
        continue
    print(f"device {i}: EGL {major.value}.{minor.value}")
    egl.eglBindAPI(EGL_OPENGL_API)
    cfg = C.c_void_p()
    ncfg = C.c_int(0)
    attrs = (C.c_int * 3)(0x3033, 0x0001, EGL_NONE)       # EGL_SURFACE_TYPE = PBUFFER
    egl.eglChooseConfig.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_int)]
    egl.eglChooseConfig(dpy, attrs, C.byref(cfg), 1, C.byref(ncfg))
    egl.eglCreateContext.restype = C.c_void_p
    egl.eglCreateContext.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    ctx = egl.eglCreateContext(dpy, cfg, None, None)
    egl.eglMakeCurrent.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    if not ctx or not egl.eglMakeCurrent(dpy, None, None, ctx):
        print(f"device {i}: no surfaceless context (configs={ncfg.value}, ctx={ctx})")
        continue
    ctx_ok = True
    print(f"device {i}: OpenGL context current")
    break
if not ctx_ok:
    print("RESULT: EGL present but no OpenGL context could be made current; cudaGraphicsGLRegisterImage cannot be exercised")
    sys.exit(0)
def glfn(name, restype, *argtypes):
    addr = egl.eglGetProcAddress(name.encode())
    if not addr:
        raise RuntimeError(f"{name} not found")
    return C.CFUNCTYPE(restype, *argtypes)(addr)


try:
    glGenTextures = glfn("glGenTextures", None, C.c_int, C.POINTER(C.c_uint))
    glBindTexture = glfn("glBindTexture", None, C.c_uint, C.c_uint)
    glTexStorage2D = glfn("glTexStorage2D", None, C.c_uint, C.c_int, C.c_uint, C.c_int, C.c_int)
    glGetError = glfn("glGetError", C.c_uint)
    glFinish = glfn("glFinish", None)
except RuntimeError as e:
    print(f"RESULT: context current but GL entry points missing: {e}")
    sys.exit(0)
tex = C.c_uint(0)
glGenTextures(1, C.byref(tex))
glBindTexture(0x0DE1, tex.value)
glTexStorage2D(0x0DE1, 1, 0x881A, 64, 32)              # GL_RGBA16F
glFinish()
print(f"GL texture {tex.value} created, glGetError = {glGetError()}")
import torch  # noqa: E402
from stable_renderer_b200.texture import Texture  # noqa: E402
t = Texture(64, 32, 4, torch.float16, gl_texture=tex.value)
data = torch.randn(32, 64, 4).half().cuda()
try:
    t.set_data(data)
    back = t.tensor(flip=False)
    print(f"RESULT: GL texture registered, written and read back through the mapped cudaArray: equal={bool(torch.equal(back, data))}")
except Exception as e:  # noqa: BLE001
    print(f"RESULT: GL context up, but the CUDA registration failed: {e}")
