"""TEST INFRASTRUCTURE ONLY — import shim that executes the reference's own hot-path files, unmodified,
from /root/reference.

The reference cannot be imported as a package in this container (taichi, PyOpenGL, pycuda, glm, PySide6,
diffusers, concurrent_log_handler are absent), so the *files* that hold the hot path are loaded one by one
with ``importlib.util.spec_from_file_location`` after ``sys.modules`` has been pre-seeded with inert
stand-ins for everything they import but do not need for arithmetic.

Nothing in the product package, the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this module:
/root/reference does not exist on the GPU box.  It is used by ``oracle/make_golden.py`` (fixture
generation) and by the ``-m "not gpu"`` differential tests, which skip when the mount is absent.

Loaded reference files (all read-only, nothing is copied):
  source/common_utils/math_utils.py
  source/engine/static/corrmap.py
  source/common_utils/stable_render_utils/corresponder.py
  legacy_codes/stable_rendering_algo/overlap/{algorithms,utils,overlap_scheduler,overlap}.py
  legacy_codes/stable_rendering_algo/data_classes/{common,correspondence_map}.py
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SRX_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "source", "common_utils", "math_utils.py"))


class _Anything:
    """Attribute-tolerant stand-in: any attribute / call / subscript yields another stand-in, and it is
    usable as a decorator (returns the decorated function unchanged)."""

    def __init__(self, name="stub"):
        self.__dict__["_name"] = name

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Anything(f"{self._name}.{item}")

    def __call__(self, *args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs and not isinstance(args[0], _Anything):
            return args[0]
        return _Anything(f"{self._name}()")

    def __getitem__(self, item):
        return _Anything(f"{self._name}[]")

    def __mro_entries__(self, bases):
        return (object,)

    def __or__(self, other):
        return self

    def __ror__(self, other):
        return self


class _StubModule(types.ModuleType):
    """Module whose missing attributes resolve to `_Anything`."""

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Anything(f"{self.__name__}.{item}")


def _stub(name, **attrs):
    m = _StubModule(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so that `import a.b` works
    sys.modules[name] = m
    return m


class _Logger:
    def _noop(self, *a, **k):
        pass

    debug = info = warn = warning = error = success = print = critical = _noop


def _load(modname: str, relpath: str, package: str | None = None):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    if package is not None:
        mod.__package__ = package
    sys.modules[modname] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    return mod


_LOADED: dict | None = None


def load_reference() -> dict:
    """Returns a dict of the loaded reference modules.  Idempotent."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    # ---- third-party stand-ins -------------------------------------------------------------------
    def _identity_decorator(*a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f

    ti = _stub("taichi", init=lambda *a, **k: None, kernel=_identity_decorator, func=_identity_decorator)
    ti.gpu = "gpu"
    for name in ("OpenGL", "OpenGL.GL", "OpenGL.error", "diffusers", "diffusers.utils", "PySide6",
                 "PySide6.QtCore", "PySide6.QtWidgets", "PySide6.QtGui", "dotenv"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["diffusers"].AutoencoderKL = type("AutoencoderKL", (), {})
    sys.modules["diffusers.utils"].PIL_INTERPOLATION = {}
    sys.modules["dotenv"].load_dotenv = lambda *a, **k: None

    # ---- reference-internal stand-ins (utilities that are off the arithmetic path) -------------------
    cu = _stub("common_utils")
    gv: dict = {}
    _stub("common_utils.global_utils",
          GetOrAddGlobalValue=lambda k, d=None: gv.setdefault(k, d),
          SetGlobalValue=lambda k, v: gv.__setitem__(k, v),
          GetGlobalValue=lambda k, d=None: gv.get(k, d),
          is_dev_mode=lambda: False, is_engine_looping=lambda: True, is_verbose_mode=lambda: False)
    lg = _Logger()
    _stub("common_utils.debug_utils", EngineLogger=lg, ComfyUILogger=lg, DefaultLogger=lg)
    _stub("common_utils.decorators", Overload=lambda f: f)
    _stub("common_utils.system_utils", is_windows=lambda: False)
    import tempfile
    from pathlib import Path

    def extract_index(file_path, i):  # same contract as common_utils/path_utils.py:173-179 (host-side file ordering)
        stem = os.path.basename(file_path).split(".")[0]
        if stem.split("_")[-1].isdigit():
            return int(stem.split("_")[-1])
        if os.path.basename(file_path).split("_")[0].isdigit():
            return int(os.path.basename(file_path).split("_")[0])
        return i

    _stub("common_utils.path_utils", TEMP_DIR=Path(tempfile.gettempdir()), MAP_OUTPUT_DIR=Path(tempfile.gettempdir()),
          RESOURCES_DIR=Path(REFERENCE_ROOT) / "resources", extract_index=extract_index)
    ds = _stub("common_utils.data_struct")
    se = _load("common_utils.data_struct.sortableElement", "source/common_utils/data_struct/sortableElement.py")
    for k in dir(se):
        if not k.startswith("_"):
            setattr(ds, k, getattr(se, k))

    eng = _stub("engine")

    class Texture:  # isinstance() target only
        pass

    class Color:
        pass

    class ResourcesObj:
        name = None

        def __attrs_post_init__(self):
            pass

        def clear(self):
            pass

        def load(self):
            pass

    # attrs needs `name` to be an attribute of the base for CorrespondMap(name=...)
    from attr import attrs, attrib

    @attrs(eq=False, repr=False)
    class ResourcesObjAttrs:
        name = attrib(default=None)

        def __attrs_post_init__(self):
            pass

        def clear(self):
            pass

        def load(self):
            pass

    _stub("engine.static", Texture=Texture, Color=Color)
    _stub("engine.static.resources_obj", ResourcesObj=ResourcesObjAttrs)
    _stub("engine.static.enums")
    _stub("engine.static.texture", Texture=Texture)

    # ---- the real reference files ---------------------------------------------------------------------
    out = {}
    out["math_utils"] = _load("common_utils.math_utils", "source/common_utils/math_utils.py", "common_utils")
    cu.math_utils = out["math_utils"]
    out["corrmap"] = _load("engine.static.corrmap", "source/engine/static/corrmap.py", "engine.static")
    sys.modules["engine.static"].IDMap = out["corrmap"].IDMap
    sys.modules["engine.static"].CorrespondMap = out["corrmap"].CorrespondMap
    sys.modules["engine.static"].UpdateMode = out["corrmap"].UpdateMode

    _stub("common_utils.stable_render_utils")
    out["corr_utils"] = _load("common_utils.stable_render_utils.corr_utils",
                              "source/common_utils/stable_render_utils/corr_utils.py",
                              "common_utils.stable_render_utils")
    out["corresponder"] = _load("common_utils.stable_render_utils.corresponder",
                                "source/common_utils/stable_render_utils/corresponder.py",
                                "common_utils.stable_render_utils")

    # legacy generation: load as package `legacy_algo` so that the relative imports resolve
    _stub("legacy_algo")
    _stub("legacy_algo.data_classes")
    _stub("legacy_algo.overlap")
    out["legacy_common"] = _load("legacy_algo.data_classes.common",
                                 "legacy_codes/stable_rendering_algo/data_classes/common.py", "legacy_algo.data_classes")
    sys.modules["legacy_algo.data_classes"].Rectangle = out["legacy_common"].Rectangle
    out["correspondence_map"] = _load("legacy_algo.data_classes.correspondence_map",
                                      "legacy_codes/stable_rendering_algo/data_classes/correspondence_map.py",
                                      "legacy_algo.data_classes")
    sys.modules["legacy_algo.data_classes"].CorrespondenceMap = out["correspondence_map"].CorrespondenceMap
    out["algorithms"] = _load("legacy_algo.overlap.algorithms",
                              "legacy_codes/stable_rendering_algo/overlap/algorithms.py", "legacy_algo.overlap")
    out["overlap_utils"] = _load("legacy_algo.overlap.utils",
                                 "legacy_codes/stable_rendering_algo/overlap/utils.py", "legacy_algo.overlap")
    out["overlap_scheduler"] = _load("legacy_algo.overlap.overlap_scheduler",
                                     "legacy_codes/stable_rendering_algo/overlap/overlap_scheduler.py", "legacy_algo.overlap")
    out["overlap"] = _load("legacy_algo.overlap.overlap",
                           "legacy_codes/stable_rendering_algo/overlap/overlap.py", "legacy_algo.overlap")
    # legacy node CorrMapLatentNoiseInitializer: needs only three names of the ComfyUI type system
    _stub("comfyUI")
    _stub("comfyUI.types", StableRenderingNode=type("StableRenderingNode", (), {}), INT=lambda *a, **k: int, LATENT=dict, torch=None)
    del sys.modules["comfyUI.types"].__dict__["torch"]
    _stub("stable_rendering")
    _stub("stable_rendering.src")
    _stub("stable_rendering.src.data_classes", CorrespondenceMap=out["correspondence_map"].CorrespondenceMap)
    out["legacy_latent_node"] = _load("legacy_nodes.latent", "legacy_codes/nodes/latent.py")
    _LOADED = out
    return out


@contextlib.contextmanager
def quiet():
    """Silences the reference's stray print() calls (corrmap.py:276,722; corresponder.py:345-347)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def _extract_function(relpath: str, name: str, namespace: dict):
    """Executes ONE top-level function of a reference file, unmodified, without importing the file (whose other
    imports — typeguard, pydantic — are absent here)."""
    import ast
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path) as f:
        src = f.read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, namespace)
            return namespace[name]
    raise KeyError(name)


def load_reference_nodes(ksampler) -> dict:
    """The reference's node files (legacy_codes/nodes/{samplers,schedulers}.py and
    source/comfyUI/stable_rendering/_nodes/samplers.py), unmodified, wired to `ksampler` in place of ComfyUI's
    `custom_ksampler` (the sampler belongs to the host application; a scripted one drives the callbacks in fixtures)."""
    import inspect
    import re
    from typing import Any, Optional
    R = load_reference()

    class _Choices:
        __args__ = ("euler", "ddim", "ddpm")

    ct = sys.modules["comfyUI.types"]
    ct.__dict__.update(
        StableRenderingNode=type("StableRenderingNode", (), {}), INT=lambda *a, **k: int, FLOAT=lambda *a, **k: float,
        LATENT=dict, MODEL=Any, EngineData=Any, SamplingCallbackContext=Any, VAEDecodeCallback=Any,
        COMFY_SAMPLERS=_Choices, COMFY_SCHEDULERS=_Choices, UIImage=lambda *a, **k: None, Optional=Optional)
    _stub("comfyUI.nodes", custom_ksampler=ksampler)
    _stub("common_utils.type_utils",
          is_empty_method=_extract_function("source/common_utils/type_utils.py", "is_empty_method",
                                            {"inspect": inspect, "re": re}))
    sru = sys.modules["common_utils.stable_render_utils"]
    sru.Corresponder = R["corresponder"].Corresponder
    sru.DefaultCorresponder = R["corresponder"].DefaultCorresponder
    sru.OverlapCorresponder = R["corresponder"].OverlapCorresponder
    _stub("stable_rendering.src.overlap", ResizeOverlap=R["overlap"].ResizeOverlap, Scheduler=R["overlap_scheduler"].Scheduler,
          overlap_algorithm_factory=R["algorithms"].overlap_algorithm_factory)
    _stub("stable_rendering.src.overlap.overlap_scheduler", Scheduler=R["overlap_scheduler"].Scheduler)
    out = {}
    out["legacy_samplers"] = _load("legacy_nodes.samplers", "legacy_codes/nodes/samplers.py")
    out["legacy_schedulers"] = _load("legacy_nodes.schedulers", "legacy_codes/nodes/schedulers.py")
    out["samplers"] = _load("ref_nodes.samplers", "source/comfyUI/stable_rendering/_nodes/samplers.py")
    return out


def post_atten_inject_body():
    """The body of ``OverlapCorresponder.post_atten_inject`` with its leading ``return origin_values`` (corresponder.py:228)
    removed: the reference's own statements for the wide-channel feature overlap, unreachable as shipped.  Returns a function
    ``f(self, block, engine_data, origin_values, layer)``."""
    import ast
    R = load_reference()
    path = os.path.join(REFERENCE_ROOT, "source/common_utils/stable_render_utils/corresponder.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "OverlapCorresponder":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "post_atten_inject":
                    assert isinstance(fn.body[0], ast.Return), "the early return is expected to be the first statement"
                    fn.body = fn.body[1:]
                    fn.returns = None
                    for a in fn.args.args:
                        a.annotation = None
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ast.fix_missing_locations(mod)
                    ns = dict(vars(R["corresponder"]))
                    exec(compile(mod, path, "exec"), ns)
                    return ns["post_atten_inject"]
    raise KeyError("post_atten_inject")


def render_manager_functions():
    """``RenderManager._save_frame_data`` (renderManager.py:877-948) and ``_wrapIdenticalGBufferTask`` (:88-133), cut out of
    the reference's file and compiled unmodified except that parameter annotations are dropped (they name engine classes).  The
    module itself cannot be imported (OpenGL context, window, managers): the two functions run against duck-typed stand-ins for the
    manager and its textures — see oracle/make_golden.py::ingest_cases.  Returns ``(save_frame_data(self), wrap_task(render_manager,
    task, shader, mesh, callback=None, save_to_temp=False))``."""
    import ast
    R = load_reference()
    path = os.path.join(REFERENCE_ROOT, "source/engine/managers/renderManager.py")
    with open(path) as f:
        tree = ast.parse(f.read())

    def strip(fn):
        fn.returns = None
        for a in fn.args.args + fn.args.kwonlyargs:
            a.annotation = None
        return fn
    found = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "_wrapIdenticalGBufferTask":
            found["wrap"] = strip(node)
        if isinstance(node, ast.ClassDef) and node.name == "RenderManager":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "_save_frame_data":
                    found["save"] = strip(fn)
    assert set(found) == {"wrap", "save"}, found.keys()
    import torch
    gl = _StubModule("OpenGL.GL")
    gl.glClear = lambda *a, **k: None
    gl.glDisable = gl.glEnable = lambda *a, **k: None
    gl.GL_DEPTH_BUFFER_BIT, gl.GL_COLOR_BUFFER_BIT, gl.GL_CULL_FACE = 0x100, 0x4000, 0x0B44
    ns = {"torch": torch, "gl": gl, "adaptive_instance_normalization": R["math_utils"].adaptive_instance_normalization}
    mod = ast.Module(body=[found["wrap"], found["save"]], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, path, "exec"), ns)
    return ns["_save_frame_data"], ns["_wrapIdenticalGBufferTask"]


def noise_node_call():
    """``CreateNoiseSequenceFromIdMap.__call__`` (source/comfyUI/stable_rendering/_nodes/loaders.py:154-271), cut out of the
    reference's file: parameter annotations dropped (ComfyUI type constructors) and the device string ``'cuda'`` replaced by
    ``'cpu'`` (five ``.to('cuda')`` calls — the build container has no GPU); every other statement runs as written, around the
    reference's real ``tensor_group_by_then_randn_init`` and ``IDMap``.  Returns ``f(self, id_map, seed, sd_version, downsample_option)``
    whose result is a dict with ``samples`` and ``noise``."""
    import ast
    import torch
    import torch.nn.functional as F
    R = load_reference()
    path = os.path.join(REFERENCE_ROOT, "source/comfyUI/stable_rendering/_nodes/loaders.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "CreateNoiseSequenceFromIdMap":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "__call__":
                    fn.returns = None
                    for a in fn.args.args:
                        a.annotation = None
                    swapped = 0
                    for sub in ast.walk(fn):
                        if isinstance(sub, ast.Constant) and sub.value == "cuda":
                            sub.value = "cpu"
                            swapped += 1
                    assert swapped >= 3, swapped
                    fn.name = "noise_node_call"
                    mod = ast.Module(body=[fn], type_ignores=[])
                    ast.fix_missing_locations(mod)
                    ns = {"torch": torch, "F": F, "LATENT": lambda **kw: dict(kw),
                          "tensor_group_by_then_randn_init": R["math_utils"].tensor_group_by_then_randn_init}
                    exec(compile(mod, path, "exec"), ns)
                    return ns["noise_node_call"]
    raise KeyError("CreateNoiseSequenceFromIdMap.__call__")


def taichi_cells_overlap_python():
    """``taichi_cells_overlap`` (corr_utils.py:110-134) executed as plain Python: the shim's taichi stand-in makes
    ``ti.kernel`` / ``ti.func`` identity decorators, so only ``ti.ndrange`` and ``ti.math.ivec2`` need real behaviour."""
    import itertools
    R = load_reference()
    ti = sys.modules["taichi"]
    ti.ndrange = lambda *ns: itertools.product(*[range(int(n)) for n in ns])
    m = _StubModule("taichi.math")
    m.ivec2 = lambda v: list(v) if isinstance(v, (list, tuple)) else v
    ti.math = m
    mod = R["corr_utils"]
    mod.ti = ti
    return mod.taichi_cells_overlap


def johnny_overlap_function():
    """``johnny_overlap.overlap`` (legacy_diffuser/modules/diffuser_pipelines/overlap/johnny_overlap.py:15-141) with ONE
    statement changed: ``beta = schedule(step, timestep, 'constant')`` (:38) omits a required argument and raises TypeError as
    shipped; it becomes ``beta = kwargs.get('beta', 0)``.  Everything else runs as written."""
    import ast
    R = load_reference()
    path = os.path.join(REFERENCE_ROOT, "legacy_codes/legacy_diffuser/modules/diffuser_pipelines/overlap/johnny_overlap.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    patched = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and getattr(node.targets[0], "id", None) == "beta" \
                and isinstance(node.value, ast.Call) and getattr(node.value.func, "id", None) == "schedule":
            node.value = ast.parse("kwargs.get('beta', 0)", mode="eval").body
            patched += 1
    assert patched == 1, "expected exactly one `beta = schedule(...)` assignment"
    ast.fix_missing_locations(tree)
    pk = "ldiff.modules.diffuser_pipelines.overlap"
    for name in ("ldiff", "ldiff.modules", "ldiff.modules.diffuser_pipelines", pk, "ldiff.modules.data_classes"):
        if name not in sys.modules:
            _stub(name)
    _stub(pk + ".overlap_scheduler", Scheduler=R["overlap_scheduler"].Scheduler)
    _stub(pk + ".utils", overlap_rate=lambda *a, **k: 0.0)
    _stub("ldiff.modules.data_classes.correspondenceMap", CorrespondenceMap=R["correspondence_map"].CorrespondenceMap)
    mod = types.ModuleType(pk + ".johnny_overlap")
    mod.__package__ = pk
    mod.__file__ = path
    sys.modules[mod.__name__] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(tree, path, "exec"), mod.__dict__)
    return mod.overlap
