"""Host-application glue for the node classes: the ComfyUI type system when it is importable, inert stand-ins otherwise.

The reference's nodes are `StableRenderingNode` subclasses whose `__call__` signature IS the node schema
(source/comfyUI/types/node_base.py:179-334,688-693): parameter names, defaults and return annotations are the contract, and
are kept verbatim here.  Inside the ComfyUI fork (`comfyUI.types` importable) the real base class and annotation helpers are
used, so the nodes register like the reference's; anywhere else the stand-ins below keep the module importable and the
callbacks testable.  The sampler itself (`custom_ksampler`) belongs to the host: it is resolved lazily, or injected with
`set_ksampler`."""
from __future__ import annotations

import inspect
import re
from typing import Any, Callable, Optional

try:  # inside the reference's ComfyUI fork
    from comfyUI.types import (COMFY_SAMPLERS, COMFY_SCHEDULERS, FLOAT, INT, LATENT, MODEL, EngineData,  # type: ignore
                               SamplingCallbackContext, StableRenderingNode, VAEDecodeCallback)
    IN_COMFY = True
except Exception:  # noqa: BLE001 - any import problem means "not inside the host application"
    IN_COMFY = False

    class StableRenderingNode:  # type: ignore[no-redef]
        """Stand-in base: the real one builds the node schema from `__call__`'s signature."""
        Category = "stable-rendering"

    def INT(*args, **kwargs):  # type: ignore[no-redef]  # noqa: N802 - annotation helper, same spelling as the host's
        return int

    def FLOAT(*args, **kwargs):  # type: ignore[no-redef]  # noqa: N802
        return float

    MODEL = LATENT = EngineData = SamplingCallbackContext = VAEDecodeCallback = Any  # type: ignore[misc,assignment]

    class _Choices:
        def __init__(self, *names):
            self.__args__ = names

    COMFY_SAMPLERS = _Choices("euler", "ddim", "ddpm")       # type: ignore[assignment]
    COMFY_SCHEDULERS = _Choices("normal", "karras")          # type: ignore[assignment]

_ksampler: Optional[Callable] = None


def set_ksampler(fn: Optional[Callable]) -> None:
    """Inject the host's `custom_ksampler` (tests inject a scripted one)."""
    global _ksampler
    _ksampler = fn


def get_ksampler() -> Callable:
    if _ksampler is not None:
        return _ksampler
    try:
        from comfyUI.nodes import custom_ksampler  # type: ignore
        return custom_ksampler
    except Exception as e:  # noqa: BLE001
        raise RuntimeError("no sampler available: these nodes run inside the reference's ComfyUI fork "
                           "(comfyUI.nodes.custom_ksampler), or inject one with nodes.set_ksampler()") from e


def is_empty_method(method) -> bool:
    """True when the method holds nothing but a docstring / `pass` — how the reference's callers probe optional
    corresponder hooks (source/common_utils/type_utils.py:444-459)."""
    try:
        source = inspect.getsource(method)
    except (OSError, TypeError):
        return False
    doc = getattr(method, "__doc__", None)
    if doc:
        source = source.replace(doc, "")
    source = re.sub(re.compile(r"(async)?\s*def\s+\w+\s*\(.*\).*?:", re.MULTILINE | re.DOTALL), "", source, count=1)
    lines = [ln.strip() for ln in source.split("\n") if ln.strip()]
    lines = [ln for ln in lines if not ln.startswith(("#", '"""', "'''")) and ln not in ("pass", "...")]
    return not lines


def dev_looping() -> bool:
    try:
        from common_utils.global_utils import is_dev_mode, is_engine_looping  # type: ignore
        return bool(is_dev_mode() and is_engine_looping())
    except Exception:  # noqa: BLE001
        return False
