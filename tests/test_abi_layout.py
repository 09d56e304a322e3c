"""The ctypes mirrors in stable-renderer_b200/_lib.py must have the memory layout of the structs in include/srx.h: a small C
program compiled with gcc prints sizeof / offsetof for every field, ctypes must agree (no GPU, no CUDA needed)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _structs():
    from stable_renderer_b200 import _lib
    return {name: obj for name, obj in vars(_lib).items()
            if isinstance(obj, type) and issubclass(obj, C.Structure) and obj is not C.Structure and name.startswith("srx_")}


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_ctypes_structs_match_the_header(tmp_path):
    structs = _structs()
    assert {"srx_plan_desc", "srx_step_args", "srx_bake_args", "srx_legacy_desc", "srx_legacy_args", "srx_noise_args", "srx_gbuffer",
            "srx_ingest_args", "srx_gbuffer_temp"} <= set(structs)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "srx.h"', 'int main(void) {']
    for name, st in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for ln in out.splitlines():
        name, field, value = ln.split()
        st = structs[name]
        if field == "size":
            assert C.sizeof(st) == int(value), f"sizeof({name}): header {value}, ctypes {C.sizeof(st)}"
        else:
            assert getattr(st, field).offset == int(value), f"offsetof({name}, {field}): header {value}, ctypes {getattr(st, field).offset}"
        seen += 1
    assert seen == sum(len(st._fields_) + 1 for st in structs.values())
