"""A/B timing helper: python tools/ab_time.py <libsrx.so> <tools script> [args...] — runs a tools/ script against another build."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stable_renderer_b200._lib as L  # noqa: E402

L.LIB_PATH = os.path.abspath(sys.argv[1])
script = sys.argv[2]
sys.argv = [script] + sys.argv[3:]
runpy.run_path(script, run_name="__main__")
