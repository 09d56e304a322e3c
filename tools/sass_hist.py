"""Opcode histogram per kernel of the in-tree library (cuobjdump works without a GPU; this .so is the one that travels):
    python tools/sass_hist.py k_overlap_fused > profiles/sass_k_overlap_fused.txt
The first block per function lists the memory / synchronisation / reduction opcodes that carry the design (bulk copies, mbarrier
waits, vector reductions, 256-bit loads, multimem), the second the 25 most frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "stable-renderer_b200", "csrc", "libsrx.so")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "k_overlap_fused")
KEY = re.compile(r"^(UBLKCP|UBLKRED|UBLKPF|SYNCS|RED|REDG|ATOM|ATOMG|ATOMS|LDG|STG|LDS|STS|LD\.|ST\.|LD$|ST$|REDUX|SHFL|MATCH|VOTE|MEMBAR|FENCE|ERRBAR|CCTL|"
                 r"BAR|UCGABAR|ELECT|NANOSLEEP|LDGSTS|LDGDEPBAR|DEPBAR|UTMA|UTCMMA|ACQBULK|CS2R|S2UR)")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
hist, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1) if pat.search(m.group(1)) else None
        if fn:
            hist[fn] = collections.Counter()
        continue
    if fn is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Za-z0-9_.]+)", line)
    if m:
        hist[fn][m.group(1)] += 1
print(f"# cuobjdump -sass {os.path.basename(lib)}: architectures {arch}; functions matching /{pat.pattern}/")
for fn, h in hist.items():
    demangled = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    print(f"\n## {demangled}   ({sum(h.values())} instructions)")
    print("# memory / synchronisation / reduction opcodes")
    for op, n in sorted(h.items(), key=lambda kv: (-kv[1], kv[0])):
        if KEY.match(op):
            print(f"{n:7d}  {op}")
    print("# most frequent opcodes")
    for op, n in h.most_common(25):
        print(f"{n:7d}  {op}")
