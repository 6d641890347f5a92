#!/bin/bash
# rebuild libcoopsearch.so (all objects) and print only register / spill lines of the kernels named by $1 (regex)
python - <<PY
import importlib, re, io, contextlib
b = importlib.import_module('cooperative-search_b200.build')
buf = io.StringIO()
try:
    with contextlib.redirect_stdout(buf):
        b.build_library(force=True, verbose=True)
except RuntimeError as e:
    print(str(e)[-5000:]); raise SystemExit(1)
out = buf.getvalue().splitlines()
pat = re.compile(r"${1:-fused_kernel|map_generic|search_kernelILi0|flight_kernelILi16ELi0}")
for i, l in enumerate(out):
    if 'Compiling entry function' in l and pat.search(l):
        print(l.split("'")[1][:110]); print('   ', out[i+2].strip()); print('   ', out[i+3].strip())
    if 'error' in l or 'warning' in l: print(l)
PY
