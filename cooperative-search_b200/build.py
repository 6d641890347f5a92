"""Builds csrc/libcoopsearch.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

Every translation unit is compiled to its own object under csrc/build/ (in parallel; flight_tpe.cu four times, once
per pair of agent counts) and the objects are linked into one shared library.  Staleness is decided by content
hashes recorded next to the objects, not by mtimes, so a copy of the tree (the GPU box) never rebuilds what it
received."""
import hashlib
import json
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(CSRC, "libcoopsearch.so")
STAMP = os.path.join(OBJ_DIR, "stamp.json")
HEADERS = ["cs_common.cuh", "cs_philox.cuh", "flight_common.cuh", "flight_map.cuh", "flight_internal.h",
           os.path.join("..", "..", "include", "coopsearch.h")]
# object name -> (source, extra flags)
UNITS = {
    "runtime": ("runtime.cu", []),
    "search": ("search.cu", []),
    "policy": ("policy.cu", []),
    "spread": ("spread.cu", []),
    "flight_host": ("flight_host.cu", []),
    "flight_hostio": ("flight_hostio.cu", []),
    "flight_lpa": ("flight_lpa.cu", []),
    "flight_aux": ("flight_aux.cu", []),
    "flight_mapk": ("flight_mapk.cu", []),
    "flight_tpe0": ("flight_tpe.cu", ["-DCS_TPE_PART=0"]),
    "flight_tpe1": ("flight_tpe.cu", ["-DCS_TPE_PART=1"]),
    "flight_tpe2": ("flight_tpe.cu", ["-DCS_TPE_PART=2"]),
    "flight_tpe3": ("flight_tpe.cu", ["-DCS_TPE_PART=3"]),
}

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",          # keep the reference's separate mul/add roundings
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libcoopsearch.so must be prebuilt (python -c 'import __graft_entry__ as g; g.build()')")
    return exe


def sources():
    return sorted({os.path.join(CSRC, src) for src, _ in UNITS.values()})


def _sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def _headers_digest():
    h = hashlib.sha256()
    for name in HEADERS + ["policy_tc.cuh"]:
        path = os.path.normpath(os.path.join(CSRC, name))
        if os.path.exists(path):
            h.update(name.encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS + [os.environ.get("CS_NVCC_EXTRA", "")]).encode())
    return h.hexdigest()


def _unit_digest(name, headers_digest):
    src, extra = UNITS[name]
    return hashlib.sha256((headers_digest + _sha(os.path.join(CSRC, src)) + " ".join(extra)).encode()).hexdigest()


def _load_stamp():
    try:
        with open(STAMP) as fh:
            return json.load(fh)
    except Exception:
        return {}


def stale_units():
    """Objects whose source, headers or flags differ from what they were built from."""
    stamp, hd = _load_stamp(), _headers_digest()
    return [n for n in UNITS if stamp.get(n) != _unit_digest(n, hd) or not os.path.exists(os.path.join(OBJ_DIR, n + ".o"))]


def is_stale():
    return not os.path.exists(LIB_PATH) or bool(stale_units()) or _load_stamp().get("__lib__") != "linked"


def build_library(force=False, verbose=False, jobs=None):
    """Compile the stale translation units (all with force=True) in parallel and link.  Returns the library path."""
    todo = list(UNITS) if force else stale_units()
    if not todo and os.path.exists(LIB_PATH) and _load_stamp().get("__lib__") == "linked":
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc, hd = _nvcc(), _headers_digest()
    extra_all = os.environ.get("CS_NVCC_EXTRA", "").split()
    logs = {}

    def compile_unit(name):
        src, extra = UNITS[name]
        cmd = [nvcc] + NVCC_FLAGS + extra_all + extra + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", os.path.join(OBJ_DIR, name + ".o")]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-6000:]))
        logs[name] = proc.stderr
        return name

    stamp = _load_stamp()
    stamp["__lib__"] = "pending"
    with ThreadPoolExecutor(max_workers=jobs or min(len(todo) or 1, os.cpu_count() or 1)) as pool:
        for name in pool.map(compile_unit, todo):
            stamp[name] = _unit_digest(name, hd)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + \
          [os.path.join(OBJ_DIR, n + ".o") for n in UNITS]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), proc.stderr[-4000:]))
    stamp["__lib__"] = "linked"
    with open(STAMP, "w") as fh:
        json.dump(stamp, fh, indent=1)
    if verbose:
        for name in todo:
            print("==== %s\n%s" % (name, logs.get(name, "")))
    return LIB_PATH
