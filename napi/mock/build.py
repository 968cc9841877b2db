"""Build the N-API addon (napi/rbk_napi.cc) against the MOCK runtime in this directory and link the harness with
librbk_knn.so:

    python napi/mock/build.py            # -> napi/mock/_build/harness_real

Test scaffolding (see node_api.h here): the real build is node-gyp against Node's own headers.
(The CPU-only variant of the harness, linked against an oracle-backed stand-in of the C ABI, is built by
tests/napi_shim/build.py - nothing outside tests/ touches oracle/.)
"""
from __future__ import annotations

import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
HERE = Path(__file__).resolve().parent
OUT = HERE / "_build"
CXX = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror"]
HEADERS = [HERE / "node_api.h", HERE / "mock_napi.h", ROOT / "include" / "rbk_knn.h"]


def stale(target: Path, sources) -> bool:
    return not target.exists() or any(Path(s).stat().st_mtime > target.stat().st_mtime for s in sources)


def run(cmd) -> None:
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(str(c) for c in cmd) + "\n" + r.stdout + r.stderr)


def build_objects() -> list[Path]:
    """addon + mock runtime + harness, compiled with -Wall -Wextra -Werror."""
    OUT.mkdir(exist_ok=True)
    objs = []
    for src, extra in ((ROOT / "napi" / "rbk_napi.cc", ["-DNODE_GYP_MODULE_NAME=rbk_knn", "-I", ROOT / "include"]),
                       (HERE / "mock_napi.cc", []), (HERE / "harness.cc", [])):
        obj = OUT / (src.stem + ".o")
        if stale(obj, [src] + HEADERS):
            run(CXX + ["-I", HERE] + extra + ["-c", src, "-o", obj])
        objs.append(obj)
    return objs


def build() -> Path:
    """The harness linked against librbk_knn.so (which must exist: python -m runbookai_b200.build)."""
    objs = build_objects()
    exe = OUT / "harness_real"
    lib = ROOT / "runbookai_b200" / "lib" / "librbk_knn.so"
    if stale(exe, objs + [lib]):
        run(["g++"] + objs + ["-o", exe, "-L", lib.parent, "-lrbk_knn", f"-Wl,-rpath,{lib.parent}", "-lpthread"])
    return exe


if __name__ == "__main__":
    print(build())
