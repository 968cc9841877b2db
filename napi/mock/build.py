"""Build the N-API addon (napi/rbk_napi.cc) against the MOCK runtime in this directory and link the harness:

    python napi/mock/build.py            # -> napi/mock/_build/harness_real  (librbk_knn.so)
                                         #    napi/mock/_build/harness_shim  (tests/napi_shim: oracle-backed, no GPU)

Test scaffolding (see node_api.h here): the real build is node-gyp against Node's own headers.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
HERE = Path(__file__).resolve().parent
OUT = HERE / "_build"
CXX = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror"]


def _stale(target: Path, sources) -> bool:
    return not target.exists() or any(Path(s).stat().st_mtime > target.stat().st_mtime for s in sources)


def _run(cmd) -> None:
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(str(c) for c in cmd) + "\n" + r.stdout + r.stderr)


def build(kind: str = "real") -> Path:
    """kind: "real" links librbk_knn.so (must exist), "shim" links the oracle-backed CPU stand-in."""
    OUT.mkdir(exist_ok=True)
    hdrs = [HERE / "node_api.h", HERE / "mock_napi.h", ROOT / "include" / "rbk_knn.h"]
    objs = []
    for src, extra in ((ROOT / "napi" / "rbk_napi.cc", ["-DNODE_GYP_MODULE_NAME=rbk_knn", "-I", ROOT / "include"]),
                       (HERE / "mock_napi.cc", []), (HERE / "harness.cc", [])):
        obj = OUT / (src.stem + ".o")
        if _stale(obj, [src] + hdrs):
            _run(CXX + ["-I", HERE] + extra + ["-c", src, "-o", obj])
        objs.append(obj)
    exe = OUT / f"harness_{kind}"
    if kind == "real":
        lib = ROOT / "runbookai_b200" / "lib" / "librbk_knn.so"
        if _stale(exe, objs + [lib]):
            _run(["g++"] + objs + ["-o", exe, "-L", lib.parent, "-lrbk_knn", f"-Wl,-rpath,{lib.parent}", "-lpthread"])
    else:
        sys.path.insert(0, str(ROOT))
        import oracle
        oracle.build()
        olib = ROOT / "oracle" / "librbk_oracle.so"
        shim = OUT / "librbk_knn_shim.so"
        shim_src = ROOT / "tests" / "napi_shim" / "rbk_shim.cc"
        if _stale(shim, [shim_src, olib] + hdrs):
            _run(CXX + ["-fPIC", "-shared", shim_src, "-o", shim, "-L", olib.parent, "-l:librbk_oracle.so",
                        f"-Wl,-rpath,{olib.parent}"])
        if _stale(exe, objs + [shim]):
            _run(["g++"] + objs + ["-o", exe, "-L", OUT, "-l:librbk_knn_shim.so", f"-Wl,-rpath,{OUT}",
                                   f"-Wl,-rpath,{olib.parent}", "-L", olib.parent, "-l:librbk_oracle.so", "-lpthread"])
    return exe


if __name__ == "__main__":
    for k in ("shim", "real"):
        print(build(k))
