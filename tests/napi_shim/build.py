"""TEST INFRASTRUCTURE: link the addon harness (napi/mock/) against rbk_shim.cc, the oracle-backed CPU stand-in of
the C ABI, so the whole addon scenario can run and be checked where there is no GPU."""
from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
HERE = Path(__file__).resolve().parent


def _mock_build():
    spec = importlib.util.spec_from_file_location("rbk_napi_mock_build", ROOT / "napi" / "mock" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build() -> Path:
    mb = _mock_build()
    objs = mb.build_objects()
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    import oracle
    oracle.build()
    olib = ROOT / "oracle" / "librbk_oracle.so"
    out = HERE / "_build"
    out.mkdir(exist_ok=True)
    shim = out / "librbk_knn_shim.so"
    if mb.stale(shim, [HERE / "rbk_shim.cc", olib] + mb.HEADERS):
        mb.run(mb.CXX + ["-fPIC", "-shared", HERE / "rbk_shim.cc", "-o", shim, "-L", olib.parent, "-l:librbk_oracle.so",
                         f"-Wl,-rpath,{olib.parent}"])
    exe = out / "harness_shim"
    if mb.stale(exe, objs + [shim]):
        mb.run(["g++"] + objs + ["-o", exe, "-L", out, "-l:librbk_knn_shim.so", f"-Wl,-rpath,{out}",
                                 f"-Wl,-rpath,{olib.parent}", "-L", olib.parent, "-l:librbk_oracle.so", "-lpthread"])
    return exe
