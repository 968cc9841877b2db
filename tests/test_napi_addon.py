"""napi/rbk_napi.cc compiled, linked and RUN - against a mock of Node's N-API (napi/mock/), because the image has no
Node.  The harness (napi/mock/harness.cc) plays ts/gpu-embedding-index.ts: construct, appendBlobs / appendF64,
overwriteF64(Batch), tombstone, count, `await search(...)` through the promise + async-work path, every error path,
clear, finalizer.  What comes back is compared with the oracle, bit for bit.

* CPU: the addon links against the real librbk_knn.so and its constructor THROWS without a GPU (no fallback); the
  whole scenario runs against an oracle-backed stand-in of the C ABI (tests/napi_shim) so the addon's own argument
  handling and result marshalling are checked where there is no GPU.
* `-m gpu`: the same scenario against the real library, one device and a device list (rbk_group behind the handle).
"""
import importlib.util
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import HAS_CUDA, ROOT


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _build_real():
    return _load("rbk_napi_mock_build", ROOT / "napi" / "mock" / "build.py").build()


def _build_shim():
    return _load("rbk_napi_shim_build", ROOT / "tests" / "napi_shim" / "build.py").build()


def _write_inputs(d: Path, devices, dim=96, n=3000, nq=9, k=24, min_score=0.05, seed=11):
    rng = np.random.default_rng(seed)
    rows = rng.standard_normal((n, dim))                       # arbitrary doubles: KEEP_F64 makes them the exact source
    rows[n // 2 + 7] = rows[5]                                 # exact ties across the two load paths
    rows[n - 3] = rows[5]
    rows[17] = 0.0                                             # a zero vector: cosine NaN, never returned
    q = rng.standard_normal((nq, dim))
    q[0] = rows[5] * 3.0                                       # cosine 1 with three rows -> slot order decides
    over_slots = rng.choice(n, 40, replace=False).astype(np.int64)
    over_slots = over_slots[~np.isin(over_slots, [5, n // 2 + 7, n - 3])]
    over_rows = rng.standard_normal((len(over_slots), dim))
    over_rows[1] = q[1] * 0.5                                  # planted hits in overwritten slots
    over_rows[0] = q[2] * 2.0
    dead = np.setdiff1d(rng.choice(n, 200, replace=False), over_slots).astype(np.int64)
    dead = dead[dead != 5]
    (d / "meta.txt").write_text(f"{dim} {n} {nq} {k} {min_score} {len(devices)} " + " ".join(map(str, devices)) + "\n")
    rows.astype("<f8").tofile(d / "rows.f64")
    q.astype("<f8").tofile(d / "queries.f64")
    over_slots.tofile(d / "over_slots.i64")
    over_rows.astype("<f8").tofile(d / "over_rows.f64")
    dead.tofile(d / "dead.i64")
    corpus = rows.copy()
    corpus[over_slots] = over_rows
    live = np.ones(n, dtype=np.uint8)
    live[dead] = 0
    return dict(corpus=corpus, live=live, q=q, k=k, min_score=min_score, n=n, nq=nq, dead=dead)


def _check_outputs(d: Path, w, oracle_mod):
    log = dict(line.split(" ", 1) for line in (d / "log.txt").read_text().strip().splitlines())
    n, nq, k = w["n"], w["nq"], w["k"]
    assert float(log["appendBlobs_first"]) == 0 and float(log["appendF64_first"]) == n // 2
    assert float(log["count"]) == n - len(w["dead"])
    slots = np.fromfile(d / "slots.i64", dtype=np.int64).reshape(nq, k)
    scores = np.fromfile(d / "scores.f64", dtype=np.float64).reshape(nq, k)
    counts = np.fromfile(d / "counts.i32", dtype=np.int32)
    for b in range(nq):
        es, ev = oracle_mod.search(w["corpus"], w["q"][b], k, w["min_score"], live=w["live"])
        assert counts[b] == len(es) and (slots[b, :len(es)] == es).all()
        assert scores[b, :len(es)].tobytes() == ev.tobytes()          # fp64 scores bit-identical
    assert list(slots[0, :3]) == [5, n // 2 + 7, n - 3]                  # equal scores: Map (slot) order
    assert 17 not in slots[:, :]                                          # the zero vector never matches
    # error paths: the reference's wording, as exceptions / a rejected promise
    assert log["err_append"] == "Vectors must have the same length"
    assert log["err_search"] == "Vectors must have the same length"
    assert log["err_tombstone"] == "slots must be a BigInt64Array"
    assert "tombstoned" in log["err_overwrite_dead"]
    assert log["err_overwrite_len"] == "Vectors must have the same length"
    assert log["repeat_identical"] == "1"
    assert float(log["count_after_clear"]) == 0 and float(log["hits_after_clear"]) == 0 and log["finalized"] == "1"


def test_addon_links_against_the_library_and_its_constructor_throws_without_a_gpu(tmp_path, native):
    exe = _build_real()          # compiles with -Wall -Wextra -Werror; every C-ABI symbol it binds resolves
    _write_inputs(tmp_path, devices=[])
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=120)
    if HAS_CUDA:
        assert r.returncode == 0, r.stderr
        return
    assert r.returncode == 3, r.stderr       # `new RbkIndex(...)` threw: no device, no CPU path
    assert "no CUDA device" in (tmp_path / "error.txt").read_text()


@pytest.mark.parametrize("devices", [[], [0]], ids=["index", "group"])
def test_addon_scenario_against_the_oracle_backed_stand_in(tmp_path, oracle_mod, devices):
    exe = _build_shim()
    w = _write_inputs(tmp_path, devices)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + ((tmp_path / "error.txt").read_text() if (tmp_path / "error.txt").exists() else "")
    _check_outputs(tmp_path, w, oracle_mod)


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [[], [0]], ids=["index", "group"])
def test_addon_scenario_on_the_gpu_matches_the_oracle(tmp_path, oracle_mod, native, devices):
    exe = _build_real()
    w = _write_inputs(tmp_path, devices, n=6000, dim=200, nq=13, k=32)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + ((tmp_path / "error.txt").read_text() if (tmp_path / "error.txt").exists() else "")
    _check_outputs(tmp_path, w, oracle_mod)
