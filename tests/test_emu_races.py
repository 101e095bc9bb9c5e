"""Data-race check of the kernels' synchronisation (no GPU): tests/emu/race_check.cpp runs the score and
gradient functions of every model (kge_models.cuh, kge_grads.cuh), the projection-tail kernels
(kge_proj.cuh), the ConvE trunk (kge_conve.cuh) and the TransH / TransD projection (kge_project.cuh)
under the host emulation — one host thread per CUDA thread, __syncthreads / __syncwarp / shuffles as the
only happens-before edges — in a ThreadSanitizer build.  A missing barrier between a shared-memory
write and another thread's read (which lock-step execution on real hardware can hide) fails here.
A control build with the block barrier compiled out must be flagged, so a silent pass means something."""
import os
import struct
import subprocess

import numpy as np
import pytest

import golden_util as gu

import emu_build

TSAN_ENV = dict(os.environ, TSAN_OPTIONS="exitcode=66 halt_on_error=1")
ALL_HEADERS = tuple(sorted(set(emu_build.MODEL_HEADERS + emu_build.PROJ_HEADERS)))


def _build(name, extra):
    return emu_build.build("race_check.cpp", name, ALL_HEADERS, extra=["-g", "-fsanitize=thread"] + extra, shared=False)


@pytest.fixture(scope="module")
def checker():
    exe = _build("race_check", [])
    probe = subprocess.run([exe], env=TSAN_ENV, capture_output=True, text=True)
    if probe.returncode != 64:     # usage exit code: the sanitizer runtime itself starts up here
        pytest.skip("ThreadSanitizer runtime unavailable in this environment: %s" % probe.stderr[-300:])
    return exe


def _run(exe, *args):
    res = subprocess.run([exe] + list(args), env=TSAN_ENV, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, "race_check %s -> rc %d\n%s" % (" ".join(args), res.returncode, res.stderr[-3000:])


@pytest.mark.parametrize("what", ["proj", "conve", "project"])
def test_projection_kernels_are_race_free(checker, what):
    _run(checker, what)


# models whose score / gradient functions keep per-group scratch in shared memory or synchronise
# lanes (__syncwarp / group_sync) — where a missing barrier could hide — plus three plain ones
RACE_MODELS = {"transr", "hole", "rescal", "slm", "ntn", "sme", "sme_bl", "kg2e", "convkb", "quate", "octonione",
               "transe", "complex", "rotate"}


def _one_case_per_model():
    seen, out = set(), []
    for name in gu.case_names():
        model = name.split("_d")[0].split("_l")[0].rstrip("_0123456789x")
        if name.startswith("pretrained") or model in seen or model not in RACE_MODELS:
            continue
        seen.add(model)
        out.append(name)
    return out


@pytest.mark.parametrize("name", _one_case_per_model())
def test_score_and_gradient_functions_are_race_free(checker, name, tmp_path):
    import oracle
    g = gu.load(name)
    om = gu.oracle_model(g)
    n = min(40, len(g["h"]))
    blob = struct.pack("<5i2f2q", oracle.MODEL_IDS[om.name], om.dim, om.rel_dim, int(om.l1_flag), len(om.tables),
                       om.margin, om.phase_scale, om.num_ent, om.num_rel)
    for t in om.tables:
        a = np.ascontiguousarray(t, dtype=np.float32)
        blob += struct.pack("<q", a.size) + a.tobytes()
    blob += struct.pack("<q", n)
    for k in ("h", "r", "t"):
        blob += np.ascontiguousarray(g[k][:n], dtype=np.int64).tobytes()
    blob += np.ascontiguousarray(g["upstream"][:n], dtype=np.float32).tobytes()
    path = tmp_path / "case.bin"
    path.write_bytes(blob)
    _run(checker, "score", str(path))


def test_detector_flags_a_missing_barrier(checker):
    """positive control: the same kernels with __syncthreads() compiled out must be reported"""
    exe = _build("race_check_nobar", ["-DCUDA_EMU_NO_BARRIERS", "-DRACE_CHECK_MINIMAL"])
    res = subprocess.run([exe, "conve"], env=TSAN_ENV, capture_output=True, text=True, timeout=900)
    assert res.returncode == 66 and "data race" in res.stderr
