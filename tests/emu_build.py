"""Build helpers for the host-emulation test libraries under tests/emu/ (g++; cached in tests/emu/_build)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU = os.path.join(HERE, "emu")
CSRC = os.path.join(ROOT, "pykg2vec_b200", "csrc")
BUILD = os.path.join(EMU, "_build")
FLAGS = ["-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-pthread", "-w"]


def _stale(out, deps):
    return not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps)


def build(src_name, out_name, headers, extra=(), shared=True):
    """compile tests/emu/<src_name> against the product's kernel headers -> tests/emu/_build/<out_name>"""
    src = os.path.join(EMU, src_name)
    out = os.path.join(BUILD, out_name)
    deps = [src, os.path.join(EMU, "cuda_runtime.h")] + [os.path.join(CSRC, h) for h in headers]
    if _stale(out, deps):
        os.makedirs(BUILD, exist_ok=True)
        cmd = ["g++"] + FLAGS + (["-fPIC", "-shared"] if shared else []) + list(extra) + \
              ["-I", EMU, "-I", CSRC, "-o", out + ".tmp", src]
        subprocess.run(cmd, check=True)
        os.replace(out + ".tmp", out)
    return out


MODEL_HEADERS = ("kge_common.cuh", "kge_models.cuh", "kge_grads.cuh", "kge_project.cuh")
PROJ_HEADERS = ("kge_common.cuh", "kge_proj.cuh", "kge_conve.cuh")


def models_lib():
    return build("emu_models.cpp", "libemu_models.so", MODEL_HEADERS)


def proj_lib():
    return build("emu_proj.cpp", "libemu_proj.so", PROJ_HEADERS)
