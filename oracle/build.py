"""Build recipe for the oracle (TEST INFRASTRUCTURE ONLY — see oracle/pn2_oracle.c header).

Two artefacts:

1. ``oracle/liboracle.so`` — the C restatement in ``oracle/pn2_oracle.c`` (always built).

2. ``oracle/_ref/*.so`` — the reference's OWN code, compiled from the sources where they lie
   under ``/root/reference`` (never copied into this repo; outputs only into the git-ignored
   ``oracle/_ref/``).  Only possible where ``/root/reference`` exists (the build container); the
   GPU box uses the prebuilt files that travel with the snapshot.

   * ``libref_sampling.so``  <- tf_ops/sampling/tf_sampling_g.cu   (nvcc -O2, sm_100a, unmodified)
   * ``libref_grouping.so``  <- tf_ops/grouping/tf_grouping_g.cu   (nvcc -O2, sm_100a, unmodified)
   * ``libref_cpu.so``       <- the TF-free CPU functions of
         tf_ops/3d_interpolation/tf_interpolate.cpp (threenn_cpu .. threeinterpolate_grad_cpu,
         the span between the "Find three nearest" comment and ``class ThreeNNOp``),
         tf_ops/grouping/test/query_ball_point.cpp (everything above ``int main``) and
         tf_ops/grouping/test/selection_sort.cpp (selection_sort_cpu only),
     streamed through a pipe into ``g++ -O2 -x c++ -`` exactly as the reference's compile
     scripts build them (no -march, no fast-math, no OpenMP): no source text is written to disk.
     The rest of those files (TensorFlow OpKernel shims, main()) cannot be built here: there is
     no TensorFlow in this image.

The reference's own build system (tf_*_compile.sh) is not run: it needs the TensorFlow include
tree and CUDA 8.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PN2_REFERENCE_ROOT", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")


def _newer(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.exists(s) and os.path.getmtime(s) <= t for s in sources)


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return False


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "pn2_oracle.c")
    out = os.path.join(HERE, "liboracle.so")
    if not force and _newer(out, src, __file__):
        return out
    # -ffp-contract=off: the only fused multiply-adds are the explicit fmaf() calls.
    # -mfma (when the build host has it) turns fmaf() into one instruction instead of a libm call.
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math"]
    if _cpu_has_fma() and os.environ.get("PN2_ORACLE_NO_MFMA") is None:
        cmd.append("-mfma")
    cmd += ["-o", out, src, "-lm"]
    subprocess.run(cmd, check=True)
    return out


def _extract(path: str, start_pat: str | None, stop_pat: str) -> str:
    """Return the text of ``path`` from the first line containing start_pat (or the top) up to,
    not including, the first later line that starts with stop_pat."""
    out, on = [], start_pat is None
    with open(path) as f:
        for line in f:
            if not on and start_pat is not None and start_pat in line:
                on = True
            if on and line.startswith(stop_pat):
                break
            if on:
                out.append(line)
    if not out:
        raise RuntimeError(f"could not locate {start_pat!r} .. {stop_pat!r} in {path}")
    return "".join(out)


def build_ref(force: bool = False) -> dict:
    """Compile the reference's own kernels/functions into oracle/_ref/. Returns {name: path}."""
    res = {}
    if not os.path.isdir(REF_ROOT):
        # GPU box: use whatever travelled with the snapshot.
        for name in ("libref_sampling.so", "libref_grouping.so", "libref_cpu.so"):
            p = os.path.join(REF_OUT, name)
            if os.path.exists(p):
                res[name] = p
        return res
    os.makedirs(REF_OUT, exist_ok=True)
    cu = {
        "libref_sampling.so": os.path.join(REF_ROOT, "tf_ops/sampling/tf_sampling_g.cu"),
        "libref_grouping.so": os.path.join(REF_ROOT, "tf_ops/grouping/tf_grouping_g.cu"),
    }
    for name, src in cu.items():
        out = os.path.join(REF_OUT, name)
        if force or not _newer(out, src, __file__):
            # the reference's own flag set (tf_sampling_compile.sh:2: nvcc ... -O2 -D GOOGLE_CUDA=1
            # -x cu -Xcompiler -fPIC) plus the arch this box needs.
            subprocess.run([NVCC, "-O2", "-DGOOGLE_CUDA=1", "-x", "cu", "-Xcompiler", "-fPIC",
                            "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, src],
                           check=True)
        res[name] = out

    interp = os.path.join(REF_ROOT, "tf_ops/3d_interpolation/tf_interpolate.cpp")
    qbp = os.path.join(REF_ROOT, "tf_ops/grouping/test/query_ball_point.cpp")
    ssort = os.path.join(REF_ROOT, "tf_ops/grouping/test/selection_sort.cpp")
    out = os.path.join(REF_OUT, "libref_cpu.so")
    if force or not _newer(out, interp, qbp, ssort, __file__):
        text = "#include <cstdio>\n#include <cstring>\n#include <cstdlib>\n#include <cmath>\n"
        text += "namespace ref_qbp {\n" + _extract(qbp, None, "int main") + "\n}\n"
        text += "namespace ref_interp {\n" + _extract(interp, "// Find three nearest", "class ThreeNNOp") + "\n}\n"
        text += "namespace ref_ssort {\n" + _extract(ssort, "// input: k (1)", "int main") + "\n}\n"
        # system headers must not be re-included inside a namespace: pull them in first (done
        # above; their include guards make the in-namespace #include lines no-ops).
        text = "#include <ctime>\n#include <string>\n#include <vector>\nusing namespace std;\n" + text
        subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-x", "c++", "-", "-o", out],
                       input=text.encode(), check=True)
    res["libref_cpu.so"] = out
    return res


def main() -> None:
    force = "--force" in sys.argv
    print("oracle:", build_oracle(force))
    for k, v in build_ref(force).items():
        print("ref:", k, "->", v)


if __name__ == "__main__":
    main()
