#!/usr/bin/env python
"""Build tests/host_emul/_build/libb200rk_emul.so: the product's own host sources (runtime / launch / executor / driver /
capi / quadrature .cu) and kernels compiled by g++ for the CPU. The only source transformation is the launch syntax:

    kernel<targs><<<grid, kThreads, 0, c->stream>>>(args);   ->   { evaluate args once; run `kernel<targs>(args)` once per
                                                                   emulated thread of a (grid x kThreads) launch }

(threaded, one host thread per CUDA thread, for the two kernels whose threads interact). CUDA runtime calls resolve to
tests/host_emul/fake_cuda/cuda_runtime.h ("device" memory = host memory, everything synchronous); jit.cu is replaced by
stand-ins (emul_lib_support.cpp). TEST INFRASTRUCTURE ONLY: it exists so that the CPU test-suite can run the GPU parity
tests' host logic + kernels against the oracle without a GPU (`pytest --host-emulation`); the product package never loads it."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "numericalnim_b200", "csrc")
BUILD = os.path.join(HERE, "_build", "emul_lib")
OUT = os.path.join(HERE, "_build", "libb200rk_emul.so")
SOURCES = ["runtime.cu", "launch.cu", "executor.cu", "driver.cu", "capi.cu", "quadrature.cu"]
THREADED = ("stage_l96_kernel", "l96_attempt_kernel", "l96_warp_attempt_kernel", "l96_rk4_kernel")   # shared-memory tile + __syncthreads (the cooperative loop goes through cudaLaunchCooperativeKernel)

LAUNCH = re.compile(r"(?P<kernel>\b\w+<[^;<>]*(?:<[^;<>]*>[^;<>]*)*>)<<<(?P<grid>.+?), (?P<threads>kThreads(?: / 2)?), 0, c->stream>>>\((?P<args>.*)\);(?P<tail>\s*(//.*)?)$")


def transform(text: str, name: str) -> str:
    out, n = [], 0
    for line in text.splitlines():
        m = LAUNCH.search(line)
        if m:
            kernel, grid, args = m.group("kernel"), m.group("grid"), m.group("args")
            fn = "emul_launch_threaded" if kernel.split("<")[0] in THREADED else "emul_launch_serial"
            body = ("{ auto emul_args_ = std::make_tuple(%s); %s((%s), %s, [&] { std::apply([](auto&... x_) { %s(x_...); }, emul_args_); }); }%s"
                    % (args, fn, grid, m.group("threads"), kernel, m.group("tail")))
            line = line[: m.start()] + body
            n += 1
        line = line.replace("cudaLaunchCooperativeKernel((void*)kernel,", "cudaLaunchCooperativeKernel(kernel,")
        out.append(line)
    if "<<<" in "\n".join(out):
        raise SystemExit(f"{name}: a kernel launch was not transformed")
    return "\n".join(out) + "\n", n


def build(force: bool = False, sanitize: str = "") -> str:
    """sanitize = "address": an AddressSanitizer build (libb200rk_emul_asan.so) — since "device" memory is host memory,
    an out-of-bounds access of a kernel or of the host code is reported like compute-sanitizer memcheck would."""
    global BUILD, OUT
    if sanitize:
        BUILD = os.path.join(HERE, "_build", "emul_lib_" + sanitize)
        OUT = os.path.join(HERE, "_build", "libb200rk_emul_%s.so" % {"address": "asan"}.get(sanitize, sanitize))
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".hpp", ".cuh", ".h"))] + \
        [os.path.join(HERE, f) for f in ("emul_lib_prelude.hpp", "emul_lib_support.cpp", "build_emul_lib.py")] + \
        [os.path.join(HERE, "fake_cuda", f) for f in os.listdir(os.path.join(HERE, "fake_cuda"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(BUILD, exist_ok=True)
    total = 0
    cpps = []
    for s in srcs:
        text, n = transform(open(s).read(), os.path.basename(s))
        total += n
        dst = os.path.join(BUILD, os.path.basename(s).replace(".cu", "_emul.cpp"))
        with open(dst, "w") as fh:
            fh.write(text)
        cpps.append(dst)
    cpps.append(os.path.join(HERE, "emul_lib_support.cpp"))
    flags = ["-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-pthread", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable",
             "-Wno-unused-but-set-variable", "-include", os.path.join(HERE, "emul_lib_prelude.hpp"),
             f"-I{os.path.join(HERE, 'fake_cuda')}", f"-I{os.path.join(HERE, 'cuda_stubs')}", f"-I{CSRC}", f"-I{HERE}"]
    if sanitize:
        flags += ["-fsanitize=" + sanitize, "-fno-omit-frame-pointer", "-g"]
    objs = []
    procs = []
    for c in cpps:
        o = os.path.join(BUILD, os.path.basename(c).replace(".cpp", ".o"))
        objs.append(o)
        procs.append((c, subprocess.Popen(["g++", *flags, "-c", c, "-o", o], stderr=subprocess.PIPE, text=True)))
    for c, p in procs:
        err = p.communicate()[1]
        if p.returncode != 0:
            raise SystemExit(f"compiling {c} failed:\n{err[-4000:]}")
    subprocess.run(["g++", "-shared", "-pthread", *(["-fsanitize=" + sanitize] if sanitize else []), "-o", OUT, *objs, "-ldl"], check=True)
    print(f"{OUT}: {total} kernel launches rewritten", file=sys.stderr)
    return OUT


def build_fake_nccl() -> str:
    """tests/host_emul/_build/libnccl_emul.so: in-process stand-in for NCCL (fake_nccl.cpp) — ranks are threads."""
    src, out = os.path.join(HERE, "fake_nccl.cpp"), os.path.join(HERE, "_build", "libnccl_emul.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-Wall", src, "-o", out], check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, sanitize="address" if "--asan" in sys.argv else ""))
