"""TEST INFRASTRUCTURE: builds the WHOLE library (csrc/*.cu, host code included) for the CPU behind tests/emu/ and loads it through
the product's own ctypes binding, so that the C ABI, the host-side sequencing of mtm_api.cu and every kernel can be exercised without
a GPU (tests/test_library_emulation.py).  The source text is the library's, with three mechanical rewrites:

  * ``kernel<<<grid, block, smem, stream>>>(args);``  ->  ``emu_launch_dyn(grid, block, smem, [&] { kernel(args); });``
  * ``extern __shared__ T name[];``                   ->  a pointer to the launch's dynamic shared memory
  * ncc_tc.cu: the PTX-wrapper section is left out (tests/emu/tcgen05_model.h defines the same functions), two
    ``fence.mbarrier_init`` lines and the ``prefetch.global.L1`` helper go.

The CUDA runtime API is the synchronous stand-in at the end of tests/emu/cuda_runtime.h (one emulated sm_100 device with
EMU_SM_COUNT SMs, device memory = host memory).  Nothing here is reachable from the product: the shared object is written to a
temporary directory and only ever loaded by tests that monkeypatch ``mtm_b200._native``.
"""
import ctypes
import os
import re
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multitemplatematching-python_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")


def _match_back(src, close, open_ch, close_ch):
    """Index of the bracket that matches the closing one at ``close`` (scanning backwards)."""
    depth = 0
    for k in range(close, -1, -1):
        if src[k] == close_ch:
            depth += 1
        elif src[k] == open_ch:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced brackets")


def _match_fwd(src, start, open_ch, close_ch):
    depth = 0
    for k in range(start, len(src)):
        if src[k] == open_ch:
            depth += 1
        elif src[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced brackets")


def _split_top(text):
    """Comma-separated pieces of ``text`` at bracket depth 0."""
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src):
    """``K<<<cfg>>>(args)`` -> ``emu_launch_dyn(grid, block, smem, [&] { K(args); })`` (K may carry template arguments)."""
    out, pos, n = "", 0, 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return out + src[pos:], n
        j = i
        if src[j - 1] == ">":                                   # template argument list of the kernel
            j = _match_back(src, j - 1, "<", ">")
        k = j
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        kernel = src[k:i]
        e = src.index(">>>", i)
        cfg = _split_top(src[i + 3:e])
        assert 2 <= len(cfg) <= 4, cfg
        p = e + 3
        while src[p].isspace():
            p += 1
        assert src[p] == "(", src[i - 40:p + 10]
        q = _match_fwd(src, p, "(", ")")
        smem = cfg[2] if len(cfg) > 2 else "0"
        out += src[pos:k] + "emu_launch_dyn(%s, %s, %s, [&] { %s%s; })" % (cfg[0], cfg[1], smem, kernel, src[p:q + 1])
        pos = q + 1
        n += 1


def host_source(name):
    """Text of csrc/<name> as the host build compiles it."""
    src = open(os.path.join(CSRC, name)).read()
    if name == "ncc_tc.cu":
        a, b = src.index("// ---------------------------------------------------------------- PTX wrappers"), src.index("// Normalise 16 consecutive")
        src = src[:a] + src[b:]
        for text, count in (('__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }', 1),
                            ('asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");', 2)):
            assert src.count(text) == count, text
            src = src.replace(text, "")
        src = src.replace('#include "mtm_internal.cuh"', '#include "mtm_internal.cuh"\n#include "tcgen05_model.h"', 1)
    src = re.sub(r"extern __shared__ (?:__align__\(\d+\) )?(\w+) (\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(emu_dyn_smem);", src)
    src = src.replace("#include <math_constants.h>", "")
    src, _ = rewrite_launches(src)
    assert "asm" not in re.sub(r"//[^\n]*", "", src) and "<<<" not in src, name
    return src


def build(out_dir, sm_count=4):
    """Compiles the library for the host into ``out_dir``; returns the path of the shared object."""
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("no g++")
    os.makedirs(out_dir, exist_ok=True)
    names = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    objs, jobs = [], []
    for name in names:
        cpp = os.path.join(out_dir, name[:-3] + "_host.cpp")
        with open(cpp, "w") as f:
            f.write(host_source(name))
        obj = cpp[:-4] + ".o"
        objs.append(obj)
        opt = ["-O3", "-march=native"] if name == "ncc_tc.cu" else ["-O1"]       # the tcgen05.mma model is the hot loop of the emulation
        jobs.append([gxx] + opt + ["-std=c++20", "-fPIC", "-w", "-DEMU_SM_COUNT=%d" % sm_count, "-I", EMU, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
                     "-c", cpp, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host build failed: %s\n%s" % (" ".join(cmd), r.stderr[-6000:]))

    with ThreadPoolExecutor(max_workers=8) as pool:
        list(pool.map(run, jobs))
    lib = os.path.join(out_dir, "libmtm_emu_TESTONLY.so")
    run([gxx, "-shared", "-o", lib] + objs)
    return lib
