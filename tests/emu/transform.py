"""TEST INFRASTRUCTURE ONLY: rewrites a copy of sprs_b200/csrc/*.cu into plain C++ that
compiles against tests/emu/cuemu.h (see that header).  The product sources are not touched.

  * `#include <cuda_runtime.h>`            -> `#include "cuemu.h"`
  * `kernel<<<grid, block, smem, s>>>(a)`  -> `cuemu::launch(cuemu::cfg(grid, block, smem, s), [&] { kernel(a); })`
  * `extern __shared__ [__align__(N)] T v[];` -> `T* v = (T*)cuemu::dyn_smem();`
  * `__shared__ __align__(N) T v[..];`     -> `alignas(N) static T v[..];`
  * csrc/ptx.cuh (the inline-PTX wrappers)  -> `#include "cuemu_ptx.h"`
  * `#ifdef __CUDACC__`                    -> `#ifdef CUEMU`
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "..", "..", "sprs_b200", "csrc")
FILES = ["api.cu", "spmv.cu", "spmm.cu", "spgemm.cu", "transpose.cu", "gen.cu", "peer.cu",
         "solver.cu", "csvec.cu", "common.cuh", "scan.cuh", "ptx.cuh"]
# diag.cu (event-timed measurement aid) and comm.cu (process rendezvous, CUDA IPC / VMM, device barrier) is not emulated: the multi-rank
# path is covered on hardware by tests/cpp/test_comm_ranks.cpp and tests/test_gpu_comm.py


def match_back(s, end):
    """s[end-1] == '>': index of the matching '<' scanning backwards."""
    depth = 0
    i = end - 1
    while i >= 0:
        if s[i] == ">":
            depth += 1
        elif s[i] == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments")


def match_paren(s, start):
    """s[start] == '(': index of the matching ')'."""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced parentheses")


def rewrite_launches(s):
    out = []
    pos = 0
    while True:
        k = s.find("<<<", pos)
        if k < 0:
            out.append(s[pos:])
            return "".join(out)
        # callee: identifier (with ::), optionally followed by <template args>
        j = k
        if s[j - 1] == ">":
            j = match_back(s, j)
        while j > 0 and (s[j - 1].isalnum() or s[j - 1] in "_:"):
            j -= 1
        callee = s[j:k]
        e = s.index(">>>", k)
        cfg = s[k + 3:e]
        a0 = e + 3
        while s[a0] in " \t\n\\":
            a0 += 1
        assert s[a0] == "(", "launch without an argument list near: " + s[k - 40:k + 40]
        a1 = match_paren(s, a0)
        args = s[a0 + 1:a1]
        out.append(s[pos:j])
        out.append("cuemu::launch(cuemu::cfg(%s), [&] { %s(%s); })" % (cfg, callee, args))
        pos = a1 + 1


def transform(name, s):
    s = s.replace("#include <cuda_runtime.h>", '#include "cuemu.h"')
    s = s.replace("__CUDACC__", "CUEMU")
    if name == "ptx.cuh":  # the inline-PTX wrappers: replaced wholesale by their stand-ins
        return '#pragma once\n#include "common.cuh"\n#include "cuemu_ptx.h"\n'
    s = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];",
               lambda m: "%s* %s = (%s*)cuemu::dyn_smem();" % (m.group(1), m.group(2), m.group(1)),
               s)
    s = re.sub(r"__shared__\s+__align__\((\d+)\)", r"alignas(\1) static", s)
    assert "asm" not in re.sub(r"//.*", "", s).replace("__asm", ""), name + ": inline asm left"
    return rewrite_launches(s)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for f in FILES:
        src = open(os.path.join(SRC, f)).read()
        dst = transform(f, src)
        # .cu -> .cpp so that g++ picks the language by itself; headers keep their names
        open(os.path.join(out_dir, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w").write(dst)
    # the public header is included as ../../include/sprs_b200.h relative to csrc
    inc = os.path.join(out_dir, "..", "..", "include")
    os.makedirs(inc, exist_ok=True)
    src_h = os.path.join(HERE, "..", "..", "include", "sprs_b200.h")
    open(os.path.join(inc, "sprs_b200.h"), "w").write(open(src_h).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "build", "csrc", "x"))
