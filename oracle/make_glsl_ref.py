#!/usr/bin/env python
"""Rewrite the reference's shader sources into C++-compilable text under oracle/_ref/ (TEST INFRASTRUCTURE — see
oracle/ref_shim/glsl_compat.h).  oracle/Makefile deletes the rewritten text again right after compiling it: nothing of
the shader text is stored in the repository or left in the tree.
usage: python oracle/make_glsl_ref.py /root/reference oracle/_ref
writes rt_glsl_gen.inc (rtcommon.glsl + restir.glsl), rt_rgen_gen.inc (rt.rgen without its binding declarations)
and tonemap_gen.inc (tonemap.frag)"""
import re
import sys

ref, out_dir = sys.argv[1], sys.argv[2]


def sequence_constructor_args(src):
    """GLSL evaluates call arguments left to right; C++ leaves the order open (g++: right to left), which would swap
    the two random numbers of `vec2(randf(seed), randf(seed))`.  Brace initialisation is sequenced left to right."""
    res, i = [], 0
    for m in re.finditer(r"\bvec[234]\(", src):
        if m.start() < i:
            continue
        depth, j = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        args = src[m.end():j - 1]
        if args.count("randf(") >= 2:
            res.append(src[i:m.end() - 1] + "{" + args + "}")
            i = j
    res.append(src[i:])
    return "".join(res)


def rewrite(text):
    text = re.sub(r"^#(extension|include|version).*$", "", text, flags=re.M)
    text = text.replace("layout(push_constant) uniform Constants", "struct Constants")
    # binding declarations: blocks `layout(...) ... { ... } name;` and one-liners `layout(...) ... name;`
    # (their storage is declared by oracle/ref_shim/glsl_ref.cpp)
    text = re.sub(r"layout\s*\([^)]*\)[^;{]*\{[^}]*\}[^;]*;", "", text, flags=re.S)
    text = re.sub(r"layout\s*\([^)]*\)[^;]*;", "", text)
    text = re.sub(r"\binout\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"\bout\s+(\w+)\s+(\w+)", r"\1& \2", text)
    # fp32 literals: GLSL decimal literals are float, C++ ones double
    text = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])", r"\1f", text)
    text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void rgen_main()", text)
    return sequence_constructor_args(text)


head = "/* generated from the reference shaders by oracle/make_glsl_ref.py — do not commit */\n"
common = "".join(open(f"{ref}/src/shaders/rt/{n}").read() + "\n" for n in ("rtcommon.glsl", "restir.glsl"))
open(f"{out_dir}/rt_glsl_gen.inc", "w").write(head + rewrite(common))
open(f"{out_dir}/rt_rgen_gen.inc", "w").write(head + rewrite(open(f"{ref}/src/shaders/rt/rt.rgen").read()))
open(f"{out_dir}/tonemap_gen.inc", "w").write(head + rewrite(open(f"{ref}/src/shaders/tonemap.frag").read()))
