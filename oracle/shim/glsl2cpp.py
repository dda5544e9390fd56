#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- turns one of the reference's GLSL programs into a C++ struct, on stdout.

    glsl2cpp.py <shader dir> <root shader file> <StructName> [-DNAME ...] [--set NAME=VALUE ...]

--set rewrites the value of a `#define NAME <value>` line of the shader text itself: the edit a maintainer of the reference makes
to pick a compile-time variant the text already contains (voxelize.gs:15-19: `#define THICKNESS THIN` / FAT).

The shader TEXT is read from the reference tree where it lies (never copied into the repo). Steps:
  1. the preprocessor defines the reference passes as source string 0 (renderer.cpp:227, servicePicking.cpp:21),
  2. textual `#include <file>` splicing, exactly as src/shaders/shader.cpp:54-94 does (no include guards),
  3. SYNTACTIC rewrites only, so that the text compiles as the body of a C++ struct against oracle/shim/glsl_emu.h
     (every rewrite is listed below; none changes an expression, a constant or the order of evaluation):
       - `#version` / `#extension` lines, `layout(...)` qualifiers, `uniform`, `writeonly`, `precision` dropped
       - `layout(std430, binding=N) buffer X { ... } Y;`  ->  `struct X { ... } Y;`      (SSBOs: members of the struct)
       - `in block { ... } In[];` / `out block { ... } Out;`                -> `struct block_in { ... } In[3];` / `struct block_out {...} Out;`
       - `out gl_PerVertex { ... };` dropped (gl_Position is a member supplied by the harness)
       - global `in T x;` / `out T x;`                                      -> `T x;`
       - parameters `in T x` -> `T x`, `out T x` / `inout T x` -> `T& x`
       - function prototypes (forward declarations) dropped: C++ class scope does not need them
       - floating literals get an `f` suffix (GLSL literals are binary32; C++ ones would be binary64)
       - `float(`/`int(` constructor casts -> `to_float(`/`to_int(` (saturating, NaN -> 0 conversion of glsl_emu.h)
       - `T name[] = { ... }` -> `T name[N] = { ... }` (C++ needs the array bound in a class)
       - the driver-tolerated `int x = texelFetch(isampler1D...)` of pathTracer.fs:85 needs no rewrite (glsl_emu.h)
"""
import os
import re
import sys


def splice_includes(code, base):
    # shader.cpp:54-94: repeatedly replace the first #include with the file's text
    while True:
        m = re.search(r"#include[^\n]*", code)
        if not m:
            return code
        inc = m.group(0)
        f = re.search(r"<([^>]+)>", inc)
        if not f:
            raise SystemExit("cannot parse " + inc)
        with open(os.path.join(base, f.group(1))) as fh:
            text = fh.read()
        code = code[:m.start()] + text + code[m.end():]


def strip_comments(code):
    code = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), code, flags=re.S)
    return re.sub(r"//[^\n]*", "", code)


def drop_prototypes(code):
    out, depth = [], 0
    for line in code.split("\n"):
        if depth == 0 and re.match(r"^\s*[A-Za-z_]\w*\s+[A-Za-z_]\w*\s*\([^;{}]*\)\s*;\s*$", line):
            out.append("")
        else:
            out.append(line)
        depth += line.count("{") - line.count("}")
    return "\n".join(out)


def rewrite(code):
    code = strip_comments(code)
    code = re.sub(r"^\s*#\s*(version|extension)[^\n]*", "", code, flags=re.M)
    # interface blocks
    code = re.sub(r"layout\s*\(\s*std430[^)]*\)\s*buffer\s+(\w+)", r"struct \1", code)
    code = re.sub(r"\bout\s+gl_PerVertex\s*\{[^}]*\}\s*;", "", code)
    code = re.sub(r"\bin\s+block\s*(\{[^}]*\})\s*(\w+)\s*\[\s*\]\s*;", r"struct block_in \1 \2[3];", code)
    code = re.sub(r"\bout\s+block\s*(\{[^}]*\})\s*(\w+)\s*;", r"struct block_out \1 \2;", code)
    code = re.sub(r"^\s*layout\s*\([^)]*\)\s*(in|out)\s*;", "", code, flags=re.M)
    code = re.sub(r"layout\s*\([^)]*\)", "", code)
    code = re.sub(r"\b(uniform|writeonly|readonly|coherent)\b", "", code)
    # global in/out declarations, then parameter qualifiers
    code = re.sub(r"^(\s*)(in|out)\s+(\w+\s+\w+\s*;)", r"\1\3", code, flags=re.M)
    code = re.sub(r"\b(?:out|inout)\s+(\w+)\s+(\w+)", r"\1& \2", code)
    code = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", code)
    code = drop_prototypes(code)
    # literals: digits '.' digits [exp] | '.' digits [exp] | digits exp, not already suffixed, not part of an identifier
    code = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)[fF]?(?![\w.])", r"\1f", code)
    code = re.sub(r"\bfloat\s*\(", "to_float(", code)
    code = re.sub(r"\bint\s*\(", "to_int(", code)
    # unsized arrays with an initialiser list: count the top-level elements
    def size_array(m):
        body, depth, n = m.group(3), 0, 1
        for ch in body:
            depth += ch in "({"
            depth -= ch in ")}"
            n += (ch == "," and depth == 0)
        return "%s %s[%d] = {%s}" % (m.group(1), m.group(2), n, body)
    code = re.sub(r"(\w+)\s+(\w+)\s*\[\s*\]\s*=\s*\{((?:[^{}]|\{[^{}]*\})*)\}", size_array, code, flags=re.S)
    return code


def main():
    base, root, name = sys.argv[1], sys.argv[2], sys.argv[3]
    defines = [a[2:] for a in sys.argv[4:] if a.startswith("-D")]
    sets = [sys.argv[i + 1].split("=", 1) for i in range(4, len(sys.argv) - 1) if sys.argv[i] == "--set"]
    with open(os.path.join(base, root)) as fh:
        code = fh.read()
    for nm, val in sets:
        code, n = re.subn(r"^(\s*#\s*define\s+%s)\s+\w+[^\n]*" % re.escape(nm), r"\1 %s" % val, code, flags=re.M)
        if n != 1:
            raise SystemExit("--set %s: expected exactly one #define line, found %d" % (nm, n))
    code = "".join("#define %s\n" % d for d in defines) + code
    code = splice_includes(code, base if base.endswith("/") else base + "/")
    body = rewrite(code)
    undefs = "".join("#undef %s\n" % d for d in defines)
    sys.stdout.write("// generated at build time from %s (reference tree); not stored in the repository\n" % root)
    sys.stdout.write("struct %s : ShaderBase {\n%s\n};\n%s" % (name, body, undefs))
    # macros defined by the shader text itself must not leak into the next program
    for m in sorted(set(re.findall(r"^\s*#\s*define\s+(\w+)", body, flags=re.M))):
        if m not in defines:
            sys.stdout.write("#undef %s\n" % m)


if __name__ == "__main__":
    main()
