#!/usr/bin/env python3
"""oracle/ref_shim/glsl_prep.py — TEST INFRASTRUCTURE.

Makes the reference's pure-arithmetic GLSL include files compilable as C++ *where they lie*: reads a file under
/root/reference/shaders and writes a token-level transliteration of it to oracle/_ref/gen/ (git-ignored build output; no reference
source is stored in the repository).  The transliteration never touches an expression's structure:

  * parameter qualifiers:  `inout T x` / `out T x` -> `T& x`,  `in T x` -> `T x`
  * float literals get an `f` suffix (every GLSL literal is fp32; in C++ an unsuffixed one would be a double)
  * vector swizzles used as values (`.xy`, `.xyz`, `.rgb`) -> member calls `.xy()` ...
  * GLSL built-ins (`pow`, `exp`, `acos`, `normalize`, `mix`, `max` ...) -> `GLSL_<name>` so that they resolve to the numerical
    contract's definitions (oracle/glsl_types.h, include/eid_detmath.h) instead of libm
  * a vector constructor whose arguments are all rand() calls, `vec3(rand(s), rand(s), rand(s))`, becomes a brace initialiser
    `vec3{rand(s), rand(s), rand(s)}`: GLSL evaluates constructor arguments left to right, C++ only guarantees that order inside braces
    (GCC evaluates parenthesised arguments right to left), and the RNG draw order is part of the result
  * (--swizzle-assign, post.frag only) a statement that assigns to the first three components of a vec4, `v.rgb = e;` / `v.xyz = e;`,
    becomes `v = vec4(e, v.w);`, and the identity swizzle `.rgba` is dropped
  * `#include`, `precision`, `#extension`, `#version` lines and a stage's interface declarations (`layout(push_constant) uniform ...`,
    `layout(local_size_x ...) in;`) are dropped; optionally only the named top-level functions are kept

usage: glsl_prep.py <in.glsl> <out.hpp> [--only name1,name2,...] [--drop name1,...] [--strip <regex of whole lines to drop>]
"""
import re
import sys

BUILTINS = ["abs", "acos", "asin", "atan", "clamp", "cos", "cross", "dot", "exp", "floor", "inverse", "isnan", "isinf", "length", "max", "min",
            "mix", "normalize", "pow", "reflect", "sin", "smoothstep", "sqrt", "tan", "intBitsToFloat", "floatBitsToInt", "uintBitsToFloat",
            "floatBitsToUint", "unpackUnorm4x8", "packUnorm4x8"]
TYPES = r"(?:float|int|uint|bool|vec2|vec3|vec4|ivec2|ivec3|uvec2|uvec3|uvec4|ivec4|mat3|mat4|State|Material|Ray|SunAndSky|DirectReservoir|IndirectReservoir|LightSample|GISample|RngStateType|sampler2D|GltfShadeMaterial|TrigLight|PuncLight|PtPayload|uimage2D|image2D|rayQueryEXT|ShadeState)"
FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def top_level_chunks(src):
    """Splits the file into top-level chunks: (function name or None, text).  A function chunk ends at its closing brace."""
    out, i, n = [], 0, len(src)
    while i < n:
        m = re.compile(r"[ \t]*(?:const\s+)?\w[\w<>]*\s+(\w+)\s*\(([^;{}]*)\)\s*\{").match(src, i)
        if m and (i == 0 or src[i - 1] == "\n"):
            depth, j = 0, m.end() - 1
            while j < n:
                if src[j] == "{":
                    depth += 1
                elif src[j] == "}":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            out.append((m.group(1), src[i:j + 1] + "\n"))
            i = j + 1
        else:
            k = src.find("\n", i)
            k = n if k < 0 else k + 1
            out.append((None, src[i:k]))
            i = k
    return out


SWIZZLE_ASSIGN = False


def transliterate(text):
    if SWIZZLE_ASSIGN:
        text = re.sub(r"\b(\w+)\.(?:rgb|xyz)\s*=(?!=)\s*([^;]+);", r"\1 = vec4(\2, \1.w);", text)
        text = re.sub(r"\.rgba\b(?!\s*\()", "", text)
    text = re.sub(r"\b(?:inout|out)\s+(" + TYPES + r")\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"\bin\s+(" + TYPES + r")\s+(\w+)", r"\1 \2", text)
    text = FLOAT_LIT.sub(lambda m: m.group(1) + "f", text)
    text = re.sub(r"\b(vec[234])\(\s*(rand\([\w.]+\)(?:\s*,\s*rand\([\w.]+\))+)\s*\)", r"\1{\2}", text)
    text = re.sub(r"\b(attr\d\.tangent)\.x\b", r"\1", text)      # `.x` of a uint (VertexAttributes.tangent): a scalar swizzle, not C++
    text = re.sub(r"\.(xy|xyz|rgb)\b(?!\s*\()", r".\1()", text)
    text = re.sub(r"(?<=[\w\)\]])\.([rgba])\b(?!\s*\()", lambda m: "." + "xyzw"["rgba".index(m.group(1))], text)   # colour component names
    text = re.sub(r"\b(" + "|".join(BUILTINS) + r")\s*\(", r"GLSL_\1(", text)
    return text


def main():
    src_path, out_path = sys.argv[1], sys.argv[2]
    only = drop = strip = None
    args = sys.argv[3:]
    while args:
        if args[0] == "--only":
            only = set(args[1].split(","))
        elif args[0] == "--drop":
            drop = set(args[1].split(","))
        elif args[0] == "--strip":
            strip = args[1]
        elif args[0] == "--swizzle-assign":
            global SWIZZLE_ASSIGN
            SWIZZLE_ASSIGN = True
            args = args[1:]
            continue
        args = args[2:]
    src = strip_comments(open(src_path).read())
    lines = [l for l in src.split("\n") if not re.match(r"\s*#\s*(include|extension|version)\b", l) and not re.match(r"\s*precision\s", l)]
    if strip:
        lines = [l for l in lines if not re.match(strip, l.strip())]
    src = "\n".join(lines)
    # interface declarations of a shader stage (bound by the wrapper as plain globals): push-constant blocks, local_size
    src = re.sub(r"layout\s*\(\s*push_constant\s*\)\s*uniform\s+\w+\s*\{[^}]*\}\s*;", "", src)
    src = re.sub(r"layout\s*\([^)]*\)\s*in\s*;", "", src)
    keep = []
    for name, text in top_level_chunks(src):
        if name is not None and ((only is not None and name not in only) or (drop is not None and name in drop)):
            continue
        keep.append(text)
    with open(out_path, "w") as f:
        f.write("// GENERATED by oracle/ref_shim/glsl_prep.py from %s — build output, do not commit\n" % src_path)
        f.write(transliterate("".join(keep)))


if __name__ == "__main__":
    main()
