#!/usr/bin/env python
"""oracle/build_ref.py — compile the reference's OWN shader sources for the CPU (TEST INFRASTRUCTURE).

Reads the GLSL files where they lie under $VXRT_REFERENCE (default /root/reference), rewrites the
GLSL-only syntax into C++ (storage qualifiers, parameter qualifiers, float literals, array constructors),
wraps each shader in a namespace nested in `glsl` (oracle/glsl_shim.h supplies types, built-ins and
samplers) and links the result with oracle/ref_driver.cpp into oracle/_ref/libvxrt_ref.so.

Nothing from the reference is written into the repository: the generated C++ lives only under
oracle/_ref/gen (git-ignored).  The shader *logic* is therefore executed unmodified; only what GL leaves
to the driver (texture filtering, conversions) comes from our shim and is documented in DESIGN.md.
"""
from __future__ import annotations

import hashlib
import itertools
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ODIR = ROOT / "oracle"
REF = Path(os.environ.get("VXRT_REFERENCE", "/root/reference"))
GEN = ODIR / "_ref" / "gen"
LIB = ODIR / "_ref" / "libvxrt_ref.so"

SHADERS = {
    # namespace name -> (file under Core/Shaders, kind)
    "ManhattanDistanceX": ("ManhattanDistanceX.comp", "compute"),
    "ManhattanDistanceY": ("ManhattanDistanceY.comp", "compute"),
    "ManhattanDistanceZ": ("ManhattanDistanceZ.comp", "compute"),
    "InitialRayTraceFrag": ("InitialRayTraceFrag.glsl", "fragment"),
    "ShadowRayTraceFrag": ("ShadowRayTraceFrag.glsl", "fragment"),
    "DiffuseRayTraceFrag": ("DiffuseRayTraceFrag.glsl", "fragment"),
    "ReflectionTraceFrag": ("ReflectionTraceFrag.glsl", "fragment"),
    "GenerateGBuffer": ("GenerateGBuffer.glsl", "fragment"),
    # SVGF chain of the diffuse GI: temporal accumulation, variance estimate, a-trous spatial filter
    "SVGFTemporal": ("SVGF/TemporalFilter.glsl", "fragment"),
    "SVGFVariance": ("SVGF/VarianceEstimate.glsl", "fragment"),
    "SVGFSpatial": ("SVGF/SpatialFilter.glsl", "fragment"),
    "SVGFPreSpatial": ("Spatial3x3Initial.glsl", "fragment"),
    # average colour per block type for the light propagation volume (Volumetrics::CreateVolume, VolumetricFloodFill.cpp:102-123)
    "LPVAverageColor": ("Volumetrics/PrecomputeAverageBlockColor.comp", "compute"),   # the 3 x 3 pass in front of the temporal filter (Pipeline.cpp:2381-2424)
    # sun-shadow denoiser
    "ShadowTemporal": ("ShadowTemporalFilter.glsl", "fragment"),
    "ShadowFilter": ("ShadowFilter.glsl", "fragment"),
    # reflection temporal filter
    "SpecularTemporal": ("SpecularTemporalFilter.glsl", "fragment"),
    "ReflectionDenoise": ("ReflectionDenoiserNew.glsl", "fragment"),
    # the colour pass composites sky / clouds / denoised GI (out of scope); only its Cook-Torrance
    # functions are compiled: they are cut out of the file by name, unmodified
    "ColorPassDirect": ("ColorPassFrag.glsl", "extract"),
}
EXTRACT = {
    "ColorPassDirect": {
        "prelude": "#define PI 3.14159265359\nuniform vec3 u_ViewerPosition;\n"
                   "vec3 FresnelSchlickRoughness(vec3 Eye, vec3 norm, vec3 F0, float roughness);\n",
        "functions": ["ndfGGX", "gaSchlickG1", "gaSchlickGGX", "CalculateDirectionalLight", "FresnelSchlickRoughness", "BasicSaturation"],
    }
}


def extract_functions(text: str, names) -> str:
    """Cut whole function definitions (all overloads) out of a GLSL file, in file order."""
    text = strip_comments(text)
    out = []
    pat = re.compile(r"^[ \t]*(?:float|vec[234]|void|int|bool)\s+(\w+)\s*\([^;{]*\)\s*\{", re.M)
    for m in pat.finditer(text):
        if m.group(1) not in names:
            continue
        depth, j = 0, m.end() - 1
        while j < len(text):
            if text[j] == "{":
                depth += 1
            elif text[j] == "}":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        out.append(text[m.start():j + 1])
    return "\n\n".join(out)


TYPES = (r"(?:float|int|uint|bool|vec[234]|ivec[234]|uvec[234]|bvec[234]|mat[34](?:x[34])?|"
         r"sampler2D|sampler3D|usampler3D|sampler2DArray|samplerCube|image3D|Ray|[A-Z]\w*)")


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def fix_float_literals(src: str) -> str:
    out = []
    for line in src.split("\n"):
        if line.lstrip().startswith("#version") or line.lstrip().startswith("#extension"):
            continue
        # 1.0  .5  1.  1e-5  1.0e5   (not already suffixed, not part of an identifier)
        line = re.sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])", r"\1f", line)
        # GLSL accepts 'lf'/'F' rarely; normalise 1.0F -> 1.0f is already valid C++
        out.append(line)
    return "\n".join(out)


def match_paren(s: str, i: int) -> int:
    """index of the ')' matching the '(' at s[i]"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    raise ValueError("unbalanced parentheses")


def fix_array_constructors(src: str) -> str:
    # TYPE[](...) / TYPE[N](...)  ->  { ... }
    pat = re.compile(r"\b(?:float|int|uint|vec[234]|ivec[234])\s*\[\s*\w*\s*\]\s*\(")
    while True:
        m = pat.search(src)
        if not m:
            break
        open_i = m.end() - 1
        close_i = match_paren(src, open_i)
        src = src[:m.start()] + "{" + src[open_i + 1:close_i] + "}" + src[close_i + 1:]
    # functions returning arrays: `float[6] f(` -> `arr<float, 6> f(`;  `float x[6] = f(...)` -> `auto x = f(...)`
    src = re.sub(r"\b(float|int|vec[234])\s*\[\s*(\d+)\s*\]\s+(\w+)\s*\(", r"arr<\1, \2> \3(", src)
    src = re.sub(r"\b(?:float|int|vec[234])\s+(\w+)\s*\[\s*\d+\s*\]\s*=\s*(?![\s{])", r"auto \1 = ", src)
    # `const vec2[4] name = {` -> `const vec2 name[4] = {`
    src = re.sub(r"\b(" + TYPES + r")\s*\[\s*(\w+)\s*\]\s+(\w+)\s*=", r"\1 \3[\2] =", src)
    return src


def wrap_call_swizzles(src: str) -> str:
    """`f(args).xy` -> `vec2(f(args).xy)`: a swizzle of a temporary must be materialised before the
    temporary dies or is copied (e.g. as an operand of ?:)."""
    pat = re.compile(r"\)\s*\.([xyzw]{2,4}|[rgba]{2,4}|[stpq]{2,4})\b(?!\s*\()")
    pos = 0
    while True:
        m = pat.search(src, pos)
        if not m:
            break
        close_i = m.start()
        depth, j = 0, close_i
        while j >= 0:
            if src[j] == ")":
                depth += 1
            elif src[j] == "(":
                depth -= 1
                if depth == 0:
                    break
            j -= 1
        k = j
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] == "_"):
            k -= 1
        name = src[k:j]
        n = len(m.group(1))
        if not name or name in ("if", "while", "for", "return", "switch"):
            pos = m.end()
            continue
        # integer-valued built-ins keep their integer type
        ctor = ("ivec%d" if name in ("textureSize", "ivec2", "ivec3", "ivec4") else "vec%d") % n
        if name in ("uvec2", "uvec3", "uvec4"):
            ctor = "uvec%d" % n
        src = src[:k] + ctor + "(" + src[k:m.end()] + ")" + src[m.end():]
        pos = m.end() + len(ctor) + 2
    return src


def fix_param_qualifiers(src: str) -> str:
    src = re.sub(r"\b(?:inout|out)\s+(" + TYPES + r")\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bconst\s+in\s+(" + TYPES + r")\s+(\w+)", r"const \1 \2", src)
    src = re.sub(r"(?<=[(,])\s*in\s+(" + TYPES + r")\s+(\w+)", r" \1 \2", src)
    return src


def fix_globals(src: str, kind: str, resets: list) -> str:
    """storage qualifiers at global scope -> plain / thread_local globals; SSBO blocks -> ssbo arrays.
    `resets` collects `name = init;` statements so every invocation starts from the declared values."""
    out = []
    depth = 0
    in_buffer = False
    lines = src.split("\n")
    for line in lines:
        stripped = line.strip()
        if depth == 0:
            m = re.match(r"layout\s*\(.*?\)\s*(?:readonly\s+|writeonly\s+|restrict\s+|coherent\s+)*buffer\s+\w+", stripped)
            if m:
                in_buffer = True
                out.append("// ssbo " + stripped.split("buffer")[1].strip())
                if "{" in stripped:
                    depth += stripped.count("{") - stripped.count("}")
                continue
            if re.match(r"layout\s*\(\s*local_size", stripped):
                out.append("")
                continue
            if re.match(r"precision\s", stripped):
                out.append("")
                continue
            m = re.match(r"layout\s*\([^)]*\)\s*out\s+(" + TYPES + r")\s+(\w+)\s*;", stripped)
            if m:
                out.append(f"thread_local {m.group(1)} {m.group(2)};")
                continue
            m = re.match(r"(?:flat\s+|smooth\s+)?(in|out)\s+(" + TYPES + r")\s+(\w+)\s*;", stripped)
            if m:
                out.append(f"thread_local {m.group(2)} {m.group(3)};")
                continue
            m = re.match(r"layout\s*\([^)]*\)\s*uniform\s+(.*)", stripped)
            if m:
                stripped = "uniform " + m.group(1)
            m = re.match(r"uniform\s+(" + TYPES + r")\s+(\w+)\s*(\[\s*\w+\s*\])?\s*(=\s*[^;]+)?;", stripped)
            if m:
                out.append(f"{m.group(1)} {m.group(2)}{m.group(3) or ''} {m.group(4) or ''};")
                continue
            # mutable plain globals (one per line in these shaders) become per-invocation state
            m = re.match(r"(" + TYPES + r")\s+(\w+)\s*(=\s*[^;{]+)?;\s*$", stripped)
            if m and not stripped.startswith(("return", "const", "struct")) and "(" not in stripped.split("=")[0]:
                out.append(f"thread_local {stripped}")
                if m.group(3):
                    resets.append(f"{m.group(2)} {m.group(3).strip()};")
                continue
        elif in_buffer and depth == 1:
            m = re.match(r"(" + TYPES + r")\s+(\w+)\s*\[\s*([^\]]*)\s*\]\s*;", stripped)
            if m:
                ty, name, size = m.group(1), m.group(2), m.group(3).strip()
                if size:
                    out.append(f"ssbo_array<{ty}, ({size})> {name};")
                else:
                    out.append(f"ssbo_unsized<{ty}> {name};")
                continue
            if stripped.startswith("}"):
                in_buffer = False
                depth = 0
                out.append("")
                continue
            if stripped in ("{", ""):
                depth = 1 if stripped == "{" else depth
                out.append("")
                continue
        if in_buffer and depth == 0 and stripped == "{":
            depth = 1
            out.append("")
            continue
        out.append(line)
        if not in_buffer:
            depth += line.count("{") - line.count("}")
    return "\n".join(out)


def transform(name: str, text: str, kind: str) -> str:
    src = strip_comments(text)
    src = fix_float_literals(src)
    src = fix_array_constructors(src)
    resets: list = []
    src = fix_globals(src, kind, resets)
    src = fix_param_qualifiers(src)
    src = wrap_call_swizzles(src)
    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void shader_main()", src)
    src = re.sub(r"\bdiscard\s*;", "{ shader_discarded = true; return; }", src)
    # mat3(mat4) constructor
    src = re.sub(r"\bmat3\s*\(\s*(u_\w+|\w+Matrix\w*)\s*\)", r"to_mat3(\1)", src)
    hdr = (f"// GENERATED by oracle/build_ref.py from the reference shader {SHADERS[name][0]} — do not commit\n"
           "#include \"../../glsl_shim.h\"\n"
           f"namespace glsl {{ namespace shader_{name} {{\n"
           "thread_local bool shader_discarded = false;\n")
    reset_fn = "\nvoid shader_reset() { shader_discarded = false; " + " ".join(resets) + " }\n"
    return hdr + src + reset_fn + "\n} }\n"


def gen_raycast_detect() -> bool:
    """World::RaycastDetect is host C++ (Core/World.cpp): its text is lifted in place into a generated translation unit
    with a minimal World (the real class drags in OpenGL), compiled against the reference's own glm.  The function has
    no return statement after its loop (undefined behaviour when nothing is hit): the generated copy returns
    ivec4(-2) there so that the no-hit case can be observed."""
    src = (REF / "Core" / "World.cpp").read_text(errors="replace")
    m = re.search(r"glm::ivec4\s+VoxelRT::World::RaycastDetect\s*\(", src)
    glm_dir = REF / "Dependencies" / "glm"
    if not m or not (glm_dir / "glm" / "glm.hpp").exists():
        return False
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        depth += {"{": 1, "}": -1}.get(src[j], 0)
        if depth == 0:
            break
        j += 1
    body = src[m.start():j] + "\n\treturn glm::ivec4(-2); // added: the reference falls off the end here\n}"
    macros = (REF / "Core" / "Macros.h").read_text(errors="replace")
    sizes = "\n".join(l for l in macros.splitlines() if re.match(r"\s*#define\s+WORLD_SIZE_[XYZ]\b", l))
    (GEN / "RaycastDetect.cpp").write_text(f"""// GENERATED from Core/World.cpp by oracle/build_ref.py -- do not commit
#include <stdint.h>
#include <math.h>
#include <algorithm>
#include <glm/glm.hpp>
{sizes}
namespace VoxelRT {{
struct Block {{ uint8_t block; }};
struct World {{
    const Block* m_WorldData;
    const Block& GetBlock(uint16_t x, uint16_t y, uint16_t z) {{ return m_WorldData[x + y * WORLD_SIZE_X + z * WORLD_SIZE_X * WORLD_SIZE_Y]; }}
    glm::ivec4 RaycastDetect(const glm::vec3& pos, const glm::vec3& dir);
}};
}}
{body}
extern "C" void vxref_raycast_detect(const uint8_t* blocks, const float* pos, const float* dir, int32_t n, int32_t* out4) {{
    VoxelRT::World w; w.m_WorldData = reinterpret_cast<const VoxelRT::Block*>(blocks);
    for (int32_t i = 0; i < n; ++i) {{
        glm::ivec4 r = w.RaycastDetect(glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), glm::vec3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]));
        out4[4 * i] = r.x; out4[4 * i + 1] = r.y; out4[4 * i + 2] = r.z; out4[4 * i + 3] = r.w;
    }}
}}
""")
    cmd = ["g++", "-std=gnu++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w", f"-I{glm_dir}", "-c", str(GEN / "RaycastDetect.cpp"),
           "-o", str(GEN / "RaycastDetect.o")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print("[build_ref] RaycastDetect: does not compile, skipped\n" + "\n".join(r.stderr.split("\n")[:30]))
        return False
    return True


def gen_swizzles():
    GEN.mkdir(parents=True, exist_ok=True)
    sets = ["xyzw", "rgba", "stpq"]
    for n in (2, 3, 4):
        lines = []
        for names in sets:
            comps = names[:n]
            for k in (2, 3, 4):
                for combo in itertools.product(range(n), repeat=k):
                    nm = "".join(comps[c] for c in combo)
                    idx = list(combo) + [-1] * (4 - k)
                    lines.append(f"swz<T, {k}, {idx[0]}, {idx[1]}, {idx[2]}, {idx[3]}> {nm};")
        (GEN / f"swizzle{n}.inc").write_text("\n".join(lines) + "\n")


def main():
    force = "--force" in sys.argv
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    if not REF.exists():
        print("reference tree not mounted; nothing to do")
        return 0
    gen_swizzles()
    sdir = REF / "Core" / "Shaders"
    driver = ODIR / "ref_driver.cpp"
    deps = [ODIR / "glsl_shim.h", ODIR / "vxo_math.h", ODIR / "vxo_texture.h", driver, Path(__file__), ROOT / "include" / "vxrt_cuda.h"]
    h = hashlib.sha256()
    for d in deps:
        h.update(d.read_bytes())
    if (REF / "Core" / "World.cpp").exists():
        h.update((REF / "Core" / "World.cpp").read_bytes())
    names = [n for n in SHADERS if (not only or n in only)]
    enabled = []
    for n in names:
        f = sdir / SHADERS[n][0]
        if f.exists():
            h.update(f.read_bytes())
            enabled.append(n)
    stamp = h.hexdigest()
    stamp_file = LIB.with_suffix(".so.stamp")
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return 0
    want = set(re.findall(r"VXREF_HAVE_(\w+)", driver.read_text()))
    objs, defs = [], []
    for n in enabled:
        if n not in want:
            continue
        cpp = GEN / f"{n}.cpp"
        text = (sdir / SHADERS[n][0]).read_text(errors="replace")
        if SHADERS[n][1] == "extract":
            text = EXTRACT[n]["prelude"] + extract_functions(text, EXTRACT[n]["functions"])
        code = transform(n, text, SHADERS[n][1])
        if n == "LPVAverageColor":   # the one buffer a shader of this set writes: a plain array instead of the read-only SSBO view
            code, k = re.subn(r"ssbo_array<vec4,\s*\(128\)>\s+BlockAverageColorData;", "vec4 BlockAverageColorData[128];", code)
            assert k == 1
        cpp.write_text(code)
        obj = GEN / f"{n}.o"
        cmd = ["g++", "-std=gnu++20", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-fopenmp", "-w", "-fpermissive",
               "-c", str(cpp), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(f"[build_ref] {n}: does not compile through the shim, skipped\n" + "\n".join(r.stderr.split("\n")[:40]))
            continue
        defs.append(f"-DVXREF_HAVE_{n}=1")
    extra = []
    if gen_raycast_detect():
        extra.append(str(GEN / "RaycastDetect.o"))
        defs.append("-DVXREF_HAVE_RaycastDetect=1")
    cmd = ["g++", "-std=gnu++20", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-fopenmp", "-w", "-shared",
           f"-I{GEN}", "-o", str(LIB), str(driver)] + extra + defs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stderr[:6000])
        return 1
    stamp_file.write_text(stamp)
    print("[build_ref] built", LIB, "with shaders:", ", ".join(d.split("HAVE_")[1].split("=")[0] for d in defs))
    return 0


if __name__ == "__main__":
    sys.exit(main())
