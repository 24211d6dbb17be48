#!/usr/bin/env python3
"""oracle/_ref/libvxrt_ref_world.so — the reference's own world producers compiled for the CPU, to pin the
oracle restatement (vxrt_oracle_world.cpp) and to generate tests/golden/world_ref.npz.  TEST INFRASTRUCTURE ONLY.

Nothing of the reference is copied into the repository: its sources are compiled where they lie under /root/reference
(FastNoise.cpp, enkimi.c, miniz.c as they are) or lifted as text into generated translation units under the git-ignored
oracle/_ref/gen (Core/WorldGenerator.cpp and Core/NBT/Importer.cpp: everything after their #include lines, behind a few
stub declarations replacing the headers that drag in OpenGL).

  * WorldGenerator.cpp: `rand` is redirected to a scripted source so that the three FastNoise seeds the function draws as
    rand() % 50000 (:217-219) can be chosen by the caller (later draws, only made when structures are on, come from an
    LCG); `srand(time(0))` becomes a no-op; the three function-local `static FastNoise` objects are made automatic so the
    function can be called with more than one seed triple per process.  World / Block are the minimal container the
    function needs (World.h:46-70: unchecked SetBlock / GetBlock over the x-fastest array).
  * VolumetricFloodFill.h/.cpp: the OpenGL calls (texture mirrors of the two host arrays) become no-ops and World is
    the read-only block array; the entry points replay the call sequences of Pipeline.cpp:1602-1611 / World.cpp:554-572 and of the
    block edit in World.cpp:273-485 around the unmodified PropogateVolume / DepropogateVolume.
  * Importer.cpp: BlockDatabase::GetIDFromMCID answers from the 256-entry table the caller passes (the oracle and the
    product build that table from blockdb.txt, vxh_blockdb_minecraft_lut).
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("VXRT_REFERENCE", "/root/reference"))
GEN = ROOT / "oracle" / "_ref" / "gen"
LIB = ROOT / "oracle" / "_ref" / "libvxrt_ref_world.so"

CXX = ["g++", "-std=gnu++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w"]
CC = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-w"]


def body_after_includes(text: str) -> str:
    lines = text.splitlines()
    last = max((i for i, l in enumerate(lines) if l.lstrip().startswith("#include")), default=-1)
    return "\n".join(lines[last + 1:])


def sizes() -> str:
    macros = (REF / "Core" / "Macros.h").read_text(errors="replace")
    return "\n".join(l for l in macros.splitlines() if re.match(r"\s*#define\s+WORLD_SIZE_[XYZ]\b", l))


def gen_worldgen() -> Path:
    src = body_after_includes((REF / "Core" / "WorldGenerator.cpp").read_text(errors="replace"))
    src, n = re.subn(r"\bstatic\s+FastNoise\b", "FastNoise", src)
    assert n == 3, "expected the three function-local FastNoise objects"
    out = GEN / "WorldGenerator.cpp"
    out.write_text(f"""// GENERATED from Core/WorldGenerator.cpp by oracle/build_ref_world.py -- do not commit
#include <stdint.h>
#include <math.h>
#include <time.h>
#include <string.h>
#include <array>
#include <random>
#include <string>
#include <vector>
#include <glm/glm.hpp>
#include <FastNoise.h>
{sizes()}
namespace VoxelRT {{
struct Block {{ uint8_t block; }};
struct World {{
    Block* m_WorldData;
    const Block& GetBlock(const glm::ivec3& p) {{ return m_WorldData[p.x + p.y * WORLD_SIZE_X + p.z * WORLD_SIZE_X * WORLD_SIZE_Y]; }}
    void SetBlock(uint16_t x, uint16_t y, uint16_t z, Block block) {{ m_WorldData[x + y * WORLD_SIZE_X + z * WORLD_SIZE_X * WORLD_SIZE_Y] = block; }}
}};
void GenerateWorld(World* world, bool gen_type, bool gen_structures);
namespace BlockDatabase {{ uint8_t GetBlockID(const std::string& name); }}
}}
static int g_script[3];
static int g_script_at;
static uint32_t g_lcg;
static int vxref_rand() {{
    if (g_script_at < 3) return g_script[g_script_at++];
    g_lcg = g_lcg * 1664525u + 1013904223u;
    return (int)((g_lcg >> 1) & 0x7fffffff);
}}
static const int32_t* g_ids;   // Grass, Dirt, Stone, Sand, oak_log, oak_leaves, Cactus, Cobblestone
uint8_t VoxelRT::BlockDatabase::GetBlockID(const std::string& name) {{
    static const char* names[8] = {{"Grass", "Dirt", "Stone", "Sand", "oak_log", "oak_leaves", "Cactus", "Cobblestone"}};
    for (int i = 0; i < 8; ++i) if (name == names[i]) return (uint8_t)g_ids[i];
    return 0;
}}
#define rand vxref_rand
#define srand(x) ((void)0)
{src}
#undef rand
#undef srand
// seeds3: the values the three rand() % 50000 draws return, in the order of the function: biome, height noise, stone
extern "C" void vxref_generate_world(uint8_t* blocks, int32_t gen_type, int32_t gen_structures, const int32_t* seeds3, const int32_t* ids8) {{
    memset(blocks, 0, (size_t)WORLD_SIZE_X * WORLD_SIZE_Y * WORLD_SIZE_Z);
    g_script[0] = seeds3[0]; g_script[1] = seeds3[1]; g_script[2] = seeds3[2]; g_script_at = 0; g_lcg = 12345u;
    g_ids = ids8;
    VoxelRT::World w; w.m_WorldData = reinterpret_cast<VoxelRT::Block*>(blocks);
    VoxelRT::GenerateWorld(&w, gen_type != 0, gen_structures != 0);
}}
// FastNoise itself, for direct checks of the restated noise
extern "C" void vxref_fastnoise_2d(int32_t seed, int32_t fractal, float frequency, int32_t octaves, const float* xy, int32_t n, float* out) {{
    FastNoise g(seed);
    g.SetNoiseType(fractal ? FastNoise::SimplexFractal : FastNoise::Simplex);
    g.SetFrequency(frequency);
    if (fractal) g.SetFractalOctaves(octaves);
    for (int32_t i = 0; i < n; ++i) out[i] = g.GetNoise(xy[2 * i], xy[2 * i + 1]);
}}
""")
    return out


def gen_importer() -> Path:
    src = body_after_includes((REF / "Core" / "NBT" / "Importer.cpp").read_text(errors="replace"))
    out = GEN / "Importer.cpp"
    out.write_text(f"""// GENERATED from Core/NBT/Importer.cpp by oracle/build_ref_world.py -- do not commit
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <array>
#include <filesystem>
#include <iostream>
#include <string>
#include <vector>
#include <glm/glm.hpp>
extern "C" {{
#include <enkimi.h>
}}
{sizes()}
static const uint8_t* g_lut;
namespace VoxelRT {{ namespace BlockDatabase {{ static uint8_t GetIDFromMCID(uint8_t id) {{ return g_lut[id]; }} }} }}
{src}
extern "C" int32_t vxref_import_world(const char* dir, const float* origin3, const uint8_t* lut256, uint8_t* out) {{
    g_lut = lut256;
    try {{
        VoxelRT::MCWorldImporter::ImportWorld(dir, out, glm::vec3(origin3[0], origin3[1], origin3[2]));
    }} catch (...) {{
        return -1;
    }}
    return 0;
}}
""")
    return out


def gen_floodfill() -> Path:
    hdr = body_after_includes((REF / "Core" / "VolumetricFloodFill.h").read_text(errors="replace"))
    src = body_after_includes((REF / "Core" / "VolumetricFloodFill.cpp").read_text(errors="replace"))
    gl_enums = ["GL_TEXTURE_3D", "GL_TEXTURE_MIN_FILTER", "GL_TEXTURE_MAG_FILTER", "GL_LINEAR", "GL_NEAREST", "GL_TEXTURE_WRAP_S",
                "GL_TEXTURE_WRAP_T", "GL_TEXTURE_WRAP_R", "GL_CLAMP_TO_EDGE", "GL_RED", "GL_UNSIGNED_BYTE", "GL_R8UI", "GL_RED_INTEGER",
                "GL_TRUE", "GL_READ_WRITE", "GL_R8", "GL_SHADER_IMAGE_ACCESS_BARRIER_BIT", "GL_SHADER_STORAGE_BUFFER", "GL_STATIC_DRAW",
                "GL_TEXTURE4", "GL_TEXTURE_2D_ARRAY"]
    gl_calls = ["glGenTextures", "glBindTexture", "glTexParameteri", "glTexImage3D", "glBindImageTexture", "glDispatchCompute",
                "glMemoryBarrier", "glGenBuffers", "glBindBuffer", "glBufferData", "glActiveTexture", "glBindBufferBase", "glFinish",
                "glTexSubImage3D"]
    out = GEN / "VolumetricFloodFill.cpp"
    out.write_text(f"""// GENERATED from Core/VolumetricFloodFill.h/.cpp by oracle/build_ref_world.py -- do not commit
#include <stdint.h>
#include <string.h>
#include <array>
#include <iostream>
#include <memory>
#include <queue>
#include <glm/glm.hpp>
{sizes()}
// the OpenGL side of the file (texture creation, per-voxel glTexSubImage3D mirrors of the host arrays) is stubbed out: the arrays are
// the authoritative state (Reupload copies them whole)
typedef unsigned GLuint; typedef float GLfloat;
enum {{ {", ".join(f"{e} = {i + 1}" for i, e in enumerate(gl_enums))} }};
{chr(10).join(f"template <class... A> static void {c}(A...) {{}}" for c in gl_calls)}
namespace GLClasses {{ struct ComputeShader {{ void CreateComputeShader(const char*) {{}} void Compile() {{}} void Use() {{}} void SetInteger(const char*, int) {{}} }}; }}
namespace VoxelRT {{
struct Block {{ uint8_t block; }};
struct World {{
    const Block* m_WorldData;
    const Block& GetBlock(const glm::ivec3& p) {{ return m_WorldData[p.x + p.y * WORLD_SIZE_X + p.z * WORLD_SIZE_X * WORLD_SIZE_Y]; }}
}};
}}
int VoxelRT_FloodFillDistanceLimit = 4;   // Pipeline.cpp:53
{hdr}
{src}
namespace {{
VoxelRT::World g_world;
bool g_created = false;
void bind(const uint8_t* blocks, int32_t limit) {{
    g_world.m_WorldData = reinterpret_cast<const VoxelRT::Block*>(blocks);
    VoxelRT_FloodFillDistanceLimit = limit;
    if (!g_created) {{ VoxelRT::Volumetrics::CreateVolume(&g_world, 0, 0); g_created = true; }}
    while (!VoxelRT::LightBFS.empty()) VoxelRT::LightBFS.pop();
    while (!VoxelRT::LightRemovalBFS.empty()) VoxelRT::LightRemovalBFS.pop();
}}
const size_t kN = (size_t)WORLD_SIZE_X * WORLD_SIZE_Y * WORLD_SIZE_Z;
}}
// World::RepropogateLPV_ (World.cpp:554-572) / the start-up sequence Pipeline.cpp:1602-1611, `iterations` PropogateVolume calls
extern "C" void vxref_lpv_repropagate(const uint8_t* blocks, const int32_t* xyz, int32_t n, int32_t limit, int32_t iterations,
                                      uint8_t* level, uint8_t* color) {{
    using namespace VoxelRT;
    bind(blocks, limit);
    Volumetrics::ClearEntireVolume();
    for (int32_t i = 0; i < n; ++i) {{
        const glm::ivec3 e(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        const uint8_t block_at = g_world.GetBlock(e).block;
        Volumetrics::AddLightToVolume(e, block_at);
    }}
    for (int32_t i = 0; i < iterations; ++i) Volumetrics::PropogateVolume();
    memcpy(level, WorldVolumetricDensityData->data(), kN);
    memcpy(color, WorldVolumetricColorData->data(), kN);
}}
// the LPV statements of the block edit in World::Raycast, in the order of World.cpp:273-333 (op 1) and :395-446 (op 0), then :482-485
extern "C" void vxref_lpv_edit(const uint8_t* blocks_after, int32_t op, int32_t x, int32_t y, int32_t z, int32_t block, int32_t emissive,
                               int32_t limit, uint8_t* level, uint8_t* color) {{
    using namespace VoxelRT;
    bind(blocks_after, limit);
    memcpy(WorldVolumetricDensityData->data(), level, kN);
    memcpy(WorldVolumetricColorData->data(), color, kN);
    const glm::vec3 position((float)x, (float)y, (float)z);
    auto& LightRemovalBFS = Volumetrics::GetLightRemovalBFSQueue();
    auto& LightPropogateBFS = Volumetrics::GetLightBFSQueue();
    const float d[6][3] = {{{{1, 0, 0}}, {{-1, 0, 0}}, {{0, 1, 0}}, {{0, -1, 0}}, {{0, 0, 1}}, {{0, 0, -1}}}};
    if (op == 1) {{
        LightRemovalBFS.push(LightRemovalNode(glm::floor(position), Volumetrics::GetLightValue(glm::ivec3(glm::floor(position)))));
        for (int k = 0; k < 6; ++k) {{
            const glm::vec3 q = glm::floor(position + glm::vec3(d[k][0], d[k][1], d[k][2]));
            LightRemovalBFS.push(LightRemovalNode(q, Volumetrics::GetLightValue(glm::ivec3(q))));
        }}
        if (emissive) Volumetrics::AddLightToVolume(glm::ivec3((int)position.x, (int)position.y, (int)position.z), (uint8_t)block);
    }} else {{
        if (emissive) {{
            LightRemovalBFS.push(LightRemovalNode(glm::floor(position), Volumetrics::GetLightValue(glm::ivec3(glm::floor(position)))));
            Volumetrics::SetLightValue(glm::ivec3(glm::floor(position)), 0, 0);
            Volumetrics::UploadLight(glm::ivec3(glm::floor(position)), 0, 0, true);
        }}
        for (int k = 0; k < 6; ++k) LightPropogateBFS.push(LightNode(glm::floor(position + glm::vec3(d[k][0], d[k][1], d[k][2]))));
    }}
    for (int it = 0; it < 4; ++it) {{
        Volumetrics::DepropogateVolume();
        Volumetrics::PropogateVolume();
    }}
    memcpy(level, WorldVolumetricDensityData->data(), kN);
    memcpy(color, WorldVolumetricColorData->data(), kN);
}}
""")
    return out


def main() -> int:
    force = "--force" in sys.argv
    if not REF.exists():
        print("reference tree not mounted; nothing to do")
        return 0
    GEN.mkdir(parents=True, exist_ok=True)
    fn = REF / "Dependencies" / "fast_noise"
    enki = REF / "Dependencies" / "enkiMI"
    glm = REF / "Dependencies" / "glm"
    inputs = [Path(__file__), fn / "FastNoise.cpp", fn / "FastNoise.h", enki / "enkimi.c", enki / "enkimi.h", enki / "miniz.c", enki / "miniz.h",
              REF / "Core" / "WorldGenerator.cpp", REF / "Core" / "NBT" / "Importer.cpp", REF / "Core" / "Macros.h",
              REF / "Core" / "VolumetricFloodFill.cpp", REF / "Core" / "VolumetricFloodFill.h"]
    if not all(p.exists() for p in inputs) or not (glm / "glm" / "glm.hpp").exists():
        print("[build_ref_world] reference sources missing; skipped")
        return 0
    h = hashlib.sha256()
    for p in inputs:
        h.update(p.read_bytes())
    stamp_file = LIB.with_suffix(".so.stamp")
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == h.hexdigest():
        return 0
    objs = []

    def compile_(cmd, src, name):
        obj = GEN / name
        r = subprocess.run(cmd + ["-c", str(src), "-o", str(obj)], capture_output=True, text=True)
        if r.returncode != 0:
            print(f"[build_ref_world] {src} does not compile:\n" + "\n".join(r.stderr.splitlines()[:40]))
            raise SystemExit(1)
        objs.append(str(obj))
    compile_(CXX + [f"-I{fn}"], fn / "FastNoise.cpp", "FastNoise.o")
    compile_(CC + [f"-I{enki}"], enki / "enkimi.c", "enkimi.o")
    compile_(CC + [f"-I{enki}"], enki / "miniz.c", "miniz.o")
    compile_(CXX + [f"-I{fn}", f"-I{glm}"], gen_worldgen(), "WorldGenerator.o")
    compile_(CXX + [f"-I{enki}", f"-I{glm}"], gen_importer(), "Importer.o")
    compile_(CXX + [f"-I{glm}"], gen_floodfill(), "VolumetricFloodFill.o")
    r = subprocess.run(["g++", "-shared", "-o", str(LIB)] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stderr[:4000])
        return 1
    stamp_file.write_text(h.hexdigest())
    print("[build_ref_world] built", LIB)
    return 0


if __name__ == "__main__":
    sys.exit(main())
