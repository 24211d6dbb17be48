/*
 * vxrt_host.h — host-side mirror of the reference's C++ producers of the hot path's inputs:
 * camera matrices (Core/FpsCamera.cpp, glm), TAA jitter (Core/TAAJitter.cpp), sun/moon direction
 * (Core/Pipeline.cpp:1731-1749), world container + raw save format (Core/World.h,
 * Core/WorldFileHandler.cpp), deterministic stand-in world generators and the
 * block database (Core/BlockDatabaseParser.cpp, Core/BlockDataSSBO.cpp).
 * Plain C ABI so the parity tests and bench (Python, ctypes) and a C++ engine can both use it.
 */
#ifndef VXRT_HOST_H
#define VXRT_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* glm::perspective(radians(fov_deg), aspect, near, far) — column-major out[16] */
void vxh_perspective(float fov_deg, float aspect, float z_near, float z_far, float* out16);
/* glm::lookAt(eye, center, up) */
void vxh_look_at(const float* eye, const float* center, const float* up, float* out16);
/* glm::inverse(mat4) */
void vxh_inverse(const float* m16, float* out16);
/* FPSCamera: front from yaw/pitch (FpsCamera.cpp:66-70), view = lookAt(pos, pos+front, (0,1,0)),
 * proj = perspective(fov, aspect, 0.1, 1000) (Player.cpp:10); outputs may be NULL.               */
void vxh_camera(const float* pos, float yaw_deg, float pitch_deg, float fov_deg, float aspect,
                float* view16, float* proj16, float* inv_view16, float* inv_proj16);
/* Halton(2,3) jitter of GenerateJitterStuff / GetTAAJitter (TAAJitter.cpp:6-41) */
void vxh_taa_jitter(int32_t frame, float* out2);
/* Pipeline.cpp:1731-1749: sun = normalize(Rz(2*sun_tick deg) * (1,1,1)), moon = (-x,-y,z),
 * stronger = (-sun.y < 0.01) ? sun : moon                                                        */
void vxh_sun_direction(float sun_tick, float* sun3, float* moon3, float* stronger3);

/* deterministic stand-in worlds; blocks is nx*ny*nz bytes, x fastest */
void vxh_gen_plains(uint32_t seed, int32_t structures, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks);
void vxh_gen_rooms(uint32_t seed, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks);
void vxh_gen_town(uint32_t seed, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks);
/* config-2 edit list: n random toggles (solid -> 0, air -> id 3) applied to `blocks` in order;
 * xyz_id receives n x {x,y,z,id}.  PRNG: mt19937(seed).                                          */
void vxh_random_edits(uint32_t seed, int32_t n, int32_t nx, int32_t ny, int32_t nz, uint8_t* blocks,
                      int32_t* xyz_id);

/* raw headerless world dump (WorldFileHandler.cpp:10-83); 0 on success */
int32_t vxh_world_save(const char* path, const uint8_t* blocks, int64_t nbytes);
int32_t vxh_world_load(const char* path, uint8_t* blocks, int64_t nbytes);

/* ---- block database (Core/BlockDatabaseParser.cpp:44-417, Core/BlockDatabase.cpp:13-104, 111-468) ----
 * kinds: 0 albedo, 1 normal, 2 pbr, 3 emissive.  faces: 0 front, 1 back, 2 top, 3 bottom, 4 left, 5 right.
 * Block ids are the 1-based order of the records; a texture's layer is its index in the sorted list of
 * unique paths of its kind (Core/GLClasses/TextureArray.cpp:12-13,56).                              */
typedef struct vxh_blockdb vxh_blockdb;
vxh_blockdb* vxh_blockdb_parse(const char* path); /* NULL if the file cannot be opened */
void vxh_blockdb_free(vxh_blockdb* db);
int32_t vxh_blockdb_block_count(const vxh_blockdb* db);
int32_t vxh_blockdb_block_id(const vxh_blockdb* db, const char* name);       /* 0 if unknown */
const char* vxh_blockdb_block_name(const vxh_blockdb* db, int32_t id);       /* "???" if unknown */
int32_t vxh_blockdb_layer_count(const vxh_blockdb* db, int32_t kind);
const char* vxh_blockdb_layer_path(const vxh_blockdb* db, int32_t kind, int32_t layer);
/* GetBlockTexture / GetBlockNormalTexture / GetBlockPBRTexture / GetBlockEmissiveTexture by id */
int32_t vxh_blockdb_texture(const vxh_blockdb* db, int32_t kind, int32_t block_id, int32_t face);
/* BlockDataSSBO::CreateBuffers (Core/BlockDataSSBO.cpp:17-35): int[6][128] */
void vxh_blockdb_table(const vxh_blockdb* db, int32_t* out6x128);
/* u_GrassBlockProps / u_CactusBlockProps layout (Core/Pipeline.cpp:2166-2186): {id, top a/n/p, front a/n/p, bottom a/n/p} */
void vxh_blockdb_face_props(const vxh_blockdb* db, const char* name, int32_t* out10);
/* BlockDatabase::GetIDFromMCID for every 8-bit Minecraft id (Core/BlockDatabase.cpp:93-103, 599-612): out256[0] = 0,
 * out256[mc_id] = the block that lists mc_id, or the id of "INVALID_BLOCK" for ids no block lists */
void vxh_blockdb_minecraft_lut(const vxh_blockdb* db, uint8_t* out256);

/* ---- Minecraft Anvil region reader: host half of MCWorldImporter::ImportWorld (Core/NBT/Importer.cpp:85-166);
 * see vxrt_mca.cpp.  A vxh_mca accumulates chunk sections in the layout vxrt_cuda_import_sections takes. ---- */
typedef struct vxh_mca vxh_mca;
vxh_mca* vxh_mca_open(void);
void vxh_mca_free(vxh_mca* m);
/* number of sections added, -1 if the file / directory cannot be opened, -2 on a short read */
int32_t vxh_mca_add_region_file(vxh_mca* m, const char* path);
int32_t vxh_mca_add_region_dir(vxh_mca* m, const char* dir);   /* every *.mca of the directory (Importer.cpp:152-159) */
int32_t vxh_mca_section_count(const vxh_mca* m);
int32_t vxh_mca_chunk_count(const vxh_mca* m);
int32_t vxh_mca_palette_section_count(const vxh_mca* m);       /* 1.13+ palette sections met and skipped */
int32_t vxh_mca_bad_chunk_count(const vxh_mca* m);             /* chunks whose stream does not inflate */
const uint8_t* vxh_mca_block_ids(const vxh_mca* m);            /* n * 4096, YZX */
const uint8_t* vxh_mca_data_nibbles(const vxh_mca* m);         /* n * 2048, low nibble first (zeros where has_data = 0) */
const uint8_t* vxh_mca_has_data(const vxh_mca* m);             /* n */
const int32_t* vxh_mca_section_origins(const vxh_mca* m);      /* 3 * n: chunk x * 16, section Y * 16, chunk z * 16 */

/* deterministic synthetic 'block textures' (the reference's PNGs do not travel to the GPU box):
 * layers*size*size RGBA8, kind-appropriate content (albedo colours, tangent-space normals,
 * roughness/metal/displacement/AO, emissive masks).                                                 */
void vxh_gen_texture_array(uint32_t seed, int32_t kind, int32_t layers, int32_t size, uint8_t* rgba);

#ifdef __cplusplus
}
#endif
#endif
