/*
 * vxrt_oracle_shade.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, see vxrt_oracle.h).
 * Hit-material fetch (GenerateGBuffer.glsl), Cook-Torrance direct term (ColorPassFrag.glsl),
 * diffuse GI (DiffuseRayTraceFrag.glsl) and reflections (ReflectionTraceFrag.glsl).
 * Citations are Core/Shaders/<file>:line of the reference.
 */
#include "vxrt_oracle.h"
#include "vxo_math.h"
#include "vxo_texture.h"
#include "vxo_grid.h"

#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace vxo;

struct vxo_scene {
    vxo_world world;
    int32_t block_data[6 * 128];
    std::vector<int32_t> blue_noise;  /* sobol[65536] ++ scramble[131072] ++ ranking[131072] */
    TexArray tex[4];
    std::vector<float> sky;
    TexCube skymap;
    /* light propagation volume + BlockAverageColorData for ApproximateGILPV (borrowed; vxo_scene_set_lpv) */
    const uint8_t* lpv_level = nullptr;
    const uint8_t* lpv_type = nullptr;
    const float* lpv_avg512 = nullptr;
};

static const float PI = 3.14159265359f;

static inline int nthreads() { return vxo_get_threads(); }
static inline void tile_rows(const vxrt_tile& t, int height, int* r0, int* r1) {
    if (t.rows <= 0) { *r0 = 0; *r1 = height; }
    else { *r0 = t.row0; *r1 = t.row0 + t.rows; if (*r1 > height) *r1 = height; }
}
static inline v3 sat3(v3 a) { return a; }
static inline v3 xyz(v4 a) { return V3(a.x, a.y, a.z); }
static inline v3 abs3(v3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline bool eq3(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
static inline v2 fract2(float a, float b) { return V2(gfract(a), gfract(b)); }

extern "C" {

vxo_scene* vxo_scene_create(const vxo_world* w) {
    vxo_scene* s = new vxo_scene();
    s->world = *w;
    for (int i = 0; i < 6 * 128; ++i) s->block_data[i] = -1;
    return s;
}
void vxo_scene_destroy(vxo_scene* s) { delete s; }
void vxo_scene_set_lpv(vxo_scene* s, const uint8_t* level, const uint8_t* block_type, const float* avg512) {
    s->lpv_level = level; s->lpv_type = block_type; s->lpv_avg512 = avg512;
}
void vxo_scene_set_block_data(vxo_scene* s, const int32_t* t) { memcpy(s->block_data, t, sizeof(s->block_data)); }
void vxo_scene_set_blue_noise(vxo_scene* s, const int32_t* d, int32_t n) { s->blue_noise.assign(d, d + n); }
void vxo_scene_set_texture_array(vxo_scene* s, int32_t kind, int32_t layers, int32_t w, int32_t h, const uint8_t* rgba) {
    texarray_build(s->tex[kind], rgba, layers, w, h, kind == VXRT_TEX_ALBEDO);
}
void vxo_scene_set_skymap(vxo_scene* s, int32_t res, const float* f) {
    s->sky.assign(f, f + (size_t)6 * res * res * 3);
    s->skymap.data = s->sky.data();
    s->skymap.res = res;
}
int32_t vxo_scene_texture_level(const vxo_scene* s, int32_t kind, int32_t level, uint8_t* out, int64_t out_bytes) {
    if (kind < 0 || kind > 3 || level < 0 || level >= (int)s->tex[kind].levels.size()) return -1;
    const std::vector<uint8_t>& l = s->tex[kind].levels[level];
    if ((int64_t)l.size() != out_bytes) return -2;
    memcpy(out, l.data(), l.size());
    return 0;
}

/* PrecomputeAverageBlockColor.comp main() (:23-54), dispatched once by Volumetrics::CreateVolume (VolumetricFloodFill.cpp:102-123):
 * BlockAverageColorData[id] = pow((five samples of the block's albedo layer at LOD 8 + five at LOD 6 / 6.5 / 6.5 / 5.5 / 5.5) / 10, 1.8);
 * (0, 0, 0, 0) for ids without an albedo layer.  The samples sit at (0.5, 0.25, 0.75, 1.0, 0.0)^2 of the layer. */
void vxo_scene_lpv_average_colors(const vxo_scene* s, float* out512) {
    static const float at[5] = {0.5f, 0.25f, 0.75f, 1.0f, 0.0f};
    static const float lod2[5] = {6.0f, 6.5f, 6.5f, 5.5f, 5.5f};
    for (int id = 0; id < 128; ++id) {
        float* o = out512 + 4 * id;
        o[0] = o[1] = o[2] = o[3] = 0.0f;
        const int layer = s->block_data[id];   /* BlockAlbedoData */
        if (layer < 0) continue;
        v3 a = V3(0.0f, 0.0f, 0.0f), b = a;
        for (int k = 0; k < 5; ++k) {
            const v4 t = texarray_sample(s->tex[0], at[k], at[k], (float)layer, 8.0f);
            a = k == 0 ? V3(t.x, t.y, t.z) : V3(a.x + t.x, a.y + t.y, a.z + t.z);
        }
        for (int k = 0; k < 5; ++k) {
            const v4 t = texarray_sample(s->tex[0], at[k], at[k], (float)layer, lod2[k]);
            b = k == 0 ? V3(t.x, t.y, t.z) : V3(b.x + t.x, b.y + t.y, b.z + t.z);
        }
        const float d = 5.0f * 2.0f;
        o[0] = powf((a.x + b.x) / d, 1.8f); o[1] = powf((a.y + b.y) / d, 1.8f); o[2] = powf((a.z + b.z) / d, 1.8f);
    }
}

}  // extern "C"

/* ------------------------------------------------------------------------------------------------
 * shared shader helpers
 * ---------------------------------------------------------------------------------------------- */

/* GetRayDirectionAt — identical in every pass (e.g. GenerateGBuffer.glsl:110-115) */
static inline v3 ray_direction_at(const float* inv_view, const float* inv_proj, v2 ss) {
    v4 clip = V4(ss.x * 2.0f - 1.0f, ss.y * 2.0f - 1.0f, -1.0f, 1.0f);
    v4 e = mat4_mul(inv_proj, clip);
    v4 r = mat4_mul(inv_view, V4(e.x, e.y, -1.0f, 0.0f));
    return V3(r.x, r.y, r.z);
}
static inline v2 pixel_uv(int px, int py, int W, int H) { return V2(((float)px + 0.5f) / (float)W, ((float)py + 0.5f) / (float)H); }

/* GetNormalFromID: idx > 5 returns `miss` ((1,1,1), or (0.5,0.5,0.5) in the GI shader: DiffuseRayTraceFrag.glsl:790-801) */
static inline v3 normal_from_id(float n, v3 miss) {
    static const v3 N[6] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}};
    int i = cvt_round(n * 10.0f);
    if (i > 5) return miss;
    return N[i];
}
/* CompareVec3 with e = 0.0125 (GenerateGBuffer.glsl:125-128) */
static inline bool cmp3(v3 a, v3 b) {
    const float e = 0.0125f;
    return fabsf(a.x - b.x) < e && fabsf(a.y - b.y) < e && fabsf(a.z - b.z) < e;
}
/* CalculateVectors (GenerateGBuffer.glsl:444-513, DiffuseRayTraceFrag.glsl:1207-1270, ReflectionTraceFrag.glsl:1388-1449) */
static inline void calculate_vectors(v3 p, v3 n, v3* tangent, v3* bitangent, v2* uv) {
    static const v3 N[6] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}};
    static const v3 T[6] = {{1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {1, 0, 0}, {0, 0, -1}, {0, 0, -1}};
    static const v3 B[6] = {{0, 1, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {0, -1, 0}, {0, -1, 0}};
    for (int i = 0; i < 6; ++i)
        if (cmp3(n, N[i])) {
            *uv = (i < 2) ? fract2(p.x, p.y) : ((i < 4) ? fract2(p.x, p.z) : fract2(p.z, p.y));
            *tangent = T[i];
            *bitangent = B[i];
            return;
        }
}
/* CalculateUV (DiffuseRayTraceFrag.glsl:1321-1359): top/bottom xz, right/left zy, front/back xy */
static inline void calculate_uv(v3 p, v3 n, v2* uv) {
    if (cmp3(n, V3(0, 1, 0)) || cmp3(n, V3(0, -1, 0))) *uv = fract2(p.x, p.z);
    else if (cmp3(n, V3(1, 0, 0)) || cmp3(n, V3(-1, 0, 0))) *uv = fract2(p.z, p.y);
    else if (cmp3(n, V3(0, 0, 1)) || cmp3(n, V3(0, 0, -1))) *uv = fract2(p.x, p.y);
}
/* ------------------------------------------------------------------------------------------------
 * Alpha-tested traversal (off by default in the engine, Pipeline.cpp:146-147)
 * ---------------------------------------------------------------------------------------------- */
/* StopRay (InitialRayTraceFrag.glsl:189-203; ShadowRayTraceFrag.glsl:105-118 flips only v and biases the LOD by -2).
 * `block` is the raw texel byte; GetBlockID clamps it to 0..127.  An N that matches no axis (RaySign == 0 on MinIdx)
 * leaves uv undefined in the shader: pinned to (0, 0).                                                       */
static inline bool stop_ray(const vxo_scene* s, v3 P, v3 N, int block, v3 viewer, float g_K, bool shadow_variant) {
    const int id = iclamp(block, 0, 127);
    if (s->block_data[4 * 128 + id] == 0) return true;
    v2 uv = V2(0.0f, 0.0f);
    calculate_uv(P, N, &uv);
    uv.y = 1.0f - uv.y;
    if (!shadow_variant) uv.x = 1.0f - uv.x;
    float D = distance(P, viewer);
    int LOD = cvt_trunc(log2f(512.0f / (1.0f / D * g_K)));
    float lod = shadow_variant ? gclamp((float)LOD - 2.0f, 0.0f, 8.0f) : gclamp((float)LOD, 0.0f, 8.0f);
    float Alpha = texarray_sample(s->tex[VXRT_TEX_ALBEDO], uv.x, uv.y, (float)s->block_data[0 * 128 + id], lod).w;
    return Alpha > 0.975f;
}

/* one DDA step (InitialRayTraceFrag.glsl:233-243 == :268-282) */
static inline void alpha_dda_step(v3& origin, v3 direction, i3 RaySign, i3 Step01, int& MinIdx) {
    i3 G = {cvt_trunc(origin.x), cvt_trunc(origin.y), cvt_trunc(origin.z)};
    v3 W = origin - V3((float)G.x, (float)G.y, (float)G.z);
    v3 inv = V3(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    v3 DF = (V3((float)Step01.x, (float)Step01.y, (float)Step01.z) - W) * inv;
    MinIdx = (DF.x < DF.y && RaySign.x != 0) ? ((DF.x < DF.z || RaySign.z == 0) ? 0 : 2) : ((DF.y < DF.z || RaySign.z == 0) ? 1 : 2);
    idx(G, MinIdx) += idx(RaySign, MinIdx);
    W = W + direction * idx(DF, MinIdx);
    idx(W, MinIdx) = (float)(1 - idx(Step01, MinIdx));
    origin = V3((float)G.x, (float)G.y, (float)G.z) + W;
    idx(origin, MinIdx) += (float)idx(RaySign, MinIdx) * 0.0001f;
}

extern "C" float vxo_traverse_alpha(const vxo_scene* s, const float origin0[3], const float dir[3], int32_t max_iter,
                                    const float viewer3[3], float g_K, int32_t shadow_variant, vxo_hit* hit) {
    const vxo_world* w = &s->world;
    const v3 viewer = V3(viewer3[0], viewer3[1], viewer3[2]);
    const v3 initial_origin = V3(origin0[0], origin0[1], origin0[2]);
    v3 origin = initial_origin;
    const v3 direction = V3(dir[0], dir[1], dir[2]);
    bool Intersection = false;
    int MinIdx = 0;
    const i3 RaySign = {gsign(direction.x), gsign(direction.y), gsign(direction.z)};
    const i3 Step01 = {(1 + RaySign.x) >> 1, (1 + RaySign.y) >> 1, (1 + RaySign.z) >> 1};
    int iters = 0, dda = 0;
    float t = -1.0f;
    int block = 0;
    v3 normal = V3(0.0f);
    bool returned = false;

    for (int itr = 0; itr < max_iter && !returned; ++itr) {
        int lx = cvt_floor(origin.x), ly = cvt_floor(origin.y), lz = cvt_floor(origin.z);
        if (!in_volume(w, lx, ly, lz)) {
            Intersection = false;
            break;
        }
        iters++;
        int k = w->df[lx + (size_t)ly * w->nx + (size_t)lz * w->nx * w->ny];
        int Euclidean = euclidean_step(k);
        if (Euclidean == 0) {
            v3 tn = V3(0.0f);
            idx(tn, MinIdx) = (float)(-idx(RaySign, MinIdx));
            int bt = get_voxel(w, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
            if (stop_ray(s, origin, tn, bt, viewer, g_K, shadow_variant != 0)) break;
            for (int i = 0; i < 4; ++i) {
                alpha_dda_step(origin, direction, RaySign, Step01, MinIdx);
                dda++;
                int b2 = get_voxel(w, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
                if (b2 > 0) {
                    v3 tn2 = V3(0.0f);
                    idx(tn2, MinIdx) = (float)(-idx(RaySign, MinIdx));
                    if (stop_ray(s, origin, tn2, b2, viewer, g_K, shadow_variant != 0)) {
                        normal = V3(0.0f);
                        idx(normal, MinIdx) = (float)(-idx(RaySign, MinIdx));
                        block = b2;
                        t = block > 0 ? distance(origin, initial_origin) : -1.0f;
                        returned = true;  /* `return` inside the loop (:251-256) */
                        break;
                    }
                }
            }
            if (returned) break;
        }
        if (Euclidean == 1) {
            alpha_dda_step(origin, direction, RaySign, Step01, MinIdx);
            dda++;
            Intersection = true;
        } else {
            /* also taken after an unresolved Euclidean == 0 block: int(0 - 1) * direction steps one unit BACK (:285-288) */
            origin = origin + (float)(Euclidean - 1) * direction;
        }
    }
    if (!returned && Intersection) {
        normal = V3(0.0f);
        idx(normal, MinIdx) = (float)(-idx(RaySign, MinIdx));
        block = get_voxel(w, cvt_floor(origin.x), cvt_floor(origin.y), cvt_floor(origin.z));
        t = block > 0 ? distance(origin, initial_origin) : -1.0f;
    }
    if (hit) {
        hit->t = t;
        hit->normal[0] = normal.x; hit->normal[1] = normal.y; hit->normal[2] = normal.z;
        hit->end[0] = origin.x; hit->end[1] = origin.y; hit->end[2] = origin.z;
        hit->block = block;
        hit->intersection = (returned || Intersection) ? 1 : 0;
        hit->min_idx = MinIdx;
        hit->iterations = iters;
        hit->dda_steps = dda;
    }
    return t;
}

/* BasicSaturation (ColorPassFrag.glsl:1228-1233) */
static inline v3 basic_saturation(v3 c, float adj) {
    float l = dot(c, V3(0.2125f, 0.7154f, 0.0721f));
    return gmix(V3(l), c, adj);
}

/* attachment views */
static inline Tex2D view_f32(const float* d, int w, int h, int ch, bool linear) { Tex2D t; t.data = d; t.w = w; t.h = h; t.ch = ch; t.linear = linear; return t; }
static std::vector<float> half_to_f32(const uint16_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = half_to_float(h[i]); return o; }
static std::vector<float> u8_to_f32(const uint8_t* h, size_t n) { std::vector<float> o(n); for (size_t i = 0; i < n; ++i) o[i] = unorm8_to_float(h[i]); return o; }

/* ------------------------------------------------------------------------------------------------
 * GenerateGBuffer.glsl
 * ---------------------------------------------------------------------------------------------- */

/* GetTextureIDs (GenerateGBuffer.glsl:522-578) */
static inline v4 gbuffer_texture_ids(const vxo_scene* s, const vxrt_gbuffer_params* p, int id, v3 n) {
    const int32_t* bd = s->block_data;
    v4 d = V4((float)bd[0 * 128 + id], (float)bd[1 * 128 + id], (float)bd[2 * 128 + id], (float)bd[3 * 128 + id]);
    const v3 TOP = V3(0, 1, 0), BOTTOM = V3(0, -1, 0), FRONT = V3(0, 0, 1), BACK = V3(0, 0, -1), LEFT = V3(-1, 0, 0), RIGHT = V3(1, 0, 0);
    const int32_t* sets[2] = {p->grass_props, p->cactus_props};
    for (int k = 0; k < 2; ++k) {
        const int32_t* q = sets[k];
        if (id == q[0]) {
            if (eq3(n, LEFT) || eq3(n, RIGHT) || eq3(n, FRONT) || eq3(n, BACK)) { d.x = (float)q[4]; d.y = (float)q[5]; d.z = (float)q[6]; }
            else if (eq3(n, TOP)) { d.x = (float)q[1]; d.y = (float)q[2]; d.z = (float)q[3]; }
            else if (eq3(n, BOTTOM)) { d.x = (float)q[7]; d.y = (float)q[8]; d.z = (float)q[9]; }
        }
    }
    return d;
}

extern "C" void vxo_generate_gbuffer(const vxo_scene* s, const vxrt_gbuffer_params* p, const float* g_inv_t, const uint8_t* g_normal,
                                     const uint8_t* g_block, int32_t gw, int32_t gh, uint16_t* albedo_h3, uint16_t* normal_h3,
                                     uint8_t* pbr_u8x4, uint8_t* texao_u8) {
    const int W = p->width, H = p->height;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
    std::vector<float> nf = u8_to_f32(g_normal, (size_t)gw * gh), bf = u8_to_f32(g_block, (size_t)gw * gh);
    const Tex2D tInvT = view_f32(g_inv_t, gw, gh, 1, true);      /* R32F LINEAR (Pipeline.cpp:1142) */
    const Tex2D tN = view_f32(nf.data(), gw, gh, 1, false), tB = view_f32(bf.data(), gw, gh, 1, false);
    const v3 cam = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads())
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t i = (size_t)py * W + px;
            v2 tc = pixel_uv(px, py, W, H);
            /* GetBlockID (:96-100) */
            int BaseID = iclamp(cvt_floor(tex2d_sample(tB, tc.x, tc.y).x * 255.0f), 0, 127);
            /* GetPositionAt (:117-122) */
            float Dist = 1.0f / tex2d_sample(tInvT, tc.x, tc.y).x;
            v3 P = cam + normalize(ray_direction_at(p->inv_view, p->inv_projection, tc)) * Dist;
            v3 oA, oN; v4 oP; float oAO;
            if (Dist < 0.0f) {
                oA = V3(0.0f); oN = V3(1.0f); oP = V4(0, 0, 0, 0); oAO = 0.0f;
            } else {
                v3 FlatNormal = normal_from_id(tex2d_sample(tN, tc.x, tc.y).x, V3(1.0f));
                v4 data = gbuffer_texture_ids(s, p, BaseID, FlatNormal);
                v2 UV = V2(1.0f, 1.0f), tUV = V2(1.0f, 1.0f);
                v3 T = V3(0.0f), B = V3(0.0f), tT, tB2;
                calculate_vectors(P, FlatNormal, &T, &B, &UV);
                calculate_vectors(P, abs3(FlatNormal), &tT, &tB2, &tUV);
                /* Parallax() with u_POM == false returns FlatUV (:338-349); lava path disabled */
                UV = tUV;
                UV = V2(1.0f - UV.x, 1.0f - UV.y);
                v3 nm = xyz(texarray_sample(s->tex[VXRT_TEX_NORMAL], UV.x, UV.y, data.y, 0.0f));
                nm = nm * 2.0f - V3(1.0f);
                nm = mat3_mul(T, B, FlatNormal, nm);
                v4 PBRMap = texarray_sample(s->tex[VXRT_TEX_PBR], UV.x, UV.y, data.z, 0.0f);
                float Emissivity = data.w > -0.5f ? texarray_sample(s->tex[VXRT_TEX_EMISSIVE], UV.x, UV.y, data.w, 0.0f).x : 0.0f;
                oN = nm;
                oP = V4(gclamp(PBRMap.x, 0.0f, 1.0f), gclamp(PBRMap.y, 0.0f, 1.0f), gclamp(PBRMap.z, 0.0f, 1.0f), gclamp(Emissivity, 0.0f, 1.0f));
                oAO = gclamp(PBRMap.w, 0.00000001f, 1.0f);
                oA = xyz(texarray_sample(s->tex[VXRT_TEX_ALBEDO], UV.x, UV.y, data.x, 0.0f));
                const float lb = 0.02f;
                oP.w *= (UV.x > lb && UV.x < 1.0f - lb && UV.y > lb && UV.y < 1.0f - lb) ? 1.0f : 0.0f;
            }
            albedo_h3[3 * i] = float_to_half(oA.x); albedo_h3[3 * i + 1] = float_to_half(oA.y); albedo_h3[3 * i + 2] = float_to_half(oA.z);
            normal_h3[3 * i] = float_to_half(oN.x); normal_h3[3 * i + 1] = float_to_half(oN.y); normal_h3[3 * i + 2] = float_to_half(oN.z);
            pbr_u8x4[4 * i] = float_to_unorm8(oP.x); pbr_u8x4[4 * i + 1] = float_to_unorm8(oP.y);
            pbr_u8x4[4 * i + 2] = float_to_unorm8(oP.z); pbr_u8x4[4 * i + 3] = float_to_unorm8(oP.w);
            texao_u8[i] = float_to_unorm8(oAO);
        }
}

/* ------------------------------------------------------------------------------------------------
 * Cook-Torrance direct term
 * ---------------------------------------------------------------------------------------------- */

/* ndfGGX / gaSchlickG1 / gaSchlickGGX / FresnelSchlickRoughness / CalculateDirectionalLight
 * (ColorPassFrag.glsl:394-451, 1201-1210) */
static inline float ndf_ggx(float cosLh, float roughness) {
    float alpha = roughness * roughness;
    float alphaSq = alpha * alpha;
    float denom = (cosLh * cosLh) * (alphaSq - 1.0f) + 1.0f;
    return alphaSq / (PI * denom * denom);
}
static inline float ga_schlick_g1(float c, float k) { return c / (c * (1.0f - k) + k); }
static inline float ga_schlick_ggx(float cosLi, float cosLo, float roughness) {
    float r = roughness + 1.0f;
    float k = (r * r) / 8.0f;
    return ga_schlick_g1(cosLi, k) * ga_schlick_g1(cosLo, k);
}
static inline v3 fresnel_schlick_roughness(v3 Eye, v3 norm, v3 F0, float roughness) {
    float cosTheta = gclamp(dot(Eye, norm), 0.00001f, 1.0f);
    float pw = powf(1.0f - cosTheta, 5.0f);
    v3 m = V3(gmax(1.0f - roughness, F0.x), gmax(1.0f - roughness, F0.y), gmax(1.0f - roughness, F0.z));
    return F0 + (m - F0) * pw;
}
static inline v3 color_directional_light(v3 viewer, v3 world_pos, v3 light_dir, v3 radiance, v3 radiance_s, v3 albedo, v3 normal,
                                         v3 pbr, float shadow) {
    const float Epsilon = 0.00001f;
    float Shadow = gmin(shadow, 1.0f);
    v3 Lo = normalize(viewer - world_pos);
    v3 N = normal;
    float cosLo = gmax(0.0f, dot(N, Lo));
    v3 F0 = gmix(V3(0.04f), albedo, pbr.y);
    v3 Li = light_dir;
    v3 Lh = normalize(Li + Lo);
    float cosLi = gmax(0.0f, dot(N, Li));
    float cosLh = gmax(0.0f, dot(N, Lh));
    v3 F = fresnel_schlick_roughness(Lo, normal, F0, pbr.x);
    float D = ndf_ggx(cosLh, pbr.x);
    float G = ga_schlick_ggx(cosLi, cosLo, pbr.x);
    v3 kd = gmix(V3(1.0f) - F, V3(0.0f), pbr.y);
    v3 diffuseBRDF = kd * albedo;
    v3 specularBRDF = (F * D * G) / gmax(Epsilon, 4.0f * cosLi * cosLo);
    specularBRDF = gclamp(specularBRDF, 0.0f, 2.0f);
    v3 Result = (diffuseBRDF * radiance * cosLi) + (specularBRDF * radiance_s * cosLi);
    return gclamp(Result, 0.0f, 2.5f) * gclamp(1.0f - Shadow, 0.0f, 1.0f);
}

extern "C" void vxo_shade_direct(const vxrt_direct_params* p, const float* g_inv_t, int32_t gw, int32_t gh, const uint16_t* albedo_h3,
                                 const uint16_t* normal_h3, const uint8_t* pbr_u8x4, const uint8_t* texao_u8, int32_t mw, int32_t mh,
                                 const uint8_t* shadow_u8, int32_t sw, int32_t sh, uint16_t* direct_h3) {
    const int W = p->width, H = p->height;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
    std::vector<float> af = half_to_f32(albedo_h3, (size_t)mw * mh * 3), nf = half_to_f32(normal_h3, (size_t)mw * mh * 3);
    std::vector<float> pf = u8_to_f32(pbr_u8x4, (size_t)mw * mh * 4), aof = u8_to_f32(texao_u8, (size_t)mw * mh), sf = u8_to_f32(shadow_u8, (size_t)sw * sh);
    const Tex2D tInvT = view_f32(g_inv_t, gw, gh, 1, true);
    const Tex2D tA = view_f32(af.data(), mw, mh, 3, true), tN = view_f32(nf.data(), mw, mh, 3, true);        /* RGB16F LINEAR (Pipeline.cpp:1148) */
    const Tex2D tP = view_f32(pf.data(), mw, mh, 4, false), tAO = view_f32(aof.data(), mw, mh, 1, false);    /* RGBA8 / R8 NEAREST */
    const Tex2D tS = view_f32(sf.data(), sw, sh, 1, true);                                                   /* R8 LINEAR (Pipeline.cpp:1200) */
    const v3 cam = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    const v3 viewer = V3(p->viewer_position[0], p->viewer_position[1], p->viewer_position[2]);
    const v3 sun = V3(p->sun_direction[0], p->sun_direction[1], p->sun_direction[2]), moon = V3(p->moon_direction[0], p->moon_direction[1], p->moon_direction[2]);
    const v3 SunColor = V3(p->sun_color[0], p->sun_color[1], p->sun_color[2]), MoonColor = V3(p->moon_color[0], p->moon_color[1], p->moon_color[2]);
    /* ColorPassFrag.glsl:776 */
    float SunVisibility = gclamp(dot(sun, V3(0.0f, 1.0f, 0.0f)) + 0.05f, 0.0f, 0.1f) * 12.0f;
    SunVisibility = 1.0f - SunVisibility;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads())
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t i = (size_t)py * W + px;
            v2 tc = pixel_uv(px, py, W, H);
            float Dist = 1.0f / tex2d_sample(tInvT, tc.x, tc.y).x; /* SamplePositionAt (:1123-1127) */
            v3 P = cam + normalize(ray_direction_at(p->inv_view, p->inv_projection, tc)) * Dist;
            v3 out = V3(0.0f);
            if (Dist > 0.0f) {
                v3 Albedo = xyz(tex2d_sample(tA, tc.x, tc.y));
                v3 N = xyz(tex2d_sample(tN, tc.x, tc.y));
                v4 PBRMap = tex2d_sample(tP, tc.x, tc.y);
                float Emissivity = PBRMap.w;
                Albedo = basic_saturation(Albedo, 1.0f - p->texture_desat_amount);      /* :812 */
                if (PBRMap.y >= 0.1f - 0.01f) Albedo = basic_saturation(Albedo, 0.9f);    /* :814-816 */
                if (p->amplify_normal_map) {                                              /* :820-825 */
                    N.x *= 1.64f; N.z *= 1.85f;
                    N = N + V3(1e-4f);
                    N = normalize(N);
                }
                float shadow = gclamp(tex2d_sample(tS, tc.x, tc.y).x, 0.0f, 1.0f);
                v3 pbr = V3(PBRMap.x, PBRMap.y, PBRMap.z);
                v3 SunDirect = color_directional_light(viewer, P, sun, SunColor, SunColor, Albedo, N, pbr, shadow);     /* :893 */
                v3 MoonDirect = color_directional_light(viewer, P, moon, MoonColor, MoonColor, Albedo, N, pbr, shadow); /* :894 */
                v3 Direct = V3(gmix(SunDirect.x, MoonDirect.x, SunVisibility * 1.0f), gmix(SunDirect.y, MoonDirect.y, SunVisibility * 1.0f),
                               gmix(SunDirect.z, MoonDirect.z, SunVisibility * 1.0f));
                Direct = ((!(Emissivity > 0.05f)) ? 1.0f : 0.0f) * Direct;                 /* :897 */
                out = gmax(Direct, 0.000001f);                                            /* :899 */
            }
            direct_h3[3 * i] = float_to_half(out.x); direct_h3[3 * i + 1] = float_to_half(out.y); direct_h3[3 * i + 2] = float_to_half(out.z);
        }
}

/* ------------------------------------------------------------------------------------------------
 * blue noise + sun colour (shared by GI and reflections)
 * ---------------------------------------------------------------------------------------------- */

/* samplerBlueNoiseErrorDistribution_128x128_OptimizedFor_2d2d2d2d_32spp (DiffuseRayTraceFrag.glsl:130-153).
 * The SSBO is sobol ++ scramble ++ ranking; reads past the end of rankingTile return 0 (pinned). */
static inline float blue_noise_1d(const vxo_scene* s, int px, int py, int sampleIndex, int sampleDimension) {
    const int32_t* sobol = s->blue_noise.data();
    const int32_t* scramble = sobol + 256 * 256;
    const int32_t* ranking = scramble + 128 * 128 * 8;
    int pi = px & 127, pj = py & 127;
    sampleIndex &= 255;
    sampleDimension &= 255;
    int ri = sampleDimension + (pi + pj * 128) * 8;
    int rank = (ri < 128 * 128 * 8) ? ranking[ri] : 0;
    int rankedSampleIndex = sampleIndex ^ rank;
    int si = sampleDimension + rankedSampleIndex * 256;
    int value = (si >= 0 && si < 256 * 256) ? sobol[si] : 0;
    value = value ^ scramble[(sampleDimension % 8) + (pi + pj * 128) * 8];
    return (0.5f + (float)value) / 256.0f;
}

/* SRGBToLinear / TemperatureToRGB (DiffuseRayTraceFrag.glsl:822-825, 862-886) */
static inline float srgb_to_linear_f(float x) { return x > 0.04045f ? powf(x * (1.0f / 1.055f) + 0.0521327f, 2.4f) : x / 12.92f; }
static inline v3 temperature_to_rgb(float K) {
    v3 c;
    K = gclamp(K, 1000.0f, 50000.0f) / 100.0f;
    if (K <= 66.0f) {
        c.x = 1.0f;
        c.y = gclamp(0.39008157876901960784f * logf(K) - 0.63184144378862745098f, 0.0f, 1.0f);
    } else {
        float t = K - 60.0f;
        c.x = gclamp(1.29293618606274509804f * powf(t, -0.1332047592f), 0.0f, 1.0f);
        c.y = gclamp(1.12989086089529411765f * powf(t, -0.0755148492f), 0.0f, 1.0f);
    }
    if (K >= 66.0f) c.z = 1.0f;
    else if (K <= 19.0f) c.z = 0.0f;
    else c.z = gclamp(0.54320678911019607843f * logf(K - 10.0f) - 1.19625408914f, 0.0f, 1.0f);
    return V3(srgb_to_linear_f(c.x), srgb_to_linear_f(c.y), srgb_to_linear_f(c.z));
}
/* SampleSunColor (DiffuseRayTraceFrag.glsl:901-908, ReflectionTraceFrag.glsl:649-656) */
static inline v3 sample_sun_color(const vxo_scene* s, v3 sun, float strength) {
    v3 c = xyz(texcube_sample(s->skymap, sun.x, sun.y, sun.z));
    c = c * temperature_to_rgb(5778.0f);
    return c * PI * 2.2f * strength;
}

/* ------------------------------------------------------------------------------------------------
 * DiffuseRayTraceFrag.glsl
 * ---------------------------------------------------------------------------------------------- */
namespace {

struct GiCtx {
    const vxo_scene* s;
    const vxrt_gi_params* p;
    int px, py;
    int CurrentBLSample;
    v3 LIGHT_COLOR, StrongerLightDirection;
    bool Moonstronger;
    float EmissivityMultiplier;
    uint64_t rays, iters, dda, hits;
};

inline float trace(GiCtx& c, v3 o, v3 d, int max_iter, vxo_hit* h) {
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    float t = vxo_traverse(&c.s->world, oo, dd, max_iter, h);
    c.rays++; c.iters += h->iterations; c.dda += h->dda_steps; c.hits += t > 0.0f ? 1 : 0;
    return t;
}

/* SampleBlueNoise2D (:807-820) */
inline v2 sample_blue_noise_2d(GiCtx& c, int Index) {
    v2 n;
    n.x = blue_noise_1d(c.s, c.px, c.py, Index, 1 + c.CurrentBLSample);
    n.y = blue_noise_1d(c.s, c.px, c.py, Index, 2 + c.CurrentBLSample);
    c.CurrentBLSample += 2;
    return n;
}
/* cosWeightedRandomHemisphereDirection (:1031-1053) */
inline v3 cos_weighted_hemisphere(GiCtx& c, v3 n) {
    v2 r = sample_blue_noise_2d(c, c.p->current_frame_mod128);
    float PI2 = 2.0f * PI;
    v3 uu = normalize(cross(n, V3(0.0f, 1.0f, 1.0f)));
    v3 vv = cross(uu, n);
    float ra = sqrtf(r.y);
    float rx = ra * cosf(PI2 * r.x);
    float ry = ra * sinf(PI2 * r.x);
    float rz = sqrtf(1.0f - r.y);
    v3 rr = rx * uu + ry * vv + rz * n;
    return normalize(rr);
}
/* InverseSchlick / DiffuseHammon (:1397-1418) */
inline float inverse_schlick(float f0, float VoH) { return 1.0f - gclamp(f0 + (1.0f - f0) * powf(1.0f - VoH, 5.0f), 0.0f, 1.0f); }
inline float diffuse_hammon(v3 normal, v3 viewDir, v3 lightDir, float roughness) {
    float nDotL = gmax(dot(normal, lightDir), 0.0f);
    if (nDotL <= 0.0f) return 0.0f;
    float nDotV = gmax(dot(normal, viewDir), 0.0f);
    float lDotV = gmax(dot(lightDir, viewDir), 0.0f);
    v3 halfWay = normalize(viewDir + lightDir);
    float nDotH = gmax(dot(normal, halfWay), 0.0f);
    float facing = lDotV * 0.5f + 0.5f;
    float singleRough = facing * (0.9f - 0.4f * facing) * ((0.5f + nDotH) * (1 / gmax(nDotH, 0.02f)));
    float singleSmooth = 1.05f * inverse_schlick(0.0f, nDotL) * inverse_schlick(0.0f, gmax(nDotV, 0.0f));
    float single = gclamp(gmix(singleSmooth, singleRough, roughness) * (1 / PI), 0.0f, 1.0f);
    float multi = 0.1159f * roughness;
    return gclamp((multi + single) * nDotL, 0.0f, 1.0f);
}
/* GetSkyColorAt (:1070-1074) */
inline v3 sky_color_at(const vxo_scene* s, v3 rd) {
    rd.y = gclamp(rd.y, 0.125f, 1.5f);
    return xyz(texcube_sample(s->skymap, rd.x, rd.y, rd.z));
}
/* RayBoxIntersect (:1274-1288) */
inline bool ray_box_intersect(v3 boxMin, v3 boxMax, v3 r0, v3 rD) {
    v3 inv = V3(1.0f / rD.x, 1.0f / rD.y, 1.0f / rD.z);
    v3 tbot = inv * (boxMin - r0), ttop = inv * (boxMax - r0);
    v3 tmin = V3(gmin(ttop.x, tbot.x), gmin(ttop.y, tbot.y), gmin(ttop.z, tbot.z));
    v3 tmax = V3(gmax(ttop.x, tbot.x), gmax(ttop.y, tbot.y), gmax(ttop.z, tbot.z));
    float t0 = gmax(gmax(tmin.x, tmin.y), gmax(tmin.x, tmin.z));
    float t1 = gmin(gmin(tmax.x, tmax.y), gmin(tmax.x, tmax.z));
    return t1 > gmax(t0, 0.0f);
}
/* GetShadowAt (:1290-1308) */
inline float gi_shadow_at(GiCtx& c, v3 pos, v3 ldir) {
    if (c.p->apply_player_shadow) {
        v3 vp = V3(c.p->viewer_position[0], c.p->viewer_position[1], c.p->viewer_position[2]);
        if (ray_box_intersect(vp + V3(0.2f, 0.0f, 0.2f), vp - V3(0.75f, 1.75f, 0.75f), pos, ldir)) return 1.0f;
    }
    vxo_hit h;
    float T = trace(c, pos, ldir, c.p->shadow_trace_length, &h);
    return T > 0.0f ? 1.0f : 0.0f;
}

/* CalculateDiffuse (:535-664) */
v4 calculate_diffuse(GiCtx& c, v3 initial_origin, v3 input_normal, v3* odir, bool* Skyhit) {
    const vxo_scene* s = c.s;
    *Skyhit = false;
    const float bias = 0.06f;
    v3 rayO = initial_origin + input_normal * bias;
    v3 rayD = cos_weighted_hemisphere(c, input_normal);
    float ao = 1.0f;
    v3 RayContribution = V3(0.0f), RayThroughput = V3(1.0f);
    for (int i = 0; i < 2; ++i) {  /* MAX_BOUNCE_LIMIT 2 (:19) */
        if (i == 0) *odir = rayD;
        vxo_hit h;
        float T = trace(c, rayO, rayD, c.p->trace_length, &h);
        v3 HitNormal = V3(h.normal[0], h.normal[1], h.normal[2]);
        int tex_ref = iclamp(h.block, 0, 127); /* clamp(int(floor(HitBlock*255)),0,127): (b/255.0f)*255.0f == b */
        bool Intersect = T > 0.0f;
        v3 IntersectionPosition = rayO + (rayD * T);
        if (Intersect && h.block > 0) {
            v2 txc = V2(0.0f, 0.0f);
            calculate_uv(IntersectionPosition, HitNormal, &txc);
            float TexA = (float)s->block_data[0 * 128 + tex_ref], TexE = (float)s->block_data[3 * 128 + tex_ref];
            v3 Albedo = xyz(texarray_sample(s->tex[VXRT_TEX_ALBEDO], txc.x, txc.y, TexA, 3.0f));
            v3 PBR = xyz(texarray_sample(s->tex[VXRT_TEX_PBR], txc.x, txc.y, TexA, 2.0f)); /* sic: albedo layer id (:578) */
            float Emmisivity = 0.0f;
            if (TexE >= 0.0f) {
                float SampledEmmisivity = texarray_sample(s->tex[VXRT_TEX_EMISSIVE], txc.x, txc.y, TexE, 0.0f).x;
                Emmisivity = SampledEmmisivity * c.EmissivityMultiplier * c.p->diffuse_light_intensity;
            }
            float NDotL = gmax(dot(HitNormal, c.StrongerLightDirection), 0.0f);
            v3 bias_shadow = HitNormal * 0.045f;
            float ShadowAt;
            if (c.Moonstronger) ShadowAt = 1.0f;
            else ShadowAt = (NDotL < 0.001f) ? 0.0f : gi_shadow_at(c, IntersectionPosition + bias_shadow, c.StrongerLightDirection);
            v3 EmmisivityColor = (Emmisivity * gmix(1.0f, 1.0f, c.p->sun_visibility)) * Albedo;
            /* SunBRDF (:499-503), CAUSTICS == false */
            v3 SUNBRDF = Albedo * diffuse_hammon(HitNormal, -rayD, c.StrongerLightDirection, PBR.x) * (c.LIGHT_COLOR * 3.5f) * (1.0f - ShadowAt) * PI;
            v3 NewDirection = cos_weighted_hemisphere(c, HitNormal);
            float CosTheta = gclamp(dot(HitNormal, NewDirection), 0.0f, 1.0f);
            float PDF = gmax(CosTheta / PI, 0.00001f);
            v3 Attenuation = V3(1.0f) * diffuse_hammon(HitNormal, -rayD, NewDirection, PBR.x); /* DiffuseRayBRDF (:527-531) */
            RayContribution = RayContribution + RayThroughput * SUNBRDF;
            RayContribution = RayContribution + EmmisivityColor * RayThroughput;
            RayThroughput = RayThroughput * (Albedo * Attenuation / PDF);
            rayD = NewDirection;
            rayO = IntersectionPosition + HitNormal * bias;
        } else {
            float x = gmix(1.0f, 1.05f, c.p->sun_visibility);
            x = gclamp(x * 1.0f * c.p->gi_sky_strength, 0.0f, 5.0f);
            v3 sky = sky_color_at(s, rayD) * x;
            RayContribution = RayContribution + sky * RayThroughput;
            *Skyhit = true;
            break;
        }
        if (i == 0) {
            const float dao = 2.0f;
            if (T < dao && T > 0.0f) ao = gmax(T / dao, 0.0f);
        }
    }
    return V4(RayContribution.x, RayContribution.y, RayContribution.z, ao);
}

/* IrridianceToSH (:766-784) */
inline void irradiance_to_sh(v3 Radiance, v3 Direction, float* o) {
    float Co = Radiance.x - Radiance.z;
    float T = Radiance.z + Co * 0.5f;
    float Cg = Radiance.y - T;
    float Y = gmax(T + Cg * 0.5f, 0.0f);
    float L00 = 0.282095f;
    float L1_1 = 0.488603f * Direction.y, L10 = 0.488603f * Direction.z, L11 = 0.488603f * Direction.x;
    o[0] = gmax(L11 * Y, -100.0f); o[1] = gmax(L1_1 * Y, -100.0f); o[2] = gmax(L10 * Y, -100.0f); o[3] = gmax(L00 * Y, -100.0f);
    o[4] = Co; o[5] = Cg;
}

}  // namespace

extern "C" void vxo_diffuse_trace(const vxo_scene* s, const vxrt_gi_params* p, const uint16_t* g_t_half, const uint8_t* g_normal,
                                  int32_t gw, int32_t gh, uint16_t* sh_h4, uint16_t* cocg_h2, uint16_t* utility_h, uint8_t* aosky_u8x2,
                                  vxrt_trace_stats* stats) {
    const int W = p->width, H = p->height;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
    std::vector<float> tf = half_to_f32(g_t_half, (size_t)gw * gh), nf = u8_to_f32(g_normal, (size_t)gw * gh);
    const Tex2D tT = view_f32(tf.data(), gw, gh, 1, true), tN = view_f32(nf.data(), gw, gh, 1, false);
    const v3 cam = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    const v3 sun = V3(p->sun_direction[0], p->sun_direction[1], p->sun_direction[2]), moon = V3(p->moon_direction[0], p->moon_direction[1], p->moon_direction[2]);
    /* main() prologue (:910-925) */
    const bool SunStronger = -sun.y < 0.01f;
    const v3 LIGHT_COLOR = SunStronger ? sample_sun_color(s, sun, p->gi_sun_strength) : V3(1.0f);
    uint64_t s_rays = 0, s_it = 0, s_dda = 0, s_hits = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(nthreads()) reduction(+ : s_rays, s_it, s_dda, s_hits)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t i = (size_t)py * W + px;
            GiCtx c;
            c.s = s; c.p = p; c.px = px; c.py = py; c.CurrentBLSample = 0;
            c.LIGHT_COLOR = LIGHT_COLOR; c.StrongerLightDirection = SunStronger ? sun : moon; c.Moonstronger = !SunStronger;
            c.EmissivityMultiplier = 12.0f;
            c.rays = c.iters = c.dda = c.hits = 0;
            v2 tc = pixel_uv(px, py, W, H);
            if (p->supersample) tc = tc + (V2(p->halton[0], p->halton[1]) * 0.75f) / V2((float)W, (float)H);
            float Dist = tex2d_sample(tT, tc.x, tc.y).x;
            v3 P = cam + normalize(ray_direction_at(p->inv_view, p->inv_projection, tc)) * Dist;
            v3 Normal = normal_from_id(tex2d_sample(tN, tc.x, tc.y).x, V3(0.5f));
            float oSH[4], oCoCg[2], oUtil = 0.0f, oAO = 1.0f, oSky = 0.0f;
            if (Dist < 0.0f) {
                /* v_RayDirection is the interpolated FBOVert.glsl output == GetRayDirectionAt(v_TexCoords) */
                v3 rd = normalize(ray_direction_at(p->inv_view, p->inv_projection, pixel_uv(px, py, W, H)));
                float SH[6];
                irradiance_to_sh(xyz(texcube_sample(s->skymap, rd.x, rd.y, rd.z)) * 2.66f, Normal, SH);
                oSH[0] = SH[0]; oSH[1] = SH[1]; oSH[2] = SH[2]; oSH[3] = SH[3]; oCoCg[0] = SH[4]; oCoCg[1] = SH[5];
            } else {
                int SPP = iclamp(p->spp, 1, 32);
                if (p->checkerboard) {
                    bool CheckerStep = cvt_trunc(((float)px + 0.5f) + ((float)py + 0.5f)) % 2 == p->current_frame % 2;
                    SPP = cvt_trunc(gmix((float)p->spp, (float)p->checker_spp, CheckerStep ? 1.0f : 0.0f));
                }
                SPP = iclamp(SPP, 1, 32);
                if (c.Moonstronger) SPP *= 2;
                v4 TotalSHy = V4(0, 0, 0, 0);
                v2 CoCg = V2(0, 0);
                v3 radiance = V3(0.0f);
                float Skyhits = 0.0f, AccumulatedAO = 0.0f;
                for (int sidx = 0; sidx < SPP; ++sidx) {
                    v3 d = V3(0.0f);
                    bool ss = false;
                    v4 x = calculate_diffuse(c, P, Normal, &d, &ss);
                    v3 xc = gclamp(V3(x.x, x.y, x.z), 0.0f, 8.0f);
                    radiance = radiance + xc;
                    AccumulatedAO += x.w;
                    float SH[6];
                    irradiance_to_sh(xc, d, SH);
                    TotalSHy = V4(TotalSHy.x + SH[0], TotalSHy.y + SH[1], TotalSHy.z + SH[2], TotalSHy.w + SH[3]);
                    CoCg = CoCg + V2(SH[4], SH[5]);
                    Skyhits += ss ? 1.0f : 0.0f;
                }
                const float n = (float)SPP;
                AccumulatedAO /= n;
                TotalSHy = V4(TotalSHy.x / n, TotalSHy.y / n, TotalSHy.z / n, TotalSHy.w / n);
                CoCg = V2(CoCg.x / n, CoCg.y / n);
                radiance = radiance / n;
                Skyhits /= n;
                oUtil = gmax(dot(radiance, V3(0.299f, 0.587f, 0.114f)), 0.01f);
                oAO = gclamp(AccumulatedAO, 0.0f, 1.0f);
                oSky = gclamp(Skyhits, 0.0f, 1.0f);
                oSH[0] = gclamp(TotalSHy.x, -100.0f, 100.0f); oSH[1] = gclamp(TotalSHy.y, -100.0f, 100.0f);
                oSH[2] = gclamp(TotalSHy.z, -100.0f, 100.0f); oSH[3] = gclamp(TotalSHy.w, -100.0f, 100.0f);
                oCoCg[0] = gclamp(CoCg.x, -100.0f, 100.0f); oCoCg[1] = gclamp(CoCg.y, -100.0f, 100.0f);
                oUtil = gclamp(oUtil, 0.001f, 64.0f);
            }
            for (int k = 0; k < 4; ++k) sh_h4[4 * i + k] = float_to_half(oSH[k]);
            cocg_h2[2 * i] = float_to_half(oCoCg[0]); cocg_h2[2 * i + 1] = float_to_half(oCoCg[1]);
            utility_h[i] = float_to_half(oUtil);
            aosky_u8x2[2 * i] = float_to_unorm8(oAO); aosky_u8x2[2 * i + 1] = float_to_unorm8(oSky);
            s_rays += c.rays; s_it += c.iters; s_dda += c.dda; s_hits += c.hits;
        }
    if (stats) { stats->rays += s_rays; stats->iterations += s_it; stats->dda_steps += s_dda; stats->hits += s_hits; }
}

/* ------------------------------------------------------------------------------------------------
 * ReflectionTraceFrag.glsl
 * ---------------------------------------------------------------------------------------------- */
namespace {

struct RfCtx {
    const vxo_scene* s;
    const vxrt_reflection_params* p;
    int px, py;
    int CurrentBLSample;
    v3 viewer;
    uint64_t rays, iters, dda, hits;
};
inline float rf_trace(RfCtx& c, v3 o, v3 d, int max_iter, vxo_hit* h) {
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    float t = vxo_traverse(&c.s->world, oo, dd, max_iter, h);
    c.rays++; c.iters += h->iterations; c.dda += h->dda_steps; c.hits += t > 0.0f ? 1 : 0;
    return t;
}
/* SampleBlueNoise2D (:606-614): dimension counter wraps at 128 */
inline v2 rf_blue_noise_2d(RfCtx& c, int Index) {
    v2 n;
    n.x = blue_noise_1d(c.s, c.px, c.py, Index, 1 + c.CurrentBLSample);
    n.y = blue_noise_1d(c.s, c.px, c.py, Index, 2 + c.CurrentBLSample);
    c.CurrentBLSample += 2;
    c.CurrentBLSample = c.CurrentBLSample % 128;
    return n;
}
/* ImportanceSampleGGX (:345-365) */
inline v3 importance_sample_ggx(v3 N, float roughness, v2 Xi) {
    float alpha = roughness * roughness;
    float alpha2 = alpha * alpha;
    float phi = 2.0f * PI * Xi.x;
    float cosTheta = sqrtf((1.0f - Xi.y) / (1.0f + (alpha2 - 1.0f) * Xi.y));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    v3 H = V3(cosf(phi) * sinTheta, sinf(phi) * sinTheta, cosTheta);
    v3 up = fabsf(N.z) < 0.999f ? V3(0.0f, 0.0f, 1.0f) : V3(1.0f, 0.0f, 0.0f);
    v3 tangent = normalize(cross(up, N));
    v3 bitangent = cross(N, tangent);
    v3 sampleVec = tangent * H.x + bitangent * H.y + N * H.z;
    return normalize(sampleVec);
}
/* GetReflectionDirection (:621-644): best (largest N.H) of three GGX samples */
inline v3 get_reflection_direction(RfCtx& c, v3 N, float R) {
    R = gmax(R, 0.05f);
    float NearestDot = -100.0f;
    v3 Best = V3(0.0f);
    for (int i = 0; i < 3; ++i) {
        v2 Xi = rf_blue_noise_2d(c, c.p->temporal ? c.p->current_frame_mod128 : 100);
        Xi = Xi * V2(0.9f, 0.65f);
        v3 H = importance_sample_ggx(N, R, Xi);
        float d = dot(H, N);
        if (d > NearestDot) { Best = H; NearestDot = d; }
    }
    return Best;
}
/* SHToIrradianceA (:452-462) and SHToIrridiance (:437-449) */
inline v3 sh_to_irradiance_a(v4 shY, v2 CoCg) {
    float Y = gmax(0.0f, 3.544905f * shY.w);
    CoCg = CoCg * (Y * 0.282095f / (shY.w + 1e-6f));
    float T = Y - CoCg.y * 0.5f;
    float G = CoCg.y + T;
    float B = T - CoCg.x * 0.5f;
    float R = B + CoCg.x;
    return V3(gmax(R, 0.0f), gmax(G, 0.0f), gmax(B, 0.0f));
}
inline v3 sh_to_irradiance(v4 shY, v2 CoCg, v3 v) {
    float x = dot(V3(shY.x, shY.y, shY.z), v);
    float Y = 2.0f * (1.023326f * x + 0.886226f * shY.w);
    Y = gmax(Y, 0.0f);
    CoCg = CoCg * (Y * 0.282095f / (shY.w + 1e-6f));
    float T = Y - CoCg.y * 0.5f;
    float G = CoCg.y + T;
    float B = T - CoCg.x * 0.5f;
    float R = B + CoCg.x;
    return V3(gmax(R, 0.0f), gmax(G, 0.0f), gmax(B, 0.0f));
}
/* G_Smith_over_NdotV, SpecularGGX (:410-435), DeriveSpecularFromDiffuseSH (:521-546) */
inline float sq(float x) { return x * x; }
inline float g_smith_over_ndotv(float roughness, float NdotV, float NdotL) {
    float alpha = sq(roughness);
    float g1 = NdotV * sqrtf(sq(alpha) + (1.0f - sq(alpha)) * sq(NdotL));
    float g2 = NdotL * sqrtf(sq(alpha) + (1.0f - sq(alpha)) * sq(NdotV));
    return 2.0f * NdotL / (g1 + g2);
}
inline float specular_ggx(v3 V, v3 L, v3 N, float roughness, float NoH_offset) {
    v3 H = normalize(L - V);
    float NoL = gmax(0.0f, dot(N, L));
    float NoV = gmax(0.0f, -dot(N, V));
    float NoH = gclamp(dot(N, H) + NoH_offset, 0.0f, 1.0f);
    if (NoL > 0.0f) {
        float G = g_smith_over_ndotv(roughness, NoV, NoL);
        float alpha = sq(gmax(roughness, 0.02f));
        float D = sq(alpha) / (PI * sq(sq(NoH) * sq(alpha) + (1.0f - sq(NoH))));
        return D * G / 4.0f;
    }
    return 0.0f;
}
inline v3 derive_specular_from_diffuse_sh(v4 SHy, v3 IndirectDiffuse, v3 Eye, v3 Normal) {
    float Roughness = 0.4f;
    v3 IncomingDir = V3(SHy.x, SHy.y, SHy.z) / SHy.w * (0.282095f / 0.488603f);
    v3 RawSpecularDir = reflect(Eye, Normal);
    float IncomingLen = length(IncomingDir);
    float Directionality = IncomingLen;
    float Scale = 1.0f;
    if (Directionality >= 1.0f) {
        IncomingDir = IncomingDir / IncomingLen;
    } else {
        v3 q = IncomingDir / (IncomingLen + 0.00001f);
        IncomingDir = V3(gmix(RawSpecularDir.x, q.x, Directionality), gmix(RawSpecularDir.y, q.y, Directionality), gmix(RawSpecularDir.z, q.z, Directionality));
        Scale = powf(Roughness + 1.0f, 3.0f);
    }
    float Sp = specular_ggx(Eye, IncomingDir, Normal, gmax(Roughness, 0.39f), 0.0f);
    v3 Integrated = powf(Sp, 1.2f) * IndirectDiffuse * 18.0f * Scale;
    if (Integrated.x != Integrated.x || isinf(Integrated.x) || Integrated.y != Integrated.y || isinf(Integrated.y) || Integrated.z != Integrated.z || isinf(Integrated.z))
        Integrated = V3(0.0f);
    return gmax(Integrated, 0.00001f);
}
/* capIntersect (:1264-1291), GetPlayerIntersect (:1301-1307) */
inline float cap_intersect(v3 ro, v3 rd, v3 pa, v3 pb, float r) {
    v3 ba = pb - pa, oa = ro - pa;
    float baba = dot(ba, ba), bard = dot(ba, rd), baoa = dot(ba, oa), rdoa = dot(rd, oa), oaoa = dot(oa, oa);
    float a = baba - bard * bard;
    float b = baba * rdoa - baoa * bard;
    float cc = baba * oaoa - baoa * baoa - r * r * baba;
    float h = b * b - a * cc;
    if (h >= 0.0f) {
        float t = (-b - sqrtf(h)) / a;
        float y = baoa + t * bard;
        if (y > 0.0f && y < baba) return t;
        v3 oc = (y <= 0.0f) ? oa : ro - pb;
        b = dot(rd, oc);
        cc = dot(oc, oc) - r * r;
        h = b * b - cc;
        if (h > 0.0f) return -b - sqrtf(h);
    }
    return -1.0f;
}
inline bool get_player_intersect(const RfCtx& c, v3 WorldPos, v3 d) {
    float x = 0.4f;
    v3 VP = c.viewer + V3(-x, -x, +x);
    float t = cap_intersect(WorldPos, d, VP, VP + V3(0.0f, 1.0f, 0.0f), 0.5f);
    return t > 0.0f;
}
/* GetShadowAt (:1327-1346): 150-iteration shadow ray */
inline float rf_shadow_at(RfCtx& c, v3 pos, v3 ldir) {
    if (get_player_intersect(c, pos, ldir)) return 1.0f;
    vxo_hit h;
    float T = rf_trace(c, pos, ldir, c.p->shadow_trace_length, &h);
    return T > 0.0f ? 1.0f : 0.0f;
}
/* the reflection pass' own CalculateDirectionalLight (:313-336): specular lobe multiplied by 0 */
inline v3 rf_directional_light(v3 viewer, v3 world_pos, v3 light_dir, v3 radiance, v3 albedo, v3 normal, v3 pbr, float shadow) {
    const float Epsilon = 0.00001f;
    float Shadow = gmin(shadow, 1.0f);
    v3 Lo = normalize(viewer - world_pos);
    v3 N = normal;
    float cosLo = gmax(0.0f, dot(N, Lo));
    v3 F0 = gmix(V3(0.04f), albedo, pbr.y);
    v3 Li = light_dir;
    v3 Lh = normalize(Li + Lo);
    float cosLi = gmax(0.0f, dot(N, Li));
    float cosLh = gmax(0.0f, dot(N, Lh));
    float fc = powf(1.0f - gmax(0.0f, dot(Lh, Lo)), 5.0f);
    v3 F = F0 + (V3(1.0f) - F0) * fc;                     /* fresnelSchlick (:307-310) */
    float D = ndf_ggx(cosLh, pbr.x);
    float G = ga_schlick_ggx(cosLi, cosLo, pbr.x);
    v3 kd = gmix(V3(1.0f) - F, V3(0.0f), pbr.y);
    v3 diffuseBRDF = kd * albedo;
    v3 specularBRDF = (F * D * G) / gmax(Epsilon, 4.0f * cosLi * cosLo);
    v3 radiance_s = radiance * 0.05f * 0.0f;
    v3 Result = (diffuseBRDF * radiance * cosLi) + (specularBRDF * radiance_s * cosLi);
    return gmax(Result, 0.0f) * gclamp(1.0f - Shadow, 0.0f, 1.0f);
}
/* SampleMoonColor (:657-665) */
inline v3 sample_moon_color(const vxo_scene* s, v3 moon, float strength) {
    v3 c = xyz(texcube_sample(s->skymap, moon.x, moon.y, moon.z));
    c = c * PI * strength;
    c = basic_saturation(c, 1.3f);
    return c * 0.42525f * strength;
}
inline bool in_thresholded_screen_space(v2 v) {
    float b = 0.032593f;
    return v.x > b && v.x < 1.0f - b && v.y > b && v.y < 1.0f - b;
}

}  // namespace

extern "C" void vxo_lpv_sample(const uint8_t* level, const uint8_t* block_type, int32_t nx, int32_t ny, int32_t nz, const float* avg512,
                               const float* points, int32_t n, const float dither[3], float* rgb_out);   /* vxrt_oracle_lpv.cpp */

/* bayer2 .. bayer32 (ReflectionTraceFrag.glsl:21-29) */
static inline float bayer2(float ax, float ay) {
    ax = floorf(ax); ay = floorf(ay);
    return gfract(ax * 0.5f + ay * (ay * 0.75f));   /* dot(a, vec2(0.5, a.y * 0.75)) */
}
static inline float bayer4(float ax, float ay) { return bayer2(0.5f * ax, 0.5f * ay) * 0.25f + bayer2(ax, ay); }
static inline float bayer8(float ax, float ay) { return bayer4(0.5f * ax, 0.5f * ay) * 0.25f + bayer2(ax, ay); }
static inline float bayer16(float ax, float ay) { return bayer8(0.5f * ax, 0.5f * ay) * 0.25f + bayer2(ax, ay); }
static inline float bayer32(float ax, float ay) { return bayer16(0.5f * ax, 0.5f * ay) * 0.25f + bayer2(ax, ay); }

/* ApproximateGILPV (ReflectionTraceFrag.glsl:673-700) with LPVDither of main() (:721-723) and SampleLPVData (vxo_lpv_sample) */
static v3 approximate_gi_lpv(const vxo_scene* s, const vxrt_reflection_params* p, const Tex2D& tAO, int px, int py, v2 vtc, bool SunStronger,
                             v3 SkyAmbientG, v3 P, v3 B) {
    const float fx = (float)px + 0.5f, fy = (float)py + 0.5f;
    const float tf = p->temporal ? 1.0f : 0.0f;
    const float b32 = bayer32(fx + ((float)p->current_frame * 0.75f) * tf, fy + ((float)p->current_frame * 0.5f) * tf);
    const float dither[3] = {b32 / 384.0f, b32 / 128.0f, b32 / 384.0f};
    const float pt[3] = {P.x, P.y, P.z};
    float rgb[3];
    vxo_lpv_sample(s->lpv_level, s->lpv_type, s->world.nx, s->world.ny, s->world.nz, s->lpv_avg512, pt, 1, dither, rgb);
    const v3 LPV = V3(rgb[0], rgb[1], rgb[2]);
    if (p->use_decoupled_gi) {
        v3 Sky = SkyAmbientG;
        const float L = dot(Sky, V3(0.2125f, 0.7154f, 0.0721f));
        Sky = gmix(V3(L), Sky, SunStronger ? 0.3f : 0.6f);
        v3 Skylighting = Sky * (tex2d_sample(tAO, vtc.x, vtc.y).y * (SunStronger ? 3.5f : 4.0f));
        Skylighting = Skylighting + V3(bayer16(fx, fy) / 512.0f);
        const v3 hi = B + V3(0.075f);
        Skylighting = V3(gclamp(Skylighting.x * 13.0f, 0.0f, hi.x), gclamp(Skylighting.y * 13.0f, 0.0f, hi.y), gclamp(Skylighting.z * 13.0f, 0.0f, hi.z));
        if (!p->screen_space_skylighting_valid) Skylighting = B;
        return Skylighting + LPV;
    }
    const v3 BaseAmbient = B * 0.9f;
    return LPV + BaseAmbient;
}

extern "C" void vxo_reflection_trace(const vxo_scene* s, const vxrt_reflection_params* p, const vxo_reflection_inputs* in,
                                     uint16_t* color_h4, uint16_t* hitdist_h, uint8_t* emissive_u8, vxrt_trace_stats* stats) {
    const int W = p->width, H = p->height;
    int r0, r1;
    tile_rows(p->tile, H, &r0, &r1);
    std::vector<float> tf = half_to_f32(in->g_t_half, (size_t)in->gw * in->gh), nf = u8_to_f32(in->g_normal, (size_t)in->gw * in->gh);
    std::vector<float> gnf = half_to_f32(in->gb_normal_h3, (size_t)in->mw * in->mh * 3), gpf = u8_to_f32(in->gb_pbr_u8x4, (size_t)in->mw * in->mh * 4);
    std::vector<float> shf = half_to_f32(in->gi_sh_h4, (size_t)in->iw * in->ih * 4), ccf = half_to_f32(in->gi_cocg_h2, (size_t)in->iw * in->ih * 2);
    std::vector<float> aof = u8_to_f32(in->gi_aosky_u8x2, (size_t)in->iw * in->ih * 2), sf = u8_to_f32(in->shadow_u8, (size_t)in->sw * in->sh);
    const Tex2D tT = view_f32(tf.data(), in->gw, in->gh, 1, true), tN = view_f32(nf.data(), in->gw, in->gh, 1, false);
    const Tex2D tGN = view_f32(gnf.data(), in->mw, in->mh, 3, true), tGP = view_f32(gpf.data(), in->mw, in->mh, 4, false);
    const Tex2D tSH = view_f32(shf.data(), in->iw, in->ih, 4, true), tCC = view_f32(ccf.data(), in->iw, in->ih, 2, true);
    const Tex2D tAO = view_f32(aof.data(), in->iw, in->ih, 2, true), tS = view_f32(sf.data(), in->sw, in->sh, 1, true);
    const v3 cam = V3(p->inv_view[12], p->inv_view[13], p->inv_view[14]);
    const v3 viewer = V3(p->viewer_position[0], p->viewer_position[1], p->viewer_position[2]);
    const v3 sun = V3(p->sun_direction[0], p->sun_direction[1], p->sun_direction[2]), moon = V3(p->moon_direction[0], p->moon_direction[1], p->moon_direction[2]);
    const v3 strong = V3(p->stronger_light_direction[0], p->stronger_light_direction[1], p->stronger_light_direction[2]);
    /* main() prologue (:722-727) */
    const v3 SAMPLED_SUN_COLOR = sample_sun_color(s, sun, p->sun_strength_modifier);
    const v3 SAMPLED_MOON_COLOR = sample_moon_color(s, moon, p->moon_strength_modifier);
    float SunVisibility = gclamp(dot(sun, V3(0.0f, 1.0f, 0.0f)) + 0.05f, 0.0f, 0.1f) * 12.0f;
    SunVisibility = 1.0f - SunVisibility;
    const v3 SAMPLED_COLOR_MIXED = gmix(SAMPLED_SUN_COLOR, SAMPLED_MOON_COLOR, SunVisibility);
    const bool SunStronger = strong.x == sun.x && strong.y == sun.y && strong.z == sun.z;   /* :809 */
    const v3 SkyAmbientG = p->lpv_gi ? xyz(texcube_sample(s->skymap, 0.0f, 1.0f, 0.0f)) : V3(0.0f);   /* :719 */
    uint64_t s_rays = 0, s_it = 0, s_dda = 0, s_hits = 0;
#pragma omp parallel for schedule(dynamic, 2) num_threads(nthreads()) reduction(+ : s_rays, s_it, s_dda, s_hits)
    for (int py = r0; py < r1; ++py)
        for (int px = 0; px < W; ++px) {
            const size_t i = (size_t)py * W + px;
            RfCtx c;
            c.s = s; c.p = p; c.px = px; c.py = py; c.CurrentBLSample = 0; c.viewer = viewer;
            c.rays = c.iters = c.dda = c.hits = 0;
            const v2 vtc = pixel_uv(px, py, W, H);
            bool CheckerStep = cvt_trunc(((float)px + 0.5f) + ((float)py + 0.5f)) % 2 == (p->current_frame % 2);
            int SPP = iclamp(p->spp, 1, 16);
            if (p->checkerboard) SPP = cvt_trunc(gmix((float)p->spp, (float)((p->spp + p->spp % 2) / 2), CheckerStep ? 1.0f : 0.0f));
            SPP = iclamp(SPP, 1, 16);
            v2 Jitter = V2(gclamp(p->halton[0] * 1.0f, -2.0f, 2.0f), gclamp(p->halton[1] * 1.0f, -2.0f, 2.0f));
            v2 tc = vtc + ((Jitter / V2((float)W, (float)H)) * (p->temporal ? 1.0f : 0.0f));
            float Dist = tex2d_sample(tT, tc.x, tc.y).x;
            v3 P = cam + normalize(ray_direction_at(p->inv_view, p->inv_projection, tc)) * Dist;
            v4 oColor; float oHit, oMask;
            bool written = false;
            if (Dist < 0.0f) {
                oColor = V4(0, 0, 0, 0); oMask = 0.0f; oHit = -1.0f; written = true;
            }
            v3 N0 = V3(0.0f), NormalMappedInitial = V3(0.0f), I = V3(0.0f), BaseIndirectDiffuse = V3(0.0f);
            v4 PBRMap = V4(0, 0, 0, 0), DiffuseSH = V4(0, 0, 0, 0);
            v2 DiffuseCoCg = V2(0, 0);
            if (!written) {
                N0 = normal_from_id(tex2d_sample(tN, tc.x, tc.y).x, V3(1.0f));
                PBRMap = tex2d_sample(tGP, vtc.x, vtc.y);
                I = normalize(P - viewer);
                P = P + N0 * 0.035f;
                NormalMappedInitial = xyz(tex2d_sample(tGN, vtc.x, vtc.y));
                DiffuseSH = tex2d_sample(tSH, vtc.x, vtc.y);
                v4 cc = tex2d_sample(tCC, vtc.x, vtc.y);
                DiffuseCoCg = V2(cc.x, cc.y);
                BaseIndirectDiffuse = sh_to_irradiance_a(DiffuseSH, DiffuseCoCg);
                bool TooRough = PBRMap.x >= 0.865f;
                if (TooRough && p->derive_from_diffuse_sh) {
                    v3 r = derive_specular_from_diffuse_sh(DiffuseSH, sh_to_irradiance(DiffuseSH, DiffuseCoCg, NormalMappedInitial), I, NormalMappedInitial);
                    oColor = V4(r.x, r.y, r.z, 0.0f); oHit = 0.5f; oMask = 0.0f; written = true;
                }
            }
            if (!written) {
                const float RoughnessAt = PBRMap.x;
                float RoughnessBias = gmix(1.0f, 0.85f, p->roughness_bias ? 1.0f : 0.0f);
                float ComputedShadow = 0.0f;
                int ShadowItr = 0;
                float AveragedHitDistance = 0.001f, TotalMeaningfulHits = 0.0f, EmissivityMask = 0.0f;
                int total_hits = 0;
                v4 TotalColor = V4(0, 0, 0, 0);
                for (int sidx = 0; sidx < SPP; ++sidx) {
                    v3 ReflectionNormal = p->rough_reflections ? get_reflection_direction(c, NormalMappedInitial, gclamp(RoughnessAt * RoughnessBias, 0.01f, 1.0f)) : NormalMappedInitial;
                    v3 R = reflect(I, ReflectionNormal);
                    vxo_hit h;
                    float T = rf_trace(c, P, R, p->trace_length, &h);
                    v3 Normal = V3(h.normal[0], h.normal[1], h.normal[2]);
                    v3 HitPosition = P + (R * T);
                    if (T > 0.0f) {
                        v2 UV = V2(0.0f, 0.0f);
                        v3 Tangent = V3(0.0f), Bitangent = V3(0.0f);
                        calculate_vectors(HitPosition, Normal, &Tangent, &Bitangent, &UV);
                        UV.y = 1.0f - UV.y;
                        int reference_id = iclamp(h.block, 0, 127);
                        bool ReprojectionSuccessful = false;
                        v2 SS = V2(-1.0f, -1.0f);
                        v3 Ambient = BaseIndirectDiffuse;
                        if (p->reproject_to_screen_space) {
                            /* ReprojectReflectionToScreenSpace (:574-586) */
                            v4 pv = mat4_mul(p->view, V4(HitPosition.x, HitPosition.y, HitPosition.z, 1.0f));
                            /* u_Projection * u_View * v evaluates (Projection*View) first (mat4*mat4), then the vector */
                            float PV[16];
                            for (int col = 0; col < 4; ++col) {
                                v4 cv = mat4_mul(p->projection, V4(p->view[4 * col], p->view[4 * col + 1], p->view[4 * col + 2], p->view[4 * col + 3]));
                                PV[4 * col] = cv.x; PV[4 * col + 1] = cv.y; PV[4 * col + 2] = cv.z; PV[4 * col + 3] = cv.w;
                            }
                            (void)pv;
                            v4 pp = mat4_mul(PV, V4(HitPosition.x, HitPosition.y, HitPosition.z, 1.0f));
                            v3 q = V3(pp.x / pp.w, pp.y / pp.w, pp.z / pp.w);
                            SS = V2(q.x * 0.5f + 0.5f, q.y * 0.5f + 0.5f);
                            float d2 = tex2d_sample(tT, SS.x, SS.y).x;
                            v3 PosAt = cam + normalize(ray_direction_at(p->inv_view, p->inv_projection, SS)) * d2;
                            v3 NormalAt = normal_from_id(tex2d_sample(tN, SS.x, SS.y).x, V3(1.0f));
                            v3 diff = abs3(PosAt - HitPosition);
                            float Error = dot(diff, diff);
                            ReprojectionSuccessful = Error < 0.095f && eq3(NormalAt, Normal) && in_thresholded_screen_space(SS);
                            if (ReprojectionSuccessful) {
                                v4 rsh = tex2d_sample(tSH, SS.x, SS.y);
                                v4 rcc = tex2d_sample(tCC, SS.x, SS.y);
                                Ambient = sh_to_irradiance_a(rsh, V2(rcc.x, rcc.y));
                                float ReprojectedVXAO = powf(tex2d_sample(tAO, SS.x, SS.y).x, 0.75f);
                                if (d2 > 0.0f) {
                                    if (distance(PosAt, cam) < 40.0f) Ambient = Ambient * ReprojectedVXAO;
                                }
                            }
                        }
                        if (p->lpv_gi && !ReprojectionSuccessful)   /* :881-883 */
                            Ambient = approximate_gi_lpv(s, p, tAO, px, py, vtc, SunStronger, SkyAmbientG, HitPosition + Normal * 0.5f, BaseIndirectDiffuse);
                        const int32_t* bd = s->block_data;
                        v4 ids = V4((float)bd[0 * 128 + reference_id], (float)bd[1 * 128 + reference_id], (float)bd[2 * 128 + reference_id], (float)bd[3 * 128 + reference_id]);
                        if (reference_id == p->grass_props[0]) {
                            const int32_t* g = p->grass_props;
                            if (eq3(Normal, V3(-1, 0, 0)) || eq3(Normal, V3(1, 0, 0)) || eq3(Normal, V3(0, 0, 1)) || eq3(Normal, V3(0, 0, -1))) { ids.x = (float)g[4]; ids.y = (float)g[5]; ids.z = (float)g[6]; }
                            else if (eq3(Normal, V3(0, 1, 0))) { ids.x = (float)g[1]; ids.y = (float)g[2]; ids.z = (float)g[3]; }
                            else if (eq3(Normal, V3(0, -1, 0))) { ids.x = (float)g[7]; ids.y = (float)g[8]; ids.z = (float)g[9]; }
                        }
                        v3 Albedo = xyz(texarray_sample(s->tex[VXRT_TEX_ALBEDO], UV.x, UV.y, ids.x, 0.0f));
                        v3 Radiance = SAMPLED_COLOR_MIXED * 0.6f;
                        v4 SampledPBR = texarray_sample(s->tex[VXRT_TEX_PBR], UV.x, UV.y, ids.z, 0.0f);
                        float AO = powf(SampledPBR.w, 2.0f);
                        bool PlayerInShadow = get_player_intersect(c, HitPosition + Normal * 0.035f, strong);
                        if (ShadowItr < (SPP / 4 > 1 ? SPP / 4 : 1)) {
                            if (!PlayerInShadow) {
                                if (ReprojectionSuccessful && p->reproject_to_screen_space && in_thresholded_screen_space(SS))
                                    ComputedShadow = tex2d_sample(tS, SS.x, SS.y).x;
                                else
                                    ComputedShadow = rf_shadow_at(c, HitPosition + Normal * 0.055f, strong);
                            } else {
                                ComputedShadow = 1.0f;
                            }
                            ShadowItr = ShadowItr + 1;
                        }
                        Ambient = (Ambient * 1.0f * gclamp(AO, 0.1f, 1.0f)) * Albedo;
                        v3 nm = xyz(texarray_sample(s->tex[VXRT_TEX_NORMAL], UV.x, UV.y, ids.y, 3.0f)) * 2.0f - V3(1.0f);
                        v3 NormalMapped = mat3_mul(Tangent, Bitangent, Normal, nm);
                        v3 DirectLighting = Ambient + rf_directional_light(viewer, HitPosition, strong, Radiance, Albedo, NormalMapped,
                                                                          V3(SampledPBR.x, SampledPBR.y, SampledPBR.z), ComputedShadow);
                        if (ids.w > -0.5f) {
                            float Emissivity = texarray_sample(s->tex[VXRT_TEX_EMISSIVE], UV.x, UV.y, ids.w, 2.0f).x;
                            if (Emissivity > 0.1f) {
                                float m = 19.0f;
                                float lbiasx = 0.02501f, lbiasy = 0.03001f;
                                Emissivity *= (UV.x > lbiasx && UV.x < 1.0f - lbiasx && UV.y > lbiasy && UV.y < 1.0f - lbiasy) ? 1.0f : 0.0f;
                                float Flicker = 1.0f;
                                DirectLighting = Albedo * gmax(Emissivity * m * Flicker, 2.0f);
                                EmissivityMask = 1.0f;
                            }
                        }
                        TotalColor = V4(TotalColor.x + DirectLighting.x, TotalColor.y + DirectLighting.y, TotalColor.z + DirectLighting.z, TotalColor.w + 1.0f);
                        AveragedHitDistance += T;
                        TotalMeaningfulHits += 1.0f;
                    } else {
                        /* GetAtmosphere (:494-502) without projected clouds */
                        v3 nv = normalize(R);
                        v3 Atmos = xyz(texcube_sample(s->skymap, nv.x, nv.y, nv.z));
                        v3 a = Atmos * gmix(1.0f, 1.175f, (PBRMap.y > 0.05f) ? 1.0f : 0.0f);
                        TotalColor = V4(TotalColor.x + a.x, TotalColor.y + a.y, TotalColor.z + a.z, TotalColor.w + 1.0f);
                    }
                    total_hits++;
                }
                AveragedHitDistance /= gmax(TotalMeaningfulHits, 0.01f);
                const float th = (float)total_hits;
                TotalColor = V4(TotalColor.x / th, TotalColor.y / th, TotalColor.z / th, TotalColor.w / th);
                oColor = V4(gclamp(TotalColor.x, 0.0000001f, 100.0f), gclamp(TotalColor.y, 0.0000001f, 100.0f), gclamp(TotalColor.z, 0.0000001f, 100.0f), gclamp(TotalColor.w, 0.0000001f, 100.0f));
                oHit = gclamp(TotalMeaningfulHits > 0.01f ? AveragedHitDistance : -1.0f, -10.0f, 200.0f);
                oMask = gclamp(EmissivityMask, 0.0f, 1.0f);
            }
            color_h4[4 * i] = float_to_half(oColor.x); color_h4[4 * i + 1] = float_to_half(oColor.y);
            color_h4[4 * i + 2] = float_to_half(oColor.z); color_h4[4 * i + 3] = float_to_half(oColor.w);
            hitdist_h[i] = float_to_half(oHit);
            emissive_u8[i] = float_to_unorm8(oMask);
            s_rays += c.rays; s_it += c.iters; s_dda += c.dda; s_hits += c.hits;
        }
    if (stats) { stats->rays += s_rays; stats->iterations += s_it; stats->dda_steps += s_dda; stats->hits += s_hits; }
}
