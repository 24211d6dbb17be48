/*
 * vxo_grid.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY): grid accessors shared by the traversal variants.
 */
#pragma once
#include <math.h>
#include "vxrt_oracle.h"
#include "vxo_math.h"

/* IsInVolume (InitialRayTraceFrag.glsl:68-77), applied to an ivec3 converted to vec3 */
static inline bool in_volume(const vxo_world* w, int x, int y, int z) {
    float px = (float)x, py = (float)y, pz = (float)z;
    if (px < 0.0f || py < 0.0f || pz < 0.0f || px > (float)(w->nx - 1) || py > (float)(w->ny - 1) ||
        pz > (float)(w->nz - 1))
        return false;
    return true;
}
/* GetVoxel (:79-87) — returns the raw texel code (id), the shader's value is id/255 */
static inline int get_voxel(const vxo_world* w, int x, int y, int z) {
    if (in_volume(w, x, y, z)) return w->blocks[x + (size_t)y * w->nx + (size_t)z * w->nx * w->ny];
    return 0;
}
/* GetDistance (:94-102) * 255, ToConservativeEuclidean (:89-92), int(floor()) (:331-333) */
static inline int euclidean_step(int k) {
    float Dist = vxo::unorm8_to_float(k) * 255.0f;
    float ce = (Dist == 1.0f) ? 1.0f : Dist * 0.57735026918f;
    return (int)floorf(ce);
}
