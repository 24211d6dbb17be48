/*
 * vxo_math.h — CPU ORACLE helpers (TEST INFRASTRUCTURE ONLY, see vxrt_oracle.h).
 * GLSL-semantics float math with the evaluation order pinned (no contraction; build with
 * -ffp-contract=off).  Association follows glm 0.9.8 (the reference's Dependencies/glm), which
 * is what the host side of the reference and the oracle/_ref shim evaluate.
 */
#ifndef VXO_MATH_H
#define VXO_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <limits.h>

namespace vxo {

struct v2 { float x, y; };
struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };
struct i3 { int x, y, z; };

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 V3(float s) { v3 r = {s, s, s}; return r; }
static inline v2 V2(float x, float y) { v2 r = {x, y}; return r; }
static inline v4 V4(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }

static inline v3 operator+(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 operator-(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 operator*(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 operator/(v3 a, v3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 operator*(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 operator*(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline v3 operator/(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 operator-(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v2 operator+(v2 a, v2 b) { return V2(a.x + b.x, a.y + b.y); }
static inline v2 operator-(v2 a, v2 b) { return V2(a.x - b.x, a.y - b.y); }
static inline v2 operator*(v2 a, v2 b) { return V2(a.x * b.x, a.y * b.y); }
static inline v2 operator*(v2 a, float s) { return V2(a.x * s, a.y * s); }
static inline v2 operator/(v2 a, v2 b) { return V2(a.x / b.x, a.y / b.y); }

static inline float& idx(v3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
static inline float idx(const v3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
static inline int& idx(i3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
static inline int idx(const i3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

/* GLSL/glm min/max/clamp/mix */
static inline float gmin(float a, float b) { return (b < a) ? b : a; }
static inline float gmax(float a, float b) { return (a < b) ? b : a; }
static inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
static inline int iclamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float gmix(float a, float b, float t) { return a * (1.0f - t) + b * t; } /* glm: x*(1-a) + y*a */
static inline v3 gmix(v3 a, v3 b, float t) { return V3(gmix(a.x, b.x, t), gmix(a.y, b.y, t), gmix(a.z, b.z, t)); }
static inline v3 gmax(v3 a, float b) { return V3(gmax(a.x, b), gmax(a.y, b), gmax(a.z, b)); }
static inline v3 gmin(v3 a, float b) { return V3(gmin(a.x, b), gmin(a.y, b), gmin(a.z, b)); }
static inline v3 gclamp(v3 a, float lo, float hi) { return V3(gclamp(a.x, lo, hi), gclamp(a.y, lo, hi), gclamp(a.z, lo, hi)); }
static inline float gfract(float x) { return x - floorf(x); }
static inline int gsign(float x) { return (x > 0.0f) - (x < 0.0f); }

static inline float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float dot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float length(v3 a) { return sqrtf(dot(a, a)); }
static inline float distance(v3 a, v3 b) { return length(b - a); } /* glm: length(p1 - p0) */
static inline v3 normalize(v3 a) { float s = 1.0f / sqrtf(dot(a, a)); return a * s; }
static inline v3 cross(v3 a, v3 b) {
    return V3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline v3 reflect(v3 I, v3 N) { return I - N * dot(N, I) * 2.0f; } /* glm: I - N * dot(N, I) * 2 */

/* column-major mat4 (glm::value_ptr layout) times vec4, glm 0.9.8 association */
static inline v4 mat4_mul(const float* m, v4 v) {
    v4 r;
    float* o = &r.x;
    for (int i = 0; i < 4; ++i) {
        float mul0 = m[0 + i] * v.x, mul1 = m[4 + i] * v.y;
        float add0 = mul0 + mul1;
        float mul2 = m[8 + i] * v.z, mul3 = m[12 + i] * v.w;
        float add1 = mul2 + mul3;
        o[i] = add0 + add1;
    }
    return r;
}
/* mat3 (columns c0,c1,c2) times vec3: glm  m[0]*v.x + m[1]*v.y + m[2]*v.z  per component */
static inline v3 mat3_mul(v3 c0, v3 c1, v3 c2, v3 v) {
    return V3(c0.x * v.x + c1.x * v.y + c2.x * v.z, c0.y * v.x + c1.y * v.y + c2.y * v.z,
              c0.z * v.x + c1.z * v.y + c2.z * v.z);
}

/* float -> int with cvt.s32.f32 semantics (saturating, NaN -> 0) */
static inline int cvt_floor(float x) {
    if (x != x) return 0;
    float f = floorf(x);
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int)f;
}
static inline int cvt_trunc(float x) {
    if (x != x) return 0;
    float f = truncf(x);
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int)f;
}
static inline int cvt_round(float x) { /* GLSL round(): implementation-defined ties; pinned half-even */
    if (x != x) return 0;
    float f = nearbyintf(x);
    if (f >= 2147483648.0f) return INT_MAX;
    if (f <= -2147483648.0f) return INT_MIN;
    return (int)f;
}

static inline float unorm8_to_float(int k) { return (float)k / 255.0f; }
static inline uint8_t float_to_unorm8(float f) {
    if (!(f > 0.0f)) return 0; /* NaN and negatives -> 0 */
    if (f >= 1.0f) return 255;
    return (uint8_t)nearbyintf(f * 255.0f);
}

static inline uint16_t float_to_half(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));
    }
    if (ax >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x33000001u) { /* < 2^-25 (or exactly 2^-25, tie -> even -> 0) */
        return (uint16_t)sign;
    }
    int e = (int)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    if (e < -14) { /* subnormal half */
        int shift = (-14 - e) + 13; /* bits to drop */
        uint32_t half_m = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half_m & 1u))) half_m++;
        return (uint16_t)(sign | half_m);
    }
    uint32_t half_e = (uint32_t)(e + 15);
    uint32_t half_m = (m >> 13) & 0x3ffu;
    uint32_t rem = m & 0x1fffu;
    uint32_t h = (half_e << 10) | half_m;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    return (uint16_t)(sign | h);
}
static inline float half_to_float(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1fu;
    uint32_t m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) {
            x = sign;
        } else {
            int shift = 0;
            while (!(m & 0x400u)) { m <<= 1; shift++; }
            m &= 0x3ffu;
            x = sign | ((uint32_t)(127 - 15 - shift + 1) << 23) | (m << 13);
        }
    } else if (e == 31) {
        x = sign | 0x7f800000u | (m << 13);
    } else {
        x = sign | ((e + 112u) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

}  // namespace vxo
#endif
