// vmath.cuh — GLSL-semantics float math for the trace kernels.
//
// The translation units that include this header are compiled with --fmad=false, and division /
// sqrt are the IEEE-correct defaults, so every expression below rounds exactly like the pinned
// evaluation order documented in DESIGN.md ("float semantics"): dot(a,b) = (ax*bx + ay*by) + az*bz,
// normalize(v) = v * (1/sqrt(dot(v,v))), mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w),
// min/max as GLSL defines them, float->int saturating with NaN -> 0.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

#define VXD __device__ __forceinline__

VXD f3 F3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
VXD f3 F3(float s) { return F3(s, s, s); }
VXD f2 F2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
VXD f4 F4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

VXD f3 operator+(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
VXD f3 operator-(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
VXD f3 operator*(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
VXD f3 operator/(f3 a, f3 b) { return F3(a.x / b.x, a.y / b.y, a.z / b.z); }
VXD f3 operator*(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
VXD f3 operator*(float s, f3 a) { return F3(s * a.x, s * a.y, s * a.z); }
VXD f3 operator/(f3 a, float s) { return F3(a.x / s, a.y / s, a.z / s); }
VXD f3 operator-(f3 a) { return F3(-a.x, -a.y, -a.z); }
VXD f2 operator+(f2 a, f2 b) { return F2(a.x + b.x, a.y + b.y); }
VXD f2 operator-(f2 a, f2 b) { return F2(a.x - b.x, a.y - b.y); }
VXD f2 operator*(f2 a, f2 b) { return F2(a.x * b.x, a.y * b.y); }
VXD f2 operator*(f2 a, float s) { return F2(a.x * s, a.y * s); }

VXD float comp(const f3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
VXD void set_comp(f3& v, int i, float s) { if (i == 0) v.x = s; else if (i == 1) v.y = s; else v.z = s; }

VXD float gmin(float a, float b) { return (b < a) ? b : a; }
VXD float gmax(float a, float b) { return (a < b) ? b : a; }
VXD float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
VXD int iclamp(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
VXD float gmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
VXD f3 gmix(f3 a, f3 b, float t) { return F3(gmix(a.x, b.x, t), gmix(a.y, b.y, t), gmix(a.z, b.z, t)); }
VXD f3 gmax(f3 a, float b) { return F3(gmax(a.x, b), gmax(a.y, b), gmax(a.z, b)); }
VXD f3 gmin(f3 a, float b) { return F3(gmin(a.x, b), gmin(a.y, b), gmin(a.z, b)); }
VXD f3 gclamp(f3 a, float lo, float hi) { return F3(gclamp(a.x, lo, hi), gclamp(a.y, lo, hi), gclamp(a.z, lo, hi)); }
VXD float gfract(float x) { return x - floorf(x); }
VXD int gsign(float x) { return (x > 0.0f) - (x < 0.0f); }

VXD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
VXD float length(f3 a) { return sqrtf(dot(a, a)); }
VXD float distance(f3 a, f3 b) { return length(b - a); }
VXD f3 normalize(f3 a) { float s = 1.0f / sqrtf(dot(a, a)); return a * s; }
VXD f3 cross(f3 a, f3 b) { return F3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
VXD f3 reflect(f3 I, f3 N) { return I - N * dot(N, I) * 2.0f; }

// column-major mat4 * vec4 with glm's association
VXD f4 mat4_mul(const float* __restrict__ m, f4 v) {
    f4 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * v.w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * v.w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w);
    r.w = (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w);
    return r;
}
VXD f3 mat3_mul(f3 c0, f3 c1, f3 c2, f3 v) {
    return F3(c0.x * v.x + c1.x * v.y + c2.x * v.z, c0.y * v.x + c1.y * v.y + c2.y * v.z,
              c0.z * v.x + c1.z * v.y + c2.z * v.z);
}

VXD int cvt_floor(float x) { return __float2int_rd(x); }  // saturating, NaN -> 0
VXD int cvt_trunc(float x) { return __float2int_rz(x); }
VXD int cvt_round(float x) { return __float2int_rn(x); }   // half-even

// float(k) / 255.0f for a byte k, bit for bit, without the division sequence: q = k * RN(1 / 255) is within an ulp, one residual step with
// two FMAs lands on the correctly rounded quotient (checked for all 256 codes with exact rational arithmetic; k and 255 are exact floats).
// The screen-space samplers convert 4 - 16 bytes per tap; ncu's source view put 4.5 % of shade_direct_kernel's instructions on the division.
VXD float unorm8_to_float(int k) {
    const float fk = (float)k, c = 1.0f / 255.0f;
    const float q = fk * c;
    return __fmaf_rn(__fmaf_rn(-q, 255.0f, fk), c, q);
}
// pow(x, n) for the integer exponents the shaders write (Schlick's (1 - c)^5 and friends) by multiplication: 1 - 3 roundings, i.e. within
// the 2 ulp of CUDA's powf that DESIGN.md 4 already states for this family, at 3 instructions instead of ~100 (ncu source view: 14.7 % of
// gi_wf_shade_kernel's instructions were powf(x, 5.0f) inside inverse_schlick).  x * x is the correctly rounded pow(x, 2).
VXD float pow2_mul(float x) { return x * x; }
VXD float pow3_mul(float x) { return (x * x) * x; }
VXD float pow5_mul(float x) { const float x2 = x * x; return (x2 * x2) * x; }
VXD uint8_t float_to_unorm8(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 1.0f) return 255;
    return (uint8_t)__float2int_rn(f * 255.0f);
}
VXD uint16_t float_to_half_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
VXD float half_bits_to_float(uint16_t h) { return __half2float(__ushort_as_half(h)); }

// GL_REPEAT on a texel index.  Almost every index is already inside [0, n) (the taps of a tile straddle the image edge only on its border), and
// `%` by a run-time n is ~20 instructions: 10 - 17 % of the shading kernels' instructions in ncu's source view.  One unsigned compare decides.
VXD int wrap_repeat(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    int m = i % n;
    return m < 0 ? m + n : m;
}
