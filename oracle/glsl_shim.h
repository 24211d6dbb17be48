/*
 * glsl_shim.h — a minimal GLSL 4.50 execution environment for C++ (TEST INFRASTRUCTURE ONLY).
 *
 * oracle/build_ref.py rewrites the reference's shader files (read in place from the reference
 * tree; nothing is copied into this repository) into C++ namespaces that compile against this
 * header, so that the reference's OWN shader source runs on the CPU and pins the oracle
 * (oracle/_ref/libvxrt_ref.so).  This header is our code: vector/matrix types with swizzles, the
 * GLSL built-ins, and sampler/image objects whose filtering follows the pinned GL behaviour of
 * DESIGN.md (the same formulas as oracle/vxo_math.h, vxo_texture.h).
 *
 * Everything lives in namespace glsl; generated shaders are nested inside it, so unqualified
 * built-in names resolve here and never to <math.h>.
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <type_traits>
#include <vector>

#include "vxo_math.h"
#include "vxo_texture.h"

namespace glsl {

typedef unsigned int uint;

template <class T, int N> struct vecn;

/* swizzle proxy: lives inside the union of its parent vector, addresses the parent's storage */
template <class T, int N, int A, int B, int C = -1, int D = -1>
struct swz {
    swz() = default;
    swz(const swz&) = delete; /* a copied proxy would lose its parent vector: build_ref.py wraps such uses */
    T* p() { return reinterpret_cast<T*>(this); }
    const T* p() const { return reinterpret_cast<const T*>(this); }
    operator vecn<T, N>() const;
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value && !(std::is_same<U, float>::value && std::is_integral<T>::value && !std::is_same<T, bool>::value)>>
    explicit operator vecn<U, N>() const { return vecn<U, N>((vecn<T, N>)(*this)); }
    /* GLSL converts integer vectors to float vectors implicitly (e.g. vec2 r = textureSize(s, 0).xy) */
    template <class U = T, class = std::enable_if_t<std::is_integral<U>::value && !std::is_same<U, bool>::value>>
    operator vecn<float, N>() const { return vecn<float, N>((vecn<T, N>)(*this)); }
    swz& operator=(const vecn<T, N>& v);
    swz& operator=(const swz& o) { return *this = (vecn<T, N>)o; }
    template <int A2, int B2, int C2, int D2> swz& operator=(const swz<T, N, A2, B2, C2, D2>& o) { return *this = (vecn<T, N>)o; }
#define GLSL_SWZ_OP(op)                                                                     \
    swz& operator op##=(const vecn<T, N>& v) { return *this = (vecn<T, N>)(*this)op v; }    \
    swz& operator op##=(T s) { return *this = (vecn<T, N>)(*this)op vecn<T, N>(s); }
    GLSL_SWZ_OP(+) GLSL_SWZ_OP(-) GLSL_SWZ_OP(*) GLSL_SWZ_OP(/)
#undef GLSL_SWZ_OP
    T& operator[](int i) { const int m[4] = {A, B, C, D}; return p()[m[i]]; }
    T operator[](int i) const { const int m[4] = {A, B, C, D}; return p()[m[i]]; }
    /* single-component reads of a swizzle, e.g. tex.xyz.x */
};

template <class T> struct vecn<T, 2> {
    union {
        struct { T x, y; };
        struct { T r, g; };
        struct { T s, t; };
#include "_ref/gen/swizzle2.inc"
    };
    vecn() : x(0), y(0) {}
    vecn(const vecn& o) : x(o.x), y(o.y) {}
    vecn& operator=(const vecn& o) { x = o.x; y = o.y; return *this; }
    explicit vecn(T a) : x(a), y(a) {}
    vecn(T a, T b) : x(a), y(b) {}
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value && std::is_same<T, float>::value>>
    vecn(const vecn<U, 2>& o) : x((T)o.x), y((T)o.y) {}
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value && !std::is_same<T, float>::value>, class = void>
    explicit vecn(const vecn<U, 2>& o);
    explicit vecn(const vecn<T, 3>& o);
    explicit vecn(const vecn<T, 4>& o);
    T& operator[](int i) { return (&x)[i]; }
    T operator[](int i) const { return (&x)[i]; }
};
template <class T> struct vecn<T, 3> {
    union {
        struct { T x, y, z; };
        struct { T r, g, b; };
        struct { T s, t, p; };
#include "_ref/gen/swizzle3.inc"
    };
    vecn() : x(0), y(0), z(0) {}
    vecn(const vecn& o) : x(o.x), y(o.y), z(o.z) {}
    vecn& operator=(const vecn& o) { x = o.x; y = o.y; z = o.z; return *this; }
    explicit vecn(T a) : x(a), y(a), z(a) {}
    vecn(T a, T b, T c) : x(a), y(b), z(c) {}
    vecn(const vecn<T, 2>& a, T c) : x(a.x), y(a.y), z(c) {}
    vecn(T a, const vecn<T, 2>& b) : x(a), y(b.x), z(b.y) {}
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value && std::is_same<T, float>::value>>
    vecn(const vecn<U, 3>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value && !std::is_same<T, float>::value>, class = void>
    explicit vecn(const vecn<U, 3>& o);
    explicit vecn(const vecn<T, 4>& o);
    T& operator[](int i) { return (&x)[i]; }
    T operator[](int i) const { return (&x)[i]; }
};
template <class T> struct vecn<T, 4> {
    union {
        struct { T x, y, z, w; };
        struct { T r, g, b, a; };
        struct { T s, t, p, q; };
#include "_ref/gen/swizzle4.inc"
    };
    vecn() : x(0), y(0), z(0), w(0) {}
    vecn(const vecn& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vecn& operator=(const vecn& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }
    explicit vecn(T a) : x(a), y(a), z(a), w(a) {}
    vecn(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    vecn(const vecn<T, 3>& a, T d) : x(a.x), y(a.y), z(a.z), w(d) {}
    vecn(T a, const vecn<T, 3>& b) : x(a), y(b.x), z(b.y), w(b.z) {}
    vecn(const vecn<T, 2>& a, T c, T d) : x(a.x), y(a.y), z(c), w(d) {}
    vecn(const vecn<T, 2>& a, const vecn<T, 2>& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    vecn(T a, T b, const vecn<T, 2>& c) : x(a), y(b), z(c.x), w(c.y) {}
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value && std::is_same<T, float>::value>>
    vecn(const vecn<U, 4>& o) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)o.w) {}
    template <class U, class = std::enable_if_t<!std::is_same<U, T>::value && !std::is_same<T, float>::value>, class = void>
    explicit vecn(const vecn<U, 4>& o);
    T& operator[](int i) { return (&x)[i]; }
    T operator[](int i) const { return (&x)[i]; }
};

/* float -> int conversions truncate with the pinned saturating semantics */
template <class T, class U> inline T conv(U v) { return (T)v; }
template <> inline int conv<int, float>(float v) { return vxo::cvt_trunc(v); }
template <> inline uint conv<uint, float>(float v) { return v <= 0.0f ? 0u : (v >= 4294967296.0f ? 0xffffffffu : (uint)v); }

template <class T> template <class U, class, class> vecn<T, 2>::vecn(const vecn<U, 2>& o) : x(conv<T, U>(o.x)), y(conv<T, U>(o.y)) {}
template <class T> template <class U, class, class> vecn<T, 3>::vecn(const vecn<U, 3>& o) : x(conv<T, U>(o.x)), y(conv<T, U>(o.y)), z(conv<T, U>(o.z)) {}
template <class T> template <class U, class, class> vecn<T, 4>::vecn(const vecn<U, 4>& o) : x(conv<T, U>(o.x)), y(conv<T, U>(o.y)), z(conv<T, U>(o.z)), w(conv<T, U>(o.w)) {}
template <class T> vecn<T, 2>::vecn(const vecn<T, 3>& o) : x(o.x), y(o.y) {}
template <class T> vecn<T, 2>::vecn(const vecn<T, 4>& o) : x(o.x), y(o.y) {}
template <class T> vecn<T, 3>::vecn(const vecn<T, 4>& o) : x(o.x), y(o.y), z(o.z) {}

template <class T, int N, int A, int B, int C, int D> swz<T, N, A, B, C, D>::operator vecn<T, N>() const {
    vecn<T, N> r;
    const int m[4] = {A, B, C, D};
    for (int i = 0; i < N; ++i) r[i] = p()[m[i]];
    return r;
}
template <class T, int N, int A, int B, int C, int D> swz<T, N, A, B, C, D>& swz<T, N, A, B, C, D>::operator=(const vecn<T, N>& v) {
    const int m[4] = {A, B, C, D};
    for (int i = 0; i < N; ++i) p()[m[i]] = v[i];
    return *this;
}

typedef vecn<float, 2> vec2; typedef vecn<float, 3> vec3; typedef vecn<float, 4> vec4;
typedef vecn<int, 2> ivec2; typedef vecn<int, 3> ivec3; typedef vecn<int, 4> ivec4;
typedef vecn<uint, 2> uvec2; typedef vecn<uint, 3> uvec3; typedef vecn<uint, 4> uvec4;
typedef vecn<bool, 2> bvec2; typedef vecn<bool, 3> bvec3; typedef vecn<bool, 4> bvec4;

/* ---- component-wise operators (non-template so swizzle proxies convert implicitly) ---- */
#define GLSL_BINOP(V, T, op)                                                                                      \
    inline V operator op(const V& a, const V& b) { V r; for (int i = 0; i < (int)(sizeof(V) / sizeof(T)); ++i) r[i] = a[i] op b[i]; return r; } \
    inline V operator op(const V& a, T b) { V r; for (int i = 0; i < (int)(sizeof(V) / sizeof(T)); ++i) r[i] = a[i] op b; return r; }         \
    inline V operator op(T a, const V& b) { V r; for (int i = 0; i < (int)(sizeof(V) / sizeof(T)); ++i) r[i] = a op b[i]; return r; }         \
    inline V& operator op##=(V& a, const V& b) { a = a op b; return a; }                                          \
    inline V& operator op##=(V& a, T b) { a = a op b; return a; }
#define GLSL_ARITH(V, T)                                                                   \
    GLSL_BINOP(V, T, +) GLSL_BINOP(V, T, -) GLSL_BINOP(V, T, *) GLSL_BINOP(V, T, /)          \
    inline V operator-(const V& a) { V r; for (int i = 0; i < (int)(sizeof(V) / sizeof(T)); ++i) r[i] = -a[i]; return r; } \
    inline bool operator==(const V& a, const V& b) { for (int i = 0; i < (int)(sizeof(V) / sizeof(T)); ++i) if (!(a[i] == b[i])) return false; return true; } \
    inline bool operator!=(const V& a, const V& b) { return !(a == b); }
GLSL_ARITH(vec2, float) GLSL_ARITH(vec3, float) GLSL_ARITH(vec4, float)
GLSL_ARITH(ivec2, int) GLSL_ARITH(ivec3, int) GLSL_ARITH(ivec4, int)
GLSL_ARITH(uvec2, uint) GLSL_ARITH(uvec3, uint) GLSL_ARITH(uvec4, uint)
/* mixed float / integer-vector arithmetic: GLSL converts the integer operand implicitly (ivecN -> vecN), e.g.
 * `vec2 TexelSize = 1.0f / textureSize(u_SH, 0)` (SVGF/TemporalFilter.glsl:146); without these overloads C++ would
 * convert the float to int instead */
#define GLSL_MIXOP(VF, VI, op)                                                                                          \
    inline VF operator op(float a, const VI& b) { VF r; for (int i = 0; i < (int)(sizeof(VI) / sizeof(int)); ++i) r[i] = a op (float)b[i]; return r; }       \
    inline VF operator op(const VI& a, float b) { VF r; for (int i = 0; i < (int)(sizeof(VI) / sizeof(int)); ++i) r[i] = (float)a[i] op b; return r; }       \
    inline VF operator op(const VF& a, const VI& b) { VF r; for (int i = 0; i < (int)(sizeof(VI) / sizeof(int)); ++i) r[i] = a[i] op (float)b[i]; return r; } \
    inline VF operator op(const VI& a, const VF& b) { VF r; for (int i = 0; i < (int)(sizeof(VI) / sizeof(int)); ++i) r[i] = (float)a[i] op b[i]; return r; }
#define GLSL_MIX(VF, VI) GLSL_MIXOP(VF, VI, +) GLSL_MIXOP(VF, VI, -) GLSL_MIXOP(VF, VI, *) GLSL_MIXOP(VF, VI, /)
GLSL_MIX(vec2, ivec2) GLSL_MIX(vec3, ivec3) GLSL_MIX(vec4, ivec4)
#define GLSL_INTOPS(V, T) GLSL_BINOP(V, T, %) GLSL_BINOP(V, T, >>) GLSL_BINOP(V, T, <<) GLSL_BINOP(V, T, &) GLSL_BINOP(V, T, |) GLSL_BINOP(V, T, ^)
GLSL_INTOPS(ivec2, int) GLSL_INTOPS(ivec3, int) GLSL_INTOPS(ivec4, int)
GLSL_INTOPS(uvec2, uint) GLSL_INTOPS(uvec3, uint) GLSL_INTOPS(uvec4, uint)
/* int scalars mixing with float vectors (GLSL converts implicitly): covered by int -> float on the
 * scalar overloads above.  vec op ivec is covered by the implicit ivec -> vec constructor.        */

/* ---- scalar built-ins (explicit so nothing falls through to <math.h>'s double versions) ---- */
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
inline float degrees(float r) { return r * 57.295779513082320876798154814105f; }
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float tan(float x) { return ::tanf(x); }
inline float asin(float x) { return ::asinf(x); }
inline float acos(float x) { return ::acosf(x); }
inline float atan(float x) { return ::atanf(x); }
inline float atan(float y, float x) { return ::atan2f(y, x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline float exp2(float x) { return ::exp2f(x); }
inline float log2(float x) { return ::log2f(x); }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float inversesqrt(float x) { return 1.0f / ::sqrtf(x); }
inline float abs(float x) { return ::fabsf(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float sign(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }
inline int sign(int x) { return (x > 0) - (x < 0); }
inline float floor(float x) { return ::floorf(x); }
inline float ceil(float x) { return ::ceilf(x); }
inline float trunc(float x) { return ::truncf(x); }
inline float round(float x) { return ::nearbyintf(x); } /* ties: pinned half-even */
inline float fract(float x) { return x - ::floorf(x); }
inline float mod(float x, float y) { return x - y * ::floorf(x / y); }
inline bool isnan(float x) { return x != x; }
inline bool isinf(float x) { return ::isinf(x); }

template <class A, class B> using arith2 = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value,
    std::conditional_t<std::is_floating_point<A>::value || std::is_floating_point<B>::value, float,
    std::conditional_t<std::is_unsigned<A>::value && std::is_unsigned<B>::value, uint, int>>>;
template <class A, class B> inline arith2<A, B> min(A a, B b) { typedef arith2<A, B> R; R x = (R)a, y = (R)b; return (y < x) ? y : x; }
template <class A, class B> inline arith2<A, B> max(A a, B b) { typedef arith2<A, B> R; R x = (R)a, y = (R)b; return (x < y) ? y : x; }
template <class A, class B, class C> inline arith2<arith2<A, B>, C> clamp(A x, B lo, C hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float mix(float a, float b, bool t) { return t ? b : a; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x) { float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f); return t * t * (3.0f - 2.0f * t); }
inline uint floatBitsToUint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline int floatBitsToInt(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float intBitsToFloat(int u) { float f; memcpy(&f, &u, 4); return f; }

/* ---- vector built-ins ---- */
#define GLSL_MAP1(V, f) inline V f(const V& a) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = f(a[i]); return r; }
#define GLSL_MAP2(V, f)                                                                                              \
    inline V f(const V& a, const V& b) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = f(a[i], b[i]); return r; }
#define GLSL_MAP2S(V, T, f)                                                                                           \
    inline V f(const V& a, T b) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = f(a[i], b); return r; }
#define GLSL_FVEC(V)                                                                                                  \
    GLSL_MAP1(V, radians) GLSL_MAP1(V, degrees) GLSL_MAP1(V, sin) GLSL_MAP1(V, cos) GLSL_MAP1(V, tan) GLSL_MAP1(V, asin)  \
    GLSL_MAP1(V, acos) GLSL_MAP1(V, atan) GLSL_MAP1(V, exp) GLSL_MAP1(V, log) GLSL_MAP1(V, exp2) GLSL_MAP1(V, log2)   \
    GLSL_MAP1(V, sqrt) GLSL_MAP1(V, inversesqrt) GLSL_MAP1(V, abs) GLSL_MAP1(V, sign) GLSL_MAP1(V, floor) GLSL_MAP1(V, ceil) \
    GLSL_MAP1(V, trunc) GLSL_MAP1(V, round) GLSL_MAP1(V, fract)                                                       \
    GLSL_MAP2(V, pow) GLSL_MAP2(V, mod) GLSL_MAP2(V, min) GLSL_MAP2(V, max) GLSL_MAP2(V, atan)                         \
    GLSL_MAP2S(V, float, mod) GLSL_MAP2S(V, float, min) GLSL_MAP2S(V, float, max) GLSL_MAP2S(V, float, pow)             \
    inline V step(const V& e, const V& x) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = step(e[i], x[i]); return r; } \
    inline V step(float e, const V& x) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = step(e, x[i]); return r; }       \
    inline V clamp(const V& x, const V& lo, const V& hi) { return min(max(x, lo), hi); }                             \
    inline V clamp(const V& x, float lo, float hi) { return min(max(x, lo), hi); }                                  \
    inline V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = mix(a[i], b[i], t); return r; } \
    inline V mix(const V& a, const V& b, const V& t) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = mix(a[i], b[i], t[i]); return r; } \
    inline V mix(const V& a, const V& b, bool t) { return t ? b : a; }                                               \
    inline V smoothstep(float e0, float e1, const V& x) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = smoothstep(e0, e1, x[i]); return r; } \
    inline V smoothstep(const V& e0, const V& e1, const V& x) { V r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = smoothstep(e0[i], e1[i], x[i]); return r; }
GLSL_FVEC(vec2) GLSL_FVEC(vec3) GLSL_FVEC(vec4)
#define GLSL_IVEC(V, T)                                                                                               \
    GLSL_MAP2(V, min) GLSL_MAP2(V, max) GLSL_MAP2S(V, T, min) GLSL_MAP2S(V, T, max)                                     \
    inline V clamp(const V& x, T lo, T hi) { return min(max(x, lo), hi); }                                           \
    inline V clamp(const V& x, const V& lo, const V& hi) { return min(max(x, lo), hi); }
GLSL_IVEC(ivec2, int) GLSL_IVEC(ivec3, int) GLSL_IVEC(ivec4, int)
GLSL_MAP1(ivec2, abs) GLSL_MAP1(ivec3, abs) GLSL_MAP1(ivec4, abs) GLSL_MAP1(ivec2, sign) GLSL_MAP1(ivec3, sign) GLSL_MAP1(ivec4, sign)

/* geometric functions: evaluation order pinned as in vxo_math.h */
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
inline float length(float a) { return ::fabsf(a); }
inline float length(const vec2& a) { return ::sqrtf(dot(a, a)); }
inline float length(const vec3& a) { return ::sqrtf(dot(a, a)); }
inline float length(const vec4& a) { return ::sqrtf(dot(a, a)); }
inline float distance(float a, float b) { return ::fabsf(b - a); }
inline float distance(const vec2& a, const vec2& b) { return length(b - a); }
inline float distance(const vec3& a, const vec3& b) { return length(b - a); }
inline vec2 normalize(const vec2& a) { return a * (1.0f / ::sqrtf(dot(a, a))); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / ::sqrtf(dot(a, a))); }
inline vec4 normalize(const vec4& a) { return a * (1.0f / ::sqrtf(dot(a, a))); }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline vec3 reflect(const vec3& I, const vec3& N) { return I - N * dot(N, I) * 2.0f; }
inline vec3 refract(const vec3& I, const vec3& N, float eta) {
    float d = dot(N, I), k = 1.0f - eta * eta * (1.0f - d * d);
    return k < 0.0f ? vec3(0.0f) : (I * eta - N * (eta * d + ::sqrtf(k)));
}
inline vec3 faceforward(const vec3& N, const vec3& I, const vec3& Nref) { return dot(Nref, I) < 0.0f ? N : -N; }

#define GLSL_REL(VB, V, name, op) inline VB name(const V& a, const V& b) { VB r; for (int i = 0; i < (int)(sizeof(V) / 4); ++i) r[i] = a[i] op b[i]; return r; }
#define GLSL_RELS(VB, V) GLSL_REL(VB, V, lessThan, <) GLSL_REL(VB, V, lessThanEqual, <=) GLSL_REL(VB, V, greaterThan, >) GLSL_REL(VB, V, greaterThanEqual, >=) GLSL_REL(VB, V, equal, ==) GLSL_REL(VB, V, notEqual, !=)
GLSL_RELS(bvec2, vec2) GLSL_RELS(bvec3, vec3) GLSL_RELS(bvec4, vec4) GLSL_RELS(bvec2, ivec2) GLSL_RELS(bvec3, ivec3) GLSL_RELS(bvec4, ivec4)
inline bool any(const bvec2& b) { return b.x || b.y; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }
inline bool any(const bvec4& b) { return b.x || b.y || b.z || b.w; }
inline bool all(const bvec2& b) { return b.x && b.y; }
inline bool all(const bvec3& b) { return b.x && b.y && b.z; }
inline bool all(const bvec4& b) { return b.x && b.y && b.z && b.w; }
inline bvec3 isnan(const vec3& v) { return bvec3(v.x != v.x, v.y != v.y, v.z != v.z); }
inline bvec4 isnan(const vec4& v) { return bvec4(v.x != v.x, v.y != v.y, v.z != v.z, v.w != v.w); }
inline bvec3 isinf(const vec3& v) { return bvec3(isinf(v.x), isinf(v.y), isinf(v.z)); }
inline float dFdx(float) { return 0.0f; }  /* screen-space derivatives are not emulated (POM only, off by default) */
inline float dFdy(float) { return 0.0f; }
inline vec2 dFdx(const vec2&) { return vec2(0.0f); }
inline vec2 dFdy(const vec2&) { return vec2(0.0f); }
inline vec3 dFdx(const vec3&) { return vec3(0.0f); }
inline vec3 dFdy(const vec3&) { return vec3(0.0f); }

/* ---- matrices (column-major) ---- */
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    mat3(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2) { c[0] = vec3(a0, a1, a2); c[1] = vec3(b0, b1, b2); c[2] = vec3(c0, c1, c2); }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float d) { c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d); }
    mat4(const vec4& a, const vec4& b, const vec4& d, const vec4& e) { c[0] = a; c[1] = b; c[2] = d; c[3] = e; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    void load(const float* m) { for (int i = 0; i < 4; ++i) c[i] = vec4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]); }
};
typedef mat3 mat3x3; typedef mat4 mat4x4;
inline mat3 to_mat3(const mat4& m) { return mat3(vec3(m[0]), vec3(m[1]), vec3(m[2])); }
/* mat4 * vec4: (m0*x + m1*y) + (m2*z + m3*w)  — the association pinned in vxo_math.h */
inline vec4 operator*(const mat4& m, const vec4& v) { return (m[0] * v.x + m[1] * v.y) + (m[2] * v.z + m[3] * v.w); }
inline vec3 operator*(const mat3& m, const vec3& v) {
    return vec3(m[0].x * v.x + m[1].x * v.y + m[2].x * v.z, m[0].y * v.x + m[1].y * v.y + m[2].y * v.z, m[0].z * v.x + m[1].z * v.y + m[2].z * v.z);
}
inline vec3 operator*(const vec3& v, const mat3& m) { return vec3(dot(v, m[0]), dot(v, m[1]), dot(v, m[2])); }
inline mat4 operator*(const mat4& a, const mat4& b) { return mat4(a * b[0], a * b[1], a * b[2], a * b[3]); }
inline mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b[0], a * b[1], a * b[2]); }
inline mat3 transpose(const mat3& m) { return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z)); }
inline mat3 inverse(const mat3& m) {
    vec3 a = m[0], b = m[1], c = m[2];
    vec3 r0 = cross(b, c), r1 = cross(c, a), r2 = cross(a, b);
    float inv = 1.0f / dot(r2, c);
    return transpose(mat3(r0 * inv, r1 * inv, r2 * inv));
}

/* fixed-size arrays as function results (GLSL `float[6] f()`) */
template <class T, int N> struct arr {
    T v[N];
    arr() {}
    arr(const T (&a)[N]) { for (int i = 0; i < N; ++i) v[i] = a[i]; }
    template <class... A, class = std::enable_if_t<sizeof...(A) == N>> arr(A... a) : v{T(a)...} {}
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
};

/* ---- resources ---- */
struct sampler3D { const uint8_t* data = nullptr; int w = 0, h = 0, d = 0; };     /* R8 unorm, NEAREST (Texture3D.cpp:20-27) */
struct image3D { uint8_t* data = nullptr; int w = 0, h = 0, d = 0; };             /* r8 image */
struct usampler3D { const uint8_t* data = nullptr; int w = 0, h = 0, d = 0; };    /* LPV block types: R8UI, NEAREST, CLAMP_TO_EDGE (VolumetricFloodFill.cpp:51-58) */
typedef vxo::Tex2D sampler2D;
struct sampler2DArray { const vxo::TexArray* t = nullptr; };
typedef vxo::TexCube samplerCube;

inline vec4 texelFetch(const sampler3D& s, const ivec3& p, int) {
    return vec4(vxo::unorm8_to_float(s.data[p.x + (size_t)p.y * s.w + (size_t)p.z * s.w * s.h]), 0.0f, 0.0f, 1.0f);
}
inline vec4 imageLoad(const image3D& s, const ivec3& p) {
    return vec4(vxo::unorm8_to_float(s.data[p.x + (size_t)p.y * s.w + (size_t)p.z * s.w * s.h]), 0.0f, 0.0f, 1.0f);
}
inline void imageStore(image3D& s, const ivec3& p, const vec4& v) {
    s.data[p.x + (size_t)p.y * s.w + (size_t)p.z * s.w * s.h] = vxo::float_to_unorm8(v.x);
}
/* filtered reads of 3-D textures only occur on the LPV path: the light-level volume is R8 unorm, LINEAR, CLAMP_TO_EDGE
 * (VolumetricFloodFill.cpp:41-48); texel centres at +0.5, weights in full float, x then y then z like the 2-D model of vxo_texture.h.
 * An unbound sampler reads 0. */
inline int clamp_texel(int i, int n) { return i < 0 ? 0 : (i > n - 1 ? n - 1 : i); }
inline vec4 texture(const sampler3D& s, const vec3& c) {
    if (!s.data) return vec4(0.0f);
    const float u = c.x * (float)s.w - 0.5f, v = c.y * (float)s.h - 0.5f, w = c.z * (float)s.d - 0.5f;
    const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
    const float a = u - fu, b = v - fv, g = w - fw;
    const int i0 = clamp_texel(vxo::cvt_floor(fu), s.w), i1 = clamp_texel(vxo::cvt_floor(fu) + 1, s.w);
    const int j0 = clamp_texel(vxo::cvt_floor(fv), s.h), j1 = clamp_texel(vxo::cvt_floor(fv) + 1, s.h);
    const int k0 = clamp_texel(vxo::cvt_floor(fw), s.d), k1 = clamp_texel(vxo::cvt_floor(fw) + 1, s.d);
    auto at = [&](int i, int j, int k) { return vxo::unorm8_to_float(s.data[i + (size_t)j * s.w + (size_t)k * s.w * s.h]); };
    auto plane = [&](int k) {
        const float top = at(i0, j0, k) * (1.0f - a) + at(i1, j0, k) * a;
        const float bot = at(i0, j1, k) * (1.0f - a) + at(i1, j1, k) * a;
        return top * (1.0f - b) + bot * b;
    };
    return vec4(plane(k0) * (1.0f - g) + plane(k1) * g, 0.0f, 0.0f, 1.0f);
}
inline uvec4 texture(const usampler3D& s, const vec3& c) {
    if (!s.data) return uvec4(0u);
    const int i = clamp_texel(vxo::cvt_floor(c.x * (float)s.w), s.w), j = clamp_texel(vxo::cvt_floor(c.y * (float)s.h), s.h);
    const int k = clamp_texel(vxo::cvt_floor(c.z * (float)s.d), s.d);
    return uvec4((unsigned)s.data[i + (size_t)j * s.w + (size_t)k * s.w * s.h], 0u, 0u, 1u);
}
inline uvec4 texelFetch(const usampler3D& s, const ivec3& p, int) {
    if (!s.data) return uvec4(0u);
    return uvec4((unsigned)s.data[p.x + (size_t)p.y * s.w + (size_t)p.z * s.w * s.h], 0u, 0u, 1u);
}
inline vec4 to4(const vxo::v4& v) { return vec4(v.x, v.y, v.z, v.w); }
inline vec4 texture(const sampler2D& s, const vec2& uv) { return to4(vxo::tex2d_sample(s, uv.x, uv.y)); }
inline vec4 textureLod(const sampler2D& s, const vec2& uv, float) { return to4(vxo::tex2d_sample(s, uv.x, uv.y)); }
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int) { return to4(vxo::tex2d_fetch(s, p.x, p.y)); }
inline ivec2 textureSize(const sampler2D& s, int) { return ivec2(s.w, s.h); }
inline vec4 textureLod(const sampler2DArray& s, const vec3& c, float lod) { return to4(vxo::texarray_sample(*s.t, c.x, c.y, c.z, lod)); }
inline vec4 texture(const sampler2DArray& s, const vec3& c) { return to4(vxo::texarray_sample(*s.t, c.x, c.y, c.z, 0.0f)); } /* implicit LOD pinned to 0 */
inline vec4 texture(const sampler2DArray& s, const vec3& c, float bias) { return to4(vxo::texarray_sample(*s.t, c.x, c.y, c.z, bias)); }
inline vec4 textureGrad(const sampler2DArray& s, const vec3& c, const vec2&, const vec2&) { return to4(vxo::texarray_sample(*s.t, c.x, c.y, c.z, 0.0f)); }
inline ivec3 textureSize(const sampler2DArray& s, int) { return ivec3(s.t->w, s.t->h, s.t->layers); }
inline vec4 texture(const samplerCube& s, const vec3& d) { return to4(vxo::texcube_sample(s, d.x, d.y, d.z)); }
inline vec4 textureLod(const samplerCube& s, const vec3& d, float) { return to4(vxo::texcube_sample(s, d.x, d.y, d.z)); }

/* SSBO arrays: out-of-bounds reads return 0 (pinned) */
template <class T, int N> struct ssbo_array {
    const T* data = nullptr;
    T operator[](int i) const { return (i >= 0 && i < N && data) ? data[i] : T(0); }
};
template <class T> struct ssbo_unsized {
    const T* data = nullptr; int n = 0;
    T operator[](int i) const { return (i >= 0 && i < n && data) ? data[i] : T(); }
};

/* per-invocation built-in variables */
extern thread_local vec4 gl_FragCoord;
extern thread_local uvec3 gl_GlobalInvocationID;

}  // namespace glsl
