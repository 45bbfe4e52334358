/* glsl_compat.h -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Just enough of GLSL 4.30 in C++ for the reference's three compute shaders (assets/shaders/voxelShared.comp, voxelLighting.comp,
 * voxelDraw.comp) to be compiled AS THEY ARE by g++: oracle/glsl/translate.py reads them where they lie under /root/reference at
 * build time, does the purely syntactic part (uniform / buffer declarations out, `out` / `inout` parameters to references, float
 * literals, `.xyz` swizzles to `.xyz()`), and wraps the text in a class; this header supplies the vector types, the operators with
 * GLSL's implicit int -> uint -> float conversions, and the built-in functions.  Nothing of the shaders is copied into the repo:
 * the generated source and the library live under oracle/_ref/ (git-ignored).
 *
 * Built-ins whose precision GLSL leaves to the implementation are defined as oracle.h N5/N6 defines them (they are what pins the
 * oracle and the CUDA kernels to each other; with them fixed, THIS file pins the hand restatement oracle/shader_cpu.c to the
 * shader text):  sin = libm sinf;  normalize(v) = v * (1 / sqrt(dot(v, v)));  dot = x*x + y*y + z*z left to right;
 * min / max = fminf / fmaxf;  round = nearest even;  reflect = I - 2 dot(N, I) N;  refract per the specification's formula;
 * mix(a, b, t) = a (1 - t) + b t;  pow = powf;  sign(NaN) = 0;  clamp(x, a, b) = min(max(x, a), b);  length = sqrt(dot);
 * matrix products add their terms left to right.  Compiled with -ffp-contract=off -fno-fast-math.
 */
#ifndef DN_GLSL_COMPAT_H
#define DN_GLSL_COMPAT_H

#include <math.h>
#include <stdint.h>

namespace glsl
{

typedef uint32_t uint;

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct uvec2; struct uvec3; struct uvec4;

struct bvec3 { bool x, y, z; };

struct uvec2 { uint x, y; uvec2() {} uvec2(uint a, uint b) : x(a), y(b) {} };
struct ivec2
{
	int x, y;
	ivec2() {}
	ivec2(int a, int b) : x(a), y(b) {}
	explicit ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}
};
struct vec2
{
	float x, y;
	vec2() {}
	vec2(float a, float b) : x(a), y(b) {}
	explicit vec2(float s) : x(s), y(s) {}
	vec2(const ivec2& v) : x((float)v.x), y((float)v.y) {} /* implicit, as in GLSL */
};

struct uvec3
{
	uint x, y, z;
	uvec3() {}
	uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
	explicit uvec3(const vec3& v);
	uvec2 xy() const { return uvec2(x, y); }
	uvec3 xyz() const { return *this; }
};
struct ivec3
{
	int x, y, z;
	ivec3() {}
	ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
	explicit ivec3(const vec3& v);
	explicit ivec3(const bvec3& m) : x(m.x), y(m.y), z(m.z) {}
	ivec3 xyz() const { return *this; }
};
struct vec3
{
	float x, y, z;
	vec3() {}
	vec3(float a, float b, float c) : x(a), y(b), z(c) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(const ivec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {} /* implicit */
	vec3(const uvec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {} /* implicit */
	explicit vec3(const bvec3& m) : x(m.x ? 1.0f : 0.0f), y(m.y ? 1.0f : 0.0f), z(m.z ? 1.0f : 0.0f) {}
	vec3(uint a, const uvec2& b) : x((float)a), y((float)b.x), z((float)b.y) {}
	vec3 xyz() const { return *this; }
	vec3 yzx() const { return vec3(y, z, x); }
	vec3 zxy() const { return vec3(z, x, y); }
	vec2 xy() const { return vec2(x, y); }
	vec2 yz() const { return vec2(y, z); }
};
inline uvec3::uvec3(const vec3& v) : x((uint)v.x), y((uint)v.y), z((uint)v.z) {}
inline ivec3::ivec3(const vec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}

struct uvec4
{
	uint x, y, z, w;
	uvec4() {}
	uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
	uvec4(const vec3& a, float d) : x((uint)a.x), y((uint)a.y), z((uint)a.z), w((uint)d) {}
	uvec4(const vec2& a, uint c, uint d) : x((uint)a.x), y((uint)a.y), z(c), w(d) {}
	uvec3 xyz() const { return uvec3(x, y, z); }
	uvec3 yzw() const { return uvec3(y, z, w); }
	uvec2 xy() const { return uvec2(x, y); }
};
struct vec4
{
	float x, y, z, w;
	vec4() {}
	vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
	vec4(const vec2& a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
	vec4(const vec3& a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
	vec3 xyz() const { return vec3(x, y, z); }
};

/* ---- vec3 ---- */
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(float s, const vec3& a) { return vec3(s + a.x, s + a.y, s + a.z); }
inline vec3 operator-(float s, const vec3& a) { return vec3(s - a.x, s - a.y, s - a.z); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, const vec3& b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, const vec3& b) { a = a - b; return a; }
inline vec3& operator*=(vec3& a, const vec3& b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* ---- vec2 / vec4 ---- */
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator-(const vec2& a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
inline vec4& operator/=(vec4& a, float s) { a = vec4(a.x / s, a.y / s, a.z / s, a.w / s); return a; }

/* ---- ivec3 ---- */
inline ivec3 operator+(const ivec3& a, const ivec3& b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ivec3 operator*(const ivec3& a, const ivec3& b) { return ivec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline ivec3 operator-(const ivec3& a) { return ivec3(-a.x, -a.y, -a.z); }
inline ivec3& operator+=(ivec3& a, const ivec3& b) { a = a + b; return a; }
/* mixed ivec3 / float forms the shaders use: GLSL converts the integer operand */
inline vec3 operator-(const ivec3& a, float s) { return vec3(a) - s; }
inline vec3 operator*(const ivec3& a, float s) { return vec3(a) * s; }

/* ---- uvec3 / uvec4 bit operations ---- */
inline uvec3 operator<<(const uvec3& a, int s) { return uvec3(a.x << s, a.y << s, a.z << s); }
inline uvec3 operator>>(const uvec3& a, int s) { return uvec3(a.x >> s, a.y >> s, a.z >> s); }
inline uvec3 operator|(const uvec3& a, const uvec3& b) { return uvec3(a.x | b.x, a.y | b.y, a.z | b.z); }
inline uvec3 operator&(const uvec3& a, uint m) { return uvec3(a.x & m, a.y & m, a.z & m); }
inline uvec4 operator&(const uvec4& a, uint m) { return uvec4(a.x & m, a.y & m, a.z & m, a.w & m); }

/* ---- 4x4 matrices, column-major (m[col * 4 + row]); products add their terms left to right ---- */
struct mat4 { float m[16]; };
inline vec4 operator*(const mat4& a, const vec4& v)
{
	float o[4];
	const float in[4] = {v.x, v.y, v.z, v.w};
	for(int r = 0; r < 4; r++)
		o[r] = a.m[0 * 4 + r] * in[0] + a.m[1 * 4 + r] * in[1] + a.m[2 * 4 + r] * in[2] + a.m[3 * 4 + r] * in[3];
	return vec4(o[0], o[1], o[2], o[3]);
}
inline mat4 operator*(const mat4& a, const mat4& b)
{
	mat4 o;
	for(int c = 0; c < 4; c++)
		for(int r = 0; r < 4; r++)
			o.m[c * 4 + r] = a.m[0 * 4 + r] * b.m[c * 4 + 0] + a.m[1 * 4 + r] * b.m[c * 4 + 1] + a.m[2 * 4 + r] * b.m[c * 4 + 2] + a.m[3 * 4 + r] * b.m[c * 4 + 3];
	return o;
}

/* ---- built-in functions ---- */
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline uint  min(uint a, uint b) { return a < b ? a : b; }
inline vec3  min(const vec3& a, const vec3& b) { return vec3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
inline vec3  max(const vec3& a, const vec3& b) { return vec3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
inline vec3  abs(const vec3& a) { return vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
inline vec3  floor(const vec3& a) { return vec3(floorf(a.x), floorf(a.y), floorf(a.z)); }
inline vec3  trunc(const vec3& a) { return vec3(truncf(a.x), truncf(a.y), truncf(a.z)); }
inline float fract(float a) { return a - floorf(a); }
inline float sin(float a) { return sinf(a); }
inline float sign1(float a) { return (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f); }
inline vec3  sign(const vec3& a) { return vec3(sign1(a.x), sign1(a.y), sign1(a.z)); }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(const vec3& a) { return sqrtf(dot(a, a)); }
inline float distance(const vec3& a, const vec3& b) { return length(a - b); }
inline vec3  normalize(const vec3& a) { const float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline vec3  reflect(const vec3& I, const vec3& N) { const float d = dot(N, I); return I - N * (2.0f * d); }
inline vec3  refract(const vec3& I, const vec3& N, float eta)
{
	const float d = dot(N, I);
	const float k = 1.0f - eta * eta * (1.0f - d * d);
	if(k < 0.0f)
		return vec3(0.0f);
	return I * eta - N * (eta * d + sqrtf(k));
}
inline vec3  mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  pow(const vec3& a, const vec3& b) { return vec3(powf(a.x, b.x), powf(a.y, b.y), powf(a.z, b.z)); }
inline vec3  clamp(const vec3& x, const vec3& lo, const vec3& hi) { return min(max(x, lo), hi); }
inline float round(float a) { return rintf(a); }
inline vec2  round(const vec2& a) { return vec2(rintf(a.x), rintf(a.y)); }
inline vec3  round(const vec3& a) { return vec3(rintf(a.x), rintf(a.y), rintf(a.z)); }
inline bvec3 lessThanEqual(const vec3& a, const vec3& b) { bvec3 r; r.x = a.x <= b.x; r.y = a.y <= b.y; r.z = a.z <= b.z; return r; }
inline int   bitCount(uint a) { return __builtin_popcount(a); }

/* ---- opaque types; nothing of them is ever sampled in the configurations the oracle covers (useCubemap = composeRasterized = false) ---- */
struct samplerCube {};
struct sampler2D {};
inline vec4  texture(const samplerCube&, const vec3&) { return vec4(0.0f, 0.0f, 0.0f, 0.0f); }
inline vec4  texture(const sampler2D&, const vec2&) { return vec4(0.0f, 0.0f, 0.0f, 0.0f); }
inline ivec2 textureSize(const sampler2D&, int) { return ivec2(1, 1); }

} // namespace glsl

#endif
