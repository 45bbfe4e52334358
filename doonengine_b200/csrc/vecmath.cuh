/* vecmath.cuh -- float3 helpers with the exact evaluation order the parity rules fix (oracle.h N6):
 * dot = x*x + y*y + z*z left to right, normalize(v) = v * (1/sqrt(dot)), reflect = I - 2*dot(N,I)*N,
 * mix(a,b,t) = a*(1-t) + b*t, min/max = IEEE minNum/maxNum.  The translation unit is compiled with
 * -fmad=false (no contraction), IEEE division and square root, no flush-to-zero.
 */
#ifndef DN_B200_VECMATH_CUH
#define DN_B200_VECMATH_CUH

#include <cuda_runtime.h>

struct f3 { float x, y, z; };
struct i3 { int x, y, z; };

/* A host harness (tests/csrc/ray_step_harness.cu) may define DNB_FN as __host__ __device__ to run the traversal code on the CPU; the read-only-load and
 * bit intrinsics go through these wrappers for the same reason (device code is unchanged by them). */
#ifndef DNB_FN
#define DNB_FN __device__ __forceinline__
#endif
#ifdef __CUDA_ARCH__
#define DNB_LDG(p) __ldg(p)
#define DNB_POPC(x) __popc(x)
#define DNB_U2F(x) __uint_as_float(x)
#define DNB_F2U(x) __float_as_uint(x)
#else
#include <string.h>
template <typename T> static inline T dnb_host_load(const T* p) { return *p; }
static inline float dnb_host_u2f(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned dnb_host_f2u(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
#define DNB_LDG(p) dnb_host_load(p)
#define DNB_POPC(x) __builtin_popcount(x)
#define DNB_U2F(x) dnb_host_u2f(x)
#define DNB_F2U(x) dnb_host_f2u(x)
#endif

DNB_FN f3 mk3(float x, float y, float z)  { f3 r; r.x = x; r.y = y; r.z = z; return r; }
DNB_FN f3 splat3(float s)                 { return mk3(s, s, s); }
DNB_FN f3 ld3(const float* p)             { return mk3(p[0], p[1], p[2]); }
DNB_FN f3 tof3(i3 a)                      { return mk3((float)a.x, (float)a.y, (float)a.z); }
DNB_FN f3 operator+(f3 a, f3 b)           { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
DNB_FN f3 operator-(f3 a, f3 b)           { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
DNB_FN f3 operator*(f3 a, f3 b)           { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
DNB_FN f3 operator*(f3 a, float s)        { return mk3(a.x * s, a.y * s, a.z * s); }
DNB_FN f3 operator+(f3 a, float s)        { return mk3(a.x + s, a.y + s, a.z + s); }
DNB_FN f3 operator-(f3 a)                 { return mk3(-a.x, -a.y, -a.z); }
DNB_FN f3 div3(f3 a, float s)             { return mk3(a.x / s, a.y / s, a.z / s); }
DNB_FN float dot3(f3 a, f3 b)             { return a.x * b.x + a.y * b.y + a.z * b.z; }
DNB_FN f3 rcp3(f3 a)                      { return mk3(1.0f / a.x, 1.0f / a.y, 1.0f / a.z); }
DNB_FN f3 abs3(f3 a)                      { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
DNB_FN f3 floor3(f3 a)                    { return mk3(floorf(a.x), floorf(a.y), floorf(a.z)); }
DNB_FN f3 trunc3(f3 a)                    { return mk3(truncf(a.x), truncf(a.y), truncf(a.z)); }
DNB_FN f3 min3v(f3 a, f3 b)               { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
DNB_FN f3 max3v(f3 a, f3 b)               { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
DNB_FN float hmin3(f3 a)                  { return fminf(fminf(a.x, a.y), a.z); }
DNB_FN float sgn(float a)                 { return (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f); }
DNB_FN f3 normalize3(f3 a)                { float inv = 1.0f / sqrtf(dot3(a, a)); return a * inv; }
DNB_FN f3 reflect3(f3 I, f3 N)            { float d = dot3(N, I); return I - N * (2.0f * d); }
DNB_FN f3 clamp01(f3 a)                   { return min3v(max3v(a, splat3(0.0f)), splat3(1.0f)); }
DNB_FN i3 toi3(f3 a)                      { i3 r; r.x = (int)a.x; r.y = (int)a.y; r.z = (int)a.z; return r; }

/* GLSL 4.30 refract(), section 8.5 */
DNB_FN f3 refract3(f3 I, f3 N, float eta)
{
	float d = dot3(N, I);
	float k = 1.0f - eta * eta * (1.0f - d * d);
	if(k < 0.0f)
		return splat3(0.0f);
	return I * eta - N * (eta * d + sqrtf(k));
}

/* column-major 4x4 times vec4, terms added left to right */
DNB_FN void mat4_mul_vec4(const float* m, const float* v, float* out)
{
#pragma unroll
	for(int r = 0; r < 4; r++)
		out[r] = m[0 * 4 + r] * v[0] + m[1 * 4 + r] * v[1] + m[2 * 4 + r] * v[2] + m[3 * 4 + r] * v[3];
}

#endif
