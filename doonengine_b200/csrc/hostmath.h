/* hostmath.h -- the handful of host-side float routines on the path.  They have to reproduce the reference's
 * results bit for bit because their outputs (view / projection / inverse matrices, normalised sun direction,
 * the per-dispatch random table) are inputs of both the kernels and the oracle.
 *
 * Arithmetic follows /root/reference/dependencies/include/QuickMath/quickmath.h (QM):
 *   rotate_euler QM:1214-1243, lookat QM:1301-1333, perspective QM:1264-1283, mat4_inv QM:1037-1124,
 *   vec3_normalize QM:538-552; DN_set_view_projection_matrices itself is voxel.c:788-810.
 * Built with -ffp-contract=off; x86-64 baseline has no FMA.
 */
#ifndef DN_B200_HOSTMATH_H
#define DN_B200_HOSTMATH_H

#include <math.h>
#include <string.h>

namespace dnb
{

struct Mat4 { float m[4][4]; }; /* m[column][row] */

inline float deg2rad(float d) { return d * 0.01745329251f; }

inline Mat4 identity()
{
	Mat4 r;
	memset(&r, 0, sizeof(r));
	for(int i = 0; i < 4; i++)
		r.m[i][i] = 1.0f;
	return r;
}

inline void normalize(float v[3])
{
	float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	if(len != 0.0f)
	{
		float inv = 1.0f / len;
		v[0] *= inv;
		v[1] *= inv;
		v[2] *= inv;
	}
	else
		v[0] = v[1] = v[2] = 0.0f;
}

inline void cross(const float a[3], const float b[3], float out[3])
{
	out[0] = (a[1] * b[2]) - (a[2] * b[1]);
	out[1] = (a[2] * b[0]) - (a[0] * b[2]);
	out[2] = (a[0] * b[1]) - (a[1] * b[0]);
}

/* each output element is the left-to-right sum over k of a[k][row] * b[col][k] */
inline Mat4 mul(const Mat4& a, const Mat4& b)
{
	Mat4 r;
	for(int c = 0; c < 4; c++)
		for(int row = 0; row < 4; row++)
			r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2] + a.m[3][row] * b.m[c][3];
	return r;
}

/* top-left 3x3 of the pitch/yaw/roll matrix, R[col][row] */
inline void euler3x3(const float orientDeg[3], float R[3][3])
{
	float rx = deg2rad(orientDeg[0]), ry = deg2rad(orientDeg[1]), rz = deg2rad(orientDeg[2]);
	float sX = sinf(rx), cX = cosf(rx), sY = sinf(ry), cY = cosf(ry), sZ = sinf(rz), cZ = cosf(rz);
	R[0][0] = cY * cZ;
	R[0][1] = cY * sZ;
	R[0][2] = -sY;
	R[1][0] = sX * sY * cZ - cX * sZ;
	R[1][1] = sX * sY * sZ + cX * cZ;
	R[1][2] = sX * cY;
	R[2][0] = cX * sY * cZ + sX * sZ;
	R[2][1] = cX * sY * sZ - sX * cZ;
	R[2][2] = cX * cY;
}

inline void mat3_mul_vec3(const float R[3][3], const float v[3], float out[3])
{
	out[0] = R[0][0] * v[0] + R[1][0] * v[1] + R[2][0] * v[2];
	out[1] = R[0][1] * v[0] + R[1][1] * v[1] + R[2][1] * v[2];
	out[2] = R[0][2] * v[0] + R[1][2] * v[1] + R[2][2] * v[2];
}

inline Mat4 lookat(const float pos[3], const float target[3])
{
	float dir[3] = {pos[0] - target[0], pos[1] - target[1], pos[2] - target[2]};
	normalize(dir);
	const float up[3] = {0.0f, 1.0f, 0.0f};
	float r[3], u[3];
	cross(up, dir, r);
	normalize(r);
	cross(dir, r, u);

	Mat4 basis = identity();
	for(int i = 0; i < 3; i++)
	{
		basis.m[i][0] = r[i];
		basis.m[i][1] = u[i];
		basis.m[i][2] = dir[i];
	}
	Mat4 shift = identity();
	shift.m[3][0] = -pos[0];
	shift.m[3][1] = -pos[1];
	shift.m[3][2] = -pos[2];
	return mul(basis, shift);
}

inline Mat4 perspective(float fovDeg, float aspect, float nearPlane, float farPlane)
{
	Mat4 P;
	memset(&P, 0, sizeof(P));
	float scale = tanf(deg2rad(fovDeg * 0.5f)) * nearPlane;
	float right = aspect * scale;
	float top = scale;
	P.m[0][0] = nearPlane / right;
	P.m[1][1] = nearPlane / top;
	P.m[2][2] = -(farPlane + nearPlane) / (farPlane - nearPlane);
	P.m[3][2] = -2.0f * farPlane * nearPlane / (farPlane - nearPlane);
	P.m[2][3] = -1.0f;
	return P;
}

/* cofactor inverse; the 2x2 sub-determinants are shared between the columns exactly as QM:1037-1124 shares them */
inline Mat4 inverse(const Mat4& M)
{
	const float a = M.m[0][0], b = M.m[0][1], c = M.m[0][2], d = M.m[0][3];
	const float e = M.m[1][0], f = M.m[1][1], g = M.m[1][2], h = M.m[1][3];
	const float i = M.m[2][0], j = M.m[2][1], k = M.m[2][2], l = M.m[2][3];
	const float m = M.m[3][0], n = M.m[3][1], o = M.m[3][2], p = M.m[3][3];
	Mat4 r;
	float t0, t1, t2, t3, t4, t5;

	t0 = k * p - o * l; t1 = j * p - n * l; t2 = j * o - n * k; t3 = i * p - m * l; t4 = i * o - m * k; t5 = i * n - m * j;
	r.m[0][0] =   f * t0 - g * t1 + h * t2;
	r.m[1][0] = -(e * t0 - g * t3 + h * t4);
	r.m[2][0] =   e * t1 - f * t3 + h * t5;
	r.m[3][0] = -(e * t2 - f * t4 + g * t5);
	r.m[0][1] = -(b * t0 - c * t1 + d * t2);
	r.m[1][1] =   a * t0 - c * t3 + d * t4;
	r.m[2][1] = -(a * t1 - b * t3 + d * t5);
	r.m[3][1] =   a * t2 - b * t4 + c * t5;

	t0 = g * p - o * h; t1 = f * p - n * h; t2 = f * o - n * g; t3 = e * p - m * h; t4 = e * o - m * g; t5 = e * n - m * f;
	r.m[0][2] =   b * t0 - c * t1 + d * t2;
	r.m[1][2] = -(a * t0 - c * t3 + d * t4);
	r.m[2][2] =   a * t1 - b * t3 + d * t5;
	r.m[3][2] = -(a * t2 - b * t4 + c * t5);

	t0 = g * l - k * h; t1 = f * l - j * h; t2 = f * k - j * g; t3 = e * l - i * h; t4 = e * k - i * g; t5 = e * j - i * f;
	r.m[0][3] = -(b * t0 - c * t1 + d * t2);
	r.m[1][3] =   a * t0 - c * t3 + d * t4;
	r.m[2][3] = -(a * t1 - b * t3 + d * t5);
	r.m[3][3] =   a * t2 - b * t4 + c * t5;

	const float det = 1.0f / (a * r.m[0][0] + b * r.m[1][0] + c * r.m[2][0] + d * r.m[3][0]);
	for(int cc = 0; cc < 4; cc++)
		for(int rr = 0; rr < 4; rr++)
			r.m[cc][rr] = r.m[cc][rr] * det;
	return r;
}

/* voxelLighting.comp:29-32 with libm sinf (oracle.h N5) */
inline float shader_rand(float seed)
{
	float s = sinf(seed) * 43758.5453f;
	return (s - floorf(s)) * 2.0f - 1.0f;
}

/* voxelLighting.comp:47-57: rejection-sample the unit ball, seed advancing by 1 per try (1024-try guard, N11) */
inline void shader_rand_unit_sphere(float seed, float out[3])
{
	for(int tries = 0; tries < 1024; tries++)
	{
		out[0] = shader_rand(seed);
		out[1] = shader_rand(seed * 2.0f);
		out[2] = shader_rand(seed * 3.0f);
		seed = seed + 1.0f;
		if(out[0] * out[0] + out[1] * out[1] + out[2] * out[2] >= 1.0f)
			continue;
		return;
	}
}

} // namespace dnb

#endif
