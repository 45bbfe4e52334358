/* DoonEngine/mathtypes.h -- the float vector / matrix TYPES that appear in the DN_* signatures.
 *
 * The reference takes these from its vendored QuickMath header
 * (/root/reference/dependencies/include/QuickMath/quickmath.h:111-163, typedef'd to DN* at :180-185).
 * Only the types cross the ABI, so only the types are declared here; they are layout-compatible:
 *   DNvec2  8 B, DNvec3 12 B, DNvec4 16 B (16-byte aligned: QuickMath unions it with __m128),
 *   DNmat3 36 B, DNmat4 64 B column-major m[col][row], 16-byte aligned and passed BY VALUE to DN_draw.
 * An application that already includes the real QuickMath (QM_MATH_H defined, prefix DN_) keeps its own
 * definitions and this header adds nothing.
 */
#ifndef DN_MATHTYPES_H
#define DN_MATHTYPES_H

#ifndef QM_MATH_H

#ifdef __cplusplus
extern "C" {
#endif

typedef union DNvec2
{
	float v[2];
	struct { float x, y; };
} DNvec2;

typedef union DNvec3
{
	float v[3];
	struct { float x, y, z; };
	struct { float r, g, b; };
} DNvec3;

typedef union
#if defined(__GNUC__)
__attribute__((aligned(16)))
#endif
DNvec4
{
	float v[4];
	struct { float x, y, z, w; };
} DNvec4;

typedef union DNmat3
{
	float m[3][3];
} DNmat3;

typedef union
#if defined(__GNUC__)
__attribute__((aligned(16)))
#endif
DNmat4
{
	float m[4][4]; /* m[column][row] */
} DNmat4;

#ifdef __cplusplus
}
#endif

#endif /* QM_MATH_H */

#endif
