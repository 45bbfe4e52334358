/* shader_cpu.c -- TEST INFRASTRUCTURE ONLY ("Oracle B").
 *
 * Scalar, strict-IEEE CPU restatement of the reference's three compute
 * shaders, statement by statement:
 *   SH = /root/reference/assets/shaders/voxelShared.comp
 *   LI = /root/reference/assets/shaders/voxelLighting.comp
 *   DR = /root/reference/assets/shaders/voxelDraw.comp
 * Every function cites the lines it follows.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp
 * so that no multiply-add is fused and libm min/max/rint keep IEEE meaning.
 * See oracle.h for the determinism rules (N1..N11) and the parity status
 * ("parity unpinned by the reference" for the GLSL arithmetic).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* small vector layer (GLSL semantics, N6)                              */

typedef struct { float x, y, z; } v3;
typedef struct { int   x, y, z; } i3;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 v3s(float s)                  { v3 r = {s, s, s}; return r; }
static inline v3 v3i(i3 a)                     { v3 r = {(float)a.x, (float)a.y, (float)a.z}; return r; }
static inline v3 vadd(v3 a, v3 b)              { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b)              { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b)              { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s)         { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 vaddf(v3 a, float s)          { return V3(a.x + s, a.y + s, a.z + s); }
static inline v3 vneg(v3 a)                    { return V3(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b)           { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 vrcp(v3 a)                    { return V3(1.0f / a.x, 1.0f / a.y, 1.0f / a.z); }
static inline v3 vabs(v3 a)                    { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline v3 vfloor(v3 a)                  { return V3(floorf(a.x), floorf(a.y), floorf(a.z)); }
static inline v3 vtrunc(v3 a)                  { return V3(truncf(a.x), truncf(a.y), truncf(a.z)); }
static inline v3 vmin(v3 a, v3 b)              { return V3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static inline v3 vmax(v3 a, v3 b)              { return V3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static inline float fsign(float a)             { return (a > 0.0f) ? 1.0f : ((a < 0.0f) ? -1.0f : 0.0f); }
static inline v3 vsign(v3 a)                   { return V3(fsign(a.x), fsign(a.y), fsign(a.z)); }
static inline v3 vnormalize(v3 a)              { float inv = 1.0f / sqrtf(vdot(a, a)); return vscale(a, inv); }
static inline v3 vreflect(v3 I, v3 N)          { float d = vdot(N, I); return vsub(I, vscale(N, 2.0f * d)); }
static inline v3 vclamp01(v3 a)                { return vmin(vmax(a, v3s(0.0f)), v3s(1.0f)); }
static inline float min3(v3 a)                 { return fminf(fminf(a.x, a.y), a.z); }
static inline i3 i3f(v3 a)                     { i3 r = {(int)a.x, (int)a.y, (int)a.z}; return r; }

/* GLSL refract() as written in the GLSL 4.30 specification, section 8.5 */
static inline v3 vrefract(v3 I, v3 N, float eta)
{
	float d = vdot(N, I);
	float k = 1.0f - eta * eta * (1.0f - d * d);
	if(k < 0.0f)
		return v3s(0.0f);
	return vsub(vscale(I, eta), vscale(N, eta * d + sqrtf(k)));
}

/* column-major 4x4 (m[col*4 + row]) times vec4, terms added left to right */
static inline void mat4_mul_vec4(const float* m, const float v[4], float out[4])
{
	for(int r = 0; r < 4; r++)
		out[r] = m[0 * 4 + r] * v[0] + m[1 * 4 + r] * v[1] + m[2 * 4 + r] * v[2] + m[3 * 4 + r] * v[3];
}

static inline void mat4_mul_mat4(const float* a, const float* b, float* out)
{
	for(int c = 0; c < 4; c++)
		for(int r = 0; r < 4; r++)
			out[c * 4 + r] = a[0 * 4 + r] * b[c * 4 + 0] + a[1 * 4 + r] * b[c * 4 + 1] + a[2 * 4 + r] * b[c * 4 + 2] + a[3 * 4 + r] * b[c * 4 + 3];
}

/* ------------------------------------------------------------------ */
/* shader-side structs (SH:19-27, SH:57-69)                             */

typedef struct
{
	v3 normal;
	uint32_t material;
	v3 albedo;
	v3 specLight;
	v3 diffuseLight;
} Voxel;

/* one shader invocation: the GLSL globals (SH:321-325, LI:62) live here */
typedef struct
{
	const OrbBuffers*  b;
	const OrbUniforms* u;
	const OrbVoxel*    voxelSnapshot;   /* N1: what rays read (== b->voxels, which is not written during a pass) */

	int      enableRefraction;  /* SH:321 */
	v3       orgRayPos;         /* SH:322 */
	uint32_t lastVoxID;         /* SH:324 */
	float    lastVoxRefract;    /* SH:325 */
	int      firstSample;       /* LI:62 */

	/* where the last opaque hit happened (instrumentation, not in the shader) */
	uint32_t hitMapIndex, hitLocalIndex, hitRecordIndex;
	int      guardTripped;      /* N11 */

	/* lighting only: specular visible-bit propagation target (N3) */
	uint8_t* propagate;
	const uint8_t* visibleSnapshot;

	OrbCounters c;
} Inv;

/* ------------------------------------------------------------------ */
/* SH:103-118                                                           */

static inline void decode_uint_RGBA(uint32_t val, uint32_t out[4])
{
	out[0] = (val >> 24) & 0xFF;
	out[1] = (val >> 16) & 0xFF;
	out[2] = (val >> 8) & 0xFF;
	out[3] = val & 0xFF;
}

static inline uint32_t encode_uint_RGBA(uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
	return ((x & 0xFF) << 24) | ((y & 0xFF) << 16) | ((z & 0xFF) << 8) | (w & 0xFF);
}

/* SH:123-138 */
static inline int in_map_bounds(const OrbUniforms* u, i3 p)
{
	return p.x >= 0 && p.y >= 0 && p.z >= 0 && (uint32_t)p.x < u->mapSize[0] && (uint32_t)p.y < u->mapSize[1] && (uint32_t)p.z < u->mapSize[2];
}

static inline int in_chunk_bounds(i3 p)
{
	return p.x < 8 && p.y < 8 && p.z < 8 && p.x >= 0 && p.y >= 0 && p.z >= 0;
}

static inline uint32_t get_map_index(const OrbUniforms* u, i3 p)
{
	return (uint32_t)p.x + u->mapSize[0] * ((uint32_t)p.y + u->mapSize[1] * (uint32_t)p.z);
}

/* SH:141-147 (N4: the lastUsed reset is kept; all writers store 0) */
static inline OrbHandle get_map_tile(Inv* inv, uint32_t index)
{
	OrbHandle* m = &inv->b->map[index];
	if(__atomic_load_n(&m->lastUsed, __ATOMIC_RELAXED) > 0)
		__atomic_store_n(&m->lastUsed, 0, __ATOMIC_RELAXED);

	OrbHandle h;
	h.flags = __atomic_load_n(&m->flags, __ATOMIC_RELAXED);
	h.lastUsed = 0;
	h.voxelIndex = m->voxelIndex;
	return h;
}

/* SH:150-169 */
static inline uint32_t get_voxel_index(const OrbBuffers* b, uint32_t mapIndex, i3 chunkPos)
{
	uint32_t localIndex = (uint32_t)chunkPos.x + 8u * ((uint32_t)chunkPos.y + 8u * (uint32_t)chunkPos.z);
	uint32_t bitMaskIndex = localIndex >> 5;

	uint32_t voxNum = (bitMaskIndex > 3) ? b->chunks[mapIndex].partialCounts[(bitMaskIndex >> 2) - 1] : 0;

	for(uint32_t i = bitMaskIndex & ~3u; i <= bitMaskIndex; i++)
	{
		uint32_t bits = b->chunks[mapIndex].bitMask[i];
		if(i == bitMaskIndex)
			bits &= (1u << (localIndex & 31)) - 1;

		voxNum += (uint32_t)__builtin_popcount(bits);
	}

	return b->map[mapIndex].voxelIndex + voxNum;
}

uint32_t orb_get_voxel_index(const OrbBuffers* buf, uint32_t mapIndex, int x, int y, int z)
{
	i3 p = {x, y, z};
	return get_voxel_index(buf, mapIndex, p);
}

/* SH:172-224, including the `<` (not `<=`) partial-count quirk at SH:181 */
static inline i3 get_voxel_position(const OrbBuffers* b, uint32_t chunk, uint32_t voxNum)
{
	uint32_t count = 0;
	uint32_t pos = 0;

	uint32_t startIndex = 0;
	for(int i = 0; i < 3; i++)
	{
		if(b->chunks[chunk].partialCounts[i] < voxNum)
		{
			count = b->chunks[chunk].partialCounts[i];
			pos += 128;
			startIndex += 4;
		}
		else
			break;
	}

	voxNum++;

	for(uint32_t i = startIndex; i < 16; i++)
	{
		if(count == voxNum)
			break;

		uint32_t bc = (uint32_t)__builtin_popcount(b->chunks[chunk].bitMask[i]);
		if(count + bc >= voxNum)
		{
			uint32_t bitNum = 0;
			while(count < voxNum)
			{
				count += (b->chunks[chunk].bitMask[i] >> bitNum) & 1;
				bitNum++;
				pos++;
			}
		}
		else
		{
			count += bc;
			pos += 32;
		}
	}
	pos--;

	i3 r;
	if(count < voxNum)
	{
		r.x = r.y = r.z = -1;
	}
	else
	{
		r.x = (int)(pos % 8);
		r.y = (int)((pos / 8) % 8);
		r.z = (int)(pos / 64);
	}
	return r;
}

void orb_get_voxel_position(const OrbBuffers* buf, uint32_t chunk, uint32_t voxNum, int out[3])
{
	i3 r = get_voxel_position(buf, chunk, voxNum);
	out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

/* SH:227-231 */
static inline int does_voxel_exist(const OrbBuffers* b, uint32_t chunk, i3 p)
{
	uint32_t index = (uint32_t)p.x + 8u * ((uint32_t)p.y + 8u * (uint32_t)p.z);
	return (b->chunks[chunk].bitMask[index >> 5] >> (index & 31)) & 1;
}

/* SH:236-256 */
static inline Voxel decompress_voxel(OrbVoxel c)
{
	Voxel res;
	uint32_t n[4], a[4], s[4], d[4];
	decode_uint_RGBA(c.normal, n);
	decode_uint_RGBA(c.albedo, a);
	decode_uint_RGBA(c.specLight, s);
	decode_uint_RGBA(c.diffuseLight, d);

	uint32_t dx = (s[2] << 8) | s[3];
	uint32_t dy = (d[0] << 8) | d[1];
	uint32_t dz = (d[2] << 8) | d[3];

	res.normal       = V3(((float)n[1] * 0.00392156862f - 0.5f) * 2.0f, ((float)n[2] * 0.00392156862f - 0.5f) * 2.0f, ((float)n[3] * 0.00392156862f - 0.5f) * 2.0f);
	res.material     = n[0];
	res.albedo       = V3((float)a[0] * 0.00392156862f, (float)a[1] * 0.00392156862f, (float)a[2] * 0.00392156862f);
	res.diffuseLight = V3((float)dx * 0.0000152590219f, (float)dy * 0.0000152590219f, (float)dz * 0.0000152590219f);
	res.specLight    = V3((float)a[3] * 0.00392156862f, (float)s[0] * 0.00392156862f, (float)s[1] * 0.00392156862f);
	return res;
}

/* SH:261-273 */
static inline void intersect_AABB(v3 invRayDir, v3 rayPos, v3 boxMin, v3 boxMax, float* tNear, float* tFar)
{
	v3 tMin = vmul(vsub(boxMin, rayPos), invRayDir);
	v3 tMax = vmul(vsub(boxMax, rayPos), invRayDir);
	v3 t1 = vmin(tMin, tMax);
	v3 t2 = vmax(tMin, tMax);
	*tNear = fmaxf(fmaxf(t1.x, t1.y), t1.z);
	*tFar  = fminf(fminf(t2.x, t2.y), t2.z);
}

/* SH:276-284 */
static inline v3 normal_AABB(v3 intersectPos, v3 boxMin, v3 boxMax)
{
	v3 c = vscale(vadd(boxMin, boxMax), 0.5f);
	v3 p = vsub(intersectPos, c);
	v3 d = vscale(vsub(boxMax, boxMin), 0.5f);
	const float bias = 1.0f + ORB_EPSILON;
	v3 q = V3(p.x / d.x * bias, p.y / d.y * bias, p.z / d.z * bias);
	return vnormalize(vtrunc(q));
}

/* SH:289-295, gradient branch only (useCubemap is out of scope) */
static inline v3 sky_color(const OrbUniforms* u, v3 rayDir)
{
	float t = rayDir.y * 0.5f + 1.0f;
	v3 bot = V3(u->skyGradientBot[0], u->skyGradientBot[1], u->skyGradientBot[2]);
	v3 top = V3(u->skyGradientTop[0], u->skyGradientTop[1], u->skyGradientTop[2]);
	return vadd(vscale(bot, 1.0f - t), vscale(top, t));
}

/* SH:300-306 */
static inline void init_DDA(v3 rayDir, v3 invRayDir, v3 rayPos, i3* pos, v3* deltaDist, i3* rayStep, v3* sideDist)
{
	*pos = i3f(vfloor(rayPos));
	*deltaDist = vabs(invRayDir);
	v3 sg = vsign(rayDir);
	*rayStep = i3f(sg);
	v3 t = vadd(vmul(sg, vsub(v3i(*pos), rayPos)), vscale(sg, 0.5f));
	*sideDist = vmul(vaddf(t, 0.5f), *deltaDist);
}

/* SH:309-317 (mask products evaluated as selects, N6) */
static inline void iterate_DDA(v3 deltaDist, i3 rayStep, v3* sideDist, i3* mapPos, v3* normal)
{
	v3 s = *sideDist;
	int mx = s.x <= fminf(s.y, s.z);
	int my = s.y <= fminf(s.z, s.x);
	int mz = s.z <= fminf(s.x, s.y);

	if(mx) sideDist->x = s.x + deltaDist.x;
	if(my) sideDist->y = s.y + deltaDist.y;
	if(mz) sideDist->z = s.z + deltaDist.z;

	if(mx) mapPos->x += rayStep.x;
	if(my) mapPos->y += rayStep.y;
	if(mz) mapPos->z += rayStep.z;

	normal->x = (mx ? 1.0f : 0.0f) * (float)(-rayStep.x);
	normal->y = (my ? 1.0f : 0.0f) * (float)(-rayStep.y);
	normal->z = (mz ? 1.0f : 0.0f) * (float)(-rayStep.z);
}

/* ------------------------------------------------------------------ */
/* SH:328-418                                                           */

static int step_chunk(Inv* inv, i3 mapPos, uint32_t mapIndex, v3* rayDir, v3* invRayDir, v3* rayPos, int ignoreFirst, float maxDepth,
                      v3* hitNormal, Voxel* voxel, v3* colorAdd, float* colorMult, int* refracted)
{
	const OrbBuffers* b = inv->b;
	const OrbUniforms* u = inv->u;
	*refracted = 0;
	inv->c.chunks++;

	i3 pos, rayStep;
	v3 deltaDist, sideDist;
	v3 lastSideDist = v3s(0.0f);
	init_DDA(*rayDir, *invRayDir, *rayPos, &pos, &deltaDist, &rayStep, &sideDist);

	uint32_t guard = 0;
	while(in_chunk_bounds(pos))
	{
		if(++guard > ORB_MAX_CHUNK_STEPS) /* N11 */
		{
			inv->guardTripped = 1;
			return 0;
		}
		inv->c.voxelSteps++;

		if(does_voxel_exist(b, mapIndex, pos) && !ignoreFirst)
		{
			uint32_t recordIndex = get_voxel_index(b, mapIndex, pos);
			OrbVoxel compressed = inv->voxelSnapshot[recordIndex];
			inv->c.records++;
			*voxel = decompress_voxel(compressed);

			OrbMaterial material = b->materials[voxel->material];
			uint32_t thisVoxID = (compressed.albedo & 0xFFFFFF00u) | (compressed.normal >> 24);

			if(material.opacity == 1.0f)
			{
				*rayPos = vadd(*rayPos, vscale(*rayDir, min3(lastSideDist) + ORB_EPSILON));
				inv->hitMapIndex = mapIndex;
				inv->hitLocalIndex = (uint32_t)pos.x + 8u * ((uint32_t)pos.y + 8u * (uint32_t)pos.z);
				inv->hitRecordIndex = recordIndex;
				return 1;
			}
			else if(inv->lastVoxID != thisVoxID)
			{
				v3 curPos = vadd(*rayPos, vscale(*rayDir, min3(lastSideDist)));
				curPos = vadd(v3i(mapPos), vscale(curPos, 0.125f));
				v3 toCurPos = vsub(curPos, inv->orgRayPos);
				if(maxDepth < 0.0f || vdot(toCurPos, toCurPos) < maxDepth * maxDepth)
				{
					v3 sun = V3(u->sunStrength[0], u->sunStrength[1], u->sunStrength[2]);
					/* colorAdd += colorMult * material.opacity * voxel.albedo * sunStrength  (left to right) */
					float cm = *colorMult * material.opacity;
					*colorAdd = vadd(*colorAdd, vmul(vscale(voxel->albedo, cm), sun));
					*colorMult = *colorMult * (1.0f - material.opacity);
				}

				float rayDist = min3(lastSideDist);
				if(inv->enableRefraction && rayDist > 0.0f)
				{
					*refracted = 1;
					*rayPos = vadd(*rayPos, vscale(*rayDir, rayDist + ORB_EPSILON));

					v3 normal = vdot(voxel->normal, *rayDir) < 0.0f ? vnormalize(voxel->normal) : *hitNormal;
					*rayDir = vrefract(*rayDir, normal, inv->lastVoxRefract / material.refractIndex);
					*invRayDir = vrcp(*rayDir);

					init_DDA(*rayDir, *invRayDir, *rayPos, &pos, &deltaDist, &rayStep, &sideDist);
					lastSideDist = v3s(0.0f);
				}

				inv->lastVoxID = thisVoxID;
				inv->lastVoxRefract = material.refractIndex;
			}
		}
		else if(inv->lastVoxID != 255)
		{
			if(inv->enableRefraction)
			{
				*refracted = 1;
				*rayPos = vadd(*rayPos, vscale(*rayDir, min3(lastSideDist) + ORB_EPSILON));

				v3 oldRayDir = *rayDir;
				v3 nneg = vneg(voxel->normal);
				v3 normal = vdot(nneg, *rayDir) < 0.0f ? vnormalize(nneg) : *hitNormal;
				*rayDir = vrefract(*rayDir, normal, inv->lastVoxRefract);
				if(rayDir->x == 0.0f && rayDir->y == 0.0f && rayDir->z == 0.0f)
					*rayDir = oldRayDir;
				*invRayDir = vrcp(*rayDir);

				init_DDA(*rayDir, *invRayDir, *rayPos, &pos, &deltaDist, &rayStep, &sideDist);
				lastSideDist = v3s(0.0f);
			}

			inv->lastVoxID = 255;
			inv->lastVoxRefract = 1.0f;
		}

		lastSideDist = sideDist;
		iterate_DDA(deltaDist, rayStep, &sideDist, &pos, hitNormal);
		ignoreFirst = 0;
	}

	return 0;
}

/* SH:421-475 */
static int step_map(Inv* inv, v3* rayDir, v3 invRayDir, v3* rayPos, int ignoreFirst, float maxDepth, v3* hitNormal, Voxel* voxel, v3* colorAdd, float* colorMult)
{
	const OrbUniforms* u = inv->u;
	int requested = 0;
	inv->c.rays++;

	*colorAdd = v3s(0.0f);
	*colorMult = 1.0f;

	i3 pos, rayStep;
	v3 deltaDist, sideDist;
	v3 lastSideDist = v3s(0.0f);
	init_DDA(*rayDir, invRayDir, *rayPos, &pos, &deltaDist, &rayStep, &sideDist);

	const uint32_t maxSteps = 4u * (u->mapSize[0] + u->mapSize[1] + u->mapSize[2]) + 256u; /* N11 */
	uint32_t guard = 0;

	while(in_map_bounds(u, pos))
	{
		if(++guard > maxSteps || inv->guardTripped)
		{
			inv->guardTripped = 1;
			return 0;
		}

		uint32_t mapIndex = get_map_index(u, pos);
		inv->c.tiles++;

		OrbHandle mapTile = get_map_tile(inv, mapIndex);
		if((mapTile.flags & 3) == 2)
		{
			v3 updatedRayPos = vadd(*rayPos, vscale(*rayDir, min3(lastSideDist) - ORB_EPSILON));
			v3 chunkRayPos = vscale(vsub(updatedRayPos, v3i(pos)), 8.0f);
			chunkRayPos = vmin(vmax(chunkRayPos, v3s(ORB_EPSILON)), v3s(8.0f - ORB_EPSILON));

			int refracted;
			if(step_chunk(inv, pos, mapIndex, rayDir, &invRayDir, &chunkRayPos, ignoreFirst, maxDepth, hitNormal, voxel, colorAdd, colorMult, &refracted))
			{
				*rayPos = vadd(v3i(pos), vscale(chunkRayPos, 0.125f));
				return 1;
			}
			if(inv->guardTripped)
				return 0;

			if(refracted)
			{
				*rayPos = vadd(v3i(pos), vscale(chunkRayPos, 0.125f));
				init_DDA(*rayDir, invRayDir, *rayPos, &pos, &deltaDist, &rayStep, &sideDist);
				lastSideDist = v3s(0.0f);
			}
		}
		else if(!requested && (mapTile.flags & 3) != 0)
		{
			__atomic_store_n(&inv->b->map[mapIndex].flags, 3u, __ATOMIC_RELAXED);
			requested = 1;
		}

		lastSideDist = sideDist;
		iterate_DDA(deltaDist, rayStep, &sideDist, &pos, hitNormal);
		ignoreFirst = 0;
	}

	return 0;
}

/* ------------------------------------------------------------------ */
/* LI:23 */
static const float spherePoints[15][3] = {
	{0.000000f, 1.000000f, 0.000000f}, {-0.379803f, 0.857143f, 0.347931f}, {0.061185f, 0.714286f, -0.697174f},
	{0.499316f, 0.571429f, 0.651270f}, {-0.889696f, 0.428571f, -0.157375f}, {0.808584f, 0.285714f, -0.514354f},
	{-0.256942f, 0.142857f, 0.955810f}, {-0.460906f, 0.000000f, -0.887449f}, {0.929687f, -0.142857f, 0.339521f},
	{-0.885815f, -0.285714f, 0.365650f}, {0.382949f, -0.428571f, -0.818338f}, {0.245607f, -0.571429f, 0.783037f},
	{-0.605521f, -0.714286f, -0.350913f}, {0.503065f, -0.857143f, -0.110596f}, {-0.000000f, -1.000000f, 0.000000f}};

/* LI:29-32 (N5) */
float orb_rand(float seed)
{
	float s = sinf(seed) * 43758.5453f;
	float fr = s - floorf(s);
	return fr * 2.0f - 1.0f;
}

/* LI:41-44 */
static inline v3 rand3(float seed)
{
	return V3(orb_rand(seed), orb_rand(seed * 2.0f), orb_rand(seed * 3.0f));
}

/* LI:47-57, with a 1024-try guard (N11) */
static v3 rand_unit_sphere(float seed)
{
	v3 point = v3s(0.0f);
	for(int tries = 0; tries < 1024; tries++)
	{
		point = rand3(seed);
		seed = seed + 1.0f;
		if(vdot(point, point) >= 1.0f)
			continue;
		return point;
	}
	return point;
}

void orb_rand_unit_sphere(float seed, float out[3])
{
	v3 p = rand_unit_sphere(seed);
	out[0] = p.x; out[1] = p.y; out[2] = p.z;
}

static inline v3 u_sunDir(const OrbUniforms* u)      { return V3(u->sunDir[0], u->sunDir[1], u->sunDir[2]); }
static inline v3 u_sunStrength(const OrbUniforms* u) { return V3(u->sunStrength[0], u->sunStrength[1], u->sunStrength[2]); }

/* LI:65-80 */
static void shadow_ray(Inv* inv, v3 rayPos, float seed, v3* color)
{
	const OrbUniforms* u = inv->u;
	Voxel hitVoxel;
	memset(&hitVoxel, 0, sizeof(hitVoxel));
	v3 colorAdd;
	float colorMult;

	v3 updatedSunDir;
	if(inv->firstSample)
		updatedSunDir = vaddf(u_sunDir(u), ORB_EPSILON);
	else
		updatedSunDir = vaddf(vnormalize(vadd(vscale(u_sunDir(u), u->shadowSoftness), rand_unit_sphere(seed))), ORB_EPSILON);

	v3 tempNormal = v3s(0.0f);
	if(!step_map(inv, &updatedSunDir, vrcp(updatedSunDir), &rayPos, 1, -1.0f, &tempNormal, &hitVoxel, &colorAdd, &colorMult))
		*color = vadd(*color, vadd(vscale(u_sunStrength(u), colorMult), colorAdd));
}

/* LI:83-146 */
static void specular_ray(Inv* inv, i3 mapPos, v3 normal, v3 rayPos, v3 rayDir, v3 albedo, uint32_t reflectType, v3* color)
{
	const OrbUniforms* u = inv->u;
	(void)normal;
	Voxel hitVoxel;
	memset(&hitVoxel, 0, sizeof(hitVoxel));
	OrbMaterial hitMaterial;

	v3 lastPos = rayPos;
	v3 multiplier = albedo;

	for(uint32_t i = 0; i < u->specularBounceLimit; i++)
	{
		v3 colorAdd;
		float colorMult;

		v3 tempNormal = v3s(0.0f);
		if(step_map(inv, &rayDir, vrcp(rayDir), &rayPos, 1, -1.0f, &tempNormal, &hitVoxel, &colorAdd, &colorMult))
		{
			i3 hitMapPos = i3f(rayPos);
			uint32_t thisMapIndex = get_map_index(u, mapPos);
			uint32_t hitMapIndex  = get_map_index(u, hitMapPos);
			if(inv->visibleSnapshot[thisMapIndex] && in_map_bounds(u, hitMapPos)) /* N3 */
				__atomic_store_n(&inv->propagate[hitMapIndex], 1, __ATOMIC_RELAXED);

			v3 dist = vabs(vsub(vfloor(vscale(rayPos, 8.0f)), vfloor(vscale(lastPos, 8.0f))));
			if(vdot(dist, dist) <= 1.0f)
				return;

			hitMaterial = inv->b->materials[hitVoxel.material];
			hitVoxel.diffuseLight = vscale(hitVoxel.diffuseLight, 1.0f - hitMaterial.specular);

			if(hitMaterial.emissive)
			{
				*color = vadd(*color, vmul(vmul(vadd(vscale(hitVoxel.albedo, colorMult), colorAdd), multiplier), albedo));
				return;
			}
			else
			{
				v3 hitColor = vmul(hitVoxel.diffuseLight, hitVoxel.albedo);
				*color = vadd(*color, vmul(vadd(vscale(hitColor, colorMult), colorAdd), multiplier));

				if(hitMaterial.specular == 0.0f)
					return;

				multiplier = vmul(multiplier, vscale(vscale(hitVoxel.albedo, colorMult), hitMaterial.specular));
				reflectType = hitMaterial.reflectType;
				lastPos = rayPos;
				rayDir = vreflect(rayDir, hitVoxel.normal);
			}
		}
		else if(vdot(rayDir, u_sunDir(u)) > 0.99f)
		{
			*color = vadd(*color, vadd(vscale(u_sunStrength(u), colorMult), colorAdd));
			return;
		}
		else
		{
			v3 base = (reflectType == 1) ? sky_color(u, rayDir) : u_sunStrength(u);
			*color = vadd(*color, vmul(vadd(vscale(base, colorMult), colorAdd), multiplier));
			return;
		}
	}
}

/* LI:149-203 */
static void diffuse_ray(Inv* inv, v3 normal, v3 rayPos, Voxel initialVoxel, float seed, v3* color)
{
	const OrbUniforms* u = inv->u;
	Voxel hitVoxel = initialVoxel;
	v3 hitNormal = normal;
	OrbMaterial hitMaterial;
	memset(&hitMaterial, 0, sizeof(hitMaterial));

	v3 newColor = v3s(1.0f);

	v3 lastPos = rayPos;
	v3 lastDir = v3s(0.0f);
	for(uint32_t i = 0; i < u->diffuseBounceLimit; i++)
	{
		v3 dir;
		if(i > 0 && (orb_rand(seed + (float)u->diffuseBounceLimit + (float)i) + 1.0f) * 0.5f < hitMaterial.specular)
			dir = vnormalize(vadd(vscale(vreflect(lastDir, hitNormal), (float)hitMaterial.shininess), rand_unit_sphere(u->time + (float)i)));
		else if(inv->firstSample)
			dir = vaddf(vnormalize(hitNormal), ORB_EPSILON);
		else
			dir = vaddf(vnormalize(vadd(hitNormal, rand_unit_sphere(seed + (float)i))), ORB_EPSILON);

		v3 colorAdd;
		float colorMult;

		v3 tempNormal = v3s(0.0f);
		int hit = step_map(inv, &dir, vrcp(dir), &rayPos, 1, -1.0f, &tempNormal, &hitVoxel, &colorAdd, &colorMult);

		hitNormal = hitVoxel.normal;
		hitMaterial = inv->b->materials[hitVoxel.material];

		if(hit)
		{
			v3 dist = vabs(vsub(vfloor(vscale(lastPos, 8.0f)), vfloor(vscale(rayPos, 8.0f))));
			if(vdot(dist, dist) < 1.0f)
				return;

			if(hitMaterial.emissive)
			{
				*color = vadd(*color, vmul(newColor, vadd(vscale(hitVoxel.albedo, colorMult), colorAdd)));
				return;
			}
			else
				newColor = vmul(newColor, vadd(vscale(hitVoxel.albedo, colorMult), colorAdd));
		}
		else
		{
			float ndl = fmaxf(vdot(dir, u_sunDir(u)), 0.0f);
			*color = vadd(*color, vadd(vscale(vmul(vscale(newColor, ndl), u_sunStrength(u)), colorMult), colorAdd));
			return;
		}

		lastDir = dir;
	}
}

/* result of one lighting invocation, committed after the pass (N1) */
typedef struct
{
	uint32_t live;
	uint32_t w1, w2, w3;
} LightOut;

/* LI:207-288 for one invocation; writes nothing to the shared buffers */
static void light_invocation(Inv* inv, uint32_t request, uint32_t lane, LightOut* out)
{
	const OrbBuffers* b = inv->b;
	const OrbUniforms* u = inv->u;

	inv->enableRefraction = 0;
	inv->lastVoxID = 255;
	inv->lastVoxRefract = 1.0f;
	inv->firstSample = 0;
	inv->guardTripped = 0;
	out->live = 0;

	uint32_t mapIndex = request >> 4;
	i3 mapPos = {b->chunks[mapIndex].pos[0], b->chunks[mapIndex].pos[1], b->chunks[mapIndex].pos[2]};
	uint32_t voxNum = lane + (request & 15) * 32;
	i3 chunkPos = get_voxel_position(b, mapIndex, voxNum);

	if(!in_chunk_bounds(chunkPos))
		return;

	inv->c.voxelsLit++;

	uint32_t voxelIndex = b->map[mapIndex].voxelIndex + voxNum;
	OrbVoxel compressed = inv->voxelSnapshot[voxelIndex];
	Voxel thisVoxel = decompress_voxel(compressed);
	OrbMaterial thisMaterial = b->materials[thisVoxel.material];
	uint32_t ns = b->chunks[mapIndex].numIndirectSamples; /* N2: pre-dispatch value */
	float indirectSamples = (float)(ns < u->maxDiffuseSamples ? ns : u->maxDiffuseSamples);

	if(indirectSamples == 0.0f)
		inv->firstSample = 1;

	v3 rayPos = vaddf(vadd(vscale(v3i(chunkPos), 0.125f), v3i(mapPos)), 0.0625f);
	rayPos = vadd(rayPos, vscale(thisVoxel.normal, 0.0625f - ORB_EPSILON));

	v3 specLight = v3s(0.0f);
	v3 diffuseLight = v3s(0.0f);

	v3 camPos = V3(u->camPos[0], u->camPos[1], u->camPos[2]);
	v3 viewDir = vsub(rayPos, camPos);
	if(thisMaterial.specular > 0.0f && vdot(viewDir, thisVoxel.normal) < 0.0f && thisMaterial.reflectType <= 1)
	{
		v3 reflected = vreflect(vnormalize(viewDir), thisVoxel.normal);

		for(int i = 0; i < 15; i++)
		{
			v3 sp = V3(spherePoints[i][0], spherePoints[i][1], spherePoints[i][2]);
			v3 specDir = vaddf(vnormalize(vadd(vscale(reflected, (float)thisMaterial.shininess), sp)), ORB_EPSILON);
			specular_ray(inv, mapPos, thisVoxel.normal, rayPos, specDir, thisVoxel.albedo, thisMaterial.reflectType, &specLight);
		}

		specLight = V3(specLight.x / 15.0f, specLight.y / 15.0f, specLight.z / 15.0f);
	}

	if(thisMaterial.specular < 1.0f)
	{
		v3 ambient = V3(u->ambientStrength[0], u->ambientStrength[1], u->ambientStrength[2]);
		for(uint32_t i = 0; i < u->numDiffuseSamples; i++)
		{
			diffuseLight = vadd(diffuseLight, ambient);
			diffuse_ray(inv, vaddf(thisVoxel.normal, ORB_EPSILON), rayPos, thisVoxel, u->time * (float)(i + 1), &diffuseLight);
			shadow_ray(inv, rayPos, u->time * (float)(i + 1 + u->numDiffuseSamples), &diffuseLight);
		}

		float denom = indirectSamples + (float)u->numDiffuseSamples;
		v3 acc = vadd(vscale(thisVoxel.diffuseLight, indirectSamples), diffuseLight);
		diffuseLight = V3(acc.x / denom, acc.y / denom, acc.z / denom);
	}

	specLight = vclamp01(specLight);
	diffuseLight = vclamp01(diffuseLight);

	uint32_t wx = (uint32_t)rintf(diffuseLight.x * 65535.0f);
	uint32_t wy = (uint32_t)rintf(diffuseLight.y * 65535.0f);
	uint32_t wz = (uint32_t)rintf(diffuseLight.z * 65535.0f);

	out->live = 1;
	out->w1 = encode_uint_RGBA((uint32_t)rintf(thisVoxel.albedo.x * 255.0f), (uint32_t)rintf(thisVoxel.albedo.y * 255.0f), (uint32_t)rintf(thisVoxel.albedo.z * 255.0f), (uint32_t)rintf(specLight.x * 255.0f));
	out->w2 = encode_uint_RGBA((uint32_t)rintf(specLight.y * 255.0f), (uint32_t)rintf(specLight.z * 255.0f), (wx >> 8) & 0xFF, wx & 0xFF);
	out->w3 = encode_uint_RGBA((wy >> 8) & 0xFF, wy & 0xFF, (wz >> 8) & 0xFF, wz & 0xFF);
}

static void counters_add(OrbCounters* d, const OrbCounters* s)
{
	d->rays += s->rays; d->tiles += s->tiles; d->chunks += s->chunks; d->voxelSteps += s->voxelSteps;
	d->records += s->records; d->voxelsLit += s->voxelsLit; d->pixels += s->pixels;
}

/* Phase 1 of a lighting dispatch: the invocations of requests [first, first+count) are evaluated against the
 * UNMODIFIED buffers (N1, N2) and their packed words are written to `staging` -- 96 words per request, indexed from
 * request 0: [32 x albedo|spec.x][32 x spec.yz|diffuse.x][32 x diffuse.yz], zero for dead lanes -- the same layout
 * the CUDA path stages (doonengine_b200/csrc/light.cu).  Visible-bit propagations (LI:101-105) are OR-ed into
 * `propagate` (one byte per tile).  Nothing in `buf` is written except map[].lastUsed (N4). */
void orb_light_compute(const OrbBuffers* buf, const OrbUniforms* u, const uint32_t* requests, size_t first, size_t count, uint32_t* staging, uint8_t* propagate, OrbCounters* counters)
{
	size_t numTiles = (size_t)u->mapSize[0] * u->mapSize[1] * u->mapSize[2];
	uint8_t* visibleSnapshot = (uint8_t*)malloc(numTiles ? numTiles : 1);
	for(size_t i = 0; i < numTiles; i++)
		visibleSnapshot[i] = (buf->map[i].flags & 4) ? 1 : 0;

	OrbCounters total;
	memset(&total, 0, sizeof(total));

	#pragma omp parallel
	{
		Inv inv;
		memset(&inv, 0, sizeof(inv));
		inv.b = buf;
		inv.u = u;
		inv.voxelSnapshot = buf->voxels;
		inv.propagate = propagate;
		inv.visibleSnapshot = visibleSnapshot;

		#pragma omp for schedule(dynamic, 4)
		for(size_t r = first; r < first + count; r++)
			for(uint32_t lane = 0; lane < 32; lane++)
			{
				LightOut o;
				light_invocation(&inv, requests[r], lane, &o);
				staging[r * 96 + lane]      = o.live ? o.w1 : 0;
				staging[r * 96 + 32 + lane] = o.live ? o.w2 : 0;
				staging[r * 96 + 64 + lane] = o.live ? o.w3 : 0;
			}

		#pragma omp critical
		counters_add(&total, &inv.c);
	}

	if(counters)
		counters_add(counters, &total);
	free(visibleSnapshot);
}

/* Phase 2: commit (N1-N3) -- records, then visible-bit clears and sample counts, then propagation.
 * A lane is live iff its voxel number is below the chunk's surface-voxel count (get_voxel_position, SH:220-223). */
void orb_light_commit(const OrbBuffers* buf, const OrbUniforms* u, const uint32_t* requests, size_t numRequests, const uint32_t* staging, const uint8_t* propagate)
{
	size_t numTiles = (size_t)u->mapSize[0] * u->mapSize[1] * u->mapSize[2];
	for(size_t r = 0; r < numRequests; r++)
	{
		uint32_t mapIndex = requests[r] >> 4;
		uint32_t numVoxels = 0;
		for(int i = 0; i < 16; i++)
			numVoxels += (uint32_t)__builtin_popcount(buf->chunks[mapIndex].bitMask[i]);
		for(uint32_t lane = 0; lane < 32; lane++)
		{
			uint32_t voxNum = lane + (requests[r] & 15) * 32;
			if(voxNum >= numVoxels)
				continue;

			uint32_t recordIndex = buf->map[mapIndex].voxelIndex + voxNum;
			buf->voxels[recordIndex].albedo = staging[r * 96 + lane];            /* LI:276 */
			buf->voxels[recordIndex].specLight = staging[r * 96 + 32 + lane];    /* LI:277 */
			buf->voxels[recordIndex].diffuseLight = staging[r * 96 + 64 + lane]; /* LI:278 */

			buf->map[mapIndex].flags &= ~4u;                 /* LI:281 */
			if(voxNum == 0)
				buf->chunks[mapIndex].numIndirectSamples++;  /* LI:284-285 */
		}
	}
	for(size_t i = 0; i < numTiles; i++)
		if(propagate[i])
			buf->map[i].flags |= 4u;                         /* LI:104-105 */
}

void orb_light(const OrbBuffers* buf, const OrbUniforms* u, const uint32_t* requests, size_t numRequests, size_t numVoxelRecords, OrbCounters* counters)
{
	(void)numVoxelRecords;
	size_t numTiles = (size_t)u->mapSize[0] * u->mapSize[1] * u->mapSize[2];
	uint32_t* staging = (uint32_t*)malloc(sizeof(uint32_t) * 96 * (numRequests ? numRequests : 1));
	uint8_t* propagate = (uint8_t*)calloc(numTiles ? numTiles : 1, 1);
	orb_light_compute(buf, u, requests, 0, numRequests, staging, propagate, counters);
	orb_light_commit(buf, u, requests, numRequests, staging, propagate);
	free(staging);
	free(propagate);
}

/* ------------------------------------------------------------------ */
/* DR:24-30, sky branch only (raster compose is out of scope) */
static inline v3 background_color(const OrbUniforms* u, v3 rayDir)
{
	v3 s = sky_color(u, rayDir);
	return V3(powf(s.x, 2.2f), powf(s.y, 2.2f), powf(s.z, 2.2f));
}

/* DR:32-61 */
static v3 voxel_color(const Inv* inv, Voxel vox, v3 colorAdd, float colorMult, v3 hitNormal)
{
	OrbMaterial material = inv->b->materials[vox.material];
	vox.specLight = vscale(vox.specLight, material.specular);
	vox.diffuseLight = vscale(vox.diffuseLight, 1.0f - material.specular);

	switch(inv->u->viewMode)
	{
	case 0:
	{
		v3 solidColor;
		if(material.emissive)
			solidColor = vox.albedo;
		else
			solidColor = vadd(vmul(vox.diffuseLight, vox.albedo), vox.specLight);
		return vadd(vscale(solidColor, colorMult), colorAdd);
	}
	case 1: return vox.albedo;
	case 2: return material.emissive ? vox.albedo : vox.diffuseLight;
	case 3: return material.emissive ? vox.albedo : vox.specLight;
	case 4: return vabs(vox.normal);
	case 5: return vabs(hitNormal);
	}
	return v3s(0.0f);
}

/* DR:63-147 for one pixel */
static void draw_invocation(Inv* inv, int px, int py, int w, int h, float* image, OrbHit* hits)
{
	const OrbUniforms* u = inv->u;

	inv->enableRefraction = 1;
	inv->lastVoxID = 255;
	inv->lastVoxRefract = 1.0f;
	inv->guardTripped = 0;

	Voxel finalVoxel;
	memset(&finalVoxel, 0, sizeof(finalVoxel));
	v3 finalColorAdd = v3s(0.0f);
	float finalColorMult = 1.0f;
	v3 finalColor;
	float finalDepth = -1.0f; /* N7 */

	float sx = (float)px / (float)w * 2.0f - 1.0f;
	float sy = (float)py / (float)h * 2.0f - 1.0f;

	float origin4[4] = {0.0f, 0.0f, 0.0f, 1.0f};
	float rp[4];
	mat4_mul_vec4(u->invViewMat, origin4, rp);
	v3 rayPos = V3(rp[0], rp[1], rp[2]);
	inv->orgRayPos = rayPos;

	float m[16], sp4[4] = {sx, sy, 0.0f, 1.0f}, rd[4];
	mat4_mul_mat4(u->invCenteredViewMat, u->invProjectionMat, m);
	mat4_mul_vec4(m, sp4, rd);
	v3 rayDir = vaddf(vnormalize(V3(rd[0], rd[1], rd[2])), ORB_EPSILON);
	v3 invRayDir = vrcp(rayDir);

	float maxDepth = -1.0f;

	v3 boxMax = V3((float)u->mapSize[0], (float)u->mapSize[1], (float)u->mapSize[2]);
	float tNear, tFar;
	intersect_AABB(invRayDir, rayPos, v3s(0.0f), boxMax, &tNear, &tFar);

	OrbHit hit = {0, 0, 0, 0};

	if(tNear > tFar || tFar < 0.0f)
	{
		finalColor = background_color(u, rayDir);
	}
	else
	{
		if(tNear > 0.0f)
			rayPos = vadd(rayPos, vscale(rayDir, tNear + ORB_EPSILON));
		v3 finalNormal = normal_AABB(rayPos, v3s(0.0f), boxMax);

		if(step_map(inv, &rayDir, invRayDir, &rayPos, 0, maxDepth, &finalNormal, &finalVoxel, &finalColorAdd, &finalColorMult))
		{
			v3 fl = vfloor(rayPos);
			if(fl.x >= 0.0f && fl.y >= 0.0f && fl.z >= 0.0f && fl.x < boxMax.x && fl.y < boxMax.y && fl.z < boxMax.z) /* reference writes out of bounds otherwise */
			{
				uint32_t mx = (uint32_t)fl.x, my = (uint32_t)fl.y, mz = (uint32_t)fl.z;
				uint32_t index = mx + u->mapSize[0] * (my + u->mapSize[1] * mz);
				__atomic_fetch_or(&inv->b->map[index].flags, 4u, __ATOMIC_RELAXED);
			}

			v3 dv = vsub(rayPos, inv->orgRayPos);
			finalDepth = sqrtf(vdot(dv, dv));

			finalColor = voxel_color(inv, finalVoxel, finalColorAdd, finalColorMult, finalNormal);

			hit.status = 2;
			hit.mapIndex = inv->hitMapIndex;
			hit.localIndex = inv->hitLocalIndex;
			hit.recordIndex = inv->hitRecordIndex;
		}
		else
		{
			finalColor = vadd(vscale(background_color(u, rayDir), finalColorMult), finalColorAdd);
			finalDepth = maxDepth;
			hit.status = 1;
		}
	}

	finalColor = V3(powf(finalColor.x, 0.4545f), powf(finalColor.y, 0.4545f), powf(finalColor.z, 0.4545f));

	float* dst = image + ((size_t)py * (size_t)w + (size_t)px) * 4;
	dst[0] = finalColor.x;
	dst[1] = finalColor.y;
	dst[2] = finalColor.z;
	dst[3] = finalDepth;
	if(hits)
		hits[(size_t)py * (size_t)w + (size_t)px] = hit;
	inv->c.pixels++;
}

void orb_draw(const OrbBuffers* buf, const OrbUniforms* u, int w, int h, float* image, OrbHit* hits, OrbCounters* counters)
{
	int gx = w / 16, gy = h / 16; /* voxel.c:879 */
	OrbCounters total;
	memset(&total, 0, sizeof(total));

	#pragma omp parallel
	{
		Inv inv;
		memset(&inv, 0, sizeof(inv));
		inv.b = buf;
		inv.u = u;
		inv.voxelSnapshot = buf->voxels;

		#pragma omp for schedule(dynamic, 1) collapse(2)
		for(int ty = 0; ty < gy; ty++)
			for(int tx = 0; tx < gx; tx++)
				for(int ly = 0; ly < 16; ly++)
					for(int lx = 0; lx < 16; lx++)
						draw_invocation(&inv, tx * 16 + lx, ty * 16 + ly, w, h, image, hits);

		#pragma omp critical
		counters_add(&total, &inv.c);
	}

	if(counters)
		counters_add(counters, &total);
}

int orb_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* bench.py --impl reference under torchrun inherits OMP_NUM_THREADS=1; the reference arm uses every host core it can,
 * so the harness sets the team size itself.  n <= 0: all online processors. */
void orb_set_num_threads(int n)
{
#ifdef _OPENMP
	if(n <= 0)
		n = omp_get_num_procs();
	omp_set_dynamic(0);
	omp_set_num_threads(n);
#else
	(void)n;
#endif
}
