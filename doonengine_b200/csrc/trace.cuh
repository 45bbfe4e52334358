/* trace.cuh -- the two-level DDA every ray of the draw and lighting kernels runs.
 *
 * Behaviour follows /root/reference/assets/shaders/voxelShared.comp:
 *   init_dda / iterate_dda   SH:300-317  (branch-free; all tied axes step together)
 *   trace_ray  outer loop    SH:421-475  step_map: one iteration per map tile
 *              inner loop    SH:328-418  step_chunk: one iteration per voxel of a resident chunk
 * The float arithmetic (sequential sideDist accumulation, entry point min(lastSideDist) -/+ EPSILON, the
 * [EPSILON, 8-EPSILON] clamp) is kept operation for operation because first-hit indices must be bit-exact.
 * What is re-designed is every memory access around it (layout.h): an empty tile costs one bit test in a
 * register-cached 4x4x4 occupancy word, a resident tile one 4-byte slot lookup, a voxel step one bit test
 * in a register-cached mask word of the 128-byte chunk slot, a hit one 16-byte record gather.
 *
 * Not implemented (out of scope): raster-depth gating (maxDepth is always < 0, so the test at
 * SH:363 is always true) and demand-stream requests (SH:462-466; the map is resident).
 */
#ifndef DN_B200_TRACE_CUH
#define DN_B200_TRACE_CUH

#include "layout.h"
#include "vecmath.cuh"

/* what survives from ray to ray inside one thread (the GLSL globals at SH:321-325 plus the `voxel` out-parameter) */
struct RayState
{
	uint32_t lastVoxID;      /* albedo|material of the transparent block the ray is inside, 255 = none */
	float    lastVoxRefract;
	uint4    vox;            /* last record fetched by any ray of this thread */
	bool     tripped;        /* a loop guard fired; all later rays of the thread miss (oracle.h N11) */
	uint32_t hitMapIndex, hitLocalIndex, hitRecord; /* where the last opaque hit happened */
};

DNB_FN void ray_state_reset(RayState& st)
{
	st.lastVoxID = 255;
	st.lastVoxRefract = 1.0f;
	st.vox = make_uint4(0, 0, 0, 0);
	st.tripped = false;
	st.hitMapIndex = st.hitLocalIndex = st.hitRecord = 0;
}

/* record decode, SH:236-256 (constants are the shader's, not exactly 1/255 and 1/65535) */
DNB_FN uint32_t vox_material(uint4 v) { return v.x >> 24; }
DNB_FN f3 vox_normal(uint4 v)
{
	return mk3(((float)((v.x >> 16) & 0xFF) * 0.00392156862f - 0.5f) * 2.0f,
	           ((float)((v.x >> 8) & 0xFF) * 0.00392156862f - 0.5f) * 2.0f,
	           ((float)(v.x & 0xFF) * 0.00392156862f - 0.5f) * 2.0f);
}
DNB_FN f3 vox_albedo(uint4 v)
{
	return mk3((float)(v.y >> 24) * 0.00392156862f, (float)((v.y >> 16) & 0xFF) * 0.00392156862f, (float)((v.y >> 8) & 0xFF) * 0.00392156862f);
}
DNB_FN f3 vox_spec(uint4 v)
{
	return mk3((float)(v.y & 0xFF) * 0.00392156862f, (float)(v.z >> 24) * 0.00392156862f, (float)((v.z >> 16) & 0xFF) * 0.00392156862f);
}
DNB_FN f3 vox_diffuse(uint4 v)
{
	return mk3((float)(v.z & 0xFFFF) * 0.0000152590219f, (float)(v.w >> 16) * 0.0000152590219f, (float)(v.w & 0xFFFF) * 0.0000152590219f);
}

DNB_FN DnbMaterial load_material(const DnbScene& S, uint32_t id)
{
	const uint4* p = reinterpret_cast<const uint4*>(S.materials + id);
	uint4 a = __ldg(p), b = __ldg(p + 1);
	DnbMaterial m;
	m.pad[0] = m.pad[1] = 0.0f;
	m.emissive = a.z;
	m.opacity = __uint_as_float(a.w);
	m.refractIndex = __uint_as_float(b.x);
	m.specular = __uint_as_float(b.y);
	m.reflectType = b.z;
	m.shininess = b.w;
	return m;
}

/* SH:289-295, gradient branch */
DNB_FN f3 sky_color(const DnbScene& S, f3 rayDir)
{
	float t = rayDir.y * 0.5f + 1.0f;
	return ld3(S.skyBot) * (1.0f - t) + ld3(S.skyTop) * t;
}

/* SH:300-306 */
DNB_FN void init_dda(f3 rayDir, f3 invRayDir, f3 rayPos, i3& pos, f3& deltaDist, i3& rayStep, f3& sideDist)
{
	pos = toi3(floor3(rayPos));
	deltaDist = abs3(invRayDir);
	f3 sg = mk3(sgn(rayDir.x), sgn(rayDir.y), sgn(rayDir.z));
	rayStep = toi3(sg);
	f3 t = sg * (tof3(pos) - rayPos) + sg * 0.5f;
	sideDist = (t + 0.5f) * deltaDist;
}

/* SH:309-317; the mask products are selects (identical unless deltaDist is infinite, oracle.h N6) */
DNB_FN void iterate_dda(f3 deltaDist, i3 rayStep, f3& sideDist, i3& pos, f3& normal)
{
	f3 s = sideDist;
	bool mx = s.x <= fminf(s.y, s.z);
	bool my = s.y <= fminf(s.z, s.x);
	bool mz = s.z <= fminf(s.x, s.y);
	if(mx) { sideDist.x = s.x + deltaDist.x; pos.x += rayStep.x; }
	if(my) { sideDist.y = s.y + deltaDist.y; pos.y += rayStep.y; }
	if(mz) { sideDist.z = s.z + deltaDist.z; pos.z += rayStep.z; }
	normal.x = (mx ? 1.0f : 0.0f) * (float)(-rayStep.x);
	normal.y = (my ? 1.0f : 0.0f) * (float)(-rayStep.y);
	normal.z = (mz ? 1.0f : 0.0f) * (float)(-rayStep.z);
}

DNB_FN bool in_map_bounds(const DnbScene& S, i3 p)
{
	return (uint32_t)p.x < S.mapSize[0] && (uint32_t)p.y < S.mapSize[1] && (uint32_t)p.z < S.mapSize[2];
}

DNB_FN bool in_chunk_bounds(i3 p)
{
	return ((uint32_t)(p.x | p.y | p.z)) < 8u;
}

#define DNB_COUNT(field) do { if(COUNT) lc.field++; } while(0)

/* step_map + step_chunk.  REFRACT: enableRefraction (true in draw, false in lighting, DR:65 / LI:209).
 * invRayDir is by value: a refraction inside a chunk updates the caller's rayDir but only this
 * function's invRayDir, exactly as the inout/in qualifiers at SH:328/421 do. */
template <bool REFRACT, bool COUNT>
DNB_FN bool trace_ray(const DnbScene& S, RayState& st, DnbCounters& lc, f3& rayDir, f3 invRayDir, f3& rayPos, bool ignoreFirst, f3& hitNormal, f3& colorAdd, float& colorMult)
{
	colorAdd = splat3(0.0f);
	colorMult = 1.0f;
	DNB_COUNT(rays);

	i3 pos, rayStep;
	f3 deltaDist, sideDist;
	f3 lastSideDist = splat3(0.0f);
	init_dda(rayDir, invRayDir, rayPos, pos, deltaDist, rayStep, sideDist);

	uint32_t guard = 0;
	uint32_t occBlock = 0xFFFFFFFFu;
	unsigned long long occWord = 0;

	while(in_map_bounds(S, pos))
	{
		if(++guard > S.maxMapSteps || st.tripped)
		{
			st.tripped = true;
			return false;
		}
		DNB_COUNT(tiles);

		uint32_t block = (uint32_t)(pos.x >> 2) + S.blocks[0] * ((uint32_t)(pos.y >> 2) + S.blocks[1] * (uint32_t)(pos.z >> 2));
		if(block != occBlock)
		{
			occBlock = block;
			occWord = __ldg(S.occ64 + block);
		}
		uint32_t bit = (uint32_t)(pos.x & 3) | ((uint32_t)(pos.y & 3) << 2) | ((uint32_t)(pos.z & 3) << 4);

		if((occWord >> bit) & 1ull)
		{
			uint32_t mapIndex = (uint32_t)pos.x + S.mapSize[0] * ((uint32_t)pos.y + S.mapSize[1] * (uint32_t)pos.z);
			uint32_t slotId = __ldg(S.tileSlot + mapIndex) - 1u;
			const DnbSlot* slot = S.slots + slotId;
			DNB_COUNT(chunks);

			/* SH:443-445: entry point in chunk-local voxel units */
			f3 entry = rayPos + rayDir * (hmin3(lastSideDist) - DNB_EPSILON);
			f3 cpos = (entry - tof3(pos)) * 8.0f;
			cpos = min3v(max3v(cpos, splat3(DNB_EPSILON)), splat3(8.0f - DNB_EPSILON));

			/* ---- step_chunk, SH:328-418 ---- */
			bool refracted = false;
			i3 p, cstep;
			f3 cdelta, cside;
			f3 clast = splat3(0.0f);
			init_dda(rayDir, invRayDir, cpos, p, cdelta, cstep, cside);

			uint32_t cguard = 0;
			uint32_t wordIdx = 0xFFFFFFFFu, word = 0;
			while(in_chunk_bounds(p))
			{
				if(++cguard > DNB_MAX_CHUNK_STEPS)
				{
					st.tripped = true;
					return false;
				}
				DNB_COUNT(voxelSteps);

				uint32_t local = (uint32_t)p.x + 8u * ((uint32_t)p.y + 8u * (uint32_t)p.z);
				if((local >> 5) != wordIdx)
				{
					wordIdx = local >> 5;
					word = __ldg(slot->mask + wordIdx);
				}

				if(((word >> (local & 31u)) & 1u) && !ignoreFirst)
				{
					/* SH:150-169 with per-word prefix counts instead of quarter counts */
					uint32_t rel = (uint32_t)__ldg(slot->prefix + wordIdx) + __popc(word & ((1u << (local & 31u)) - 1u));
					uint4 rec = __ldg(S.records + (__ldg(&slot->voxelBase) + rel));
					DNB_COUNT(records);
					st.vox = rec;

					DnbMaterial material = load_material(S, rec.x >> 24);
					uint32_t thisVoxID = (rec.y & 0xFFFFFF00u) | (rec.x >> 24);

					if(material.opacity == 1.0f)
					{
						cpos = cpos + rayDir * (hmin3(clast) + DNB_EPSILON);
						rayPos = tof3(pos) + cpos * 0.125f; /* SH:451 */
						st.hitMapIndex = mapIndex;
						st.hitLocalIndex = local;
						st.hitRecord = rel;
						return true;
					}
					else if(st.lastVoxID != thisVoxID)
					{
						/* SH:365-366, maxDepth < 0 */
						float cm = colorMult * material.opacity;
						colorAdd = colorAdd + (vox_albedo(rec) * cm) * ld3(S.sunStrength);
						colorMult = colorMult * (1.0f - material.opacity);

						if(REFRACT)
						{
							float rayDist = hmin3(clast);
							if(rayDist > 0.0f)
							{
								refracted = true;
								cpos = cpos + rayDir * (rayDist + DNB_EPSILON);
								f3 vn = vox_normal(rec);
								f3 n = dot3(vn, rayDir) < 0.0f ? normalize3(vn) : hitNormal;
								rayDir = refract3(rayDir, n, st.lastVoxRefract / material.refractIndex);
								invRayDir = rcp3(rayDir);
								init_dda(rayDir, invRayDir, cpos, p, cdelta, cstep, cside);
								clast = splat3(0.0f);
							}
						}

						st.lastVoxID = thisVoxID;
						st.lastVoxRefract = material.refractIndex;
					}
				}
				else if(st.lastVoxID != 255u)
				{
					if(REFRACT)
					{
						refracted = true;
						cpos = cpos + rayDir * (hmin3(clast) + DNB_EPSILON);
						f3 oldDir = rayDir;
						f3 nn = -vox_normal(st.vox);
						f3 n = dot3(nn, rayDir) < 0.0f ? normalize3(nn) : hitNormal;
						rayDir = refract3(rayDir, n, st.lastVoxRefract);
						if(rayDir.x == 0.0f && rayDir.y == 0.0f && rayDir.z == 0.0f)
							rayDir = oldDir;
						invRayDir = rcp3(rayDir);
						init_dda(rayDir, invRayDir, cpos, p, cdelta, cstep, cside);
						clast = splat3(0.0f);
					}
					st.lastVoxID = 255u;
					st.lastVoxRefract = 1.0f;
				}

				clast = cside;
				iterate_dda(cdelta, cstep, cside, p, hitNormal);
				ignoreFirst = false;
			}
			/* ---- end step_chunk ---- */

			if(REFRACT && refracted)
			{
				rayPos = tof3(pos) + cpos * 0.125f;
				init_dda(rayDir, invRayDir, rayPos, pos, deltaDist, rayStep, sideDist);
				lastSideDist = splat3(0.0f);
			}
		}

		lastSideDist = sideDist;
		iterate_dda(deltaDist, rayStep, sideDist, pos, hitNormal);
		ignoreFirst = false;
	}

	return false;
}

#endif
