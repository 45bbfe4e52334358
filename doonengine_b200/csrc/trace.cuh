/* trace.cuh -- the two-level DDA every ray of the draw and lighting kernels runs.
 *
 * Behaviour follows /root/reference/assets/shaders/voxelShared.comp:
 *   init_dda / iterate_dda   SH:300-317  (branch-free; all tied axes step together)
 *   trace_ray  outer loop    SH:421-475  step_map: one iteration per map tile
 *              inner loop    SH:328-418  step_chunk: one iteration per voxel of a resident chunk
 * The float arithmetic (sequential sideDist accumulation, entry point min(lastSideDist) -/+ EPSILON, the
 * [EPSILON, 8-EPSILON] clamp) is kept operation for operation because first-hit indices must be bit-exact.
 * What is re-designed is everything around it (layout.h, DESIGN.md "trace"):
 *   - an empty tile costs one bit test in a register-cached 4x4x4 occupancy word; map bounds are only checked when
 *     the ray changes block (border blocks are padded with empty tiles), and a fully empty block is crossed by a
 *     tight loop that runs nothing but the DDA recurrence;
 *   - a resident tile costs one 4-byte slot lookup, a voxel step one bit test in a register-cached mask word of
 *     the 128-byte chunk slot, a hit one 16-byte record gather;
 *   - lastSideDist is carried as its minimum (the only way it is ever used), and the face normal of the last step
 *     is kept as a 3-bit mask and only materialised where the draw pass needs it (the kernels were ALU-pipe bound).
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
	uint4 a = DNB_LDG(p), b = DNB_LDG(p + 1);
	DnbMaterial m;
	m.pad[0] = m.pad[1] = 0.0f;
	m.emissive = a.z;
	m.opacity = DNB_U2F(a.w);
	m.refractIndex = DNB_U2F(b.x);
	m.specular = DNB_U2F(b.y);
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

struct Dda
{
	i3 pos, step;
	f3 delta, side;
};

DNB_FN int isgn(float a) { return (a > 0.0f) ? 1 : ((a < 0.0f) ? -1 : 0); }

/* SH:300-306.  floor(rayPos) is used as a float directly instead of converting the integer cell back
 * (identical for |rayPos| < 2^24; saves three conversions on the XU pipe per ray segment) */
DNB_FN void init_dda(f3 rayDir, f3 invRayDir, f3 rayPos, Dda& d)
{
	const f3 cell = floor3(rayPos);
	d.pos = toi3(cell);
	d.delta = abs3(invRayDir);
	const f3 sg = mk3(sgn(rayDir.x), sgn(rayDir.y), sgn(rayDir.z));
	d.step.x = isgn(rayDir.x);
	d.step.y = isgn(rayDir.y);
	d.step.z = isgn(rayDir.z);
	const f3 t = sg * (cell - rayPos) + sg * 0.5f;
	d.side = (t + 0.5f) * d.delta;
}

/* SH:309-317: every axis whose sideDist is minimal steps (ties step together).  Returns the 3-bit axis mask and
 * writes min(sideDist) BEFORE the step, which is min(lastSideDist) of the reference (SH:354,360,370,394,443). */
DNB_FN uint32_t iterate_dda(Dda& d, float& tLast)
{
	const f3 s = d.side;
	const float myz = fminf(s.y, s.z);
	const bool mx = s.x <= myz;
	const bool my = s.y <= fminf(s.z, s.x);
	const bool mz = s.z <= fminf(s.x, s.y);
	tLast = fminf(s.x, myz);
	if(mx) { d.side.x = s.x + d.delta.x; d.pos.x += d.step.x; }
	if(my) { d.side.y = s.y + d.delta.y; d.pos.y += d.step.y; }
	if(mz) { d.side.z = s.z + d.delta.z; d.pos.z += d.step.z; }
	return (mx ? 1u : 0u) | (my ? 2u : 0u) | (mz ? 4u : 0u);
}

/* normal = vec3(mask) * -rayStep (SH:316), materialised on demand */
DNB_FN f3 normal_of(uint32_t mask, i3 step)
{
	return mk3(((mask & 1u) ? 1.0f : 0.0f) * (float)(-step.x), ((mask & 2u) ? 1.0f : 0.0f) * (float)(-step.y), ((mask & 4u) ? 1.0f : 0.0f) * (float)(-step.z));
}

DNB_FN bool in_map_bounds(const DnbScene& S, i3 p)
{
	return (uint32_t)p.x < S.mapSize[0] && (uint32_t)p.y < S.mapSize[1] && (uint32_t)p.z < S.mapSize[2];
}

DNB_FN bool in_chunk_bounds(i3 p)
{
	return ((uint32_t)(p.x | p.y | p.z)) < 8u;
}

/* Exact chunk cull.  The voxel-level DDA only ever moves each coordinate in its ray's direction, so once the cell is past the
 * bounding box of the chunk's surface voxels on an axis -- beyond its maximum going up, below its minimum going down -- no voxel
 * of the chunk can be hit any more, and nothing of the voxel-level state is read after the chunk has been left without a hit.
 * Instead of testing that, the cell is stored SHIFTED per axis (by 7 - max going up, by -min going down; DnbSlot.bbox holds both
 * ready-made) so that the ordinary chunk-bounds test of the loop fires exactly there; the voxel index subtracts the shift again.
 * A ray that enters the chunk already past the box never runs the loop at all.  Only empty voxels are skipped, and an empty
 * voxel matters only while the ray is inside a transparent block (lastVoxID != 255: the state is reset there, and the draw pass
 * refracts), so the shift is used only while lastVoxID == 255 and is taken back the moment a transparent voxel is met.
 * Returns the packed shift: bits 0-3 / 4-7 / 8-11 = x / y / z offsets as 4-bit two's complement. */
DNB_FN uint32_t cull_offsets(uint32_t bbox, i3 step, i3& pos, uint32_t& bias)
{
	const int ox = step.x > 0 ? (int)(bbox & 7u) : -(int)((bbox >> 9) & 7u);
	const int oy = step.y > 0 ? (int)((bbox >> 3) & 7u) : -(int)((bbox >> 12) & 7u);
	const int oz = step.z > 0 ? (int)((bbox >> 6) & 7u) : -(int)((bbox >> 15) & 7u);
	pos.x += ox; pos.y += oy; pos.z += oz;
	bias = (uint32_t)(ox + 8 * oy + 64 * oz);
	return ((uint32_t)ox & 15u) | (((uint32_t)oy & 15u) << 4) | (((uint32_t)oz & 15u) << 8);
}

DNB_FN void cull_undo(uint32_t& offp, i3& pos, uint32_t& bias)
{
	pos.x -= ((int)(offp << 28)) >> 28;
	pos.y -= ((int)(offp << 24)) >> 28;
	pos.z -= ((int)(offp << 20)) >> 28;
	offp = 0;
	bias = 0;
}

#define DNB_COUNT(field) do { if(COUNT) lc.field++; } while(0)

/* step_map + step_chunk.  REFRACT: enableRefraction (true in draw, false in lighting, DR:65 / LI:209); only then is
 * hitNormal read or written.  invRayDir is by value: a refraction inside a chunk updates the caller's rayDir but only
 * this function's invRayDir, exactly as the inout/in qualifiers at SH:328/421 do.
 * COUNT: instrumented build -- exact per-tile bounds checks and counters, no empty-block fast path.
 * OCCLUSION: the caller only asks WHETHER the ray hits an opaque voxel (the shadow ray, LI:77-79): in a chunk whose materials are all
 * opaque the voxel's bit answers that, and neither the record nor the hit position is produced (st.vox, st.hit*, rayPos are left
 * alone; the shadow ray reads none of them, and every ray that follows sets what it reads). */
/* 1 = test the early-out at every block change, as before (A/B builds) */
#ifndef DNB_EARLYOUT_EVERY_BLOCK
#define DNB_EARLYOUT_EVERY_BLOCK 0
#endif
#ifndef DNB_OCCLUSION_RAYS
#define DNB_OCCLUSION_RAYS 1
#endif
template <bool REFRACT, bool COUNT, bool OCCLUSION = false>
DNB_FN bool trace_ray(const DnbScene& S, RayState& st, DnbCounters& lc, f3& rayDir, f3 invRayDir, f3& rayPos, bool ignoreFirst, f3& hitNormal, f3& colorAdd, float& colorMult)
{
	colorAdd = splat3(0.0f);
	colorMult = 1.0f;
	DNB_COUNT(rays);

	Dda m;
	init_dda(rayDir, invRayDir, rayPos, m);
	float tLast = 0.0f;           /* min(lastSideDist), vec3(0) before the first step (SH:432) */

	/* where the current face normal comes from: 0 = caller's value, 1 = last map-level step, 2 = last voxel-level step */
	uint32_t nmask = 0, nsrc = 0;
	Dda c;                        /* voxel-level DDA of the chunk being crossed */
	c.step.x = c.step.y = c.step.z = 0;

	uint32_t guard = 0;
	i3 blk = {0x40000000, 0x40000000, 0x40000000}; /* base tile of the cached occupancy block (never a real one) */
	unsigned long long occWord = 0;

	for(;;)
	{
		/* ---- phase A: advance tile by tile to the next resident chunk.  Kept as its own loop ("while-while"
		 * traversal) so that the lanes of a warp search together and then cross their chunks together, instead of
		 * every lane that reaches a chunk stalling the lanes that are still stepping over empty tiles. ---- */
		bool found = false;
		for(;;)
		{
			if((uint32_t)((m.pos.x ^ blk.x) | (m.pos.y ^ blk.y) | (m.pos.z ^ blk.z)) > 3u)
			{
				/* entered another 4x4x4 block: the only place the map bounds are tested */
				if(!in_map_bounds(S, m.pos))
					break;
#if DNB_EARLYOUT_EVERY_BLOCK
				if(!COUNT && ((m.pos.x > S.occMax[0] && m.step.x >= 0) || (m.pos.x < S.occMin[0] && m.step.x <= 0) ||
				              (m.pos.y > S.occMax[1] && m.step.y >= 0) || (m.pos.y < S.occMin[1] && m.step.y <= 0) ||
				              (m.pos.z > S.occMax[2] && m.step.z >= 0) || (m.pos.z < S.occMin[2] && m.step.z <= 0)))
					break;
#endif
				blk.x = m.pos.x & ~3; blk.y = m.pos.y & ~3; blk.z = m.pos.z & ~3;
				occWord = DNB_LDG(S.occ64 + ((uint32_t)(m.pos.x >> 2) + S.blocks[0] * ((uint32_t)(m.pos.y >> 2) + S.blocks[1] * (uint32_t)(m.pos.z >> 2))));
			}
			else if(COUNT && !in_map_bounds(S, m.pos))
				break;

			if(++guard > S.maxMapSteps || st.tripped)
			{
				st.tripped = true;
				break;
			}
			DNB_COUNT(tiles);

			if(!COUNT && occWord == 0ull)
			{
				/* exact early-out: past the bounding box of everything resident and moving away from it, the ray can only miss, and
				 * nothing of its DDA state is used after a miss.  Tested in EMPTY blocks only -- a block with anything in it lies
				 * inside the box or straddles its border, and a ray that is past the border there reaches an empty block next --
				 * so rays inside the populated region never pay for the twelve comparisons */
				if(!DNB_EARLYOUT_EVERY_BLOCK &&
				   ((m.pos.x > S.occMax[0] && m.step.x >= 0) || (m.pos.x < S.occMin[0] && m.step.x <= 0) ||
				    (m.pos.y > S.occMax[1] && m.step.y >= 0) || (m.pos.y < S.occMin[1] && m.step.y <= 0) ||
				    (m.pos.z > S.occMax[2] && m.step.z >= 0) || (m.pos.z < S.occMin[2] && m.step.z <= 0)))
					break;
				/* the whole block is empty (or padding outside the map): run the bare DDA recurrence until the ray leaves it */
				do
				{
					nmask = iterate_dda(m, tLast);
					guard++;
				} while((uint32_t)((m.pos.x ^ blk.x) | (m.pos.y ^ blk.y) | (m.pos.z ^ blk.z)) <= 3u && guard <= S.maxMapSteps);
				nsrc = 1;
				ignoreFirst = false;
				continue;
			}

			const uint32_t bit = (uint32_t)(m.pos.x & 3) | ((uint32_t)(m.pos.y & 3) << 2) | ((uint32_t)(m.pos.z & 3) << 4);
			if((occWord >> bit) & 1ull)
			{
				found = true;
				break;
			}
			nmask = iterate_dda(m, tLast);
			nsrc = 1;
			ignoreFirst = false;
		}
		if(!found)
			break;

		/* ---- phase B: cross the resident chunk voxel by voxel ---- */
		{
			const uint32_t mapIndex = (uint32_t)m.pos.x + S.mapSize[0] * ((uint32_t)m.pos.y + S.mapSize[1] * (uint32_t)m.pos.z);
			const DnbSlot* slot = S.slots + (DNB_LDG(S.tileSlot + mapIndex) - 1u);
			DNB_COUNT(chunks);

			/* SH:443-445: entry point in chunk-local voxel units */
			const f3 tile = tof3(m.pos);
			const f3 entry = rayPos + rayDir * (tLast - DNB_EPSILON);
			f3 cpos = (entry - tile) * 8.0f;
			cpos = min3v(max3v(cpos, splat3(DNB_EPSILON)), splat3(8.0f - DNB_EPSILON));

			/* ---- step_chunk, SH:328-418 ---- */
			bool refracted = false;
			init_dda(rayDir, invRayDir, cpos, c);
			float ctLast = 0.0f;

			uint32_t cguard = 0;
			uint32_t wordIdx = 0xFFFFFFFFu, word = 0;
			uint32_t bias = 0, offp = 0; /* exact chunk cull (see cull_offsets): c.pos is shifted by the offsets in offp */
			/* the two loads that depend on the slot index leave together: the mask word of the entry cell (needed unless the ray
			 * enters beyond the culled box) and the bounding-box word, whose top bit also says whether every material of the chunk
			 * is opaque (layout.h DNB_BBOX_OPAQUE) -- then a set voxel bit IS an opaque hit (SH:351) and the material need not be
			 * looked at inside the loop */
			uint32_t bbox = 0;
			if(!COUNT)
			{
				wordIdx = ((uint32_t)c.pos.x + 8u * ((uint32_t)c.pos.y + 8u * (uint32_t)c.pos.z)) >> 5;
				word = DNB_LDG(slot->mask + wordIdx);
				bbox = DNB_LDG(&slot->bbox);
				if(st.lastVoxID == 255u)
					offp = cull_offsets(bbox, c.step, c.pos, bias);
			}
			const bool chunkOpaque = (bbox & DNB_BBOX_OPAQUE) != 0u;
			while(in_chunk_bounds(c.pos))
			{
				if(++cguard > DNB_MAX_CHUNK_STEPS)
				{
					st.tripped = true;
					return false;
				}
				DNB_COUNT(voxelSteps);

				const uint32_t local = ((uint32_t)c.pos.x + 8u * ((uint32_t)c.pos.y + 8u * (uint32_t)c.pos.z)) - bias;
				if((local >> 5) != wordIdx)
				{
					wordIdx = local >> 5;
					word = DNB_LDG(slot->mask + wordIdx);
				}

				if(((word >> (local & 31u)) & 1u) && !ignoreFirst)
				{
					if(OCCLUSION && DNB_OCCLUSION_RAYS && !COUNT && chunkOpaque)
						return true;
					/* SH:150-169 with per-word prefix counts instead of quarter counts */
					const uint32_t rel = (uint32_t)DNB_LDG(slot->prefix + wordIdx) + DNB_POPC(word & ((1u << (local & 31u)) - 1u));
					const uint4 rec = DNB_LDG(S.records + (DNB_LDG(&slot->voxelBase) + rel));
					DNB_COUNT(records);
					st.vox = rec;

					DnbMaterial material;
					material.opacity = 1.0f;
					if(!chunkOpaque)
						material = load_material(S, rec.x >> 24);
					const uint32_t thisVoxID = (rec.y & 0xFFFFFF00u) | (rec.x >> 24);

					if(material.opacity == 1.0f)
					{
						cpos = cpos + rayDir * (ctLast + DNB_EPSILON);
						rayPos = tile + cpos * 0.125f; /* SH:451 */
						st.hitMapIndex = mapIndex;
						st.hitLocalIndex = local;
						st.hitRecord = rel;
						if(REFRACT && nsrc)
							hitNormal = normal_of(nmask, nsrc == 2u ? c.step : m.step);
						return true;
					}
					else if(st.lastVoxID != thisVoxID)
					{
						if(offp)
							cull_undo(offp, c.pos, bias); /* inside a transparent block every empty voxel counts */
						/* SH:365-366, maxDepth < 0 */
						const float cm = colorMult * material.opacity;
						colorAdd = colorAdd + (vox_albedo(rec) * cm) * ld3(S.sunStrength);
						colorMult = colorMult * (1.0f - material.opacity);

						if(REFRACT)
						{
							const float rayDist = ctLast;
							if(rayDist > 0.0f)
							{
								refracted = true;
								cpos = cpos + rayDir * (rayDist + DNB_EPSILON);
								const f3 vn = vox_normal(rec);
								const f3 faceNormal = nsrc ? normal_of(nmask, nsrc == 2u ? c.step : m.step) : hitNormal;
								const f3 n = dot3(vn, rayDir) < 0.0f ? normalize3(vn) : faceNormal;
								rayDir = refract3(rayDir, n, st.lastVoxRefract / material.refractIndex);
								invRayDir = rcp3(rayDir);
								if(nsrc)
								{
									/* the pending normal refers to the step vectors about to be replaced */
									hitNormal = faceNormal;
									nsrc = 0;
								}
								init_dda(rayDir, invRayDir, cpos, c);
								ctLast = 0.0f;
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
						cpos = cpos + rayDir * (ctLast + DNB_EPSILON);
						const f3 oldDir = rayDir;
						const f3 nn = -vox_normal(st.vox);
						const f3 faceNormal = nsrc ? normal_of(nmask, nsrc == 2u ? c.step : m.step) : hitNormal;
						const f3 n = dot3(nn, rayDir) < 0.0f ? normalize3(nn) : faceNormal;
						rayDir = refract3(rayDir, n, st.lastVoxRefract);
						if(rayDir.x == 0.0f && rayDir.y == 0.0f && rayDir.z == 0.0f)
							rayDir = oldDir;
						invRayDir = rcp3(rayDir);
						if(nsrc)
						{
							hitNormal = faceNormal;
							nsrc = 0;
						}
						init_dda(rayDir, invRayDir, cpos, c);
						ctLast = 0.0f;
					}
					st.lastVoxID = 255u;
					st.lastVoxRefract = 1.0f;
				}

				nmask = iterate_dda(c, ctLast);
				nsrc = 2;
				ignoreFirst = false;
			}
			/* ---- end step_chunk ---- */

			if(REFRACT && refracted)
			{
				if(nsrc)
				{
					hitNormal = normal_of(nmask, nsrc == 2u ? c.step : m.step);
					nsrc = 0;
				}
				rayPos = tile + cpos * 0.125f;
				init_dda(rayDir, invRayDir, rayPos, m);
				tLast = 0.0f;
			}
			else if(REFRACT && nsrc == 2u)
			{
				/* the voxel-level step vector goes out of use: pin the normal it produced */
				hitNormal = normal_of(nmask, c.step);
				nsrc = 0;
			}
		}

		nmask = iterate_dda(m, tLast);
		nsrc = 1;
		ignoreFirst = false;
	}

	if(REFRACT && nsrc)
		hitNormal = normal_of(nmask, nsrc == 2u ? c.step : m.step);
	return false;
}

#endif
