/* ray_step.cuh -- the two-level DDA of trace.cuh as ONE loop body per lane ("lock-step" traversal of the wavefront step kernel).
 *
 * trace.cuh trace_ray<false, false> -- step_map + step_chunk of voxelShared.comp:328-475, lighting rays: no refraction -- is a loop
 * nest (tiles, voxels of a chunk).  Run by a warp, every lane waits for the slowest lane of every loop; the persistent and the first
 * wavefront kernel replaced the nest by per-lane states served in phases chosen by vote, and ncu showed where that ends up
 * (profiles/r1_v3_wave.md): 18 % of all warp instructions are the vote / phase-choice scaffolding, the phases run with 8-10 lanes,
 * and chunk entry / record fetch run with 2-6 lanes WITH the rest of the warp waiting on their dependent loads.
 *
 * Here a lane's ray advances by exactly one cell per call of ray_iter(), whatever that takes: leaving a chunk, changing the cached
 * 4x4x4 block of tiles (or the cached 8x8 z-layer of the chunk's mask), entering a chunk and testing its first voxel are all folded
 * into the same iteration, in a fixed order, as plain `if`s.  A warp therefore pays one pass over this body per cell its rays
 * advance -- no votes, no phases -- and every lane is busy in every iteration; finished lanes are refilled by the caller.
 * The two DDA levels share one register set (cell, sideDist, tLast, guard, cached occupancy word); the tile level's set is parked
 * while a chunk is crossed (lighting rays never refract, so both levels use the same delta / step vectors).
 *
 * Deferred hits: when the chunk's slot says that every material it holds is opaque (DNB_BBOX_OPAQUE, maintained by the host against
 * the material table in force, layout.h), a set, non-ignored voxel bit IS the hit (SH:351: opacity == 1.0): the lane ends its ray
 * without touching the record or the material (two dependent gathers that used to run with ~2 active lanes) and leaves
 * (slot, local index, mask word) for the serve kernel, which fetches the record with full warps.
 *
 * The float arithmetic and its order are those of trace.cuh, operation for operation; tests/test_ray_step.py runs both on the CPU
 * (compiled for the host from these headers) and compares every output bit on 80 000 rays incl. glass.
 */
#ifndef DN_B200_RAY_STEP_CUH
#define DN_B200_RAY_STEP_CUH

#include "trace.cuh"

struct RayLane
{
	/* the ray segment */
	f3       dir, rayPos;      /* rayPos: origin of the segment; replaced by the hit position on a hit */
	f3       delta;            /* |1 / dir| */
	i3       step;
	/* the active level */
	uint32_t lv;               /* 0 = tiles, 1 = voxels of the chunk being crossed */
	i3       pos;              /* cell (voxel level: shifted by the cull offsets) */
	f3       side;
	float    tl;               /* min(sideDist) before the last step */
	uint32_t g;                /* loop guard of the level */
	i3       blk;              /* tile level: base of the cached 4x4x4 block; voxel level: blk.z = cached z-layer (shifted like pos) */
	unsigned long long word;   /* occupancy of the cached block / layer */
	/* the tile level, parked while lv == 1 */
	i3       mpos, mblk;
	f3       mside;
	float    mtl;
	uint32_t mg;
	unsigned long long mword;
	/* the chunk being crossed */
	uint32_t slot;             /* index into S.slots */
	uint32_t mapIndex;
	f3       cpos;             /* entry point in chunk-local voxel units */
	i3       off;              /* cull offsets in force (0 = none) */
	bool     chunkOpaque;
	/* what the ray carries and returns */
	bool     ignoreFirst, hit, deferred, tripped;
	uint32_t lastVoxID;
	float    lastVoxRefract;
	uint4    vox;              /* record hit -- or, for a deferred hit, {slot, local index, mask word of the voxel, 0} */
	uint32_t hitLocal, hitRecord;
	f3       colorAdd;
	float    colorMult;
};

/* the segment's direction / origin / carried state are set by the caller; this is the prologue of trace_ray */
DNB_FN void ray_begin(RayLane& L, f3 inv)
{
	Dda m;
	init_dda(L.dir, inv, L.rayPos, m);
	L.delta = m.delta;
	L.step = m.step;
	L.pos = m.pos;
	L.side = m.side;
	L.tl = 0.0f;
	L.g = 0;
	L.lv = 0;
	L.blk.x = L.blk.y = L.blk.z = 0x40000000;
	L.word = 0;
	L.off.x = L.off.y = L.off.z = 0;
	L.colorAdd = splat3(0.0f);
	L.colorMult = 1.0f;
	L.hit = false;
	L.deferred = false;
}

DNB_FN void ray_advance(RayLane& L)
{
	const f3 s = L.side;
	const float myz = fminf(s.y, s.z);
	const bool mx = s.x <= myz;
	const bool my = s.y <= fminf(s.z, s.x);
	const bool mz = s.z <= fminf(s.x, s.y);
	L.tl = fminf(s.x, myz);
	if(mx) { L.side.x = s.x + L.delta.x; L.pos.x += L.step.x; }
	if(my) { L.side.y = s.y + L.delta.y; L.pos.y += L.step.y; }
	if(mz) { L.side.z = s.z + L.delta.z; L.pos.z += L.step.z; }
	L.ignoreFirst = false;
}

/* one iteration: returns true when the ray has ended (L.hit says how).  DEFER: allow deferred hits (see above). */
template <bool DEFER>
DNB_FN bool ray_iter(const DnbScene& S, RayLane& L)
{
	/* ---- left the chunk (or its culled box) without a hit: back to the tile level, one tile on (trace.cuh:403-405) ---- */
	if(L.lv != 0u && !in_chunk_bounds(L.pos))
	{
		L.pos = L.mpos; L.side = L.mside; L.tl = L.mtl; L.g = L.mg; L.blk = L.mblk; L.word = L.mword;
		L.off.x = L.off.y = L.off.z = 0;
		L.lv = 0;
		ray_advance(L);
	}

	if(L.lv == 0u)
	{
		/* ---- tile level: trace.cuh phase A ---- */
		if((uint32_t)((L.pos.x ^ L.blk.x) | (L.pos.y ^ L.blk.y) | (L.pos.z ^ L.blk.z)) > 3u)
		{
			const i3 p = L.pos;
			if(!in_map_bounds(S, p) ||
			   (p.x > S.occMax[0] && L.step.x >= 0) || (p.x < S.occMin[0] && L.step.x <= 0) ||
			   (p.y > S.occMax[1] && L.step.y >= 0) || (p.y < S.occMin[1] && L.step.y <= 0) ||
			   (p.z > S.occMax[2] && L.step.z >= 0) || (p.z < S.occMin[2] && L.step.z <= 0))
				return true; /* miss */
			L.blk.x = p.x & ~3; L.blk.y = p.y & ~3; L.blk.z = p.z & ~3;
			L.word = DNB_LDG(S.occ64 + ((uint32_t)(p.x >> 2) + S.blocks[0] * ((uint32_t)(p.y >> 2) + S.blocks[1] * (uint32_t)(p.z >> 2))));
		}
		if(++L.g > S.maxMapSteps || L.tripped)
		{
			L.tripped = true;
			return true;
		}
		const uint32_t bit = (uint32_t)(L.pos.x & 3) | ((uint32_t)(L.pos.y & 3) << 2) | ((uint32_t)(L.pos.z & 3) << 4);
		if(!((L.word >> bit) & 1ull))
		{
			ray_advance(L);
			return false;
		}

		/* ---- the tile holds a chunk: SH:443-445 + step_chunk's prologue (trace.cuh:258-277) ---- */
		L.mapIndex = (uint32_t)L.pos.x + S.mapSize[0] * ((uint32_t)L.pos.y + S.mapSize[1] * (uint32_t)L.pos.z);
		L.slot = DNB_LDG(S.tileSlot + L.mapIndex) - 1u;
		const DnbSlot* slot = S.slots + L.slot;
		const f3 tile = tof3(L.pos);
		const f3 entry = L.rayPos + L.dir * (L.tl - DNB_EPSILON);
		f3 cpos = (entry - tile) * 8.0f;
		cpos = min3v(max3v(cpos, splat3(DNB_EPSILON)), splat3(8.0f - DNB_EPSILON));
		L.cpos = cpos;
		const f3 cell = floor3(cpos);
		const i3 ci = toi3(cell);
		/* both loads that depend on the slot index leave together: the layer word of the entry cell (needed unless the ray enters
		 * beyond the culled box) and the bounding-box word */
		const unsigned long long layer = DNB_LDG(reinterpret_cast<const unsigned long long*>(slot->mask) + (uint32_t)ci.z);
		const uint32_t bbox = DNB_LDG(&slot->bbox);
		const f3 sg = mk3(sgn(L.dir.x), sgn(L.dir.y), sgn(L.dir.z));
		const f3 t = sg * (cell - cpos) + sg * 0.5f;

		L.mpos = L.pos; L.mside = L.side; L.mtl = L.tl; L.mg = L.g; L.mblk = L.blk; L.mword = L.word;
		L.pos = ci;
		L.side = (t + 0.5f) * L.delta;
		L.tl = 0.0f;
		L.g = 0;
		L.lv = 1;
		L.chunkOpaque = (bbox & DNB_BBOX_OPAQUE) != 0u;
		if(L.lastVoxID == 255u)
		{
			L.off.x = L.step.x > 0 ? (int)(bbox & 7u) : -(int)((bbox >> 9) & 7u);
			L.off.y = L.step.y > 0 ? (int)((bbox >> 3) & 7u) : -(int)((bbox >> 12) & 7u);
			L.off.z = L.step.z > 0 ? (int)((bbox >> 6) & 7u) : -(int)((bbox >> 15) & 7u);
			L.pos.x += L.off.x; L.pos.y += L.off.y; L.pos.z += L.off.z;
		}
		if(!in_chunk_bounds(L.pos))
			return false; /* entered beyond the box: nothing of this chunk can be hit; the next iteration leaves it */
		L.blk.z = L.pos.z;
		L.word = layer;
	}
	else if(L.pos.z != L.blk.z)
	{
		/* next z-layer of the chunk */
		L.blk.z = L.pos.z;
		L.word = DNB_LDG(reinterpret_cast<const unsigned long long*>(S.slots[L.slot].mask) + (uint32_t)(L.pos.z - L.off.z));
	}

	/* ---- voxel level: one iteration of step_chunk's loop (trace.cuh:278-381) ---- */
	if(++L.g > DNB_MAX_CHUNK_STEPS)
	{
		L.tripped = true;
		return true;
	}
	const uint32_t xi = (uint32_t)(L.pos.x - L.off.x), yi = (uint32_t)(L.pos.y - L.off.y);
	const uint32_t idx = xi | (yi << 3);
	if(((L.word >> idx) & 1ull) && !L.ignoreFirst)
	{
		const uint32_t zi = (uint32_t)(L.pos.z - L.off.z);
		const uint32_t local = idx + 64u * zi;
		const uint32_t word32 = (uint32_t)(L.word >> (idx & 32u));
		if(DEFER && L.chunkOpaque)
		{
			/* every material of this chunk is opaque: this IS the hit (SH:351-356); the record is fetched by the serve kernel */
			const f3 cpos = L.cpos + L.dir * (L.tl + DNB_EPSILON);
			L.rayPos = tof3(L.mpos) + cpos * 0.125f;
			L.vox = make_uint4(L.slot, local, word32, 0u);
			L.hitLocal = local;
			L.hit = true;
			L.deferred = true;
			return true;
		}
		const DnbSlot* slot = S.slots + L.slot;
		const uint32_t rel = (uint32_t)DNB_LDG(slot->prefix + (local >> 5)) + DNB_POPC(word32 & ((1u << (local & 31u)) - 1u));
		const uint4 rec = DNB_LDG(S.records + (DNB_LDG(&slot->voxelBase) + rel));
		L.vox = rec;
		const DnbMaterial material = load_material(S, rec.x >> 24);
		const uint32_t thisVoxID = (rec.y & 0xFFFFFF00u) | (rec.x >> 24);
		if(material.opacity == 1.0f)
		{
			const f3 cpos = L.cpos + L.dir * (L.tl + DNB_EPSILON);
			L.rayPos = tof3(L.mpos) + cpos * 0.125f;
			L.hitLocal = local;
			L.hitRecord = rel;
			L.hit = true;
			return true;
		}
		if(L.lastVoxID != thisVoxID)
		{
			/* inside a transparent block every empty voxel counts: take the cull offsets back (trace.cuh cull_undo) */
			L.pos.x -= L.off.x; L.pos.y -= L.off.y; L.pos.z -= L.off.z;
			L.off.x = L.off.y = L.off.z = 0;
			L.blk.z = L.pos.z;
			const float cm = L.colorMult * material.opacity;
			L.colorAdd = L.colorAdd + (vox_albedo(rec) * cm) * ld3(S.sunStrength);
			L.colorMult = L.colorMult * (1.0f - material.opacity);
			L.lastVoxID = thisVoxID;
			L.lastVoxRefract = material.refractIndex;
		}
	}
	else if(L.lastVoxID != 255u)
	{
		L.lastVoxID = 255u;
		L.lastVoxRefract = 1.0f;
	}
	ray_advance(L);
	return false;
}

/* the record of a deferred hit: what the serve kernel does with RayLane.vox = {slot, local, mask word} */
DNB_FN uint4 ray_deferred_record(const DnbScene& S, uint4 d, uint32_t* relOut)
{
	const DnbSlot* slot = S.slots + d.x;
	const uint32_t local = d.y;
	const uint32_t rel = (uint32_t)DNB_LDG(slot->prefix + (local >> 5)) + DNB_POPC(d.z & ((1u << (local & 31u)) - 1u));
	if(relOut)
		*relOut = rel;
	return DNB_LDG(S.records + (DNB_LDG(&slot->voxelBase) + rel));
}

#endif
