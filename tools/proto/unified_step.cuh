/* unified_step.cuh -- PROTOTYPE (not part of libdoon_b200.so): the two levels of the DDA as ONE stepping routine.
 *
 * Why (profiles/r1_v3_wave.md): in dn_wave_step_kernel the tile level and the voxel level take turns inside a warp (8-10 of 32
 * lanes each), and the rare expensive branches -- entering a chunk, fetching a record -- run with 4-6 lanes.  Here a lane's ACTIVE
 * level lives in one set of registers (cell, sideDist, tLast, guard, the 64-bit occupancy word of the block it is in and that
 * block's base), the tile level's set is parked while a chunk is crossed, and the cheap step is the same instructions for both
 * levels: in-block test with per-level masks, guard, one bit test in the cached word, the DDA recurrence.  At the tile level a
 * block is 4x4x4 tiles (the occ64 word); at the voxel level it is one z-layer of the chunk (8x8x1 voxels = mask words 2z, 2z+1 read
 * as one 64-bit word).  Everything else is an EVENT a lane waits in until enough lanes want the same one:
 *     BOUNDARY  left the cached block: map bounds / next occ64 word -- or next z-layer of the chunk / chunk exit
 *     ENTER     the tile's bit is set: entry point, voxel-level DDA start, exact chunk cull (trace.cuh cull_offsets)
 *     HIT       the voxel's bit is set: record fetch, opaque -> the ray ends, transparent -> accumulate and go on
 * The float arithmetic and its order are those of trace.cuh trace_ray<false, false> (lighting rays: no refraction), which the host
 * harness (step_harness.cu, tests/test_unified_step_proto.py) checks bit for bit, ray state included.
 */
#ifndef DN_B200_UNIFIED_STEP_CUH
#define DN_B200_UNIFIED_STEP_CUH

#include "trace.cuh"

enum : uint32_t { U_STEP = 0, U_BOUNDARY = 1, U_ENTER = 2, U_HIT = 3, U_END = 4 };

struct UniLane
{
	/* the ray segment */
	f3       dir, inv, rayPos; /* rayPos: origin, replaced by the hit position on a hit */
	f3       delta;
	i3       step;
	/* the active level */
	uint32_t lv;               /* 0 = tiles, 1 = voxels of the chunk being crossed */
	i3       pos;              /* cell (voxel level: shifted by the cull offsets) */
	f3       side;
	float    tl;               /* min(sideDist) before the last step */
	uint32_t g;                /* loop guard of the level */
	i3       blk;              /* base of the cached block */
	uint32_t mskxy, sh1, mskz; /* level parameters: in-block masks and bit-index layout (3, 2, 3 tiles; 7, 3, 0 voxels) */
	unsigned long long word;   /* occupancy of the cached block */
	/* the tile level, parked while lv == 1 */
	i3       mpos, mblk;
	f3       mside;
	float    mtl;
	uint32_t mg;
	unsigned long long mword;
	/* the chunk being crossed */
	const DnbSlot* slot;
	uint32_t mapIndex;
	f3       tile, cpos;
	i3       off;              /* cull offsets in force (0 = none) */
	/* what the ray carries */
	bool     ignoreFirst, hit;
	RayState st;
	f3       colorAdd;
	float    colorMult;
};

DNB_FN void uni_level_params(UniLane& L, uint32_t lv)
{
	L.lv = lv;
	L.mskxy = lv ? 7u : 3u;
	L.sh1 = lv ? 3u : 2u;
	L.mskz = lv ? 0u : 3u;
}

/* flat_start_ray / the prologue of trace_ray */
DNB_FN void uni_start(UniLane& L, uint32_t& state)
{
	Dda m;
	init_dda(L.dir, L.inv, L.rayPos, m);
	L.delta = m.delta;
	L.step = m.step;
	L.pos = m.pos;
	L.side = m.side;
	L.tl = 0.0f;
	L.g = 0;
	L.blk.x = L.blk.y = L.blk.z = 0x40000000;
	L.word = 0;
	uni_level_params(L, 0);
	L.off.x = L.off.y = L.off.z = 0;
	L.colorAdd = splat3(0.0f);
	L.colorMult = 1.0f;
	L.hit = false;
	state = U_STEP;
}

DNB_FN void uni_iterate(UniLane& L)
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
}

DNB_FN bool uni_in_block(const UniLane& L)
{
	return ((((uint32_t)(L.pos.x ^ L.blk.x) | (uint32_t)(L.pos.y ^ L.blk.y)) & ~L.mskxy) | ((uint32_t)(L.pos.z ^ L.blk.z) & ~L.mskz)) == 0u;
}

/* one cheap iteration, the same instructions at both levels */
DNB_FN void uni_step(const DnbScene& S, UniLane& L, uint32_t& state)
{
	if(!uni_in_block(L))
	{
		state = U_BOUNDARY;
		return;
	}
	const uint32_t limit = L.lv ? (uint32_t)DNB_MAX_CHUNK_STEPS : S.maxMapSteps;
	if(++L.g > limit || (L.lv == 0u && L.st.tripped))
	{
		L.st.tripped = true;
		state = U_END; /* miss */
		return;
	}

	if(L.lv == 0u && L.word == 0ull)
	{
		/* an empty 4x4x4 block of tiles: the bare recurrence until the ray leaves it (trace.cuh, same guard accounting) */
		do
		{
			uni_iterate(L);
			L.g++;
		} while(uni_in_block(L) && L.g <= S.maxMapSteps);
		L.ignoreFirst = false;
		return;
	}

	const uint32_t xi = (uint32_t)(L.pos.x - L.off.x), yi = (uint32_t)(L.pos.y - L.off.y), zi = (uint32_t)(L.pos.z - L.off.z);
	const uint32_t idx = (xi & L.mskxy) | ((yi & L.mskxy) << L.sh1) | ((zi & L.mskz) << 4);
	const bool bit = (L.word >> idx) & 1ull;
	if(bit && (L.lv == 0u || !L.ignoreFirst))
	{
		state = L.lv ? U_HIT : U_ENTER;
		return;
	}
	if(L.lv != 0u && L.st.lastVoxID != 255u)
	{
		L.st.lastVoxID = 255u;
		L.st.lastVoxRefract = 1.0f;
	}
	uni_iterate(L);
	L.ignoreFirst = false;
}

/* left the cached block */
DNB_FN void uni_boundary(const DnbScene& S, UniLane& L, uint32_t& state)
{
	if(L.lv == 0u)
	{
		const i3 p = L.pos;
		if(!in_map_bounds(S, p) ||
		   (p.x > S.occMax[0] && L.step.x >= 0) || (p.x < S.occMin[0] && L.step.x <= 0) ||
		   (p.y > S.occMax[1] && L.step.y >= 0) || (p.y < S.occMin[1] && L.step.y <= 0) ||
		   (p.z > S.occMax[2] && L.step.z >= 0) || (p.z < S.occMin[2] && L.step.z <= 0))
		{
			state = U_END; /* miss */
			return;
		}
		L.blk.x = p.x & ~3; L.blk.y = p.y & ~3; L.blk.z = p.z & ~3;
		L.word = DNB_LDG(S.occ64 + ((uint32_t)(p.x >> 2) + S.blocks[0] * ((uint32_t)(p.y >> 2) + S.blocks[1] * (uint32_t)(p.z >> 2))));
		state = U_STEP;
		return;
	}
	if(!in_chunk_bounds(L.pos))
	{
		/* left the chunk (or its culled box) without a hit: back to the tile level, one tile on */
		L.pos = L.mpos; L.side = L.mside; L.tl = L.mtl; L.g = L.mg; L.blk = L.mblk; L.word = L.mword;
		L.off.x = L.off.y = L.off.z = 0;
		uni_level_params(L, 0);
		uni_iterate(L);
		L.ignoreFirst = false;
		state = U_STEP;
		return;
	}
	/* next z-layer of the chunk */
	L.blk.z = L.pos.z;
	L.word = DNB_LDG(reinterpret_cast<const unsigned long long*>(L.slot->mask) + (uint32_t)(L.pos.z - L.off.z));
	state = U_STEP;
}

/* the tile holds a chunk: SH:443-445 + step_chunk's prologue (flat_tile_step's chunk branch) */
DNB_FN void uni_enter(const DnbScene& S, UniLane& L, uint32_t& state)
{
	L.mapIndex = (uint32_t)L.pos.x + S.mapSize[0] * ((uint32_t)L.pos.y + S.mapSize[1] * (uint32_t)L.pos.z);
	L.slot = S.slots + (DNB_LDG(S.tileSlot + L.mapIndex) - 1u);
	L.tile = tof3(L.pos);
	const f3 entry = L.rayPos + L.dir * (L.tl - DNB_EPSILON);
	f3 cpos = (entry - L.tile) * 8.0f;
	cpos = min3v(max3v(cpos, splat3(DNB_EPSILON)), splat3(8.0f - DNB_EPSILON));
	L.cpos = cpos;
	const f3 cell = floor3(cpos);
	const f3 sg = mk3(sgn(L.dir.x), sgn(L.dir.y), sgn(L.dir.z));
	const f3 t = sg * (cell - cpos) + sg * 0.5f;

	L.mpos = L.pos; L.mside = L.side; L.mtl = L.tl; L.mg = L.g; L.mblk = L.blk; L.mword = L.word;
	L.pos = toi3(cell);
	L.side = (t + 0.5f) * L.delta;
	L.tl = 0.0f;
	L.g = 0;
	uni_level_params(L, 1);
	L.off.x = L.off.y = L.off.z = 0;
	if(L.st.lastVoxID == 255u)
	{
		const uint32_t bbox = DNB_LDG(&L.slot->bbox);
		L.off.x = L.step.x > 0 ? (int)(bbox & 7u) : -(int)((bbox >> 9) & 7u);
		L.off.y = L.step.y > 0 ? (int)((bbox >> 3) & 7u) : -(int)((bbox >> 12) & 7u);
		L.off.z = L.step.z > 0 ? (int)((bbox >> 6) & 7u) : -(int)((bbox >> 15) & 7u);
		L.pos.x += L.off.x; L.pos.y += L.off.y; L.pos.z += L.off.z;
	}
	L.blk.x = L.blk.y = 0;
	if(in_chunk_bounds(L.pos))
	{
		L.blk.z = L.pos.z;
		L.word = DNB_LDG(reinterpret_cast<const unsigned long long*>(L.slot->mask) + (uint32_t)(L.pos.z - L.off.z));
	}
	else
		L.blk.z = 0x40000000; /* entered beyond the box: the first step goes to BOUNDARY and from there back to the tiles */
	state = U_STEP;
}

/* the voxel's bit is set (and it is not the ignored first voxel): flat_vox_step's record branch */
DNB_FN void uni_hit(const DnbScene& S, UniLane& L, uint32_t& state)
{
	const uint32_t xi = (uint32_t)(L.pos.x - L.off.x), yi = (uint32_t)(L.pos.y - L.off.y), zi = (uint32_t)(L.pos.z - L.off.z);
	const uint32_t local = xi + 8u * (yi + 8u * zi);
	const uint32_t wordIdx = local >> 5;
	const uint32_t word32 = (uint32_t)(L.word >> (local & 32u));
	const uint32_t rel = (uint32_t)DNB_LDG(L.slot->prefix + wordIdx) + DNB_POPC(word32 & ((1u << (local & 31u)) - 1u));
	const uint4 rec = DNB_LDG(S.records + (DNB_LDG(&L.slot->voxelBase) + rel));
	L.st.vox = rec;
	const DnbMaterial material = load_material(S, rec.x >> 24);
	const uint32_t thisVoxID = (rec.y & 0xFFFFFF00u) | (rec.x >> 24);

	if(material.opacity == 1.0f)
	{
		const f3 cpos = L.cpos + L.dir * (L.tl + DNB_EPSILON);
		L.rayPos = L.tile + cpos * 0.125f;
		L.st.hitMapIndex = L.mapIndex;
		L.st.hitLocalIndex = local;
		L.st.hitRecord = rel;
		L.hit = true;
		state = U_END;
		return;
	}
	if(L.st.lastVoxID != thisVoxID)
	{
		/* inside a transparent block every empty voxel counts: take the cull offsets back */
		L.pos.x -= L.off.x; L.pos.y -= L.off.y; L.pos.z -= L.off.z;
		L.off.x = L.off.y = L.off.z = 0;
		L.blk.z = L.pos.z;
		const float cm = L.colorMult * material.opacity;
		L.colorAdd = L.colorAdd + (vox_albedo(rec) * cm) * ld3(S.sunStrength);
		L.colorMult = L.colorMult * (1.0f - material.opacity);
		L.st.lastVoxID = thisVoxID;
		L.st.lastVoxRefract = material.refractIndex;
	}
	uni_iterate(L);
	L.ignoreFirst = false;
	state = U_STEP;
}

#endif
