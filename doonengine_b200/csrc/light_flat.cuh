/* light_flat.cuh (included at the end of light.cu, whose __constant__ tables it shares) -- the per-voxel lighting update as a warp-persistent state machine ("flat" lighting kernel).
 *
 * Same arithmetic as light.cu / trace.cuh (voxelLighting.comp LI:65-288 over voxelShared.comp SH:300-475), float
 * operation for float operation; what changes is WHO executes WHEN.  In dn_light_kernel a lane owns one voxel and the
 * warp walks the nested loops (rays -> tiles -> voxels) together, so every loop runs until its slowest lane is done:
 * ncu showed 13-15 of 32 lanes active on the terrain map and far fewer on sparse maps, where one ray of a warp crosses
 * a hundred tiles while the others hit a neighbour at once.  Here every lane carries the complete state of its voxel
 * AND of the ray it is tracing, and is always in one of five states:
 *
 *     TILE   stepping the tile-level DDA                       (phase A of trace_ray)
 *     ENTER  has found a resident chunk: slot lookup, entry cell, bounding-box offsets (~95 instructions, two dependent loads)
 *     VOX    stepping the voxel-level DDA inside a chunk       (phase B of trace_ray)
 *     END    its ray has ended: fetch the record it hit (opaque hits leave that to this phase), shade it, advance the voxel's
 *            ray schedule (specular rays, diffuse bounces, shadow ray), start the next ray, or store the voxel's result ...
 *     FETCH  ... and take the next voxel from a global work counter (persistent warps, dynamic fetch)
 *
 * Each trip round the warp's loop takes a census of the lanes' states (four ballots) and runs ONE phase:
 *   serve   (END / FETCH lanes) once `endLanes` lanes wait for it, or nothing else can run, or the waiting lanes have sat
 *           through `patience` stepping iterations: ray set-up and shading is by far the most expensive phase, so it should
 *           run with as many lanes as possible, but a few very long rays must not hold everybody up;
 *   ENTER   when it is the most populated state;
 *   TILE or VOX burst   the stepping phase with more lanes, until fewer than keep/8 of its lanes are still stepping.
 * A lane that finishes early is not idle until the warp's slowest ray ends; it waits only until enough other lanes want
 * the same phase.  Nothing about a voxel's own sequence of operations changes, so the staged words are bit-identical to
 * dn_light_kernel's (tests/test_parity_gpu.py runs every lighting test against every kernel).
 *
 * The rule that emerged in round 2 (profiles/r2_light.md sections 8-9): whatever is rare per step and dear per occurrence must be a
 * batched phase, not code inline in a step where it runs with the 2-4 lanes that happen to need it in the same iteration (chunk
 * entry, the record -> material fetch of a hit), and phases must be long, because the census costs as much as a step.  Defaults
 * (light.cu): endLanes 20, patience 48, keep 2/8; sparse 1024^3 dispatch 143 -> 106 ms, 9.7 -> 13.3 active lanes per instruction.
 *
 * Measured (B200, profiles/r2_light.md section 6 and r2_bench_*.json): the fastest kernel on the sparse map (111 vs ~210 ms per
 * dispatch at 1024^3) and on the edit-stream map; slower than dn_light_kernel on the terrain map (0.91 vs 0.42 ms), whose rays are a
 * handful of steps long and end together, and on late frames of the dense map.  The host times the candidates on live dispatches,
 * split between them, and runs the fastest (engine.cpp pick_light_kernels).
 *
 * The ray-persistent state of one shader invocation (lastVoxID / lastVoxRefract / voxel, SH:321-325) is per lane and
 * reset per voxel, as one invocation lights one voxel.  Lighting never refracts (LI:209), so the chunk-level DDA shares
 * the tile-level DDA's step and delta vectors.
 */
/* opaque hits leave their record to the END phase (flat_hit_record); 0 = fetch it inside the voxel step, as before (A/B builds) */
#ifndef FLAT_DEFER_RECORD
#define FLAT_DEFER_RECORD 1
#endif
/* two steps per census of a stepping phase (A/B builds) */
#ifndef FLAT_UNROLL2
#define FLAT_UNROLL2 0
#endif

enum : uint32_t { ST_FETCH = 0, ST_TILE = 1, ST_VOX = 2, ST_END = 3, ST_DONE = 4, ST_ENTER = 5 };
enum : uint32_t { RAY_SPEC = 0, RAY_DIFFUSE = 1, RAY_SHADOW = 2 };

struct FlatLane
{
	/* ---- the voxel being lit ---- */
	uint4    rec;            /* its record */
	f3       origin;         /* ray origin, LI:231-232 */
	f3       spec, diff;     /* accumulators */
	float    indirectSamples;
	size_t   at;             /* staging word index of this voxel: 96 * request + lane-in-request */
	bool     firstSample, sourceVisible;
	uint32_t kind, idx, seg; /* ray schedule: kind, specular ray / diffuse sample index, segment (bounce) */
	/* ---- path state: specular (lastPos, multiplier, reflectType) or diffuse (newColor, lastDir) ---- */
	f3       pa, pb;
	uint32_t reflectType;
	/* ---- survives from ray to ray of this voxel ---- */
	RayState st;
	/* ---- the ray segment being traced ---- */
	f3       dir, inv, pos;  /* pos = rayPos: segment origin, replaced by the hit position on a hit */
	Dda      m;              /* tile-level DDA; m.delta / m.step also serve the voxel level */
	float    tLast;
	bool     ignoreFirst, hit;
	uint32_t guard;
	i3       blk;
	unsigned long long occWord;
	f3       colorAdd;
	float    colorMult;
	/* ---- the chunk being crossed ---- */
	i3       cp;
	f3       cside, cpos, tile;
	float    ctLast;
	const DnbSlot* slot;
	uint32_t wordIdx, word, cguard, mapIndex;
	uint32_t cbias, coffp;   /* exact chunk cull (trace.cuh cull_offsets): cp is shifted by the offsets in coffp */
	bool     chunkOpaque;    /* every material of the chunk is opaque (layout.h DNB_BBOX_OPAQUE): a set voxel bit is an opaque hit */
};

/* ---------------------------------------------------------------------------------------------------------------- */
/* ray segment: start, one tile step, one voxel step (trace.cuh trace_ray<false,false>, unrolled into steps)          */

/* starts tracing the segment whose direction the schedule functions below have left in L.dir (they return true for "a ray
 * is ready"); kept as ONE call site in the kernel so that lanes with different kinds of ray converge for it */
DNB_FN void flat_start_ray(FlatLane& L, uint32_t& state)
{
	const f3 dir = L.dir;
	L.inv = rcp3(dir);
	L.colorAdd = splat3(0.0f);
	L.colorMult = 1.0f;
	init_dda(L.dir, L.inv, L.pos, L.m);
	L.tLast = 0.0f;
	L.ignoreFirst = true; /* every lighting ray starts inside the voxel it leaves (LI:77,94,172) */
	L.guard = 0;
	L.blk.x = L.blk.y = L.blk.z = 0x40000000;
	L.occWord = 0;
	L.hit = false;
	state = ST_TILE;
}

/* chunk entry, SH:443-445 and step_chunk's prologue: the dearest operation of a ray segment (~95 instructions, two dependent loads).
 * Inside the tile step it ran with the ~4 lanes that happened to find a chunk in the same iteration; as a phase of its own it runs
 * when it is the most populated state of the warp. */
DNB_FN void flat_enter_chunk(const DnbScene& S, FlatLane& L, uint32_t& state)
{
	Dda& m = L.m;
	/* resident chunk: SH:443-445, then step_chunk's prologue */
	L.mapIndex = (uint32_t)m.pos.x + S.mapSize[0] * ((uint32_t)m.pos.y + S.mapSize[1] * (uint32_t)m.pos.z);
	L.slot = S.slots + (__ldg(S.tileSlot + L.mapIndex) - 1u);
	L.tile = tof3(m.pos);
	const f3 entry = L.pos + L.dir * (L.tLast - DNB_EPSILON);
	f3 cpos = (entry - L.tile) * 8.0f;
	cpos = min3v(max3v(cpos, splat3(DNB_EPSILON)), splat3(8.0f - DNB_EPSILON));
	L.cpos = cpos;
	/* init_dda(rayDir, invRayDir, cpos, c): delta and step equal the tile level's */
	const f3 cell = floor3(cpos);
	L.cp = toi3(cell);
	const f3 sg = mk3(sgn(L.dir.x), sgn(L.dir.y), sgn(L.dir.z));
	const f3 t = sg * (cell - cpos) + sg * 0.5f;
	L.cside = (t + 0.5f) * m.delta;
	L.ctLast = 0.0f;
	L.cguard = 0;
	/* both loads that depend on the slot index leave together (the entry cell's mask word and the bounding-box word, whose top
	 * bit says whether every material of the chunk is opaque: trace.cuh) */
	L.wordIdx = ((uint32_t)L.cp.x + 8u * ((uint32_t)L.cp.y + 8u * (uint32_t)L.cp.z)) >> 5;
	L.word = __ldg(L.slot->mask + L.wordIdx);
	const uint32_t bbox = __ldg(&L.slot->bbox);
	L.chunkOpaque = (bbox & DNB_BBOX_OPAQUE) != 0u;
	L.cbias = 0;
	L.coffp = 0;
	if(L.st.lastVoxID == 255u)
		L.coffp = cull_offsets(bbox, m.step, L.cp, L.cbias);
	state = ST_VOX;
}

/* one iteration of trace_ray's phase-A loop */
DNB_FN void flat_tile_step(const DnbScene& S, FlatLane& L, uint32_t& state)
{
	Dda& m = L.m;
	if((uint32_t)((m.pos.x ^ L.blk.x) | (m.pos.y ^ L.blk.y) | (m.pos.z ^ L.blk.z)) > 3u)
	{
		if(!in_map_bounds(S, m.pos)
#if DNB_EARLYOUT_EVERY_BLOCK
		   || (m.pos.x > S.occMax[0] && m.step.x >= 0) || (m.pos.x < S.occMin[0] && m.step.x <= 0) ||
		   (m.pos.y > S.occMax[1] && m.step.y >= 0) || (m.pos.y < S.occMin[1] && m.step.y <= 0) ||
		   (m.pos.z > S.occMax[2] && m.step.z >= 0) || (m.pos.z < S.occMin[2] && m.step.z <= 0)
#endif
		  )
		{
			state = ST_END; /* miss */
			return;
		}
		L.blk.x = m.pos.x & ~3; L.blk.y = m.pos.y & ~3; L.blk.z = m.pos.z & ~3;
		L.occWord = __ldg(S.occ64 + ((uint32_t)(m.pos.x >> 2) + S.blocks[0] * ((uint32_t)(m.pos.y >> 2) + S.blocks[1] * (uint32_t)(m.pos.z >> 2))));
	}

	if(++L.guard > S.maxMapSteps || L.st.tripped)
	{
		L.st.tripped = true;
		state = ST_END;
		return;
	}

	if(L.occWord == 0ull)
	{
		/* exact early-out (trace.cuh): past the bounding box of everything resident and moving away; tested in empty blocks only */
		if(!DNB_EARLYOUT_EVERY_BLOCK &&
		   ((m.pos.x > S.occMax[0] && m.step.x >= 0) || (m.pos.x < S.occMin[0] && m.step.x <= 0) ||
		    (m.pos.y > S.occMax[1] && m.step.y >= 0) || (m.pos.y < S.occMin[1] && m.step.y <= 0) ||
		    (m.pos.z > S.occMax[2] && m.step.z >= 0) || (m.pos.z < S.occMin[2] && m.step.z <= 0)))
		{
			state = ST_END; /* miss */
			return;
		}
		/* empty block: the bare recurrence until the ray leaves it */
		do
		{
			iterate_dda(m, L.tLast);
			L.guard++;
		} while((uint32_t)((m.pos.x ^ L.blk.x) | (m.pos.y ^ L.blk.y) | (m.pos.z ^ L.blk.z)) <= 3u && L.guard <= S.maxMapSteps);
		L.ignoreFirst = false;
		return;
	}

	const uint32_t bit = (uint32_t)(m.pos.x & 3) | ((uint32_t)(m.pos.y & 3) << 2) | ((uint32_t)(m.pos.z & 3) << 4);
	if((L.occWord >> bit) & 1ull)
	{
		/* resident chunk: entered as a phase of its own (flat_enter_chunk), by all the lanes that have found one */
		state = ST_ENTER;
		return;
	}
	iterate_dda(m, L.tLast);
	L.ignoreFirst = false;
}

/* one iteration of step_chunk's loop (SH:339-415), or the exit from it */
DNB_FN void flat_vox_step(const DnbScene& S, FlatLane& L, uint32_t& state)
{
	if(!in_chunk_bounds(L.cp))
	{
		/* left the chunk without a hit: next tile */
		iterate_dda(L.m, L.tLast);
		L.ignoreFirst = false;
		state = ST_TILE;
		return;
	}
	if(++L.cguard > DNB_MAX_CHUNK_STEPS)
	{
		L.st.tripped = true;
		state = ST_END;
		return;
	}

	const uint32_t local = ((uint32_t)L.cp.x + 8u * ((uint32_t)L.cp.y + 8u * (uint32_t)L.cp.z)) - L.cbias;
	if((local >> 5) != L.wordIdx)
	{
		L.wordIdx = local >> 5;
		L.word = __ldg(L.slot->mask + L.wordIdx);
	}

	if(((L.word >> (local & 31u)) & 1u) && !L.ignoreFirst)
	{
		if(FLAT_DEFER_RECORD && L.chunkOpaque)
		{
			/* every material of this chunk is opaque: the set bit IS the hit (SH:351).  The record -- prefix count, record base, record:
			 * a chain of dependent loads that ran here with ~2 lanes -- is fetched by the END phase (flat_hit_record), and only by the
			 * ray kinds that look at it (a shadow ray does not) */
			const f3 cpos = L.cpos + L.dir * (L.ctLast + DNB_EPSILON);
			L.pos = L.tile + cpos * 0.125f;
			L.st.hitMapIndex = L.mapIndex;
			L.st.hitLocalIndex = local;
			L.hit = true;
			state = ST_END;
			return;
		}
		const uint32_t rel = (uint32_t)__ldg(L.slot->prefix + L.wordIdx) + __popc(L.word & ((1u << (local & 31u)) - 1u));
		const uint4 rec = __ldg(S.records + (__ldg(&L.slot->voxelBase) + rel));
		L.st.vox = rec;
		DnbMaterial material;
		material.opacity = 1.0f;
		if(FLAT_DEFER_RECORD || !L.chunkOpaque)
			material = load_material(S, rec.x >> 24);
		const uint32_t thisVoxID = (rec.y & 0xFFFFFF00u) | (rec.x >> 24);

		if(material.opacity == 1.0f)
		{
			const f3 cpos = L.cpos + L.dir * (L.ctLast + DNB_EPSILON);
			L.pos = L.tile + cpos * 0.125f;
			L.st.hitMapIndex = L.mapIndex;
			L.st.hitLocalIndex = local;
			L.st.hitRecord = rel;
			L.hit = true;
			state = ST_END;
			return;
		}
		else if(L.st.lastVoxID != thisVoxID)
		{
			if(L.coffp)
				cull_undo(L.coffp, L.cp, L.cbias);
			const float cm = L.colorMult * material.opacity;
			L.colorAdd = L.colorAdd + (vox_albedo(rec) * cm) * ld3(S.sunStrength);
			L.colorMult = L.colorMult * (1.0f - material.opacity);
			L.st.lastVoxID = thisVoxID;
			L.st.lastVoxRefract = material.refractIndex;
		}
	}
	else if(L.st.lastVoxID != 255u)
	{
		L.st.lastVoxID = 255u;
		L.st.lastVoxRefract = 1.0f;
	}

	/* iterate_dda on the voxel level with the shared delta / step */
	{
		const f3 s = L.cside;
		const float myz = fminf(s.y, s.z);
		const bool mx = s.x <= myz;
		const bool my = s.y <= fminf(s.z, s.x);
		const bool mz = s.z <= fminf(s.x, s.y);
		L.ctLast = fminf(s.x, myz);
		if(mx) { L.cside.x = s.x + L.m.delta.x; L.cp.x += L.m.step.x; }
		if(my) { L.cside.y = s.y + L.m.delta.y; L.cp.y += L.m.step.y; }
		if(mz) { L.cside.z = s.z + L.m.delta.z; L.cp.z += L.m.step.z; }
	}
	L.ignoreFirst = false;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* the voxel's ray schedule (LI:207-288): which ray comes next, and what a finished ray adds                          */

DNB_FN void flat_stage(const DnbStagingTargets& T, size_t at, uint32_t w1, uint32_t w2, uint32_t w3)
{
	for(uint32_t p = 0; p < T.count; p++)
	{
		uint32_t* out = T.dst[p] + at;
		out[0] = w1;
		out[32] = w2;
		out[64] = w3;
	}
}

/* LI:266-278: clamp, quantise, store the three staged words; the lane is free for another voxel */
DNB_FN bool flat_finish(const DnbStagingTargets& T, FlatLane& L, uint32_t& state)
{
	const f3 specLight = clamp01(L.spec);
	const f3 diffuseLight = clamp01(L.diff);
	const f3 albedo = vox_albedo(L.rec);
	const uint32_t wx = (uint32_t)rintf(diffuseLight.x * 65535.0f);
	const uint32_t wy = (uint32_t)rintf(diffuseLight.y * 65535.0f);
	const uint32_t wz = (uint32_t)rintf(diffuseLight.z * 65535.0f);
	flat_stage(T, L.at,
	           encode_rgba((uint32_t)rintf(albedo.x * 255.0f), (uint32_t)rintf(albedo.y * 255.0f), (uint32_t)rintf(albedo.z * 255.0f), (uint32_t)rintf(specLight.x * 255.0f)),
	           encode_rgba((uint32_t)rintf(specLight.y * 255.0f), (uint32_t)rintf(specLight.z * 255.0f), (wx >> 8) & 0xFFu, wx & 0xFFu),
	           encode_rgba((wy >> 8) & 0xFFu, wy & 0xFFu, (wz >> 8) & 0xFFu, wz & 0xFFu));
	state = ST_FETCH;
	return false;
}

/* specular ray L.idx of the voxel: LI:244-250 + the prologue of specular_ray (LI:86-90) */
DNB_FN bool flat_start_spec(FlatLane& L, const DnbMaterial& material)
{
	const f3 normal = vox_normal(L.rec);
	const f3 viewDir = L.origin - ld3(c_light.camPos);
	const f3 reflected = reflect3(normalize3(viewDir), normal);
	const f3 specDir = normalize3(reflected * (float)material.shininess + ld3(c_spherePoints[L.idx])) + DNB_EPSILON;
	L.kind = RAY_SPEC;
	L.seg = 0;
	L.pos = L.origin;
	L.pa = L.origin;           /* lastPos */
	L.pb = vox_albedo(L.rec);  /* multiplier */
	L.reflectType = material.reflectType;
	L.dir = specDir;
	return true;
}

/* next segment of diffuse sample L.idx (LI:159-172); false when the bounce limit is exhausted.  For seg > 0 the hit normal
 * and material are those of the record the previous segment hit (L.st.vox), as LI:176-177 leave them. */
DNB_FN bool flat_start_diffuse_segment(const DnbScene& S, FlatLane& L)
{
	if(L.seg >= c_light.diffuseBounceLimit)
		return false;
	f3 hitNormal;
	float hitSpecular = 0.0f;
	uint32_t hitShininess = 0;
	if(L.seg == 0)
		hitNormal = vox_normal(L.rec) + DNB_EPSILON; /* the `normal + EPSILON` argument at LI:259 */
	else
	{
		hitNormal = vox_normal(L.st.vox);
		const DnbMaterial hm = load_material(S, vox_material(L.st.vox));
		hitSpecular = hm.specular;
		hitShininess = hm.shininess;
	}
	f3 dir;
	if(L.seg > 0 && c_light.glossyChoice[L.idx][L.seg] < hitSpecular)
		dir = normalize3(reflect3(L.pb, hitNormal) * (float)hitShininess + ld3(c_light.glossyBall[L.seg]));
	else if(L.firstSample)
		dir = normalize3(hitNormal) + DNB_EPSILON;
	else
		dir = normalize3(hitNormal + ld3(c_light.diffuseBall[L.idx][L.seg])) + DNB_EPSILON;
	L.dir = dir;
	return true;
}

/* LI:65-77 */
DNB_FN bool flat_start_shadow(FlatLane& L)
{
	const f3 sunDir = ld3(c_light.sunDir);
	f3 dir;
	if(L.firstSample)
		dir = sunDir + DNB_EPSILON;
	else
		dir = normalize3(sunDir * c_light.shadowSoftness + ld3(c_light.shadowBall[L.idx])) + DNB_EPSILON;
	L.kind = RAY_SHADOW;
	L.pos = L.origin;
	L.dir = dir;
	return true;
}

/* diffuse sample L.idx: LI:258 + the prologue of diffuse_ray (LI:151-157) */
DNB_FN bool flat_start_sample(const DnbScene& S, FlatLane& L)
{
	L.diff = L.diff + ld3(S.ambient);
	L.st.vox = L.rec;
	L.kind = RAY_DIFFUSE;
	L.seg = 0;
	L.pa = splat3(1.0f); /* newColor */
	L.pb = splat3(0.0f); /* lastDir */
	L.pos = L.origin;
	if(!flat_start_diffuse_segment(S, L))
		flat_start_shadow(L);
	return true;
}

/* LI:254-264 once the specular phase is over */
DNB_FN bool flat_begin_diffuse(const DnbScene& S, const DnbStagingTargets& T, FlatLane& L, const DnbMaterial& material, uint32_t& state)
{
	if(material.specular < 1.0f)
	{
		if(c_light.numDiffuseSamples > 0)
		{
			L.idx = 0;
			return flat_start_sample(S, L);
		}
		L.diff = div3(vox_diffuse(L.rec) * L.indirectSamples + L.diff, L.indirectSamples + (float)c_light.numDiffuseSamples);
	}
	return flat_finish(T, L, state);
}

/* a ray segment has ended (L.hit says how): specular_ray LI:96-145, diffuse_ray LI:174-202, shadow_ray LI:77-79 after their trace */
/* the record of the voxel the segment has hit; hits in all-opaque chunks left it unfetched (flat_vox_step) */
DNB_FN uint4 flat_hit_record(const DnbScene& S, FlatLane& L)
{
	if(FLAT_DEFER_RECORD && L.chunkOpaque)
	{
		const uint32_t local = L.st.hitLocalIndex;
		const uint32_t rel = (uint32_t)__ldg(L.slot->prefix + L.wordIdx) + __popc(L.word & ((1u << (local & 31u)) - 1u));
		L.st.hitRecord = rel;
		L.st.vox = __ldg(S.records + (__ldg(&L.slot->voxelBase) + rel));
	}
	return L.st.vox;
}

DNB_FN bool flat_ray_ended(const DnbScene& S, const DnbStagingTargets& T, FlatLane& L, uint32_t& state)
{
	const f3 sunDir = ld3(c_light.sunDir);
	/* the record and the material of the voxel that was hit, for the ray kinds that look at them (LI:107-109 specular, LI:176-178
	 * diffuse: not when the hit is the neighbouring voxel): fetched HERE, for all lanes of the phase together -- two dependent loads
	 * once, not once per kind of ray */
	uint4 rec = make_uint4(0u, 0u, 0u, 0u);
	DnbMaterial hm = {};
	bool shade = false;
	if(L.hit && L.kind != RAY_SHADOW)
	{
		const f3 from = L.kind == RAY_SPEC ? L.pa : L.origin;
		const f3 dist = abs3(floor3(L.pos * 8.0f) - floor3(from * 8.0f));
		const float d2 = dot3(dist, dist);
		shade = L.kind == RAY_SPEC ? !(d2 <= 1.0f) : !(d2 < 1.0f);
		if(shade)
		{
			rec = flat_hit_record(S, L);
			hm = load_material(S, vox_material(rec));
		}
	}
	if(L.kind == RAY_SPEC)
	{
		if(L.hit)
		{
			const i3 hp = toi3(L.pos);
			if(L.sourceVisible && in_map_bounds(S, hp))
			{
				const uint32_t hitIndex = (uint32_t)hp.x + S.mapSize[0] * ((uint32_t)hp.y + S.mapSize[1] * (uint32_t)hp.z);
				const uint32_t bit = 1u << (hitIndex & 31u);
				if(!(__ldcg(S.propagate + (hitIndex >> 5)) & bit))
					atomicOr(S.propagate + (hitIndex >> 5), bit);
			}
			if(shade)
			{
				const f3 hitAlbedo = vox_albedo(rec);
				const f3 hitDiffuse = vox_diffuse(rec) * (1.0f - hm.specular);
				if(hm.emissive)
					L.spec = L.spec + ((hitAlbedo * L.colorMult + L.colorAdd) * L.pb) * vox_albedo(L.rec);
				else
				{
					const f3 hitColor = hitDiffuse * hitAlbedo;
					L.spec = L.spec + (hitColor * L.colorMult + L.colorAdd) * L.pb;
					if(hm.specular != 0.0f)
					{
						L.pb = L.pb * ((hitAlbedo * L.colorMult) * hm.specular);
						L.reflectType = hm.reflectType;
						L.pa = L.pos;
						const f3 dir = reflect3(L.dir, vox_normal(rec));
						if(++L.seg < c_light.specularBounceLimit)
						{
							L.dir = dir;
							return true;
						}
					}
				}
			}
		}
		else if(dot3(L.dir, sunDir) > 0.99f)
			L.spec = L.spec + (ld3(S.sunStrength) * L.colorMult + L.colorAdd);
		else
		{
			const f3 base = (L.reflectType == 1u) ? sky_color(S, L.dir) : ld3(S.sunStrength);
			L.spec = L.spec + (base * L.colorMult + L.colorAdd) * L.pb;
		}

		/* this specular ray is over */
		const DnbMaterial material = load_material(S, vox_material(L.rec));
		if(++L.idx < 15u)
			return flat_start_spec(L, material);
		L.spec = div3(L.spec, 15.0f);
		return flat_begin_diffuse(S, T, L, material, state);
	}
	else if(L.kind == RAY_DIFFUSE)
	{
		if(L.hit)
		{
			if(shade)
			{
				const f3 through = vox_albedo(rec) * L.colorMult + L.colorAdd;
				if(hm.emissive)
					L.diff = L.diff + L.pa * through;
				else
				{
					L.pa = L.pa * through;
					L.pb = L.dir; /* lastDir */
					L.seg++;
					if(flat_start_diffuse_segment(S, L))
						return true;
				}
			}
		}
		else
		{
			const float ndl = fmaxf(dot3(L.dir, sunDir), 0.0f);
			L.diff = L.diff + (((L.pa * ndl) * ld3(S.sunStrength)) * L.colorMult + L.colorAdd);
		}
		return flat_start_shadow(L);
	}
	else
	{
		if(!L.hit)
			L.diff = L.diff + (ld3(S.sunStrength) * L.colorMult + L.colorAdd);
		if(++L.idx < c_light.numDiffuseSamples)
			return flat_start_sample(S, L);
		L.diff = div3(vox_diffuse(L.rec) * L.indirectSamples + L.diff, L.indirectSamples + (float)c_light.numDiffuseSamples);
		return flat_finish(T, L, state);
	}
}

/* n-th set bit of the chunk's surface mask, read from the slot in global memory (cf. nth_voxel in light.cu) */
DNB_FN int flat_nth_voxel(const DnbSlot* slot, uint32_t voxNum)
{
	if(voxNum >= __ldg(&slot->numVoxels))
		return -1;
	const uint4* pp = reinterpret_cast<const uint4*>(slot->prefix);
	const uint4 a = __ldg(pp), b = __ldg(pp + 1);
	const uint32_t pw[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
	int w = 0;
	uint32_t before = 0;
#pragma unroll
	for(int i = 1; i < 16; i++)
	{
		const uint32_t p = (pw[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
		if(p <= voxNum)
		{
			w = i;
			before = p;
		}
	}
	return w * 32 + (int)__fns(__ldg(slot->mask + w), 0, (int)(voxNum - before) + 1);
}

/* work item j of this launch -> (request, lane); sets the voxel up and starts its first ray (LI:207-251).
 * Leaves the lane in ST_FETCH when the item holds no voxel (tail of a chunk's last group, removed chunk, end of list). */
DNB_FN bool flat_setup_voxel(const DnbScene& S, const DnbStagingTargets& T, const uint32_t* __restrict__ requests, uint32_t numRequests, uint32_t firstCta, uint32_t ctaStride,
                             uint32_t j, FlatLane& L, uint32_t& state)
{
	const uint32_t r = (firstCta + (j >> 7) * ctaStride) * 4u + ((j >> 5) & 3u);
	if(r >= numRequests)
		return false;
	const uint32_t request = __ldg(requests + r);
	const uint32_t mapIndex = request >> 4;
	L.at = (size_t)r * 96u + (j & 31u);

	const uint32_t slotId = __ldg(S.tileSlot + mapIndex) - 1u;
	const DnbSlot* slot = S.slots + slotId;
	const uint32_t voxNum = (j & 31u) + (request & 15u) * 32u;
	const int local = slotId == 0xFFFFFFFFu ? -1 : flat_nth_voxel(slot, voxNum);
	if(local < 0)
	{
		flat_stage(T, L.at, 0, 0, 0);
		return false;
	}

	ray_state_reset(L.st);
	L.sourceVisible = (__ldg(S.visible + (mapIndex >> 5)) >> (mapIndex & 31u)) & 1u;
	const i3 chunkPos = {local & 7, (local >> 3) & 7, local >> 6};
	const i3 mapPos = {__ldg(&slot->pos[0]), __ldg(&slot->pos[1]), __ldg(&slot->pos[2])};
	L.rec = __ldg(S.records + (__ldg(&slot->voxelBase) + voxNum));
	const f3 normal = vox_normal(L.rec);
	const DnbMaterial material = load_material(S, vox_material(L.rec));

	const uint32_t ns = __ldg(&slot->numSamples);
	L.indirectSamples = (float)(ns < c_light.maxDiffuseSamples ? ns : c_light.maxDiffuseSamples);
	L.firstSample = L.indirectSamples == 0.0f;

	f3 rayPos = (tof3(chunkPos) * 0.125f + tof3(mapPos)) + 0.0625f;
	L.origin = rayPos + normal * (0.0625f - DNB_EPSILON);
	L.spec = splat3(0.0f);
	L.diff = splat3(0.0f);
	L.idx = 0;

	const f3 viewDir = L.origin - ld3(c_light.camPos);
	if(material.specular > 0.0f && dot3(viewDir, normal) < 0.0f && material.reflectType <= 1u && c_light.specularBounceLimit > 0u)
		return flat_start_spec(L, material);
	return flat_begin_diffuse(S, T, L, material, state); /* (a zero bounce limit leaves specLight = 0 / 15 = 0) */
}

#define FLAT_WARPS 4

/* 5 CTAs per SM = 96 registers with ~100 bytes of spills: 8 % faster than 119 registers / 4 CTAs on B200 (sparse map 90 -> 82.5 ms) */
#ifndef FLAT_MIN_BLOCKS
#define FLAT_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(FLAT_WARPS * 32, FLAT_MIN_BLOCKS) dn_light_flat_kernel(DnbScene S, const uint32_t* __restrict__ requests, DnbWork W,
                                                                        uint32_t* __restrict__ workCounter, DnbStagingTargets T, DnbFlatTuning K)
{
	/* the request count is read on the device (layout.h DnbWork); a work item = one lane of a request */
	const uint32_t numRequests = work_requests(W);
	const uint32_t firstCta = W.firstCta, ctaStride = W.ctaStride;
	const uint32_t totalItems = work_ctas(W, numRequests) * 128u;
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t ltMask = (1u << lane) - 1u;
	FlatLane L;
	uint32_t state = ST_FETCH;
	L.hit = false;
	L.kind = RAY_SHADOW;
	int waited = 0;

	for(;;)
	{
		const uint32_t mT = __ballot_sync(0xFFFFFFFFu, state == ST_TILE);
		const uint32_t mV = __ballot_sync(0xFFFFFFFFu, state == ST_VOX);
		const uint32_t mE = __ballot_sync(0xFFFFFFFFu, state == ST_END || state == ST_FETCH);
		const uint32_t mN = __ballot_sync(0xFFFFFFFFu, state == ST_ENTER);
		const int nT = __popc(mT), nV = __popc(mV), nE = __popc(mE), nN = __popc(mN);
		if(nT + nV + nE + nN == 0)
			break;

		if(nE >= K.endLanes || nT + nV + nN == 0 || (nE > 0 && (waited >= K.patience || (K.endMax && nE >= nT && nE >= nV && nE >= nN))))
		{
			/* shade finished rays and pick the next ones; hand free lanes new voxels; then start all new rays together */
			waited = 0;
			bool start = false;
			if(state == ST_END)
				start = flat_ray_ended(S, T, L, state);
			/* lanes without a voxel take the next work items.  Two rounds, because an item can turn out to hold no voxel (tail of a
			 * chunk's last group).  `need` -- not `state` -- says who still fetches: a lane that has just been given a voxel stays in
			 * ST_FETCH until flat_start_ray below, and must NOT take (and thereby drop) a second item. */
			bool need = state == ST_FETCH;
#pragma unroll 1
			for(int round = 0; round < 2; round++)
			{
				const uint32_t mF = __ballot_sync(0xFFFFFFFFu, need);
				if(mF == 0u)
					break;
				uint32_t base = 0;
				const int leader = __ffs(mF) - 1;
				if((int)lane == leader)
					base = atomicAdd(workCounter, (uint32_t)__popc(mF));
				base = __shfl_sync(0xFFFFFFFFu, base, leader);
				if(need)
				{
					const uint32_t j = base + (uint32_t)__popc(mF & ltMask);
					if(j < totalItems)
					{
						start = flat_setup_voxel(S, T, requests, numRequests, firstCta, ctaStride, j, L, state);
						need = !start;
					}
					else
					{
						state = ST_DONE;
						need = false;
					}
				}
			}
			if(start)
				flat_start_ray(L, state);
		}
		else if(nN > 0 && nN >= nT && nN >= nV)
		{
			if(state == ST_ENTER)
				flat_enter_chunk(S, L, state);
			waited++;
		}
		else if(nT >= nV)
		{
			const int keep = (K.keep * nT + 7) >> 3;
#pragma unroll 1
			for(int it = 0; it < K.budget; it++)
			{
				if(state == ST_TILE)
					flat_tile_step(S, L, state);
				waited++;
#if FLAT_UNROLL2
				if(state == ST_TILE)
					flat_tile_step(S, L, state);
				waited++;
#endif
				if(K.run ? it + 1 >= K.run : __popc(__ballot_sync(0xFFFFFFFFu, state == ST_TILE)) < keep)
					break;
			}
		}
		else
		{
			const int keep = (K.keep * nV + 7) >> 3;
#pragma unroll 1
			for(int it = 0; it < K.budget; it++)
			{
				if(state == ST_VOX)
					flat_vox_step(S, L, state);
				waited++;
#if FLAT_UNROLL2
				if(state == ST_VOX)
					flat_vox_step(S, L, state);
				waited++;
#endif
				if(K.run ? it + 1 >= K.run : __popc(__ballot_sync(0xFFFFFFFFu, state == ST_VOX)) < keep)
					break;
			}
		}
	}
}
