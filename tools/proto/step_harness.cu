/* step_harness.cu -- runs ray segments on the CPU through (a) trace.cuh trace_ray<false, false>, the traversal every lighting kernel of
 * the library uses, and (b) the unified stepping prototype (unified_step.cuh), on a scene assembled in host memory, and returns both
 * results for a bit-for-bit comparison (tests/test_unified_step_proto.py).  Built with nvcc for the HOST only; nothing runs on a GPU. */
#define DNB_FN __host__ __device__ inline
#include "unified_step.cuh"

struct RayIn
{
	float    dir[3], pos[3];
	uint32_t ignoreFirst, lastVoxID;
	float    lastVoxRefract;
	uint32_t pad;
};

struct RayOut
{
	uint32_t hit, tripped, lastVoxID, hitMapIndex, hitLocalIndex, hitRecord;
	float    lastVoxRefract, colorMult;
	float    pos[3], colorAdd[3];
	uint32_t vox[4];
	uint32_t steps; /* unified: cheap-step calls; reference: 0 */
	uint32_t events;
};

static void fill(RayOut& o, bool hit, const RayState& st, f3 pos, f3 colorAdd, float colorMult)
{
	memset(&o, 0, sizeof(o));
	o.hit = hit;
	o.tripped = st.tripped;
	o.lastVoxID = st.lastVoxID;
	o.lastVoxRefract = st.lastVoxRefract;
	o.hitMapIndex = hit ? st.hitMapIndex : 0;
	o.hitLocalIndex = hit ? st.hitLocalIndex : 0;
	o.hitRecord = hit ? st.hitRecord : 0;
	o.colorMult = colorMult;
	o.pos[0] = pos.x; o.pos[1] = pos.y; o.pos[2] = pos.z;
	o.colorAdd[0] = colorAdd.x; o.colorAdd[1] = colorAdd.y; o.colorAdd[2] = colorAdd.z;
	o.vox[0] = st.vox.x; o.vox[1] = st.vox.y; o.vox[2] = st.vox.z; o.vox[3] = st.vox.w;
}

extern "C" int harness_run(const DnbScene* scene, const RayIn* rays, uint32_t count, RayOut* ref, RayOut* uni)
{
	const DnbScene S = *scene;
	for(uint32_t i = 0; i < count; i++)
	{
		const RayIn& r = rays[i];
		const f3 dir = mk3(r.dir[0], r.dir[1], r.dir[2]);
		const f3 origin = mk3(r.pos[0], r.pos[1], r.pos[2]);
		{
			RayState st;
			ray_state_reset(st);
			st.lastVoxID = r.lastVoxID;
			st.lastVoxRefract = r.lastVoxRefract;
			DnbCounters lc;
			memset(&lc, 0, sizeof(lc));
			f3 d = dir, p = origin, n = splat3(0.0f), colorAdd;
			float colorMult;
			const bool hit = trace_ray<false, false>(S, st, lc, d, rcp3(d), p, r.ignoreFirst != 0, n, colorAdd, colorMult);
			fill(ref[i], hit, st, p, colorAdd, colorMult);
		}
		{
			UniLane L;
			ray_state_reset(L.st);
			L.st.lastVoxID = r.lastVoxID;
			L.st.lastVoxRefract = r.lastVoxRefract;
			L.dir = dir;
			L.inv = rcp3(dir);
			L.rayPos = origin;
			L.ignoreFirst = r.ignoreFirst != 0;
			uint32_t state, steps = 0, events = 0;
			uni_start(L, state);
			while(state != U_END)
			{
				switch(state)
				{
				case U_STEP: uni_step(S, L, state); steps++; break;
				case U_BOUNDARY: uni_boundary(S, L, state); events++; break;
				case U_ENTER: uni_enter(S, L, state); events++; break;
				default: uni_hit(S, L, state); events++; break;
				}
				if(steps + events > 50000000u)
					return -1;
			}
			fill(uni[i], L.hit, L.st, L.rayPos, L.colorAdd, L.colorMult);
			uni[i].steps = steps;
			uni[i].events = events;
		}
	}
	return 0;
}

/* the unified stepper's per-ray sequence of operations, for the warp-scheduling model (tools/proto/schedule_model.py):
 * 0 cheap step at the tile level, 1 cheap step at the voxel level, 2 one iteration of the empty-block fast loop, 3 tile-level block
 * change, 4 next z-layer of a chunk, 5 chunk exit, 6 chunk entry, 7 opaque hit, 8 transparent voxel; 255-terminated, `cap` bytes per ray */
extern "C" int harness_trace(const DnbScene* scene, const RayIn* rays, uint32_t count, uint8_t* ops, uint32_t cap)
{
	const DnbScene S = *scene;
	for(uint32_t i = 0; i < count; i++)
	{
		const RayIn& r = rays[i];
		uint8_t* out = ops + (size_t)i * cap;
		uint32_t n = 0;
		auto put = [&](uint8_t c) { if(n + 1 < cap) out[n++] = c; };
		UniLane L;
		ray_state_reset(L.st);
		L.st.lastVoxID = r.lastVoxID;
		L.st.lastVoxRefract = r.lastVoxRefract;
		L.dir = mk3(r.dir[0], r.dir[1], r.dir[2]);
		L.inv = rcp3(L.dir);
		L.rayPos = mk3(r.pos[0], r.pos[1], r.pos[2]);
		L.ignoreFirst = r.ignoreFirst != 0;
		uint32_t state, guard = 0;
		uni_start(L, state);
		while(state != U_END && ++guard < 1000000u)
		{
			const uint32_t lv = L.lv;
			if(state == U_STEP)
			{
				const uint32_t g0 = L.g;
				uni_step(S, L, state);
				if(state == U_STEP || state == U_END)
				{
					/* a completed iteration (or a guard trip); the fast loop shows as g advancing by more than one */
					put(lv ? 1 : 0);
					for(uint32_t k = g0 + 2; k <= L.g && lv == 0; k++)
						put(2);
				}
			}
			else if(state == U_BOUNDARY)
			{
				const bool leaves = lv != 0 && !in_chunk_bounds(L.pos);
				uni_boundary(S, L, state);
				put(lv == 0 ? 3 : (leaves ? 5 : 4));
			}
			else if(state == U_ENTER)
			{
				uni_enter(S, L, state);
				put(6);
			}
			else
			{
				uni_hit(S, L, state);
				put(state == U_END ? 7 : 8);
			}
		}
		out[n] = 255;
	}
	return 0;
}

extern "C" size_t harness_sizes(int which) { return which == 0 ? sizeof(DnbScene) : which == 1 ? sizeof(RayIn) : sizeof(RayOut); }
