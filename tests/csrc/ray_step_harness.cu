/* ray_step_harness.cu -- runs ray segments on the CPU through (a) trace.cuh trace_ray<false, false>, the traversal of the warp-per-request
 * and persistent lighting kernels (step_map + step_chunk of voxelShared.comp:328-475), and (b) / (c) the lock-step iteration of
 * ray_step.cuh that the wavefront step kernel runs -- (b) fetching records itself, (c) with deferred hits resolved the way the serve
 * kernel resolves them -- on a scene assembled in host memory, and returns all three for a bit-for-bit comparison
 * (tests/test_ray_step.py).  Built with nvcc for the HOST only; nothing runs on a GPU. */
#define DNB_FN __host__ __device__ inline
#include "ray_step.cuh"

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
	uint32_t iterations; /* lock-step: ray_iter calls; reference: 0 */
	uint32_t deferred;
};

static void fill(RayOut& o, bool hit, bool tripped, uint32_t lastVoxID, float lastVoxRefract, uint32_t hitMapIndex, uint32_t hitLocal, uint32_t hitRecord, uint4 vox, f3 pos, f3 colorAdd, float colorMult)
{
	memset(&o, 0, sizeof(o));
	o.hit = hit;
	o.tripped = tripped;
	o.lastVoxID = lastVoxID;
	o.lastVoxRefract = lastVoxRefract;
	o.hitMapIndex = hit ? hitMapIndex : 0;
	o.hitLocalIndex = hit ? hitLocal : 0;
	o.hitRecord = hit ? hitRecord : 0;
	o.colorMult = colorMult;
	o.pos[0] = pos.x; o.pos[1] = pos.y; o.pos[2] = pos.z;
	o.colorAdd[0] = colorAdd.x; o.colorAdd[1] = colorAdd.y; o.colorAdd[2] = colorAdd.z;
	o.vox[0] = vox.x; o.vox[1] = vox.y; o.vox[2] = vox.z; o.vox[3] = vox.w;
}

template <bool DEFER> static int run_lock_step(const DnbScene& S, const RayIn& r, RayOut& out)
{
	RayLane L;
	memset(&L, 0, sizeof(L));
	L.dir = mk3(r.dir[0], r.dir[1], r.dir[2]);
	L.rayPos = mk3(r.pos[0], r.pos[1], r.pos[2]);
	L.lastVoxID = r.lastVoxID;
	L.lastVoxRefract = r.lastVoxRefract;
	L.tripped = false;
	L.ignoreFirst = r.ignoreFirst != 0;
	L.vox = make_uint4(0, 0, 0, 0);
	L.hitLocal = L.hitRecord = 0;
	ray_begin(L, rcp3(L.dir));
	uint32_t n = 0;
	while(!ray_iter<DEFER>(S, L))
		if(++n > 50000000u)
			return -1;
	uint32_t deferred = 0;
	if(L.hit && L.deferred)
	{
		deferred = 1;
		L.vox = ray_deferred_record(S, L.vox, &L.hitRecord); /* what dn_wave_serve_kernel does */
	}
	fill(out, L.hit, L.tripped, L.lastVoxID, L.lastVoxRefract, L.mapIndex, L.hitLocal, L.hitRecord, L.vox, L.rayPos, L.colorAdd, L.colorMult);
	out.iterations = n + 1;
	out.deferred = deferred;
	return 0;
}

extern "C" int harness_run(const DnbScene* scene, const RayIn* rays, uint32_t count, RayOut* ref, RayOut* plain, RayOut* deferred)
{
	const DnbScene S = *scene;
	for(uint32_t i = 0; i < count; i++)
	{
		const RayIn& r = rays[i];
		{
			RayState st;
			ray_state_reset(st);
			st.lastVoxID = r.lastVoxID;
			st.lastVoxRefract = r.lastVoxRefract;
			DnbCounters lc;
			memset(&lc, 0, sizeof(lc));
			f3 d = mk3(r.dir[0], r.dir[1], r.dir[2]), p = mk3(r.pos[0], r.pos[1], r.pos[2]), n = splat3(0.0f), colorAdd;
			float colorMult;
			const bool hit = trace_ray<false, false>(S, st, lc, d, rcp3(d), p, r.ignoreFirst != 0, n, colorAdd, colorMult);
			fill(ref[i], hit, st.tripped, st.lastVoxID, st.lastVoxRefract, st.hitMapIndex, st.hitLocalIndex, st.hitRecord, st.vox, p, colorAdd, colorMult);
		}
		if(run_lock_step<false>(S, r, plain[i]) || run_lock_step<true>(S, r, deferred[i]))
			return -1;
	}
	return 0;
}

extern "C" size_t harness_sizes(int which) { return which == 0 ? sizeof(DnbScene) : which == 1 ? sizeof(RayIn) : sizeof(RayOut); }
