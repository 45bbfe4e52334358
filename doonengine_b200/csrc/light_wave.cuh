/* light_wave.cuh (included at the end of light.cu, after light_flat.cuh whose per-voxel functions it reuses) -- the lighting
 * update as a WAVEFRONT: the voxel contexts live in global memory and two kernels alternate over a pool of P context slots.
 *
 *   dn_wave_serve_kernel   one thread per slot, every lane busy: shade the ray segment that just ended (flat_ray_ended: LI:96-145,
 *                          174-202, 77-79), advance the voxel's ray schedule, store a finished voxel's three staged words and
 *                          immediately set the slot up with the next voxel of the dispatch (flat_setup_voxel, LI:207-251), then
 *                          prepare the next ray segment completely (direction, reciprocal, tile-level DDA start) and store it
 *   dn_wave_step_kernel    persistent warps; every lane traces one prepared ray segment with the two-state machine TILE / VOX
 *                          (flat_tile_step / flat_vox_step = trace_ray<false,false>, trace.cuh) and, when its ray ends, stores
 *                          the result and takes the next slot of its warp's range -- a lane never waits for shading
 *
 * Why: in dn_light_flat_kernel (one context per lane, in registers) ncu attributes ~40 % of all issue slots to ray set-up and
 * shading executed with 2-4 active lanes, and the stepping phases run with 7-9 lanes because half the warp is waiting for that
 * service (profiles/r1_v3_wave.md).  Moving the contexts to memory costs ~0.3 KB of traffic per ray segment and buys full warps
 * for shading and twice the lanes for stepping.  Nothing about a voxel's own sequence of operations changes -- the same
 * functions run in the same order on the same values -- so the staged words are bit-identical to both other kernels
 * (tests/test_parity_gpu.py runs every lighting test against all three).
 *
 * Slot layout: structure of arrays, 15 planes of P uint4 (240 bytes per slot), so that both kernels move whole 16-byte words
 * and the serve kernel's accesses are fully coalesced.
 */
#include "ray_step.cuh"

enum : uint32_t
{
	WV_REC = 0,   /* the voxel's record */
	WV_ORG = 1,   /* origin.xyz, indirectSamples */
	WV_SPEC = 2,  /* spec.xyz, work item index */
	WV_DIFF = 3,  /* diff.xyz, schedule word (WS_*) */
	WV_PA = 4,    /* pa.xyz */
	WV_PB = 5,    /* pb.xyz */
	WR_DIR = 6,   /* ray: dir.xyz, lastVoxID */
	WR_POS = 7,   /* ray: origin of the segment, lastVoxRefract */
	WR_INV = 8,   /* ray: 1/dir, flags (WF_READY | WF_TRIPPED) */
	WR_SIDE = 9,  /* ray: tile-level sideDist */
	WR_CELL = 10, /* ray: tile-level cell (int) */
	WH_POS = 11,  /* result: hit position (or the origin), flags (WF_HIT | WF_TRIPPED | WF_INSIDE) */
	WH_COL = 12,  /* result: colorAdd.xyz, colorMult */
	WH_VOX = 13,  /* result: record hit (written / read only on a hit); with WF_DEFERRED: {slot, local index, mask word} -- the serve kernel fetches the record (ray_step.cuh) */
	WH_ST = 14,   /* result: lastVoxID, lastVoxRefract (written / read only with WF_INSIDE) */
	WAVE_PLANES = 15
};
enum : uint32_t { WF_READY = 1u, WF_TRIPPED = 2u, WF_HIT = 1u, WF_INSIDE = 4u, WF_DEFERRED = 8u };
/* schedule word: kind[0:2) idx[2:8) seg[8:16) firstSample[16] sourceVisible[17] reflectType[18:26) active[31] */
#define WS_ACTIVE 0x80000000u

DNB_FN uint4 f3w(f3 a, uint32_t w) { return make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), w); }
DNB_FN f3 xyz_of(uint4 v) { return mk3(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z)); }

#define WAVE_SERVE_THREADS 256
__global__ void __launch_bounds__(WAVE_SERVE_THREADS) dn_wave_serve_kernel(DnbScene S, const uint32_t* __restrict__ requests, uint32_t numRequests, uint32_t firstCta, uint32_t ctaStride, uint32_t totalItems,
                                                            uint32_t* __restrict__ workCounter, DnbStagingTargets T, uint4* __restrict__ ctx, uint32_t P, uint32_t* __restrict__ activeNow,
                                                            uint32_t* __restrict__ activeNext)
{
	__shared__ uint32_t s_warpCount[WAVE_SERVE_THREADS / 32], s_base;
	const uint32_t i = blockIdx.x * WAVE_SERVE_THREADS + threadIdx.x; /* the grid covers the pool exactly (P is a multiple of 256) */
	const uint32_t lane = threadIdx.x & 31u;
	const uint32_t ltMask = (1u << lane) - 1u;
	if(i == 0)
	{
		*activeNext = 0;     /* the counter the NEXT pass adds to (stream order keeps the passes apart) */
		workCounter[3] = 0;  /* slot cursor of the step kernel that follows this pass */
	}
#define PL(p) ctx[(size_t)(p) * P + i]

	FlatLane L;
	uint32_t state = ST_FETCH, item = 0;
	bool start = false;

	const uint4 vDiff = PL(WV_DIFF);
	if(vDiff.w & WS_ACTIVE)
	{
		const uint32_t sched = vDiff.w;
		const uint4 vRec = PL(WV_REC), vOrg = PL(WV_ORG), vSpec = PL(WV_SPEC), vPa = PL(WV_PA), vPb = PL(WV_PB), rDir = PL(WR_DIR), hPos = PL(WH_POS), hCol = PL(WH_COL);
		L.rec = vRec;
		L.origin = xyz_of(vOrg);
		L.indirectSamples = __uint_as_float(vOrg.w);
		L.spec = xyz_of(vSpec);
		item = vSpec.w;
		L.diff = xyz_of(vDiff);
		L.kind = sched & 3u;
		L.idx = (sched >> 2) & 63u;
		L.seg = (sched >> 8) & 255u;
		L.firstSample = (sched >> 16) & 1u;
		L.sourceVisible = (sched >> 17) & 1u;
		L.reflectType = (sched >> 18) & 255u;
		L.pa = xyz_of(vPa);
		L.pb = xyz_of(vPb);
		L.dir = xyz_of(rDir);
		L.pos = xyz_of(hPos);
		L.hit = hPos.w & WF_HIT;
		L.colorAdd = xyz_of(hCol);
		L.colorMult = __uint_as_float(hCol.w);
		L.st.tripped = (hPos.w & WF_TRIPPED) != 0u;
		L.st.lastVoxID = 255u;
		L.st.lastVoxRefract = 1.0f;
		L.st.vox = make_uint4(0, 0, 0, 0);
		L.st.hitMapIndex = L.st.hitLocalIndex = L.st.hitRecord = 0;
		if(L.hit)
		{
			L.st.vox = PL(WH_VOX);
			if(hPos.w & WF_DEFERRED)
				L.st.vox = ray_deferred_record(S, L.st.vox, nullptr); /* the step kernel ended the ray on the voxel's bit alone */
		}
		if(hPos.w & WF_INSIDE)
		{
			const uint4 hSt = PL(WH_ST);
			L.st.lastVoxID = hSt.x;
			L.st.lastVoxRefract = __uint_as_float(hSt.y);
		}
		const uint32_t r = (firstCta + (item >> 7) * ctaStride) * 4u + ((item >> 5) & 3u);
		L.at = (size_t)r * 96u + (item & 31u);
		state = ST_END;
		start = flat_ray_ended(S, T, L, state); /* false + ST_FETCH: the voxel is finished and its words are staged */
	}

	/* free slots take the next voxels of the dispatch (two rounds: an item can turn out to hold no voxel).  ONE atomic per CTA
	 * and round: with one per warp the 10^5 same-address atomics of a pass were a quarter of the kernel's stall samples.
	 * `need` -- not `state` -- says who still fetches: a slot that has just been given a voxel keeps state ST_FETCH (nothing sets it
	 * before the store below) and must NOT take, and thereby drop, a second item. */
	bool need = state == ST_FETCH;
#pragma unroll 1
	for(int round = 0; round < 2; round++)
	{
		const uint32_t mF = __ballot_sync(0xFFFFFFFFu, need);
		if(lane == 0)
			s_warpCount[threadIdx.x >> 5] = (uint32_t)__popc(mF);
		__syncthreads();
		uint32_t before = 0, total = 0;
#pragma unroll
		for(uint32_t w = 0; w < WAVE_SERVE_THREADS / 32; w++)
		{
			const uint32_t c = s_warpCount[w];
			before += w < (threadIdx.x >> 5) ? c : 0u;
			total += c;
		}
		if(total == 0u)
			break; /* uniform over the CTA */
		if(threadIdx.x == 0)
		{
			uint32_t base = *reinterpret_cast<volatile uint32_t*>(workCounter);
			if(base < totalItems) /* once the list is exhausted the counter stops moving (it would wrap after 2^32 idle passes otherwise) */
				base = atomicAdd(workCounter, total);
			s_base = base;
		}
		__syncthreads();
		if(need)
		{
			const uint32_t j = s_base + before + (uint32_t)__popc(mF & ltMask);
			if(j < totalItems)
			{
				item = j;
				start = flat_setup_voxel(S, T, requests, numRequests, firstCta, ctaStride, j, L, state);
				need = !start;
			}
			else
			{
				state = ST_DONE;
				need = false;
			}
		}
		__syncthreads(); /* s_warpCount / s_base are rewritten by the next round */
	}

	if(start)
	{
		/* flat_start_ray, with its result stored instead of kept */
		const f3 inv = rcp3(L.dir);
		Dda m;
		init_dda(L.dir, inv, L.pos, m);
		PL(WV_REC) = L.rec;
		PL(WV_ORG) = f3w(L.origin, __float_as_uint(L.indirectSamples));
		PL(WV_SPEC) = f3w(L.spec, item);
		PL(WV_DIFF) = f3w(L.diff, WS_ACTIVE | (L.kind & 3u) | ((L.idx & 63u) << 2) | ((L.seg & 255u) << 8) | (L.firstSample ? 1u << 16 : 0u) | (L.sourceVisible ? 1u << 17 : 0u) | ((L.reflectType & 255u) << 18));
		PL(WV_PA) = f3w(L.pa, 0);
		PL(WV_PB) = f3w(L.pb, 0);
		PL(WR_DIR) = f3w(L.dir, L.st.lastVoxID);
		PL(WR_POS) = f3w(L.pos, __float_as_uint(L.st.lastVoxRefract));
		PL(WR_INV) = f3w(inv, WF_READY | (L.st.tripped ? WF_TRIPPED : 0u));
		PL(WR_SIDE) = f3w(m.side, 0);
		PL(WR_CELL) = make_uint4((uint32_t)m.pos.x, (uint32_t)m.pos.y, (uint32_t)m.pos.z, 0);
	}
	else if(vDiff.w & WS_ACTIVE)
	{
		/* the slot goes idle (an idle slot stays as it is: both words are already zero) */
		PL(WV_DIFF) = make_uint4(0, 0, 0, 0);
		PL(WR_INV) = make_uint4(0, 0, 0, 0);
	}
#undef PL

	const uint32_t mA = __ballot_sync(0xFFFFFFFFu, start);
	if(lane == 0 && mA)
		atomicAdd(activeNow, (uint32_t)__popc(mA));
}

#ifndef WAVE_MIN_BLOCKS
#define WAVE_MIN_BLOCKS 6
#endif
#define WAVE_GRAB 32u /* slots a warp takes from the pass's cursor at a time: one per lane */

/* the five ray planes of 32 consecutive slots -> one of the warp's two shared-memory buffers (cp.async, 16 bytes per lane and plane:
 * each plane row is one fully coalesced 512-byte read); slots past the end of the pool read slot P-1 again and are never used */
DNB_FN void wave_prefetch(uint4* __restrict__ buf, const uint4* __restrict__ ctx, uint32_t P, uint32_t base, uint32_t lane)
{
	const uint32_t idx = base + lane < P ? base + lane : P - 1u;
#pragma unroll
	for(uint32_t p = 0; p < 5u; p++)
		__pipeline_memcpy_async(buf + p * 32u + lane, ctx + (size_t)(WR_DIR + p) * P + idx, sizeof(uint4));
	__pipeline_commit();
}

__global__ void __launch_bounds__(128, WAVE_MIN_BLOCKS) dn_wave_step_kernel(DnbScene S, uint4* __restrict__ ctx, uint32_t P, uint32_t* __restrict__ cursor, DnbFlatTuning K)
{
	/* per warp: two buffers of 32 slots x 5 ray planes (2 x 2.5 KB); while the lanes work through one range of slots the next one is
	 * already on its way, so handing an idle lane its next ray costs five shared-memory reads instead of a round trip to HBM */
	__shared__ uint4 s_rays[4][2][5 * 32];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t ltMask = (1u << lane) - 1u;

	/* current range [next, end) lives in buffer `cur`; `preBase` (if < P) is the range in flight into the other buffer */
	uint32_t cur = 0, next = 0, end = 0, preBase = 0xFFFFFFFFu;
	{
		uint32_t base = 0;
		if(lane == 0)
			base = atomicAdd(cursor, WAVE_GRAB);
		preBase = __shfl_sync(0xFFFFFFFFu, base, 0);
		if(preBase >= P)
			return;
		wave_prefetch(s_rays[warp][1], ctx, P, preBase, lane);
	}

	FlatLane L;
	uint32_t state = ST_FETCH, slot = 0;
	L.hit = false;

	for(;;)
	{
		const uint32_t mT = __ballot_sync(0xFFFFFFFFu, state == ST_TILE);
		const uint32_t mV = __ballot_sync(0xFFFFFFFFu, state == ST_VOX);
		const int nT = __popc(mT), nV = __popc(mV), nF = 32 - nT - nV;

		if(nF > 0)
		{
			if(next >= end && preBase < P)
			{
				/* switch to the prefetched range and start fetching the one after it */
				__pipeline_wait_prior(0);
				__syncwarp();
				cur ^= 1u;
				next = preBase;
				end = preBase + WAVE_GRAB < P ? preBase + WAVE_GRAB : P;
				uint32_t base = 0;
				if(lane == 0)
					base = atomicAdd(cursor, WAVE_GRAB);
				preBase = __shfl_sync(0xFFFFFFFFu, base, 0);
				if(preBase < P)
					wave_prefetch(s_rays[warp][cur ^ 1u], ctx, P, preBase, lane);
			}
			if(next < end)
			{
				/* idle lanes take the next slots of the range */
				const uint32_t mF = ~(mT | mV);
				if(state == ST_FETCH)
				{
					const uint32_t idx = next + (uint32_t)__popc(mF & ltMask);
					if(idx < end)
					{
						const uint4* row = s_rays[warp][cur] + (idx & 31u);
						const uint4 rInv = row[2 * 32];
						if(rInv.w & WF_READY)
						{
							const uint4 rDir = row[0], rPos = row[1 * 32], rSide = row[3 * 32], rCell = row[4 * 32];
							slot = idx;
							L.dir = xyz_of(rDir);
							L.inv = xyz_of(rInv);
							L.pos = xyz_of(rPos);
							L.st.lastVoxID = rDir.w;
							L.st.lastVoxRefract = __uint_as_float(rPos.w);
							L.st.tripped = (rInv.w & WF_TRIPPED) != 0u;
							L.st.vox = make_uint4(0, 0, 0, 0);
							L.m.pos.x = (int)rCell.x; L.m.pos.y = (int)rCell.y; L.m.pos.z = (int)rCell.z;
							L.m.side = xyz_of(rSide);
							L.m.delta = abs3(L.inv);
							L.m.step.x = isgn(L.dir.x); L.m.step.y = isgn(L.dir.y); L.m.step.z = isgn(L.dir.z);
							L.colorAdd = splat3(0.0f);
							L.colorMult = 1.0f;
							L.tLast = 0.0f;
							L.ignoreFirst = true;
							L.guard = 0;
							L.blk.x = L.blk.y = L.blk.z = 0x40000000;
							L.occWord = 0;
							L.hit = false;
							state = ST_TILE;
						}
					}
				}
				next += (uint32_t)nF;
				if(nT + nV == 0)
					continue; /* nobody was stepping: count again */
			}
			else if(nT + nV == 0)
				break; /* no range left (the cursor is past the pool) and every lane is idle */
		}

		/* the stepping phase with more lanes, until a quarter of them has left it */
		const uint32_t mT2 = __ballot_sync(0xFFFFFFFFu, state == ST_TILE);
		const uint32_t mV2 = __ballot_sync(0xFFFFFFFFu, state == ST_VOX);
		const int cT = __popc(mT2), cV = __popc(mV2);
		if(cT >= cV)
		{
			const int keep = (K.endLanes * cT + 7) >> 3;
#pragma unroll 1
			for(int it = 0; it < K.budget; it++)
			{
				if(state == ST_TILE)
					flat_tile_step(S, L, state);
				if(__popc(__ballot_sync(0xFFFFFFFFu, state == ST_TILE)) < keep)
					break;
			}
		}
		else
		{
			const int keep = (K.endLanes * cV + 7) >> 3;
#pragma unroll 1
			for(int it = 0; it < K.budget; it++)
			{
				if(state == ST_VOX)
					flat_vox_step(S, L, state);
				if(__popc(__ballot_sync(0xFFFFFFFFu, state == ST_VOX)) < keep)
					break;
			}
		}

		if(state == ST_END)
		{
			/* the segment is over: its result goes back to the slot, the lane is free */
			const bool inside = L.st.lastVoxID != 255u;
			ctx[(size_t)WH_POS * P + slot] = f3w(L.pos, (L.hit ? WF_HIT : 0u) | (L.st.tripped ? WF_TRIPPED : 0u) | (inside ? WF_INSIDE : 0u));
			ctx[(size_t)WH_COL * P + slot] = f3w(L.colorAdd, __float_as_uint(L.colorMult));
			if(L.hit)
				ctx[(size_t)WH_VOX * P + slot] = L.st.vox;
			if(inside)
				ctx[(size_t)WH_ST * P + slot] = make_uint4(L.st.lastVoxID, __float_as_uint(L.st.lastVoxRefract), 0, 0);
			state = ST_FETCH;
		}
	}
}

/* ---- the lock-step stepping kernel (ray_step.cuh): same slot planes, same per-warp double-buffered cp.async prefetch and dynamic
 * ranges as dn_wave_step_kernel; every lane advances its ray by one cell per trip through ONE loop body (chunk exit, block / layer
 * change, chunk entry, voxel test, step -- no votes, no phases), hits in all-opaque chunks end the ray without touching the record,
 * idle lanes are refilled once `refillMin` of them have gathered (or nothing else runs). ---- */
#ifndef WAVE2_MIN_BLOCKS
#define WAVE2_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(128, WAVE2_MIN_BLOCKS) dn_wave_step2_kernel(DnbScene S, uint4* __restrict__ ctx, uint32_t P, uint32_t* __restrict__ cursor, int refillMin)
{
	__shared__ uint4 s_rays[4][2][5 * 32];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t ltMask = (1u << lane) - 1u;

	uint32_t cur = 0, next = 0, end = 0, preBase = 0xFFFFFFFFu;
	{
		uint32_t base = 0;
		if(lane == 0)
			base = atomicAdd(cursor, WAVE_GRAB);
		preBase = __shfl_sync(0xFFFFFFFFu, base, 0);
		if(preBase >= P)
			return;
		wave_prefetch(s_rays[warp][1], ctx, P, preBase, lane);
	}

	RayLane L;
	bool idle = true;
	uint32_t slot = 0;
	L.hit = false;

	for(;;)
	{
		const uint32_t mI = __ballot_sync(0xFFFFFFFFu, idle);
		if(mI != 0u && (__popc(mI) >= refillMin || mI == 0xFFFFFFFFu))
		{
			if(next >= end && preBase < P)
			{
				/* switch to the prefetched range and start fetching the one after it */
				__pipeline_wait_prior(0);
				__syncwarp();
				cur ^= 1u;
				next = preBase;
				end = preBase + WAVE_GRAB < P ? preBase + WAVE_GRAB : P;
				uint32_t base = 0;
				if(lane == 0)
					base = atomicAdd(cursor, WAVE_GRAB);
				preBase = __shfl_sync(0xFFFFFFFFu, base, 0);
				if(preBase < P)
					wave_prefetch(s_rays[warp][cur ^ 1u], ctx, P, preBase, lane);
			}
			if(next < end)
			{
				if(idle)
				{
					const uint32_t idx = next + (uint32_t)__popc(mI & ltMask);
					if(idx < end)
					{
						const uint4* row = s_rays[warp][cur] + (idx & 31u);
						const uint4 rInv = row[2 * 32];
						if(rInv.w & WF_READY)
						{
							const uint4 rDir = row[0], rPos = row[1 * 32], rSide = row[3 * 32], rCell = row[4 * 32];
							slot = idx;
							L.dir = xyz_of(rDir);
							L.rayPos = xyz_of(rPos);
							L.lastVoxID = rDir.w;
							L.lastVoxRefract = __uint_as_float(rPos.w);
							L.tripped = (rInv.w & WF_TRIPPED) != 0u;
							/* ray_begin with the serve kernel's precomputed tile-level DDA start */
							L.delta = abs3(xyz_of(rInv));
							L.step.x = isgn(L.dir.x); L.step.y = isgn(L.dir.y); L.step.z = isgn(L.dir.z);
							L.pos.x = (int)rCell.x; L.pos.y = (int)rCell.y; L.pos.z = (int)rCell.z;
							L.side = xyz_of(rSide);
							L.tl = 0.0f;
							L.g = 0;
							L.lv = 0;
							L.blk.x = L.blk.y = L.blk.z = 0x40000000;
							L.word = 0;
							L.off.x = L.off.y = L.off.z = 0;
							L.colorAdd = splat3(0.0f);
							L.colorMult = 1.0f;
							L.ignoreFirst = true;
							L.hit = false;
							L.deferred = false;
							L.vox = make_uint4(0, 0, 0, 0);
							idle = false;
						}
					}
				}
				next += (uint32_t)__popc(mI);
			}
			else if(mI == 0xFFFFFFFFu)
				break; /* no range left (the cursor is past the pool) and every lane is idle */
		}

		if(!idle && ray_iter<true>(S, L))
		{
			/* the segment is over: its result goes back to the slot, the lane is free */
			const bool inside = L.lastVoxID != 255u;
			ctx[(size_t)WH_POS * P + slot] = f3w(L.rayPos, (L.hit ? WF_HIT : 0u) | (L.tripped ? WF_TRIPPED : 0u) | (inside ? WF_INSIDE : 0u) | (L.deferred ? WF_DEFERRED : 0u));
			ctx[(size_t)WH_COL * P + slot] = f3w(L.colorAdd, __float_as_uint(L.colorMult));
			if(L.hit)
				ctx[(size_t)WH_VOX * P + slot] = L.vox;
			if(inside)
				ctx[(size_t)WH_ST * P + slot] = make_uint4(L.lastVoxID, __float_as_uint(L.lastVoxRefract), 0, 0);
			idle = true;
		}
	}
}

/* host side of one wavefront dispatch.  Passes are queued without waiting; every pass copies its count of live slots to a
 * pinned ring and the host looks at the count of the pass WAVE_LAG passes back before queueing another, so the device never
 * runs dry and at most WAVE_LAG empty passes are queued after the last voxel has finished. */
#define WAVE_LAG 3
struct DnbWaveHost
{
	uint32_t*   pinned = nullptr; /* ring of 8 counts */
	cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	int         stepCtas = 0, step2Ctas = 0;
};
static DnbWaveHost g_wave;

extern "C" size_t dnb_wave_slot_bytes(void) { return (size_t)WAVE_PLANES * sizeof(uint4); }

/* ctx: WAVE_PLANES * P uint4, P a multiple of 256; counters: 4 device words (work counter, two live-slot counters used alternately, slot cursor of the step kernel) */
extern "C" cudaError_t dnb_launch_light_wave(const DnbScene* scene, const uint32_t* requests, uint32_t numRequests, uint32_t firstCta, uint32_t ctaStride, uint32_t numCtas,
                                             const DnbStagingTargets* targets, uint4* ctx, uint32_t P, uint32_t* counters, uint32_t* passesOut, cudaStream_t stream)
{
	if(passesOut)
		*passesOut = 0;
	if(numRequests == 0 || numCtas == 0 || P == 0)
		return cudaSuccess;
	cudaError_t e;
	if(!g_wave.pinned)
	{
		if((e = cudaMallocHost((void**)&g_wave.pinned, 8 * sizeof(uint32_t))) != cudaSuccess)
			return e;
		for(int i = 0; i < 8; i++)
			if((e = cudaEventCreateWithFlags(&g_wave.ev[i], cudaEventDisableTiming)) != cudaSuccess)
				return e;
		int dev = 0, sms = 148, perSm = WAVE_MIN_BLOCKS;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, dn_wave_step_kernel, 128, 0) != cudaSuccess || perSm < 1)
			perSm = WAVE_MIN_BLOCKS;
		g_wave.stepCtas = sms * perSm;
		int perSm2 = WAVE2_MIN_BLOCKS;
		if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm2, dn_wave_step2_kernel, 128, 0) != cudaSuccess || perSm2 < 1)
			perSm2 = WAVE2_MIN_BLOCKS;
		g_wave.step2Ctas = sms * perSm2;
	}
	DnbFlatTuning tuning;
	{
		auto knob = [](const char* name, int dflt) { const char* v = getenv(name); return v && atoi(v) > 0 ? atoi(v) : dflt; };
		tuning.budget = knob("DN_B200_WAVE_BUDGET", 24);
		tuning.endLanes = knob("DN_B200_WAVE_KEEP", 4); /* a stepping burst ends once fewer than keep/8 of its lanes are still in the phase */
		tuning.patience = 0;
	}

	/* every slot idle, counters zero */
	if((e = cudaMemsetAsync(ctx + (size_t)WV_DIFF * P, 0, (size_t)P * sizeof(uint4), stream)) != cudaSuccess)
		return e;
	if((e = cudaMemsetAsync(ctx + (size_t)WR_INV * P, 0, (size_t)P * sizeof(uint4), stream)) != cudaSuccess)
		return e;
	if((e = cudaMemsetAsync(counters, 0, 4 * sizeof(uint32_t), stream)) != cudaSuccess)
		return e;

	/* which stepping kernel: 2 = lock-step (default), 1 = the two-phase kernel it replaced (kept for A/B runs: $DN_B200_WAVE_STEP=1) */
	static const int stepKind = getenv("DN_B200_WAVE_STEP") && atoi(getenv("DN_B200_WAVE_STEP")) == 1 ? 1 : 2;
	static const int refillMin = [] { const char* v = getenv("DN_B200_WAVE_REFILL"); return v && atoi(v) > 0 ? atoi(v) : 4; }();
	const uint32_t totalItems = numCtas * 128u;
	const uint32_t stepCtas = std::min<uint32_t>((uint32_t)(stepKind == 2 ? g_wave.step2Ctas : g_wave.stepCtas), (P + 4u * WAVE_GRAB - 1u) / (4u * WAVE_GRAB));

	static const bool trace = getenv("DN_B200_WAVE_TRACE") != nullptr;
	uint32_t pass = 0;
	for(;; pass++)
	{
		if(pass >= WAVE_LAG)
		{
			const uint32_t look = pass - WAVE_LAG;
			if((e = cudaEventSynchronize(g_wave.ev[look & 7u])) != cudaSuccess)
				return e;
			if(trace)
				fprintf(stderr, "wave pass %u: %u live slots of %u\n", look, g_wave.pinned[look & 7u], P);
			if(g_wave.pinned[look & 7u] == 0u)
				break; /* that pass left no live slot: every voxel of the dispatch is staged */
		}
		uint32_t* now = counters + 1 + (pass & 1u);
		uint32_t* nxt = counters + 1 + ((pass + 1u) & 1u);
		{ DNB_LAUNCHED(1); dn_wave_serve_kernel<<<P / WAVE_SERVE_THREADS, WAVE_SERVE_THREADS, 0, stream>>>(*scene, requests, numRequests, firstCta, ctaStride, totalItems, counters, *targets, ctx, P, now, nxt); }
		if((e = cudaMemcpyAsync(&g_wave.pinned[pass & 7u], now, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream)) != cudaSuccess)
			return e;
		if((e = cudaEventRecord(g_wave.ev[pass & 7u], stream)) != cudaSuccess)
			return e;
		if(stepKind == 2)
			{ DNB_LAUNCHED(1); dn_wave_step2_kernel<<<stepCtas, 128, 0, stream>>>(*scene, ctx, P, counters + 3, refillMin); }
		else
			{ DNB_LAUNCHED(1); dn_wave_step_kernel<<<stepCtas, 128, 0, stream>>>(*scene, ctx, P, counters + 3, tuning); }
		if((e = cudaGetLastError()) != cudaSuccess)
			return e;
	}
	if(passesOut)
		*passesOut = pass;
	return cudaSuccess;
}
