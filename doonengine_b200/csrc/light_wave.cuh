/* light_wave.cuh (included at the end of light.cu, after light_flat.cuh whose per-voxel functions it reuses) -- the lighting
 * update as a WAVEFRONT: the voxel contexts live in global memory and two kernels alternate over a pool of P context slots.
 *
 *   dn_wave_serve_kernel   one thread per slot, full warps: shade the ray segment that just ended (flat_ray_ended: LI:96-145,
 *                          174-202, 77-79), advance the voxel's ray schedule, store a finished voxel's three staged words and
 *                          immediately set the slot up with the next voxel of the dispatch (flat_setup_voxel, LI:207-251), then
 *                          prepare the next ray segment completely (direction, reciprocal, tile-level DDA start) and store it;
 *                          it also fetches the record of a DEFERRED hit (ray_step.cuh) -- with full warps
 *   dn_wave_step_kernel    persistent warps; every lane advances one prepared ray segment by one cell per trip through the
 *                          lock-step loop body of ray_step.cuh (= trace_ray<false,false>, trace.cuh) and, when its ray ends,
 *                          stores the result and takes the next live slot -- a lane never waits for shading
 *
 * Why a wavefront: in dn_light_flat_kernel (one context per lane, in registers) ncu attributes ~40 % of all issue slots to ray set-up
 * and shading executed with 2-4 active lanes (profiles/r1_v3_wave.md).  Nothing about a voxel's own sequence of operations changes --
 * the same functions run in the same order on the same values -- so the staged words are bit-identical to both other kernels
 * (tests/test_parity_gpu.py runs every lighting test against all three, with the staging array poisoned in between).
 *
 * LIVE ENTRIES.  A voxel with 15 specular rays of two segments needs ~35 passes after the last voxel of the dispatch has been
 * fetched, so most passes of a dispatch run with a nearly empty pool (c3s: 24 full passes, ~70 draining ones).  Scanning all P slots
 * in those made the drain as expensive as the full passes.  Each serve pass therefore writes a list of LIVE ENTRIES -- (first slot,
 * 32-bit mask) for every warp-sized group of slots that holds at least one ray to trace; the step kernel walks that list (coalesced
 * prefetch of a group's ray planes, lanes take the mask's slots), and once the work counter is exhausted ("drain mode", entered by
 * the host a few passes late, which is harmless) the serve kernel walks the previous pass's list too instead of the whole pool.
 *
 * Slot layout: structure of arrays, 15 planes of P uint4 (240 bytes per slot) so that both kernels move whole 16-byte words and the
 * serve kernel's accesses are fully coalesced, plus a 16th plane that holds the two entry lists.
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
	WX_LISTS = 15, /* not per slot: two lists of P / 32 live entries (uint2: first slot, mask) */
	WAVE_PLANES = 16
};
/* device counters of a dispatch (8 words): */
enum : uint32_t { WC_WORK = 0, WC_LIVE0 = 1, WC_LIVE1 = 2, WC_CURSOR = 3, WC_ENTRIES0 = 4, WC_ENTRIES1 = 5, WC_PREV_ENTRIES = 6, WC_WORDS = 8 };
enum : uint32_t { WF_READY = 1u, WF_TRIPPED = 2u, WF_HIT = 1u, WF_INSIDE = 4u, WF_DEFERRED = 8u };
/* schedule word: kind[0:2) idx[2:8) seg[8:16) firstSample[16] sourceVisible[17] reflectType[18:26) active[31] */
#define WS_ACTIVE 0x80000000u

DNB_FN uint4 f3w(f3 a, uint32_t w) { return make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), w); }
DNB_FN f3 xyz_of(uint4 v) { return mk3(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z)); }

#define WAVE_SERVE_THREADS 256
/* pass: parity selects the counters / entry list written (pass & 1) and read (the other one).  drain != 0: the work counter is
 * exhausted -- only the slots of the previous pass's live entries are visited (warp w of the grid takes entry w) and nothing is
 * fetched; otherwise the grid covers the pool (P is a multiple of 256) and free slots take new voxels. */
__global__ void __launch_bounds__(WAVE_SERVE_THREADS) dn_wave_serve_kernel(DnbScene S, const uint32_t* __restrict__ requests, DnbWork W,
                                                            uint32_t* __restrict__ counters, DnbStagingTargets T, uint4* __restrict__ ctx, uint32_t P, uint32_t pass, uint32_t drain)
{
	/* the request count is read on the device (layout.h DnbWork); a work item = one lane of a request */
	const uint32_t numRequests = work_requests(W);
	const uint32_t firstCta = W.firstCta, ctaStride = W.ctaStride;
	const uint32_t totalItems = work_ctas(W, numRequests) * 128u;
	__shared__ uint32_t s_warpCount[WAVE_SERVE_THREADS / 32], s_base, s_liveEntries, s_liveSlots, s_entryBase;
	const uint32_t lane = threadIdx.x & 31u, warpInCta = threadIdx.x >> 5;
	const uint32_t ltMask = (1u << lane) - 1u;
	const uint32_t par = pass & 1u;
	uint32_t* const workCounter = counters + WC_WORK;
	uint32_t* const liveNow = counters + WC_LIVE0 + par;
	uint32_t* const entriesNow = counters + WC_ENTRIES0 + par;
	uint2* const listNow = reinterpret_cast<uint2*>(ctx + (size_t)WX_LISTS * P) + (size_t)par * (P / 32u);
	const uint2* const listPrev = reinterpret_cast<const uint2*>(ctx + (size_t)WX_LISTS * P) + (size_t)(par ^ 1u) * (P / 32u);
	if(blockIdx.x == 0 && threadIdx.x == 0)
	{
		counters[WC_LIVE0 + (par ^ 1u)] = 0;    /* the counters the NEXT pass adds to (stream order keeps the passes apart) */
		counters[WC_ENTRIES0 + (par ^ 1u)] = 0;
		counters[WC_CURSOR] = 0;                /* entry cursor of the step kernel that follows this pass */
	}
	if(threadIdx.x == 0)
		s_liveEntries = s_liveSlots = 0;

	/* which slot */
	uint32_t i;
	bool mine = true;
	if(drain)
	{
		/* warp e of the grid serves the slots of entry e of the previous pass's list; its length was left in counters[WC_PREV_ENTRIES]
		 * by the step kernel that walked it (the alternating counter itself is being cleared for the next pass right now) */
		const uint32_t e = blockIdx.x * (WAVE_SERVE_THREADS / 32) + warpInCta;
		i = 0;
		mine = false;
		if(e < counters[WC_PREV_ENTRIES])
		{
			const uint2 ent = __ldg(listPrev + e);
			i = ent.x + lane;
			mine = (ent.y >> lane) & 1u;
		}
	}
	else
		i = blockIdx.x * WAVE_SERVE_THREADS + threadIdx.x;
#define PL(p) ctx[(size_t)(p) * P + i]

	FlatLane L;
	uint32_t state = ST_DONE, item = 0;
	bool start = false;
	bool wasActive = false;

	if(mine)
	{
		state = ST_FETCH;
		const uint4 vDiff = PL(WV_DIFF);
		wasActive = (vDiff.w & WS_ACTIVE) != 0u;
		if(wasActive)
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
			L.chunkOpaque = false; /* (flat_hit_record: the hit's record is settled here, not by the persistent kernel's deferred fetch) */
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
	}

	/* free slots take the next voxels of the dispatch (two rounds: an item can turn out to hold no voxel).  ONE atomic per CTA
	 * and round: with one per warp the 10^5 same-address atomics of a pass were a quarter of the kernel's stall samples.
	 * `need` -- not `state` -- says who still fetches: a slot that has just been given a voxel keeps state ST_FETCH (nothing sets it
	 * before the store below) and must NOT take, and thereby drop, a second item (round 1's kernels did: every voxel fetched while
	 * work remained was lost, hidden by stale staging rows).  Drain mode: the list is exhausted, nothing to fetch. */
	bool need = state == ST_FETCH && !drain;
	if(!drain)
	{
#pragma unroll 1
		for(int round = 0; round < 2; round++)
		{
			const uint32_t mF = __ballot_sync(0xFFFFFFFFu, need);
			if(lane == 0)
				s_warpCount[warpInCta] = (uint32_t)__popc(mF);
			__syncthreads();
			uint32_t before = 0, total = 0;
#pragma unroll
			for(uint32_t w = 0; w < WAVE_SERVE_THREADS / 32; w++)
			{
				const uint32_t c = s_warpCount[w];
				before += w < warpInCta ? c : 0u;
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
	else if(wasActive)
	{
		/* the slot goes idle (an idle slot stays as it is: both words are already zero) */
		PL(WV_DIFF) = make_uint4(0, 0, 0, 0);
		PL(WR_INV) = make_uint4(0, 0, 0, 0);
	}
#undef PL

	/* this warp's live entry.  One atomic per CTA for the list position and the two counters. */
	const uint32_t mA = __ballot_sync(0xFFFFFFFFu, start);
	__syncthreads();
	uint32_t myEntry = 0;
	if(lane == 0 && mA)
	{
		myEntry = atomicAdd(&s_liveEntries, 1u);
		atomicAdd(&s_liveSlots, (uint32_t)__popc(mA));
	}
	__syncthreads();
	if(threadIdx.x == 0 && s_liveEntries)
	{
		s_entryBase = atomicAdd(entriesNow, s_liveEntries);
		atomicAdd(liveNow, s_liveSlots);
	}
	__syncthreads();
	if(lane == 0 && mA)
		listNow[s_entryBase + myEntry] = make_uint2(i, mA); /* lane 0's slot is the first of the warp's 32 (fill mode: i is a multiple of 32; drain mode: the entry's base) */
}

#ifndef WAVE_MIN_BLOCKS
#define WAVE_MIN_BLOCKS 5
#endif

/* the five ray planes of the 32 consecutive slots of a live entry -> one of the warp's two shared-memory buffers (cp.async, 16 bytes
 * per lane and plane: each plane row is one fully coalesced 512-byte read) */
DNB_FN void wave_prefetch(uint4* __restrict__ buf, const uint4* __restrict__ ctx, uint32_t P, uint32_t base, uint32_t lane)
{
	const uint32_t idx = base + lane < P ? base + lane : P - 1u;
#pragma unroll
	for(uint32_t p = 0; p < 5u; p++)
		__pipeline_memcpy_async(buf + p * 32u + lane, ctx + (size_t)(WR_DIR + p) * P + idx, sizeof(uint4));
	__pipeline_commit();
}

/* persistent warps over the pass's live entries.  Per warp: two buffers of 32 slots x 5 ray planes (2 x 2.5 KB); while the lanes
 * work through one entry the next one's planes are already on their way, so handing an idle lane its next ray costs five
 * shared-memory reads instead of a round trip to HBM.  Every lane advances its ray by one cell per trip through ONE loop body
 * (ray_step.cuh: chunk exit, block / layer change, chunk entry, voxel test, step -- no votes, no phases); hits in all-opaque chunks
 * end the ray without touching the record; idle lanes are refilled once `refillMin` of them have gathered or nothing else runs. */
__global__ void __launch_bounds__(128, WAVE_MIN_BLOCKS) dn_wave_step_kernel(DnbScene S, uint4* __restrict__ ctx, uint32_t P, uint32_t* __restrict__ counters, uint32_t pass, int refillMin)
{
	__shared__ uint4 s_rays[4][2][5 * 32];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t ltMask = (1u << lane) - 1u;
	const uint32_t par = pass & 1u;
	const uint32_t numEntries = counters[WC_ENTRIES0 + par];
	const uint2* const list = reinterpret_cast<const uint2*>(ctx + (size_t)WX_LISTS * P) + (size_t)par * (P / 32u);
	uint32_t* const cursor = counters + WC_CURSOR;
	if(blockIdx.x == 0 && threadIdx.x == 0)
		counters[WC_PREV_ENTRIES] = numEntries; /* for a drain-mode serve pass (see there) */

	/* current entry: slots curBase + (bits of curMask), planes in buffer `cur`; (preBase, preMask) is in flight into the other buffer */
	uint32_t cur = 0, curBase = 0, curMask = 0, preBase = 0, preMask = 0;
	bool more = true; /* the cursor has not run past the list yet */
	{
		uint32_t e = 0;
		if(lane == 0)
			e = atomicAdd(cursor, 1u);
		e = __shfl_sync(0xFFFFFFFFu, e, 0);
		if(e >= numEntries)
			return;
		const uint2 ent = __ldg(list + e);
		preBase = ent.x;
		preMask = ent.y;
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
			if(curMask == 0u && preMask != 0u)
			{
				/* switch to the prefetched entry and start fetching the one after it */
				__pipeline_wait_prior(0);
				__syncwarp();
				cur ^= 1u;
				curBase = preBase;
				curMask = preMask;
				preMask = 0u;
				if(more)
				{
					uint32_t e = 0;
					if(lane == 0)
						e = atomicAdd(cursor, 1u);
					e = __shfl_sync(0xFFFFFFFFu, e, 0);
					if(e < numEntries)
					{
						const uint2 ent = __ldg(list + e);
						preBase = ent.x;
						preMask = ent.y;
						wave_prefetch(s_rays[warp][cur ^ 1u], ctx, P, preBase, lane);
					}
					else
						more = false;
				}
			}
			if(curMask != 0u)
			{
				/* the k-th idle lane takes the k-th remaining slot of the entry */
				const uint32_t k = (uint32_t)__popc(mI & ltMask);
				const uint32_t avail = (uint32_t)__popc(curMask), want = (uint32_t)__popc(mI);
				if(idle && k < avail)
				{
					const uint32_t bit = __fns(curMask, 0, (int)k + 1);
					const uint4* row = s_rays[warp][cur] + bit;
					const uint4 rDir = row[0], rPos = row[1 * 32], rInv = row[2 * 32], rSide = row[3 * 32], rCell = row[4 * 32];
					slot = curBase + bit;
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
				/* drop the slots just handed out from the mask */
				if(want >= avail)
					curMask = 0u;
				else
					curMask &= ~((2u << __fns(curMask, 0, (int)want)) - 1u);
				continue; /* count again: an entry may have held fewer slots than there were idle lanes */
			}
			if(mI == 0xFFFFFFFFu)
				break; /* no entry left and every lane is idle */
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

/* host side of one wavefront dispatch.  Passes are queued without waiting; the step kernel of every pass publishes {live slots, live
 * entries, work counter} straight into a pinned ring (zero-copy stores: no copy-engine transfer that could queue behind a framebuffer
 * read-back) and the host looks at the pass WAVE_LAG passes back before queueing another, so the device never runs dry and at most
 * WAVE_LAG empty passes are queued after the last voxel has finished.  The same look-back switches the serve kernel to drain mode once
 * the work counter is exhausted and sizes both grids by the (from then on shrinking) number of live entries. */
#define WAVE_LAG 3
struct DnbWaveHost
{
	uint32_t*   pinned = nullptr; /* ring of 8 x {live slots, live entries, work counter, -} */
	cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	int         stepCtas = 0;
};
static DnbWaveHost g_wave;

__global__ void dn_wave_publish_kernel(const uint32_t* __restrict__ counters, DnbWork W, uint32_t pass, uint32_t* __restrict__ hostRing)
{
	uint32_t* out = hostRing + (pass & 7u) * 4u;
	out[0] = counters[WC_LIVE0 + (pass & 1u)];
	out[1] = counters[WC_ENTRIES0 + (pass & 1u)];
	out[2] = counters[WC_WORK];
	out[3] = work_ctas(W, work_requests(W)) * 128u; /* the dispatch's work items: the list is exhausted once the work counter has reached this */
	__threadfence_system();
}

extern "C" size_t dnb_wave_slot_bytes(void) { return (size_t)WAVE_PLANES * sizeof(uint4); }

/* ctx: WAVE_PLANES * P uint4, P a multiple of 256; counters: WC_WORDS (8) device words */
extern "C" cudaError_t dnb_launch_light_wave(const DnbScene* scene, const uint32_t* requests, const DnbWork* work,
                                             const DnbStagingTargets* targets, uint4* ctx, uint32_t P, uint32_t* counters, uint32_t* passesOut, cudaStream_t stream)
{
	if(passesOut)
		*passesOut = 0;
	if(P == 0)
		return cudaSuccess;
	cudaError_t e;
	if(!g_wave.pinned)
	{
		if((e = cudaHostAlloc((void**)&g_wave.pinned, 8 * 4 * sizeof(uint32_t), cudaHostAllocMapped)) != cudaSuccess)
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
	}
	uint32_t* ringDev = nullptr;
	if((e = cudaHostGetDevicePointer((void**)&ringDev, g_wave.pinned, 0)) != cudaSuccess)
		return e;

	/* every slot idle, counters zero */
	if((e = cudaMemsetAsync(ctx + (size_t)WV_DIFF * P, 0, (size_t)P * sizeof(uint4), stream)) != cudaSuccess)
		return e;
	if((e = cudaMemsetAsync(ctx + (size_t)WR_INV * P, 0, (size_t)P * sizeof(uint4), stream)) != cudaSuccess)
		return e;
	if((e = cudaMemsetAsync(counters, 0, WC_WORDS * sizeof(uint32_t), stream)) != cudaSuccess)
		return e;

	static const int refillMin = [] { const char* v = getenv("DN_B200_WAVE_REFILL"); return v && atoi(v) > 0 ? atoi(v) : 4; }();

	static const bool trace = getenv("DN_B200_WAVE_TRACE") != nullptr;
	uint32_t pass = 0;
	bool drain = false;
	uint32_t entryBound = P / 32u; /* upper bound of the number of live entries the next pass can meet */
	for(;; pass++)
	{
		if(pass >= WAVE_LAG)
		{
			const uint32_t look = pass - WAVE_LAG;
			if((e = cudaEventSynchronize(g_wave.ev[look & 7u])) != cudaSuccess)
				return e;
			const volatile uint32_t* seen = g_wave.pinned + (look & 7u) * 4u;
			if(trace)
				fprintf(stderr, "wave pass %u: %u live slots in %u entries of %u, work counter %u of %u%s\n", look, seen[0], seen[1], P / 32u, seen[2], seen[3], drain ? " (drain mode)" : "");
			if(seen[0] == 0u)
				break; /* that pass left no live slot: every voxel of the dispatch is staged */
			if(seen[2] >= seen[3])
			{
				/* the list was exhausted by then: no voxel has started since, so the live entries can only have become fewer */
				drain = true;
				entryBound = std::min<uint32_t>(entryBound, (uint32_t)seen[1]);
			}
		}
		const uint32_t serveCtas = drain ? (entryBound + WAVE_SERVE_THREADS / 32 - 1) / (WAVE_SERVE_THREADS / 32) : P / WAVE_SERVE_THREADS;
		if(serveCtas > 0)
			{ DNB_LAUNCHED(1); dn_wave_serve_kernel<<<serveCtas, WAVE_SERVE_THREADS, 0, stream>>>(*scene, requests, *work, counters, *targets, ctx, P, pass, drain ? 1u : 0u); }
		const uint32_t stepCtas = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)g_wave.stepCtas, (entryBound + 3u) / 4u));
		{ DNB_LAUNCHED(1); dn_wave_step_kernel<<<stepCtas, 128, 0, stream>>>(*scene, ctx, P, counters, pass, refillMin); }
		{ DNB_LAUNCHED(1); dn_wave_publish_kernel<<<1, 1, 0, stream>>>(counters, *work, pass, ringDev); }
		if((e = cudaEventRecord(g_wave.ev[pass & 7u], stream)) != cudaSuccess)
			return e;
		if((e = cudaGetLastError()) != cudaSuccess)
			return e;
	}
	if(passesOut)
		*passesOut = pass;
	return cudaSuccess;
}
