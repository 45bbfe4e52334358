/* wave_step_unified.cu -- PROTOTYPE kernel (not built into libdoon_b200.so; compiled here only to read its register / instruction
 * statistics): dn_wave_step_kernel of csrc/light_wave.cuh with the unified stepping routine of unified_step.cuh.  Same slot planes, same
 * per-warp double-buffered cp.async prefetch, same dynamic ranges; what changes is the lane state machine:
 *
 *     STEP      cheap step, identical code for the tile level and the voxel level  -> most lanes, run in bursts
 *     BOUNDARY / ENTER / HIT   events; a lane waits until its event is the most populated one (or nothing else can run)
 *     END       store the result, become idle; idle lanes are refilled from the shared-memory range every trip
 *
 * Correctness of the per-lane functions: tests/test_unified_step_proto.py (bit-identical to trace_ray<false,false> on the CPU).
 * Build for statistics:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I doonengine_b200/csrc -I include
 *                        -Xptxas -v -c tools/proto/wave_step_unified.cu -o /dev/null */
#include "unified_step.cuh"
#include <cuda_pipeline.h>

/* plane numbering of csrc/light_wave.cuh */
enum : uint32_t { WR_DIR = 6, WR_POS = 7, WR_INV = 8, WR_SIDE = 9, WR_CELL = 10, WH_POS = 11, WH_COL = 12, WH_VOX = 13, WH_ST = 14 };
enum : uint32_t { WF_READY = 1u, WF_TRIPPED = 2u, WF_HIT = 1u, WF_INSIDE = 4u };
enum : uint32_t { U_IDLE = 5 };
#define WAVE_GRAB 32u

DNB_FN uint4 f3w(f3 a, uint32_t w) { return make_uint4(__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), w); }
DNB_FN f3 xyz_of(uint4 v) { return mk3(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z)); }

DNB_FN void wave_prefetch(uint4* __restrict__ buf, const uint4* __restrict__ ctx, uint32_t P, uint32_t base, uint32_t lane)
{
	const uint32_t idx = base + lane < P ? base + lane : P - 1u;
#pragma unroll
	for(uint32_t p = 0; p < 5u; p++)
		__pipeline_memcpy_async(buf + p * 32u + lane, ctx + (size_t)(WR_DIR + p) * P + idx, sizeof(uint4));
	__pipeline_commit();
}

__global__ void __launch_bounds__(128, 5) dn_wave_step_unified_kernel(DnbScene S, uint4* __restrict__ ctx, uint32_t P, uint32_t* __restrict__ cursor, int budget, int keepEighths)
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

	UniLane L;
	uint32_t state = U_IDLE, slot = 0;
	L.hit = false;

	for(;;)
	{
		const uint32_t mI = __ballot_sync(0xFFFFFFFFu, state == U_IDLE);
		if(mI)
		{
			if(next >= end && preBase < P)
			{
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
				if(state == U_IDLE)
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
							L.inv = xyz_of(rInv);
							L.rayPos = xyz_of(rPos);
							ray_state_reset(L.st);
							L.st.lastVoxID = rDir.w;
							L.st.lastVoxRefract = __uint_as_float(rPos.w);
							L.st.tripped = (rInv.w & WF_TRIPPED) != 0u;
							/* uni_start with the serve kernel's precomputed tile-level DDA */
							L.delta = abs3(L.inv);
							L.step.x = isgn(L.dir.x); L.step.y = isgn(L.dir.y); L.step.z = isgn(L.dir.z);
							L.pos.x = (int)rCell.x; L.pos.y = (int)rCell.y; L.pos.z = (int)rCell.z;
							L.side = xyz_of(rSide);
							L.tl = 0.0f;
							L.g = 0;
							L.blk.x = L.blk.y = L.blk.z = 0x40000000;
							L.word = 0;
							uni_level_params(L, 0);
							L.off.x = L.off.y = L.off.z = 0;
							L.colorAdd = splat3(0.0f);
							L.colorMult = 1.0f;
							L.ignoreFirst = true;
							L.hit = false;
							state = U_STEP;
						}
					}
				}
				next += (uint32_t)__popc(mI);
			}
		}

		const int nS = __popc(__ballot_sync(0xFFFFFFFFu, state == U_STEP));
		const uint32_t mB = __ballot_sync(0xFFFFFFFFu, state == U_BOUNDARY);
		const uint32_t mE = __ballot_sync(0xFFFFFFFFu, state == U_ENTER);
		const uint32_t mH = __ballot_sync(0xFFFFFFFFu, state == U_HIT);
		const int nB = __popc(mB), nE = __popc(mE), nH = __popc(mH);
		if(nS + nB + nE + nH == 0)
		{
			if(next >= end && preBase >= P)
				break;
			continue;
		}

		const int nEvMax = nB > nE ? (nB > nH ? nB : nH) : (nE > nH ? nE : nH);
		if(nS >= nEvMax)
		{
			const int keep = (keepEighths * nS + 7) >> 3;
#pragma unroll 1
			for(int it = 0; it < budget; it++)
			{
				if(state == U_STEP)
					uni_step(S, L, state);
				if(__popc(__ballot_sync(0xFFFFFFFFu, state == U_STEP)) < keep)
					break;
			}
		}
		else if(nB == nEvMax)
		{
			if(state == U_BOUNDARY)
				uni_boundary(S, L, state);
		}
		else if(nE == nEvMax)
		{
			if(state == U_ENTER)
				uni_enter(S, L, state);
		}
		else if(state == U_HIT)
			uni_hit(S, L, state);

		if(state == U_END)
		{
			const bool inside = L.st.lastVoxID != 255u;
			ctx[(size_t)WH_POS * P + slot] = f3w(L.rayPos, (L.hit ? WF_HIT : 0u) | (L.st.tripped ? WF_TRIPPED : 0u) | (inside ? WF_INSIDE : 0u));
			ctx[(size_t)WH_COL * P + slot] = f3w(L.colorAdd, __float_as_uint(L.colorMult));
			if(L.hit)
				ctx[(size_t)WH_VOX * P + slot] = L.st.vox;
			if(inside)
				ctx[(size_t)WH_ST * P + slot] = make_uint4(L.st.lastVoxID, __float_as_uint(L.st.lastVoxRefract), 0, 0);
			state = U_IDLE;
		}
	}
}
