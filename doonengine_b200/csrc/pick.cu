/* pick.cu -- batched picking rays on the device map: DN_step_map (reference voxel.c:1195-1272) for many rays at once.
 *
 * DN_step_map walks ONE ray over the CPU map: a single-level voxel DDA that steps exactly one axis per iteration (strict `<`
 * comparisons, x before y before z on ties -- unlike the shaders' DDA, which steps all tied axes), tests `maxSteps` cells
 * beginning with the one the ray starts in, and reports the first cell that holds a voxel, plus the face it was entered through.
 * Here one thread runs one ray with the same float recurrence (same initial sideDist expression, evaluated in double where the
 * C expression promotes to double) against the device map: an empty tile costs a bit test in the register-cached 4x4x4
 * occupancy word, a resident tile one test in the chunk's surface mask.
 *
 * The device map holds SURFACE voxels only (chunk-local culling, voxel.c:1391-1394).  That cannot change the first cell found:
 * the walk moves face to face, a culled voxel has six solid opaque neighbours inside its chunk, so it can only be entered from
 * another solid voxel, and the first solid voxel met after an empty cell is never culled.  The one exception -- a ray that
 * STARTS inside a culled voxel -- is settled by the host from the CPU map (engine.cpp DN_b200_step_map_batch), which also reads
 * the hit voxel's contents there (the device records hold linearised albedo, voxel.c:1441-1447).
 */
#include "kernels.h"
#include "vecmath.cuh"

/* out[i] = {cell.x, cell.y, cell.z, code}: code bit 0 = hit, bits 1-2 = axis of the last step (3 = no step taken), bit 3 = that step was positive */
__global__ void __launch_bounds__(128) dn_pick_kernel(DnbScene S, const float* __restrict__ dirs, const float* __restrict__ origins, uint32_t count, int maxSteps, int4* __restrict__ out)
{
	const uint32_t i = blockIdx.x * 128u + threadIdx.x;
	if(i >= count)
		return;

	int cell[3], step[3];
	float delta[3], side[3];
#pragma unroll
	for(int a = 0; a < 3; a++)
	{
		const float d = __ldg(dirs + 3 * (size_t)i + a);
		const float p = __ldg(origins + 3 * (size_t)i + a) * 8.0f; /* chunk units -> voxels */
		const float inv = 1.0f / d;
		const int sg = d > 0.0f ? 1 : (d < 0.0f ? -1 : 0);
		cell[a] = (int)floorf(p);
		delta[a] = fabsf(inv);
		step[a] = sg;
		/* (float)((sg * (cell - p) + (sg * 0.5) + 0.5) * delta): float product, then double sums and product, as C evaluates it */
		const float t = (float)sg * ((float)cell[a] - p);
		side[a] = (float)((((double)t + (double)sg * 0.5) + 0.5) * (double)delta[a]);
	}

	int blk[3] = {0x40000000, 0x40000000, 0x40000000};
	unsigned long long occWord = 0;
	uint32_t code = 3u << 1;
	for(int n = 0; n < maxSteps; n++)
	{
		if((cell[0] | cell[1] | cell[2]) >= 0)
		{
			const int tx = cell[0] >> 3, ty = cell[1] >> 3, tz = cell[2] >> 3;
			if((uint32_t)tx < S.mapSize[0] && (uint32_t)ty < S.mapSize[1] && (uint32_t)tz < S.mapSize[2])
			{
				if(((tx ^ blk[0]) | (ty ^ blk[1]) | (tz ^ blk[2])) > 3)
				{
					blk[0] = tx & ~3; blk[1] = ty & ~3; blk[2] = tz & ~3;
					occWord = __ldg(S.occ64 + ((uint32_t)(tx >> 2) + S.blocks[0] * ((uint32_t)(ty >> 2) + S.blocks[1] * (uint32_t)(tz >> 2))));
				}
				const uint32_t bit = (uint32_t)(tx & 3) | ((uint32_t)(ty & 3) << 2) | ((uint32_t)(tz & 3) << 4);
				if((occWord >> bit) & 1ull)
				{
					const uint32_t mapIndex = (uint32_t)tx + S.mapSize[0] * ((uint32_t)ty + S.mapSize[1] * (uint32_t)tz);
					const DnbSlot* slot = S.slots + (__ldg(S.tileSlot + mapIndex) - 1u);
					const uint32_t local = (uint32_t)(cell[0] & 7) + 8u * ((uint32_t)(cell[1] & 7) + 8u * (uint32_t)(cell[2] & 7));
					if((__ldg(slot->mask + (local >> 5)) >> (local & 31u)) & 1u)
					{
						out[i] = make_int4(cell[0], cell[1], cell[2], (int)(code | 1u));
						return;
					}
				}
			}
		}

		int axis;
		if(side[0] < side[1])
			axis = side[0] < side[2] ? 0 : 2;
		else
			axis = side[1] < side[2] ? 1 : 2;
		/* (static indexing keeps the arrays in registers) */
		if(axis == 0) { side[0] += delta[0]; cell[0] += step[0]; code = (0u << 1) | (step[0] > 0 ? 8u : 0u) | (step[0] == 0 ? 16u : 0u); }
		else if(axis == 1) { side[1] += delta[1]; cell[1] += step[1]; code = (1u << 1) | (step[1] > 0 ? 8u : 0u) | (step[1] == 0 ? 16u : 0u); }
		else { side[2] += delta[2]; cell[2] += step[2]; code = (2u << 1) | (step[2] > 0 ? 8u : 0u) | (step[2] == 0 ? 16u : 0u); }
	}
	out[i] = make_int4(cell[0], cell[1], cell[2], (int)code);
}

extern "C" cudaError_t dnb_launch_pick(const DnbScene* scene, const float* dirs, const float* origins, uint32_t count, int maxSteps, int4* out, cudaStream_t stream)
{
	if(count == 0)
		return cudaSuccess;
	{ DNB_LAUNCHED(1); dn_pick_kernel<<<(count + 127u) / 128u, 128, 0, stream>>>(*scene, dirs, origins, count, maxSteps, out); }
	return cudaGetLastError();
}
