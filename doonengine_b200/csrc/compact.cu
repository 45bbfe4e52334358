/* compact.cu -- visible-chunk collection: builds the lighting-request list on the device.
 *
 * Replaces the reference's per-frame CPU loop over every map tile behind a blocking glMapBuffer
 * (voxel.c:738-762 calling _DN_request_chunk_lighting, voxel.c:1463-1489).  Output is identical: for every tile in
 * ascending flat index that is resident, visible and selected by the lighting split
 * (mapIndex % split == frameNum, or the chunk has pending edits), ceil(numVoxels/32) words
 * (mapIndex << 4) | group, groups ascending.
 *
 * Input is the 1-bit-per-tile visible bitmap, so a 2048^3-voxel map (16.7 M tiles) is a 2 MB scan instead of
 * a 201 MB handle read-back.  Three launches:
 *   dn_compact_kernel<false>  one thread per tile, one warp per bitmap word: lane l tests bit l (lane order = tile
 *                             order), gathers the slot's voxel count, and a shuffle prefix sum gives the per-word
 *                             request count.  Per-CTA totals (256 tiles) go to blockCounts[].
 *   dn_scan_blocks_kernel     one CTA: exclusive scan of blockCounts[] and the grand total.
 *   dn_compact_kernel<true>   same walk, writing the request words at their final offsets.
 * The scan kernel also stores the total into a pinned host word (DNvolume::numLightingRequests is public, and the host sizes
 * the lighting dispatch from it).  Traffic: numTiles/8 bytes of bitmap + 4 B tileSlot + 8 B slot fields per VISIBLE tile
 * + 4 B per request; HBM/L2-streaming bound.
 */
#include "kernels.h"

#define COMPACT_WARPS 8
#define COMPACT_TILES_PER_CTA (COMPACT_WARPS * 32)

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, uint32_t lane)
{
#pragma unroll
	for(int d = 1; d < 32; d <<= 1)
	{
		uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, d);
		if(lane >= (uint32_t)d)
			v += n;
	}
	return v;
}

/* one thread per tile, one warp per bitmap word: every dependent gather (bitmap word -> slot id -> voxel count) of a CTA is
 * in flight at once, so a pass costs about two memory latencies however many tiles are visible */
template <bool WRITE>
__global__ void __launch_bounds__(COMPACT_WARPS * 32) dn_compact_kernel(const uint32_t* __restrict__ visible, const uint32_t* __restrict__ forced, const uint32_t* __restrict__ tileSlot,
                                                                        const DnbSlot* __restrict__ slots, uint32_t numTiles, uint32_t split, uint32_t frameNum,
                                                                        uint32_t* __restrict__ blockCounts, const uint32_t* __restrict__ blockOffsets, uint32_t* __restrict__ requests)
{
	__shared__ uint32_t s_warpTotal[COMPACT_WARPS];

	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t tile = blockIdx.x * COMPACT_TILES_PER_CTA + threadIdx.x;
	const uint32_t word = (tile - lane) < numTiles ? __ldg(visible + (tile >> 5)) : 0u; /* same address for the whole warp */

	uint32_t groups = 0;
	if(((word >> lane) & 1u) && tile < numTiles)
	{
		const uint32_t slotId = __ldg(tileSlot + tile);
		if(slotId != 0u)
		{
			bool selected = (tile % split) == frameNum;
			if(!selected && forced)
				selected = (__ldg(forced + (tile >> 5)) >> lane) & 1u;
			if(selected)
				groups = (__ldg(&slots[slotId - 1u].numVoxels) + 31u) / 32u;
		}
	}

	const uint32_t incl = warp_inclusive_scan(groups, lane);
	if(lane == 31)
		s_warpTotal[warp] = incl;
	__syncthreads();

	if(WRITE)
	{
		uint32_t at = __ldg(blockOffsets + blockIdx.x) + incl - groups;
		for(uint32_t w = 0; w < warp; w++)
			at += s_warpTotal[w];
		for(uint32_t g = 0; g < groups; g++)
			requests[at + g] = (tile << 4) | g;
	}
	else if(threadIdx.x == 0)
	{
		uint32_t sum = 0;
		for(int w = 0; w < COMPACT_WARPS; w++)
			sum += s_warpTotal[w];
		blockCounts[blockIdx.x] = sum;
	}
}

/* exclusive scan of blockCounts[0..n) by one CTA of 1024 threads; total -> *grandTotal */
__global__ void __launch_bounds__(1024) dn_scan_blocks_kernel(const uint32_t* __restrict__ blockCounts, uint32_t* __restrict__ blockOffsets, uint32_t n, uint32_t* __restrict__ grandTotal,
                                                              volatile uint32_t* hostTotal)
{
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if(threadIdx.x == 0)
		s_carry = 0;
	__syncthreads();

	for(uint32_t base = 0; base < n; base += 1024)
	{
		const uint32_t i = base + threadIdx.x;
		const uint32_t v = i < n ? __ldg(blockCounts + i) : 0u;
		uint32_t incl = warp_inclusive_scan(v, lane);
		if(lane == 31)
			s_warp[warp] = incl;
		__syncthreads();
		if(warp == 0)
		{
			uint32_t w = s_warp[lane];
			uint32_t wi = warp_inclusive_scan(w, lane);
			s_warp[lane] = wi - w;
		}
		__syncthreads();
		const uint32_t carry = s_carry;
		if(i < n)
			blockOffsets[i] = carry + s_warp[warp] + incl - v;
		__syncthreads();
		if(threadIdx.x == 1023)
			s_carry = carry + s_warp[31] + incl;
		__syncthreads();
	}
	if(threadIdx.x == 0)
	{
		*grandTotal = s_carry;
		if(hostTotal)
		{
			/* straight into the host's pinned word (zero-copy store): a 4-byte cudaMemcpyAsync here queued behind the 33 MB
			 * framebuffer read-back on the copy engine and kept DN_sync_gpu waiting 0.5 ms per frame */
			*hostTotal = s_carry;
			__threadfence_system();
		}
	}
}

/* forced[] |= bit(tile) for a list of tiles whose chunk has pending edits (voxel.c:1470) */
__global__ void dn_set_bits_kernel(uint32_t* __restrict__ bits, const uint32_t* __restrict__ tiles, uint32_t n)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < n)
		atomicOr(bits + (tiles[i] >> 5), 1u << (tiles[i] & 31u));
}

/* dst[] |= src[] (merging a visible bitmap gathered from another GPU) */
__global__ void dn_or_bits_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t words)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i < words)
	{
		const uint32_t s = src[i];
		if(s)
			dst[i] |= s;
	}
}

extern "C" cudaError_t dnb_launch_or_bits(uint32_t* dst, const uint32_t* src, uint32_t words, cudaStream_t stream)
{
	if(words == 0)
		return cudaSuccess;
	{ DNB_LAUNCHED(1); dn_or_bits_kernel<<<(words + 255) / 256, 256, 0, stream>>>(dst, src, words); }
	return cudaGetLastError();
}

extern "C" uint32_t dnb_compact_num_blocks(uint32_t numTiles)
{
	return (numTiles + COMPACT_TILES_PER_CTA - 1) / COMPACT_TILES_PER_CTA;
}

extern "C" cudaError_t dnb_launch_compact_count(const DnbScene* scene, const uint32_t* forced, uint32_t split, uint32_t frameNum, uint32_t* blockCounts, uint32_t* blockOffsets,
                                                uint32_t* grandTotal, uint32_t* hostTotal, cudaStream_t stream)
{
	const uint32_t blocks = dnb_compact_num_blocks(scene->numTiles);
	{ DNB_LAUNCHED(1); dn_compact_kernel<false><<<blocks, COMPACT_WARPS * 32, 0, stream>>>(scene->visible, forced, scene->tileSlot, scene->slots, scene->numTiles, split, frameNum, blockCounts, nullptr, nullptr); }
	{ DNB_LAUNCHED(1); dn_scan_blocks_kernel<<<1, 1024, 0, stream>>>(blockCounts, blockOffsets, blocks, grandTotal, hostTotal); }
	return cudaGetLastError();
}

extern "C" cudaError_t dnb_launch_compact_write(const DnbScene* scene, const uint32_t* forced, uint32_t split, uint32_t frameNum, const uint32_t* blockOffsets, uint32_t* requests,
                                                cudaStream_t stream)
{
	const uint32_t blocks = dnb_compact_num_blocks(scene->numTiles);
	{ DNB_LAUNCHED(1); dn_compact_kernel<true><<<blocks, COMPACT_WARPS * 32, 0, stream>>>(scene->visible, forced, scene->tileSlot, scene->slots, scene->numTiles, split, frameNum, nullptr, blockOffsets, requests); }
	return cudaGetLastError();
}

extern "C" cudaError_t dnb_launch_set_bits(uint32_t* bits, const uint32_t* tiles, uint32_t n, cudaStream_t stream)
{
	if(n == 0)
		return cudaSuccess;
	{ DNB_LAUNCHED(1); dn_set_bits_kernel<<<(n + 255) / 256, 256, 0, stream>>>(bits, tiles, n); }
	return cudaGetLastError();
}
