/* upload.cu -- applies a batch of edited chunks to the device-resident map.
 *
 * The reference re-uploads an edited chunk with two glBufferSubData calls per chunk (header voxel.c:1548-1552,
 * records voxel.c:1636-1637) from inside its all-tiles loop.  Here the host packs every dirty chunk of a
 * DN_sync_gpu call into ONE pinned blob (item table + 128-byte slot headers + 16-byte records), sends it with a
 * single cudaMemcpyAsync on the upload stream, and this kernel scatters it: one warp per chunk, the header as one
 * coalesced 128-byte store, the records as coalesced 16-byte-per-lane stores.  It also maintains the traversal
 * structures the reference keeps in map[i].flags: tileSlot, the 4x4x4 occupancy word, and the visible bit
 * (cleared on (re)upload and on removal, as `flags = 2` / `flags = 0` do at voxel.c:1525 / :1504).
 * Pure streaming: bytes moved = 16 + 128 + 16 n per chunk, bound by PCIe for the copy and HBM for the scatter.
 */
#include "kernels.h"

__global__ void __launch_bounds__(256) dn_scatter_chunks_kernel(const DnbUploadItem* __restrict__ items, const DnbSlot* __restrict__ headers, const uint4* __restrict__ blobRecords, uint32_t numItems,
                                                                uint32_t sx, uint32_t sy, uint32_t bx, uint32_t by,
                                                                uint32_t* __restrict__ tileSlot, unsigned long long* __restrict__ occ64, uint32_t* __restrict__ visible,
                                                                DnbSlot* __restrict__ slots, uint4* __restrict__ records)
{
	const uint32_t i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
	if(i >= numItems)
		return;

	const DnbUploadItem item = items[i];
	const uint32_t tile = item.mapIndex;
	const uint32_t x = tile % sx, y = (tile / sx) % sy, z = tile / (sx * sy);
	const uint32_t block = (x >> 2) + bx * ((y >> 2) + by * (z >> 2));
	const unsigned long long occBit = 1ull << ((x & 3u) | ((y & 3u) << 2) | ((z & 3u) << 4));

	if(item.slotPlus1 != 0u)
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(headers + i);
		const uint32_t word = src[lane];
		reinterpret_cast<uint32_t*>(slots + (item.slotPlus1 - 1u))[lane] = word;

		const uint32_t base = headers[i].voxelBase, n = headers[i].numVoxels;
		for(uint32_t k = lane; k < n; k += 32)
			records[base + k] = blobRecords[item.recordOffset + k];

		if(lane == 0)
		{
			tileSlot[tile] = item.slotPlus1;
			atomicOr(occ64 + block, occBit);
			atomicAnd(visible + (tile >> 5), ~(1u << (tile & 31u)));
		}
	}
	else if(lane == 0)
	{
		tileSlot[tile] = 0u;
		atomicAnd(occ64 + block, ~occBit);
		atomicAnd(visible + (tile >> 5), ~(1u << (tile & 31u)));
	}
}

extern "C" cudaError_t dnb_launch_scatter(const DnbUploadItem* items, const DnbSlot* headers, const uint4* blobRecords, uint32_t numItems, const uint32_t mapSize[3], const uint32_t blocks[3],
                                          uint32_t* tileSlot, unsigned long long* occ64, uint32_t* visible, DnbSlot* slots, uint4* records, cudaStream_t stream)
{
	if(numItems == 0)
		return cudaSuccess;
	{ DNB_LAUNCHED(1); dn_scatter_chunks_kernel<<<(numItems + 7) / 8, 256, 0, stream>>>(items, headers, blobRecords, numItems, mapSize[0], mapSize[1], blocks[0], blocks[1], tileSlot, occ64, visible, slots, records); }
	return cudaGetLastError();
}

/* the slots' "every material of this chunk is opaque" flag against a new material table (layout.h DNB_BBOX_OPAQUE): one thread per slot */
struct DnbOpaqueBits { uint32_t w[8]; };
__global__ void __launch_bounds__(256) dn_refresh_opaque_kernel(DnbSlot* __restrict__ slots, uint32_t numSlots, DnbOpaqueBits bits)
{
	const uint32_t i = blockIdx.x * 256 + threadIdx.x;
	if(i >= numSlots)
		return;
	const uint32_t bbox = slots[i].bbox, ids = slots[i].matIds;
	bool all = !(bbox & DNB_BBOX_MIXED);
#pragma unroll
	for(int k = 0; k < 4; k++)
	{
		const uint32_t m = (ids >> (8 * k)) & 0xFFu;
		if(m != 0xFFu)
		{
			/* static indexing of the parameter words */
			uint32_t word = 0;
#pragma unroll
			for(int w = 0; w < 8; w++)
				if((m >> 5) == (uint32_t)w)
					word = bits.w[w];
			all = all && ((word >> (m & 31u)) & 1u);
		}
	}
	const uint32_t fresh = (bbox & ~DNB_BBOX_OPAQUE) | (all ? DNB_BBOX_OPAQUE : 0u);
	if(fresh != bbox)
		slots[i].bbox = fresh;
}

extern "C" cudaError_t dnb_launch_refresh_opaque(DnbSlot* slots, uint32_t numSlots, const uint32_t opaqueBits[8], cudaStream_t stream)
{
	if(numSlots == 0)
		return cudaSuccess;
	DnbOpaqueBits b;
	for(int w = 0; w < 8; w++)
		b.w[w] = opaqueBits[w];
	{ DNB_LAUNCHED(1); dn_refresh_opaque_kernel<<<(numSlots + 255) / 256, 256, 0, stream>>>(slots, numSlots, b); }
	return cudaGetLastError();
}
