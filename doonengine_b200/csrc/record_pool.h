/* record_pool.h -- allocator of the voxel-record pool: a buddy system over 512-record blocks (host only, no CUDA).
 *
 * Nodes are 16..512 records, the reference's node sizes (voxel.c:1557-1559).  The reference splits a larger free node when no
 * exact fit exists (voxel.c:1604-1630) and merges equal free neighbours while bubbling them to the end of the pool with
 * device-side copies (voxel.c:1642-1694, at most 10 per frame), all by O(nodes) scans of `gpuVoxelLayout`.  Here every node lives
 * inside an aligned 512-record block and has ONE possible merge partner, its buddy (start ^ size): releasing a node merges it
 * upwards for as long as the buddy is free too, acquiring one splits the smallest free node that is large enough.  Same pool
 * economy under streams of edits that shift chunks between size classes (freed small nodes become large ones again and vice
 * versa), no record ever moves, O(1) per operation.
 *
 * nodeFree[start / 16] = size class of the FREE node starting there, 0xFF otherwise.  The per-class free lists are lazy: an entry
 * counts only while nodeFree agrees with it, which makes taking a buddy out of the middle of a list free.
 */
#ifndef DN_B200_RECORD_POOL_H
#define DN_B200_RECORD_POOL_H

#include <stddef.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace dnb
{

struct RecordPool
{
	enum { NUM_CLASSES = 6 }; /* 16, 32, 64, 128, 256, 512 records */

	std::vector<uint32_t> freeLists[NUM_CLASSES];
	std::vector<uint8_t>  nodeFree;
	size_t top = 0;            /* records handed to the allocator so far, a multiple of 512 */
	size_t usedNodes = 0, freeNodeCount = 0;
	uint64_t splits = 0, merges = 0;

	static int size_class(uint32_t n)
	{
		int c = 0;
		uint32_t size = 16;
		while(size < n) { size <<= 1; c++; }
		return c;
	}

	void clear()
	{
		for(int c = 0; c < NUM_CLASSES; c++)
			freeLists[c].clear();
		nodeFree.clear();
		top = 0;
		usedNodes = freeNodeCount = 0;
	}

	bool is_free(uint32_t start, int cls) const
	{
		const size_t at = start >> 4;
		return at < nodeFree.size() && nodeFree[at] == (uint8_t)cls;
	}

	/* first record of a node of 16 << cls records */
	uint32_t acquire(int cls)
	{
		uint32_t start;
		int from = cls;
		while(from < NUM_CLASSES && !pop(from, &start))
			from++;
		if(from >= NUM_CLASSES)
		{
			/* nothing free is large enough: open a new 512-record block at the top of the pool */
			start = (uint32_t)top;
			top += (size_t)16 << (NUM_CLASSES - 1);
			from = NUM_CLASSES - 1;
		}
		while(from > cls)
		{
			/* split (voxel.c:1604-1630): the upper half stays free */
			from--;
			mark_free(start + (16u << from), from);
			splits++;
		}
		usedNodes++;
		return start;
	}

	void release(uint32_t start, int cls)
	{
		usedNodes--;
		/* merge with the buddy while it is free and of the same size (voxel.c:1678-1688 merges equal free neighbours) */
		while(cls < NUM_CLASSES - 1)
		{
			const uint32_t buddy = start ^ (16u << cls);
			if(!is_free(buddy, cls))
				break;
			nodeFree[buddy >> 4] = 0xFF; /* its list entry goes stale */
			freeNodeCount--;
			merges++;
			start = std::min(start, buddy);
			cls++;
		}
		mark_free(start, cls);
	}

private:
	void mark_free(uint32_t start, int cls)
	{
		const size_t at = start >> 4;
		if(at >= nodeFree.size())
			nodeFree.resize(std::max(at + 1, nodeFree.size() * 2), 0xFF);
		nodeFree[at] = (uint8_t)cls;
		freeLists[cls].push_back(start);
		freeNodeCount++;
	}

	bool pop(int cls, uint32_t* start)
	{
		std::vector<uint32_t>& list = freeLists[cls];
		while(!list.empty())
		{
			const uint32_t s = list.back();
			list.pop_back();
			if(is_free(s, cls))
			{
				nodeFree[s >> 4] = 0xFF;
				freeNodeCount--;
				*start = s;
				return true;
			}
		}
		return false;
	}
};

} // namespace dnb

#endif
