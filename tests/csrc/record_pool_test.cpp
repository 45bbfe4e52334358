/* host harness for csrc/record_pool.h (tests/test_host_hygiene.py): random acquire / release streams, checked against a
 * brute-force model -- live nodes never overlap, free + live nodes tile the pool exactly, releasing everything merges the pool
 * back into whole 512-record blocks, and a stream of edits that moves chunks between size classes does not grow the pool. */
#include "record_pool.h"
#include <stdio.h>
#include <stdlib.h>
#include <map>

using dnb::RecordPool;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd()
{
	rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
	return (uint32_t)(rng_state >> 32);
}

static int check_tiling(const RecordPool& p, const std::map<uint32_t, int>& live)
{
	/* every 16-record granule of [0, top) belongs to exactly one node, live or free */
	std::vector<uint8_t> owner(p.top / 16, 0);
	size_t freeSeen = 0;
	for(auto& kv : live)
		for(uint32_t g = kv.first / 16; g < (kv.first + (16u << kv.second)) / 16; g++)
		{
			if(g >= owner.size() || owner[g]) return 1;
			owner[g] = 1;
		}
	for(size_t at = 0; at < p.nodeFree.size(); at++)
		if(p.nodeFree[at] != 0xFF)
		{
			freeSeen++;
			if((at * 16) % (16u << p.nodeFree[at])) return 2; /* nodes are aligned to their size */
			for(size_t g = at; g < at + (1u << p.nodeFree[at]); g++)
			{
				if(g >= owner.size() || owner[g]) return 3;
				owner[g] = 2;
			}
		}
	for(uint8_t o : owner)
		if(!o) return 4;
	if(freeSeen != p.freeNodeCount || live.size() != p.usedNodes) return 5;
	return 0;
}

int main()
{
	RecordPool p;
	std::map<uint32_t, int> live;
	std::vector<uint32_t> starts;
	/* 1. random churn */
	for(int it = 0; it < 200000; it++)
	{
		if(live.empty() || (rnd() % 100 < 55 && live.size() < 5000))
		{
			const uint32_t n = 1 + rnd() % 512;
			const int cls = RecordPool::size_class(n);
			if((16u << cls) < n || (cls > 0 && (16u << (cls - 1)) >= n)) { printf("FAIL size_class(%u) = %d\n", n, cls); return 1; }
			const uint32_t s = p.acquire(cls);
			if(live.count(s)) { printf("FAIL node %u handed out twice\n", s); return 1; }
			live[s] = cls;
			starts.push_back(s);
		}
		else
		{
			const size_t k = rnd() % starts.size();
			const uint32_t s = starts[k];
			starts[k] = starts.back();
			starts.pop_back();
			p.release(s, live[s]);
			live.erase(s);
		}
		if(it % 5000 == 0)
		{
			const int e = check_tiling(p, live);
			if(e) { printf("FAIL tiling check %d at iteration %d\n", e, it); return 1; }
		}
	}
	const size_t topAfterChurn = p.top;
	/* 2. release everything: the pool must merge back into whole 512-record blocks */
	for(uint32_t s : starts)
		p.release(s, live[s]);
	live.clear();
	starts.clear();
	if(check_tiling(p, live)) { printf("FAIL tiling after full release\n"); return 1; }
	if(p.freeNodeCount != p.top / 512) { printf("FAIL %zu free nodes for %zu blocks: buddies did not merge\n", p.freeNodeCount, p.top / 512); return 1; }
	/* 3. class-shifting edit stream (the case an allocator without merge / split leaks on): a fresh pool, 2000 chunks, each edit
	 * moves one chunk to a random other size class.  The live footprint is bounded, so the pool must stop growing once it has
	 * warmed up, and stay within a small factor of what is live. */
	RecordPool q;
	std::vector<std::pair<uint32_t, int>> chunks;
	for(int i = 0; i < 2000; i++)
	{
		const int cls = rnd() % 6;
		chunks.push_back({q.acquire(cls), cls});
	}
	size_t topWarm = 0, liveMax = 0;
	for(int it = 0; it < 400000; it++)
	{
		auto& c = chunks[rnd() % chunks.size()];
		q.release(c.first, c.second);
		c.second = rnd() % 6;
		c.first = q.acquire(c.second);
		if(it == 100000)
			topWarm = q.top;
		if(it % 1000 == 0)
		{
			size_t liveNow = 0;
			for(auto& k : chunks)
				liveNow += 16u << k.second;
			liveMax = std::max(liveMax, liveNow);
		}
	}
	for(auto& c : chunks)
		live[c.first] = c.second;
	if(check_tiling(q, live)) { printf("FAIL tiling after the edit stream\n"); return 1; }
	if(q.top > topWarm + topWarm / 8) { printf("FAIL pool grew from %zu to %zu records under a bounded edit stream\n", topWarm, q.top); return 1; }
	if(q.top > 2 * liveMax) { printf("FAIL pool of %zu records for at most %zu live ones\n", q.top, liveMax); return 1; }
	if(q.merges == 0 || q.splits == 0) { printf("FAIL no merges / splits happened\n"); return 1; }
	p.splits += q.splits; p.merges += q.merges; p.top = q.top;
	printf("OK churn top %zu, edit stream top %zu (warm %zu, live max %zu), %llu splits, %llu merges\n", topAfterChurn, p.top, topWarm, liveMax, (unsigned long long)p.splits, (unsigned long long)p.merges);
	return 0;
}
