/* scenegen.c -- native generators for the two synthetic maps of BASELINE.json whose full sizes are too large to build
 * through numpy: config 3 (sparse balls, 256^3 tiles) and config 5 (dense corridors, 128^3 tiles).
 *
 * Bench / test infrastructure, not part of the drop-in path: built as its own shared object (libdoon_scenes.so) and fed
 * to any engine through the bulk chunk call (DN_b200_set_chunks, or set_chunk of the oracle).  Bit-identical to
 * doonengine_b200/scenes.py sparse_balls / dense_corridors (tests/test_abi_host.py compares them): only 32-bit integer
 * hashing happens here; every float-derived quantity (ball radius table, packed normals) is computed by scenes.py and
 * passed in as a 512-entry table, so the two cannot diverge.
 *
 * A chunk is 512 x (normal word, albedo word) in the DNchunk order [x][y][z] (reference voxel.h:60-72).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t pcg_hash(uint32_t x)
{
	const uint32_t state = x * 747796405u + 2891336453u;
	const uint32_t shift = ((state >> 28) + 4u) & 31u;
	const uint32_t word = ((state >> shift) ^ state) * 277803737u;
	return (word >> 22) ^ word;
}

static inline uint32_t hash3(uint32_t x, uint32_t y, uint32_t z, uint32_t seed)
{
	return pcg_hash((x * 73856093u) ^ (y * 19349663u) ^ (z * 83492791u) ^ seed);
}

static inline uint32_t albedo_of(uint32_t hh)
{
	const uint32_t r = 32u + (hh & 0xFFu) % 209u, g = 32u + ((hh >> 8) & 0xFFu) % 209u, b = 32u + ((hh >> 16) & 0xFFu) % 209u;
	return (r << 24) | (g << 16) | (b << 8);
}

/* scenes.sparse_balls: tiles of slab z in [z0, z1), z-major then y then x.
 * dist[512]: distance of voxel [x][y][z] from the chunk centre (float32, from scenes.py); normal24[512]: its packed normal bytes.
 * pos: 3 x int32 per chunk, vox: 1024 x uint32 per chunk; both may be NULL to only count.  Returns the chunk count of the slab
 * (nothing is written past `cap` chunks). */
size_t dnscene_sparse_balls(const uint32_t tiles[3], uint32_t threshold, uint32_t seed, uint32_t z0, uint32_t z1, const float* dist, const uint32_t* normal24,
                            int32_t* pos, uint32_t* vox, size_t cap)
{
	const uint32_t tx = tiles[0], ty = tiles[1];
	if(z1 <= z0)
		return 0;
	const size_t rows = (size_t)(z1 - z0) * ty;
	size_t* rowStart = (size_t*)malloc((rows + 1) * sizeof(size_t));
	if(!rowStart)
		return 0;

#pragma omp parallel for schedule(static)
	for(long long r = 0; r < (long long)rows; r++)
	{
		const uint32_t cz = z0 + (uint32_t)(r / ty), cy = (uint32_t)(r % ty);
		size_t n = 0;
		for(uint32_t cx = 0; cx < tx; cx++)
			n += hash3(cx, cy, cz, seed) < threshold;
		rowStart[r + 1] = n;
	}
	rowStart[0] = 0;
	for(size_t r = 0; r < rows; r++)
		rowStart[r + 1] += rowStart[r];
	const size_t total = rowStart[rows];

	if(pos && vox)
	{
#pragma omp parallel for schedule(dynamic, 4)
		for(long long r = 0; r < (long long)rows; r++)
		{
			const uint32_t cz = z0 + (uint32_t)(r / ty), cy = (uint32_t)(r % ty);
			size_t at = rowStart[r];
			for(uint32_t cx = 0; cx < tx; cx++)
			{
				const uint32_t h = hash3(cx, cy, cz, seed);
				if(h >= threshold)
					continue;
				if(at >= cap)
					break;
				const uint32_t hv = pcg_hash(h);
				const float radius = 2.5f + 0.5f * (float)(hv & 3u);
				const uint32_t kind = (hv >> 2) % 10u;
				const uint32_t mat = kind < 7u ? 0u : (kind < 9u ? 3u : 2u);
				int32_t* p = pos + at * 3;
				p[0] = (int32_t)cx; p[1] = (int32_t)cy; p[2] = (int32_t)cz;
				uint32_t* v = vox + at * 1024;
				for(uint32_t k = 0; k < 512; k++)
				{
					v[2 * k] = dist[k] <= radius ? ((mat << 24) | normal24[k]) : 0xFFFFFFFFu;
					v[2 * k + 1] = albedo_of(hash3(k, hv & 0xFFFFu, 7u, seed));
				}
				at++;
			}
		}
	}
	free(rowStart);
	return total;
}

/* scenes.dense_corridors: tiles of slab z in [z0, z1); every voxel solid with `material`, except corridor tiles */
size_t dnscene_dense_corridors(const uint32_t tiles[3], uint32_t period, uint32_t material, uint32_t seed, uint32_t z0, uint32_t z1, const uint32_t* normal24,
                               int32_t* pos, uint32_t* vox, size_t cap)
{
	const uint32_t tx = tiles[0], ty = tiles[1];
	if(z1 <= z0 || period == 0)
		return 0;
	const size_t rows = (size_t)(z1 - z0) * ty;
	size_t* rowStart = (size_t*)malloc((rows + 1) * sizeof(size_t));
	if(!rowStart)
		return 0;
	rowStart[0] = 0;
	for(size_t r = 0; r < rows; r++)
	{
		const uint32_t cz = z0 + (uint32_t)(r / ty), cy = (uint32_t)(r % ty);
		const uint32_t yz = (cy % period == 1u) + (cz % period == 1u);
		size_t n = 0;
		for(uint32_t cx = 0; cx < tx; cx++)
			n += ((cx % period == 1u) + yz) < 2u;
		rowStart[r + 1] = rowStart[r] + n;
	}
	const size_t total = rowStart[rows];

	if(pos && vox)
	{
#pragma omp parallel for schedule(dynamic, 4)
		for(long long r = 0; r < (long long)rows; r++)
		{
			const uint32_t cz = z0 + (uint32_t)(r / ty), cy = (uint32_t)(r % ty);
			const uint32_t yz = (cy % period == 1u) + (cz % period == 1u);
			size_t at = rowStart[r];
			for(uint32_t cx = 0; cx < tx; cx++)
			{
				if(((cx % period == 1u) + yz) >= 2u)
					continue;
				if(at >= cap)
					break;
				int32_t* p = pos + at * 3;
				p[0] = (int32_t)cx; p[1] = (int32_t)cy; p[2] = (int32_t)cz;
				uint32_t* v = vox + at * 1024;
				for(uint32_t k = 0; k < 512; k++)
				{
					v[2 * k] = (material << 24) | normal24[k];
					v[2 * k + 1] = albedo_of(hash3(k, cx + 131u * cy, cz, seed));
				}
				at++;
			}
		}
	}
	free(rowStart);
	return total;
}
