/* volume_host.cpp -- the CPU-side half of the DN_* API: volume lifecycle, the chunked host map, voxel edits,
 * the .voxvol file format and the small utilities.  Behaviour follows /root/reference/src/DoonEngine/voxel.c;
 * each function cites the lines it re-hosts.  The one structural change: every call that alters a tile records
 * it in a touched-tile list, so DN_sync_gpu reconciles only what changed instead of looping over the whole map
 * (voxel.c:738-762).
 */
#include "engine.h"
#include "hostmath.h"

#include <math.h>
#include <new>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

extern "C" void (*g_DN_message_callback)(DNmessageType, DNmessageSeverity, const char*) = nullptr;

namespace dnb
{

void report(DNmessageType type, DNmessageSeverity severity, const char* fmt, ...)
{
	if(!g_DN_message_callback)
		return;
	char text[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(text, sizeof(text), fmt, ap);
	va_end(ap);
	g_DN_message_callback(type, severity, text);
}

void touch_tile(VolumeImpl* v, size_t mapIndex)
{
	if(mapIndex >= v->touchedFlag.size() || v->touchedFlag[mapIndex])
		return;
	v->touchedFlag[mapIndex] = 1;
	v->touched.push_back((uint32_t)mapIndex);
}

/* voxel.c:1353-1363; numVoxelsGpu deliberately survives, as upstream */
static void clear_chunk(DNvolume* vol, size_t index)
{
	DNchunk* c = &vol->chunks[index];
	c->pos.x = c->pos.y = c->pos.z = -1;
	c->updated = false;
	c->numVoxels = 0;
	for(int x = 0; x < DN_CHUNK_SIZE; x++)
		for(int y = 0; y < DN_CHUNK_SIZE; y++)
			for(int z = 0; z < DN_CHUNK_SIZE; z++)
			{
				c->voxels[x][y][z].normal = UINT32_MAX;
				c->voxels[x][y][z].albedo = 0; /* (upstream leaves the albedo word of an empty voxel as malloc returned it; nobody reads it, but saved files and test comparisons should not depend on heap garbage) */
			}
}

} // namespace dnb

using namespace dnb;

/* ------------------------------------------------------------------------------------------------ */
/* lifecycle: voxel.c:165-295                                                                         */

extern "C" DNvolume* DN_create_volume(DNuvec3 mapSize, unsigned int minChunks)
{
	void* mem = DN_MALLOC(sizeof(VolumeImpl));
	if(!mem)
	{
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_FATAL, "failed to allocate memory for volume");
		return NULL;
	}
	VolumeImpl* v = new(mem) VolumeImpl();
	DNvolume* vol = &v->pub;
	memset(vol, 0, sizeof(DNvolume));
	memset(&v->stats, 0, sizeof(v->stats));
	v->magic = VOLUME_MAGIC;

	const size_t tiles = (size_t)mapSize.x * mapSize.y * mapSize.z;
	size_t numChunks = tiles < minChunks ? tiles : minChunks; /* voxel.c:181-182 */
	if(numChunks == 0)
		numChunks = 1;

	vol->mapSize = mapSize;
	vol->map = (DNchunkHandle*)DN_MALLOC(sizeof(DNchunkHandle) * (tiles ? tiles : 1));
	vol->chunks = (DNchunk*)DN_MALLOC(sizeof(DNchunk) * numChunks);
	vol->materials = (DNmaterial*)DN_MALLOC(sizeof(DNmaterial) * DN_MAX_MATERIALS);
	vol->lightingRequests = (GLuint*)DN_MALLOC(sizeof(GLuint) * numChunks);
	vol->gpuVoxelLayout = NULL; /* mirrored on demand, see DN_b200 docs */
	if(!vol->map || !vol->chunks || !vol->materials || !vol->lightingRequests)
	{
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_FATAL, "failed to allocate host memory for a %ux%ux%u map", mapSize.x, mapSize.y, mapSize.z);
		DN_delete_volume(vol);
		return NULL;
	}

	for(size_t i = 0; i < tiles; i++)
	{
		vol->map[i].flag = 0;
		vol->map[i].chunkIndex = 0;
	}
	vol->chunkCap = 0;
	for(size_t i = 0; i < numChunks; i++)
	{
		clear_chunk(vol, i);
		vol->chunks[i].numVoxelsGpu = 0;
	}
	memset(vol->materials, 0, sizeof(DNmaterial) * DN_MAX_MATERIALS);

	vol->chunkCap = numChunks;
	vol->nextChunk = 0;
	vol->voxelCap = DN_CHUNK_LENGTH * numChunks / 2; /* voxel.c:190 */
	vol->numVoxelNodes = 0;
	vol->numLightingRequests = 0;
	vol->lightingRequestCap = numChunks;

	/* defaults, voxel.c:261-278 */
	vol->camPos.x = vol->camPos.y = vol->camPos.z = 0.0f;
	vol->camOrient.x = vol->camOrient.y = vol->camOrient.z = 0.0f;
	vol->camFOV = 90.0f;
	vol->camViewMode = 0;
	vol->sunDir.x = vol->sunDir.y = vol->sunDir.z = 1.0f;
	vol->sunStrength.x = vol->sunStrength.y = vol->sunStrength.z = 0.6f;
	vol->ambientLightStrength.x = vol->ambientLightStrength.y = vol->ambientLightStrength.z = 0.01f;
	vol->diffuseBounceLimit = 5;
	vol->specBounceLimit = 2;
	vol->shadowSoftness = 10.0f;
	vol->useCubemap = false;
	vol->glCubemapTex = 0;
	vol->skyGradientBot.x = 0.71f; vol->skyGradientBot.y = 0.85f; vol->skyGradientBot.z = 0.90f;
	vol->skyGradientTop.x = 0.00f; vol->skyGradientTop.y = 0.45f; vol->skyGradientTop.z = 0.74f;
	vol->frameNum = 0;
	vol->lastTime = 123.456f;

	v->tileSlotHost.assign(tiles, 0u);
	v->touchedFlag.assign(tiles, 0);

	/* Without a successful DN_init the volume is HOST-ONLY: the map can be edited, loaded and saved (offline
	 * tooling), but DN_sync_gpu / DN_draw / DN_update_lighting refuse to run.  There is no CPU rendering path. */
	if(ctx().ready)
	{
		if(!device_create(v))
		{
			DN_delete_volume(vol);
			return NULL;
		}
	}
	else
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_NOTE, "DN_create_volume without DN_init: host-only volume, nothing can be drawn or lit");
	return vol;
}

extern "C" void DN_delete_volume(DNvolume* vol)
{
	if(!vol)
		return;
	VolumeImpl* v = impl_of(vol);
	device_destroy(v);
	DN_FREE(vol->map);
	DN_FREE(vol->chunks);
	DN_FREE(vol->materials);
	DN_FREE(vol->lightingRequests);
	DN_FREE(vol->gpuVoxelLayout);
	v->~VolumeImpl();
	DN_FREE(v);
}

/* ------------------------------------------------------------------------------------------------ */
/* .voxvol: per-chunk palette + run-length codec (voxel.c:301-518), file layout (voxel.c:520-654)      */

namespace
{

struct Rgb8 { uint8_t x, y, z; };

inline int find_in_palette(const Rgb8* pal, int n, Rgb8 item)
{
	for(int i = 0; i < n; i++)
		if(pal[i].x == item.x && pal[i].y == item.y && pal[i].z == item.z)
			return i;
	return -1;
}

inline DNivec3 local_pos(int i) { DNivec3 p = {i % DN_CHUNK_SIZE, (i / DN_CHUNK_SIZE) % DN_CHUNK_SIZE, i / (DN_CHUNK_SIZE * DN_CHUNK_SIZE)}; return p; }
inline Rgb8 normal_bytes(uint32_t w) { Rgb8 r = {(uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w}; return r; }
inline Rgb8 albedo_bytes(uint32_t w) { Rgb8 r = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8)}; return r; }

/* voxel.c:301-430; returns the encoded size */
uint16_t encode_chunk(const DNchunk* chunk, DNvolume* vol, uint8_t* out)
{
	uint8_t* p = out;
	memcpy(p, &chunk->pos, sizeof(DNivec3));
	p += sizeof(DNivec3);
	if(!DN_in_map_bounds(vol, chunk->pos))
		return (uint16_t)sizeof(DNivec3);

	/* palettes are used only while they stay below numVoxels/2 entries */
	const int limit = (int)(chunk->numVoxels / 2);
	int numNormal = 0, numAlbedo = 0;
	Rgb8 normalPal[DN_CHUNK_LENGTH / 2], albedoPal[DN_CHUNK_LENGTH / 2];
	for(int i = 0; i < DN_CHUNK_LENGTH; i++)
	{
		DNivec3 q = local_pos(i);
		const DNcompressedVoxel& vx = chunk->voxels[q.x][q.y][q.z];
		if((vx.normal >> 24) == DN_MATERIAL_EMPTY)
			continue;
		Rgb8 n = normal_bytes(vx.normal), a = albedo_bytes(vx.albedo);
		if(numNormal < limit && find_in_palette(normalPal, numNormal, n) < 0)
			normalPal[numNormal++] = n;
		if(numAlbedo < limit && find_in_palette(albedoPal, numAlbedo, a) < 0)
			albedoPal[numAlbedo++] = a;
	}
	if(numNormal >= limit)
		numNormal = 0;
	if(numAlbedo >= limit)
		numAlbedo = 0;

	*p++ = (uint8_t)numNormal;
	memcpy(p, normalPal, 3 * (size_t)numNormal);
	p += 3 * (size_t)numNormal;
	*p++ = (uint8_t)numAlbedo;
	memcpy(p, albedoPal, 3 * (size_t)numAlbedo);
	p += 3 * (size_t)numAlbedo;

	/* runs of equal material, at most 255 long; solid voxels carry normal+albedo (palette index or 3 raw bytes) */
	int i = 0;
	while(i < DN_CHUNK_LENGTH)
	{
		DNivec3 q = local_pos(i);
		const uint8_t material = (uint8_t)(chunk->voxels[q.x][q.y][q.z].normal >> 24);
		*p++ = material;
		uint8_t* runLength = p++;
		uint8_t num = 0;
		int j = i;
		for(; j < DN_CHUNK_LENGTH; j++)
		{
			DNivec3 q2 = local_pos(j);
			const DNcompressedVoxel& vx = chunk->voxels[q2.x][q2.y][q2.z];
			if(num >= UINT8_MAX || (uint8_t)(vx.normal >> 24) != material)
				break;
			num++;
			if(material == DN_MATERIAL_EMPTY)
				continue;
			Rgb8 n = normal_bytes(vx.normal), a = albedo_bytes(vx.albedo);
			if(numNormal > 0)
				*p++ = (uint8_t)find_in_palette(normalPal, numNormal, n);
			else { *p++ = n.x; *p++ = n.y; *p++ = n.z; }
			if(numAlbedo > 0)
				*p++ = (uint8_t)find_in_palette(albedoPal, numAlbedo, a);
			else { *p++ = a.x; *p++ = a.y; *p++ = a.z; }
		}
		*runLength = num;
		i = j;
	}
	return (uint16_t)(p - out);
}

/* voxel.c:433-518, with an end pointer: a malformed or truncated stream stops the decode (false) instead of reading past the
 * record (the reference trusts the file: voxel.c:547-553 reads a file-supplied size into a fixed buffer and walks it unchecked) */
bool decode_chunk(const uint8_t* in, const uint8_t* end, DNvolume* vol, DNchunk* chunk)
{
	chunk->updated = false;
	chunk->numVoxels = 0;
	chunk->numVoxelsGpu = 0;
	if((size_t)(end - in) < sizeof(DNivec3))
	{
		chunk->pos = {-1, -1, -1};
		return false;
	}
	memcpy(&chunk->pos, in, sizeof(DNivec3));
	in += sizeof(DNivec3);
	if(!DN_in_map_bounds(vol, chunk->pos))
		return true;

#define NEED(n) do { if((size_t)(end - in) < (size_t)(n)) return false; } while(0)
	NEED(1);
	const uint8_t numNormal = *in++;
	const uint8_t* normalPal = in;
	NEED(3 * (size_t)numNormal + 1);
	in += 3 * (size_t)numNormal;
	const uint8_t numAlbedo = *in++;
	const uint8_t* albedoPal = in;
	NEED(3 * (size_t)numAlbedo);
	in += 3 * (size_t)numAlbedo;

	int done = 0;
	while(done < DN_CHUNK_LENGTH)
	{
		NEED(2);
		const uint8_t material = *in++;
		const uint8_t num = *in++;
		for(int i = done; i < done + num && i < DN_CHUNK_LENGTH; i++)
		{
			DNivec3 q = local_pos(i);
			DNcompressedVoxel& vx = chunk->voxels[q.x][q.y][q.z];
			if(material == DN_MATERIAL_EMPTY)
			{
				vx.normal = UINT32_MAX;
				continue;
			}
			const uint8_t* n;
			if(numNormal > 0)
			{
				NEED(1);
				const uint8_t k = *in++;
				if(k >= numNormal)
					return false; /* palette index past the palette */
				n = normalPal + 3 * (size_t)k;
			}
			else { NEED(3); n = in; in += 3; }
			const uint8_t* a;
			if(numAlbedo > 0)
			{
				NEED(1);
				const uint8_t k = *in++;
				if(k >= numAlbedo)
					return false;
				a = albedoPal + 3 * (size_t)k;
			}
			else { NEED(3); a = in; in += 3; }
			vx.normal = ((uint32_t)material << 24) | ((uint32_t)n[0] << 16) | ((uint32_t)n[1] << 8) | n[2];
			vx.albedo = ((uint32_t)a[0] << 24) | ((uint32_t)a[1] << 16) | ((uint32_t)a[2] << 8);
			chunk->numVoxels++;
		}
		if(num == 0)
			return false; /* malformed stream: a zero-length run would never terminate */
		done += num;
	}
#undef NEED
	return true;
}

} // namespace

extern "C" DNvolume* DN_load_volume(const char* filePath, unsigned int minChunks)
{
	FILE* f = fopen(filePath, "rb");
	if(!f)
	{
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "failed to open file \"%s\" for reading", filePath);
		return NULL;
	}

	DNuvec3 mapSize;
	uint64_t chunkCap = 0;
	if(fread(&mapSize, sizeof(DNuvec3), 1, f) != 1 || fread(&chunkCap, sizeof(uint64_t), 1, f) != 1)
	{
		fclose(f);
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "file \"%s\" is truncated", filePath);
		return NULL;
	}
	/* the header is untrusted: every chunk record takes at least 2 bytes of the file, and the map must be addressable by a
	 * 28-bit tile index (voxelLighting.comp:9) -- refuse before allocating anything from these numbers */
	long fileBytes = -1;
	{
		const long here = ftell(f);
		if(here >= 0 && fseek(f, 0, SEEK_END) == 0)
		{
			fileBytes = ftell(f);
			fseek(f, here, SEEK_SET);
		}
	}
	const unsigned long long tiles = (unsigned long long)mapSize.x * mapSize.y * mapSize.z;
	if(mapSize.x == 0 || mapSize.y == 0 || mapSize.z == 0 || mapSize.x > (1u << 20) || mapSize.y > (1u << 20) || mapSize.z > (1u << 20) || tiles > (1ull << 28) ||
	   (fileBytes >= 0 && chunkCap > (uint64_t)fileBytes / 2))
	{
		fclose(f);
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "file \"%s\" has an implausible header (map %ux%ux%u, %llu chunk records, %ld bytes)", filePath, mapSize.x, mapSize.y, mapSize.z,
		       (unsigned long long)chunkCap, fileBytes);
		return NULL;
	}
	DNvolume* vol = DN_create_volume(mapSize, minChunks);
	if(!vol || !DN_set_max_chunks(vol, (size_t)chunkCap))
	{
		fclose(f);
		if(vol)
			DN_delete_volume(vol);
		return NULL;
	}
	VolumeImpl* v = impl_of(vol);

	std::vector<uint8_t> buf(UINT16_MAX); /* a record's size field is 16 bits: whatever the file says fits */
	bool ok = true, malformed = false;
	for(size_t i = 0; i < (size_t)chunkCap && ok; i++)
	{
		uint16_t size;
		ok = fread(&size, sizeof(uint16_t), 1, f) == 1 && fread(buf.data(), 1, size, f) == size;
		if(!ok)
			break;
		if(!decode_chunk(buf.data(), buf.data() + size, vol, &vol->chunks[i]))
		{
			/* a damaged record: drop the chunk rather than keep half of it */
			malformed = true;
			vol->chunks[i].pos = {-1, -1, -1};
			vol->chunks[i].numVoxels = 0;
			continue;
		}
		if(DN_in_map_bounds(vol, vol->chunks[i].pos))
		{
			const size_t mapIndex = DN_FLATTEN_INDEX(vol->chunks[i].pos, mapSize);
			vol->map[mapIndex].flag = 1;
			vol->map[mapIndex].chunkIndex = (uint32_t)i;
			touch_tile(v, mapIndex);
		}
	}
	if(malformed)
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "file \"%s\" holds malformed chunk records; those chunks were dropped", filePath);

	ok = ok && fread(vol->materials, sizeof(DNmaterial), DN_MAX_MATERIALS, f) == DN_MAX_MATERIALS;
	ok = ok && fread(&vol->camPos, sizeof(DNvec3), 1, f) == 1 && fread(&vol->camOrient, sizeof(DNvec3), 1, f) == 1;
	ok = ok && fread(&vol->camFOV, sizeof(float), 1, f) == 1 && fread(&vol->camViewMode, sizeof(uint32_t), 1, f) == 1;
	ok = ok && fread(&vol->sunDir, sizeof(DNvec3), 1, f) == 1 && fread(&vol->sunStrength, sizeof(DNvec3), 1, f) == 1;
	ok = ok && fread(&vol->ambientLightStrength, sizeof(DNvec3), 1, f) == 1;
	ok = ok && fread(&vol->diffuseBounceLimit, sizeof(uint32_t), 1, f) == 1 && fread(&vol->specBounceLimit, sizeof(uint32_t), 1, f) == 1;
	ok = ok && fread(&vol->shadowSoftness, sizeof(float), 1, f) == 1;
	ok = ok && fread(&vol->skyGradientBot, sizeof(DNvec3), 1, f) == 1 && fread(&vol->skyGradientTop, sizeof(DNvec3), 1, f) == 1;
	fclose(f);
	if(!ok)
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "file \"%s\" is truncated; loaded what was there", filePath);
	return vol;
}

extern "C" bool DN_save_volume(const char* filePath, DNvolume* vol)
{
	FILE* f = fopen(filePath, "wb");
	if(!f)
	{
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "failed to open file \"%s\" for writing", filePath);
		return false;
	}
	const uint64_t chunkCap = vol->chunkCap;
	fwrite(&vol->mapSize, sizeof(DNuvec3), 1, f);
	fwrite(&chunkCap, sizeof(uint64_t), 1, f);

	std::vector<uint8_t> buf(sizeof(DNchunk) * 2);
	for(size_t i = 0; i < vol->chunkCap; i++)
	{
		const uint16_t size = encode_chunk(&vol->chunks[i], vol, buf.data());
		fwrite(&size, sizeof(uint16_t), 1, f);
		fwrite(buf.data(), 1, size, f);
	}

	fwrite(vol->materials, sizeof(DNmaterial), DN_MAX_MATERIALS, f);
	fwrite(&vol->camPos, sizeof(DNvec3), 1, f);
	fwrite(&vol->camOrient, sizeof(DNvec3), 1, f);
	fwrite(&vol->camFOV, sizeof(float), 1, f);
	fwrite(&vol->camViewMode, sizeof(uint32_t), 1, f);
	fwrite(&vol->sunDir, sizeof(DNvec3), 1, f);
	fwrite(&vol->sunStrength, sizeof(DNvec3), 1, f);
	fwrite(&vol->ambientLightStrength, sizeof(DNvec3), 1, f);
	fwrite(&vol->diffuseBounceLimit, sizeof(uint32_t), 1, f);
	fwrite(&vol->specBounceLimit, sizeof(uint32_t), 1, f);
	fwrite(&vol->shadowSoftness, sizeof(float), 1, f);
	fwrite(&vol->skyGradientBot, sizeof(DNvec3), 1, f);
	const bool ok = fwrite(&vol->skyGradientTop, sizeof(DNvec3), 1, f) == 1;
	fclose(f);
	return ok;
}

/* ------------------------------------------------------------------------------------------------ */
/* chunk slots of the host map: voxel.c:659-714                                                       */

extern "C" int DN_add_chunk(DNvolume* vol, DNivec3 pos)
{
	VolumeImpl* v = impl_of(vol);
	const size_t mapIndex = DN_FLATTEN_INDEX(pos, vol->mapSize);

	/* circular search for a free slot starting at nextChunk */
	size_t i = vol->nextChunk < vol->chunkCap ? vol->nextChunk : 0;
	const size_t start = i;
	bool found = false;
	do
	{
		if(!DN_in_map_bounds(vol, vol->chunks[i].pos))
		{
			found = true;
			break;
		}
		if(++i >= vol->chunkCap)
			i = 0;
	} while(i != start);

	if(!found)
	{
		size_t newCap = vol->chunkCap * 2;
		const size_t tiles = num_tiles(vol);
		if(newCap > tiles)
			newCap = tiles;
		if(newCap <= vol->chunkCap)
			newCap = vol->chunkCap + 1;
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_NOTE, "automatically resizing chunk memory to accomodate %zu chunks (%zu bytes)", newCap, newCap * sizeof(DNchunk));
		i = vol->chunkCap;
		if(!DN_set_max_chunks(vol, newCap))
			return -1;
	}

	vol->map[mapIndex].chunkIndex = (uint32_t)i;
	vol->map[mapIndex].flag = 1;
	vol->chunks[i].pos = pos;
	vol->nextChunk = (i == vol->chunkCap - 1) ? 0 : i + 1;
	touch_tile(v, mapIndex);
	return (int)i;
}

extern "C" void DN_remove_chunk(DNvolume* vol, DNivec3 pos)
{
	VolumeImpl* v = impl_of(vol);
	const size_t mapIndex = DN_FLATTEN_INDEX(pos, vol->mapSize);
	vol->map[mapIndex].flag = 0;
	vol->nextChunk = vol->map[mapIndex].chunkIndex;
	clear_chunk(vol, vol->map[mapIndex].chunkIndex);
	touch_tile(v, mapIndex);
}

/* ------------------------------------------------------------------------------------------------ */
/* capacities: voxel.c:957-1096                                                                       */

extern "C" bool DN_set_max_chunks(DNvolume* vol, size_t num)
{
	if(num == 0)
		num = 1;
	DNchunk* grown = (DNchunk*)DN_REALLOC(vol->chunks, sizeof(DNchunk) * num);
	if(!grown)
	{
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_ERROR, "failed to reallocate memory for chunks");
		return false;
	}
	vol->chunks = grown;
	for(size_t i = vol->chunkCap; i < num; i++)
	{
		clear_chunk(vol, i);
		vol->chunks[i].numVoxelsGpu = 0;
	}
	vol->chunkCap = num;
	if(vol->nextChunk >= num)
		vol->nextChunk = 0;
	return true;
}

extern "C" bool DN_set_max_lighting_requests(DNvolume* vol, size_t num)
{
	GLuint* grown = (GLuint*)DN_REALLOC(vol->lightingRequests, sizeof(GLuint) * (num ? num : 1));
	if(!grown)
	{
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_ERROR, "failed to reallocate memory for lighting requests");
		return false;
	}
	vol->lightingRequests = grown;
	vol->lightingRequestCap = num;
	return true;
}

extern "C" bool DN_set_map_size(DNvolume* vol, DNuvec3 size)
{
	VolumeImpl* v = impl_of(vol);
	const size_t newTiles = (size_t)size.x * size.y * size.z;
	DNchunkHandle* newMap = (DNchunkHandle*)DN_MALLOC(sizeof(DNchunkHandle) * (newTiles ? newTiles : 1));
	if(!newMap)
	{
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_ERROR, "failed to reallocate memory for map");
		return false;
	}

	/* tiles inside both the old and the new box keep their chunk (voxel.c:968-978) */
	for(uint32_t z = 0; z < size.z; z++)
		for(uint32_t y = 0; y < size.y; y++)
			for(uint32_t x = 0; x < size.x; x++)
			{
				DNivec3 pos = {(int32_t)x, (int32_t)y, (int32_t)z};
				const size_t newIndex = DN_FLATTEN_INDEX(pos, size);
				if(DN_in_map_bounds(vol, pos))
					newMap[newIndex] = vol->map[DN_FLATTEN_INDEX(pos, vol->mapSize)];
				else
				{
					newMap[newIndex].flag = 0;
					newMap[newIndex].chunkIndex = 0;
				}
			}
	DN_FREE(vol->map);
	vol->map = newMap;
	vol->mapSize = size;

	/* chunks that fell outside are dropped (voxel.c:985-987) */
	for(size_t i = 0; i < vol->chunkCap; i++)
		if(!DN_in_map_bounds(vol, vol->chunks[i].pos))
			clear_chunk(vol, i);

	/* the device copy is rebuilt from scratch at the next writing sync; accumulated lighting restarts, as it
	 * does upstream where the chunk buffer is re-created uninitialised (voxel.c:1000-1002) */
	device_destroy(v);
	v->tileSlotHost.assign(newTiles, 0u);
	v->touchedFlag.assign(newTiles, 0);
	v->touched.clear();
	v->freeSlots.clear();
	v->slotTop = 0;
	v->slotNodeStart.clear();
	v->slotNodeClass.clear();
	v->slotNumVoxels.clear();
	v->pool.clear();
	v->slotTile.clear();
	v->residentGroups = 0;
	v->stats.residentChunks = 0;
	v->stats.residentRecords = 0;
	DN_FREE(vol->gpuVoxelLayout);
	vol->gpuVoxelLayout = NULL;
	for(int a = 0; a < 3; a++)
	{
		v->occMin[a] = 0x3FFFFFFF;
		v->occMax[a] = -0x3FFFFFFF;
	}
	vol->numVoxelNodes = 0;
	vol->numLightingRequests = 0;
	v->requestsValid = 0;
	if(ctx().ready && !device_create(v))
		return false;
	for(size_t i = 0; i < newTiles; i++)
		if(vol->map[i].flag != 0)
			touch_tile(v, i);
	return true;
}

/* ------------------------------------------------------------------------------------------------ */
/* voxel access: voxel.c:1101-1193                                                                    */

extern "C" bool DN_in_map_bounds(DNvolume* vol, DNivec3 pos)
{
	return pos.x >= 0 && pos.y >= 0 && pos.z >= 0 && (uint32_t)pos.x < vol->mapSize.x && (uint32_t)pos.y < vol->mapSize.y && (uint32_t)pos.z < vol->mapSize.z;
}

extern "C" bool DN_in_chunk_bounds(DNivec3 pos)
{
	return pos.x < DN_CHUNK_SIZE && pos.y < DN_CHUNK_SIZE && pos.z < DN_CHUNK_SIZE && pos.x >= 0 && pos.y >= 0 && pos.z >= 0;
}

extern "C" DNcompressedVoxel DN_get_compressed_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos)
{
	return vol->chunks[vol->map[DN_FLATTEN_INDEX(mapPos, vol->mapSize)].chunkIndex].voxels[chunkPos.x][chunkPos.y][chunkPos.z];
}

extern "C" DNvoxel DN_get_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos)
{
	return DN_decompress_voxel(DN_get_compressed_voxel(vol, mapPos, chunkPos));
}

extern "C" void DN_set_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos, DNvoxel voxel)
{
	DN_set_compressed_voxel(vol, mapPos, chunkPos, DN_compress_voxel(voxel));
}

extern "C" void DN_set_compressed_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos, DNcompressedVoxel voxel)
{
	VolumeImpl* v = impl_of(vol);
	const size_t mapIndex = DN_FLATTEN_INDEX(mapPos, vol->mapSize);
	const uint32_t newMat = voxel.normal >> 24;

	if(vol->map[mapIndex].flag == 0)
	{
		if(newMat == DN_MATERIAL_EMPTY) /* empty voxel into an empty tile: nothing to do */
			return;
		if(DN_add_chunk(vol, mapPos) < 0)
			return;
	}

	DNchunk* chunk = &vol->chunks[vol->map[mapIndex].chunkIndex];
	DNcompressedVoxel* dst = &chunk->voxels[chunkPos.x][chunkPos.y][chunkPos.z];
	const uint32_t oldMat = dst->normal >> 24;

	if(oldMat == DN_MATERIAL_EMPTY && newMat != DN_MATERIAL_EMPTY)
		chunk->numVoxels++;
	else if(oldMat != DN_MATERIAL_EMPTY && newMat == DN_MATERIAL_EMPTY)
	{
		chunk->numVoxels--;
		if(chunk->numVoxels == 0) /* last voxel gone: the chunk is released, the voxel itself is not rewritten */
		{
			DN_remove_chunk(vol, mapPos);
			return;
		}
	}

	*dst = voxel;
	chunk->updated = true;
	touch_tile(v, mapIndex);
}

extern "C" void DN_remove_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos)
{
	VolumeImpl* v = impl_of(vol);
	const size_t mapIndex = DN_FLATTEN_INDEX(mapPos, vol->mapSize);
	if(vol->map[mapIndex].flag == 0) /* upstream dereferences a stale chunk index here; there is nothing to remove */
		return;
	DNchunk* chunk = &vol->chunks[vol->map[mapIndex].chunkIndex];
	DNcompressedVoxel* dst = &chunk->voxels[chunkPos.x][chunkPos.y][chunkPos.z];
	if((dst->normal >> 24) != DN_MATERIAL_EMPTY)
	{
		chunk->numVoxels--;
		if(chunk->numVoxels == 0)
		{
			DN_remove_chunk(vol, mapPos);
			return;
		}
	}
	dst->normal = UINT32_MAX;
	chunk->updated = true;
	touch_tile(v, mapIndex);
}

extern "C" bool DN_does_chunk_exist(DNvolume* vol, DNivec3 pos)
{
	return vol->map[DN_FLATTEN_INDEX(pos, vol->mapSize)].flag >= 1;
}

extern "C" bool DN_does_voxel_exist(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos)
{
	return (DN_get_compressed_voxel(vol, mapPos, chunkPos).normal >> 24) != DN_MATERIAL_EMPTY;
}

/* CPU picking ray over the host map: single-level voxel DDA with strict `<` axis choice (voxel.c:1195-1272) */
extern "C" bool DN_step_map(DNvolume* vol, DNvec3 rayDir, DNvec3 rayPos, int maxSteps, DNivec3* hitPos, DNvoxel* hitVoxel, DNivec3* hitNormal)
{
	hitNormal->x = hitNormal->y = hitNormal->z = -1000;

	float p[3] = {rayPos.x * DN_CHUNK_SIZE, rayPos.y * DN_CHUNK_SIZE, rayPos.z * DN_CHUNK_SIZE};
	const float d[3] = {rayDir.x, rayDir.y, rayDir.z};
	int cell[3], step[3];
	float delta[3], side[3];
	for(int a = 0; a < 3; a++)
	{
		const float inv = 1 / d[a];
		const int sg = d[a] > 0 ? 1 : (d[a] < 0 ? -1 : 0);
		cell[a] = (int)floor(p[a]);
		delta[a] = fabsf(inv);
		step[a] = sg;
		side[a] = (float)((sg * (cell[a] - p[a]) + (sg * 0.5) + 0.5) * delta[a]);
	}

	for(int n = 0; n < maxSteps; n++)
	{
		DNivec3 pos = {cell[0], cell[1], cell[2]}, mapPos, chunkPos;
		DN_separate_position(pos, &mapPos, &chunkPos);
		if(pos.x < 0) mapPos.x--;
		if(pos.y < 0) mapPos.y--;
		if(pos.z < 0) mapPos.z--;

		if(DN_in_map_bounds(vol, mapPos) && DN_does_chunk_exist(vol, mapPos) && DN_does_voxel_exist(vol, mapPos, chunkPos))
		{
			*hitVoxel = DN_get_voxel(vol, mapPos, chunkPos);
			*hitPos = pos;
			return true;
		}

		int axis;
		if(side[0] < side[1])
			axis = side[0] < side[2] ? 0 : 2;
		else
			axis = side[1] < side[2] ? 1 : 2;
		side[axis] += delta[axis];
		cell[axis] += step[axis];
		hitNormal->x = hitNormal->y = hitNormal->z = 0;
		(&hitNormal->x)[axis] = -step[axis];
	}
	return false;
}

/* ------------------------------------------------------------------------------------------------ */
/* utility: voxel.c:1277-1316, matrices voxel.c:788-810                                               */

extern "C" void DN_separate_position(DNivec3 pos, DNivec3* mapPos, DNivec3* chunkPos)
{
	mapPos->x = pos.x / DN_CHUNK_SIZE; mapPos->y = pos.y / DN_CHUNK_SIZE; mapPos->z = pos.z / DN_CHUNK_SIZE;
	chunkPos->x = pos.x % DN_CHUNK_SIZE; chunkPos->y = pos.y % DN_CHUNK_SIZE; chunkPos->z = pos.z % DN_CHUNK_SIZE;
}

extern "C" DNvec3 DN_cam_dir(DNvec3 orient)
{
	float R[3][3], out[3];
	const float o[3] = {orient.x, orient.y, orient.z}, fwd[3] = {0.0f, 0.0f, 1.0f};
	euler3x3(o, R);
	mat3_mul_vec3(R, fwd, out);
	DNvec3 r;
	r.x = out[0]; r.y = out[1]; r.z = out[2];
	return r;
}

extern "C" DNcompressedVoxel DN_compress_voxel(DNvoxel voxel)
{
	DNcompressedVoxel res;
	uint32_t n[3];
	const float in[3] = {voxel.normal.x, voxel.normal.y, voxel.normal.z};
	for(int i = 0; i < 3; i++)
	{
		float c = in[i] < 1.0f ? in[i] : 1.0f;
		c = c > -1.0f ? c : -1.0f;
		n[i] = (uint32_t)(((int)(c * 255.0f) + 255) / 2);
	}
	res.normal = ((uint32_t)voxel.material << 24) | (n[0] << 16) | (n[1] << 8) | n[2];
	res.albedo = ((uint32_t)voxel.albedo.r << 24) | ((uint32_t)voxel.albedo.g << 16) | ((uint32_t)voxel.albedo.b << 8);
	return res;
}

extern "C" DNvoxel DN_decompress_voxel(DNcompressedVoxel voxel)
{
	DNvoxel res;
	const float inv255 = 1.0f / 255.0f;
	const int nx = (int)((voxel.normal >> 16) & 0xFF) * 2 - 255, ny = (int)((voxel.normal >> 8) & 0xFF) * 2 - 255, nz = (int)(voxel.normal & 0xFF) * 2 - 255;
	res.normal.x = (float)nx * inv255;
	res.normal.y = (float)ny * inv255;
	res.normal.z = (float)nz * inv255;
	res.material = (uint8_t)(voxel.normal >> 24);
	res.albedo.r = (uint8_t)(voxel.albedo >> 24);
	res.albedo.g = (uint8_t)(voxel.albedo >> 16);
	res.albedo.b = (uint8_t)(voxel.albedo >> 8);
	return res;
}

extern "C" void DN_set_view_projection_matrices(DNvolume* vol, float aspectRatio, float nearPlane, float farPlane, DNmat4* view, DNmat4* projection)
{
	float R[3][3];
	const float orient[3] = {vol->camOrient.x, vol->camOrient.y, vol->camOrient.z};
	euler3x3(orient, R);

	/* the camera looks along +z rotated by the orientation; the length of `front` does not matter to lookat */
	float f;
	if(aspectRatio < 1.0f)
		f = aspectRatio / tanf(deg2rad(vol->camFOV * 0.5f));
	else
		f = 1.0f / tanf(deg2rad(vol->camFOV * 0.5f));
	const float fwd[3] = {0.0f, 0.0f, f};
	float front[3];
	mat3_mul_vec3(R, fwd, front);

	const float pos[3] = {vol->camPos.x, vol->camPos.y, vol->camPos.z};
	const float target[3] = {pos[0] + front[0], pos[1] + front[1], pos[2] + front[2]};
	const Mat4 V = lookat(pos, target);
	const Mat4 P = perspective(vol->camFOV, 1.0f / aspectRatio, nearPlane, farPlane);
	memcpy(view->m, V.m, sizeof(V.m));
	memcpy(projection->m, P.m, sizeof(P.m));
}

/* `count` DN_set_compressed_voxel calls in one: voxel positions in VOXEL units (split like DN_separate_position); a voxel whose
 * material is DN_MATERIAL_EMPTY is removed (DN_remove_voxel).  Positions outside the map are skipped.  Returns the edits applied. */
extern "C" size_t DN_b200_set_voxels(DNvolume* vol, size_t count, const DNivec3* positions, const DNcompressedVoxel* voxels)
{
	/* The edits of a stream land all over a map of hundreds of megabytes: every one of them misses the cache three times in a row
	 * (tile handle -> chunk header -> voxel).  The list is known in advance, so the handle of edit i + 16 and the chunk lines of edit
	 * i + 8 are prefetched while edit i is applied (prefetches are hints: a chunk array that moves in between costs nothing). */
	const size_t AHEAD = 16;
	auto tile_of = [&](size_t i, DNivec3* chunkPos) -> long long
	{
		const DNivec3 p = positions[i];
		if(p.x < 0 || p.y < 0 || p.z < 0)
			return -1;
		DNivec3 mapPos;
		DN_separate_position(p, &mapPos, chunkPos);
		if(!DN_in_map_bounds(vol, mapPos))
			return -1;
		return (long long)DN_FLATTEN_INDEX(mapPos, vol->mapSize);
	};
	size_t applied = 0;
	for(size_t i = 0; i < count; i++)
	{
		DNivec3 cp;
		if(i + AHEAD < count)
		{
			const long long t = tile_of(i + AHEAD, &cp);
			if(t >= 0)
				__builtin_prefetch(&vol->map[t], 0, 1);
		}
		if(i + AHEAD / 2 < count)
		{
			const long long t = tile_of(i + AHEAD / 2, &cp);
			if(t >= 0 && vol->map[t].flag != 0 && vol->map[t].chunkIndex < vol->chunkCap)
			{
				const DNchunk* ch = &vol->chunks[vol->map[t].chunkIndex];
				__builtin_prefetch(ch, 1, 1);
				__builtin_prefetch(&ch->voxels[cp.x][cp.y][cp.z], 1, 1);
			}
		}
		DNivec3 mapPos, chunkPos;
		if(positions[i].x < 0 || positions[i].y < 0 || positions[i].z < 0)
			continue;
		DN_separate_position(positions[i], &mapPos, &chunkPos);
		if(!DN_in_map_bounds(vol, mapPos))
			continue;
		if((voxels[i].normal >> 24) == DN_MATERIAL_EMPTY)
			DN_remove_voxel(vol, mapPos, chunkPos);
		else
			DN_set_compressed_voxel(vol, mapPos, chunkPos, voxels[i]);
		applied++;
	}
	return applied;
}

/* `count` whole chunks in one call: the bulk form of 512 DN_set_compressed_voxel calls per chunk (voxel.c:1126-1160), as the map
 * loader does it (voxel.c:560-640).  voxels = count x [8][8][8] DNcompressedVoxel in the chunk's own [x][y][z] order; a chunk without a
 * single solid voxel removes the tile's chunk.  Tiles outside the map are skipped.  Returns the number of chunks now present. */
extern "C" size_t DN_b200_set_chunks(DNvolume* vol, size_t count, const DNivec3* mapPositions, const DNcompressedVoxel* voxels)
{
	size_t present = 0;
	for(size_t i = 0; i < count; i++)
	{
		const DNivec3 mp = mapPositions[i];
		if(!DN_in_map_bounds(vol, mp))
			continue;
		const DNcompressedVoxel* src = voxels + i * 512;
		uint32_t solid = 0;
		for(int k = 0; k < 512; k++)
			solid += (src[k].normal >> 24) != DN_MATERIAL_EMPTY;
		const size_t mapIndex = DN_FLATTEN_INDEX(mp, vol->mapSize);
		if(solid == 0)
		{
			if(vol->map[mapIndex].flag != 0)
				DN_remove_chunk(vol, mp);
			continue;
		}
		if(vol->map[mapIndex].flag == 0 && DN_add_chunk(vol, mp) < 0)
			break;
		DNchunk* chunk = &vol->chunks[vol->map[mapIndex].chunkIndex];
		DNcompressedVoxel* dst = &chunk->voxels[0][0][0];
		for(int k = 0; k < 512; k++)
		{
			dst[k] = src[k];
			if((src[k].normal >> 24) == DN_MATERIAL_EMPTY)
				dst[k].normal = UINT32_MAX; /* what DN_remove_voxel leaves behind (voxel.c:1178) */
		}
		chunk->numVoxels = solid;
		chunk->updated = true;
		touch_tile(impl_of(vol), mapIndex);
		present++;
	}
	return present;
}

extern "C" void DN_b200_touch_tile(DNvolume* vol, DNivec3 mapPos)
{
	if(DN_in_map_bounds(vol, mapPos))
		touch_tile(impl_of(vol), DN_FLATTEN_INDEX(mapPos, vol->mapSize));
}

extern "C" void DN_b200_rescan(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	const size_t tiles = num_tiles(vol);
	for(size_t i = 0; i < tiles; i++)
		if(vol->map[i].flag != 0 || v->tileSlotHost[i] != 0)
			touch_tile(v, i);
}
