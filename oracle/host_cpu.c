/* host_cpu.c -- TEST INFRASTRUCTURE ONLY (host half of the oracle).
 *
 * A CPU restatement of the HOST side of the reference's hot path, so the
 * oracle is complete without /root/reference (which does not exist on the GPU
 * box).  V.c = /root/reference/src/DoonEngine/voxel.c,
 * QM = /root/reference/dependencies/include/QuickMath/quickmath.h.
 *
 *   orh_load_voxvol      V.c:433-518 (chunk codec), V.c:520-593 (file layout)
 *   orh_set_voxel        V.c:659-714, V.c:1126-1161
 *   orh_pack_chunk       V.c:1391-1461 (surface culling, bit mask, partial counts, albedo linearisation)
 *   orh_sync             V.c:719-786, V.c:1463-1536 driven in RESIDENT MODE:
 *                        a tile that reaches state 1 is requested at once, i.e. the
 *                        "a ray touched it" step (voxelShared.comp:462-466) is assumed.
 *   orh_view_projection  V.c:788-810 with QM:1214-1243, 1264-1283, 1301-1333
 *   orh_draw             V.c:812-881 (uniforms) -> orb_draw
 *   orh_update_lighting  V.c:883-952 (uniforms) -> orb_light
 *
 * Pinned against Oracle A (oracle/_ref: the reference's own voxel.c behind the
 * fake-GL shim) by tests/test_oracle_vs_reference.py: chunk headers, record
 * bytes, request list and order, matrices and uniforms must be identical.
 *
 * One deliberate difference: the voxel-pool allocator.  The reference places
 * records with an O(nodes) LRU scan + bubble compaction (V.c:1554-1694); the
 * oracle uses a bump allocator with the same power-of-two node sizes.  Record
 * CONTENTS and their order inside a chunk are identical; only the base
 * `voxelIndex` differs, so comparisons are made relative to it.
 */
#include "oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* the reference's host chunk, voxel.h:56-64 (4120 bytes, voxels[x][y][z]) */
typedef struct OrhChunk
{
	int32_t  pos[3];
	uint8_t  updated;
	uint32_t numVoxels;
	uint32_t numVoxelsGpu;
	uint32_t voxels[8][8][8][2]; /* [..][0] = normal word, [..][1] = albedo word (voxel.h:49-53) */
} OrhChunk;

typedef struct OrhVolume
{
	uint32_t mapSize[3];

	/* CPU side (voxel.h:66-71, 114-116) */
	uint8_t*  cpuFlag;
	uint32_t* cpuChunkIndex;
	OrhChunk* chunks;
	size_t    chunkCap;
	size_t    nextChunk;
	OrbMaterial materials[256];

	/* public parameters (voxel.h:120-141) */
	float    camPos[3], camOrient[3], camFOV;
	uint32_t camViewMode;
	float    sunDir[3], sunStrength[3], ambientLightStrength[3];
	uint32_t diffuseBounceLimit, specBounceLimit;
	float    shadowSoftness;
	float    skyGradientBot[3], skyGradientTop[3];
	uint32_t frameNum;
	float    lastTime;

	/* "GPU" side in the reference layout */
	OrbHandle* map;
	OrbChunk*  gchunks;
	OrbVoxel*  voxels;
	size_t     voxelCap, voxelTop;
	uint32_t*  requests;
	size_t     numRequests, requestCap;

	OrbCounters drawCounters, lightCounters;
	size_t uploadBytes;
} OrhVolume;

static size_t num_tiles(const OrhVolume* v) { return (size_t)v->mapSize[0] * v->mapSize[1] * v->mapSize[2]; }

static int orh_in_map_bounds(const OrhVolume* v, const int32_t p[3])
{
	return p[0] >= 0 && p[1] >= 0 && p[2] >= 0 && (uint32_t)p[0] < v->mapSize[0] && (uint32_t)p[1] < v->mapSize[1] && (uint32_t)p[2] < v->mapSize[2];
}

/* V.c:1353-1363 */
static void clear_chunk(OrhChunk* c)
{
	c->pos[0] = c->pos[1] = c->pos[2] = -1;
	c->updated = 0;
	c->numVoxels = 0;
	for(int x = 0; x < 8; x++)
		for(int y = 0; y < 8; y++)
			for(int z = 0; z < 8; z++)
				c->voxels[x][y][z][0] = UINT32_MAX;
}

/* V.c:1012-1029 */
static int set_max_chunks(OrhVolume* v, size_t num)
{
	OrhChunk* n = (OrhChunk*)realloc(v->chunks, sizeof(OrhChunk) * num);
	if(!n)
		return 0;
	v->chunks = n;
	for(size_t i = v->chunkCap; i < num; i++)
	{
		clear_chunk(&v->chunks[i]);
		v->chunks[i].numVoxelsGpu = 0;
	}
	v->chunkCap = num;
	return 1;
}

/* V.c:165-281 (defaults V.c:261-278) */
OrhVolume* orh_create(uint32_t sx, uint32_t sy, uint32_t sz, uint32_t minChunks)
{
	OrhVolume* v = (OrhVolume*)calloc(1, sizeof(OrhVolume));
	v->mapSize[0] = sx; v->mapSize[1] = sy; v->mapSize[2] = sz;
	size_t tiles = num_tiles(v);
	size_t numChunks = tiles < minChunks ? tiles : minChunks;
	if(numChunks == 0)
		numChunks = 1;

	v->cpuFlag = (uint8_t*)calloc(tiles, 1);
	v->cpuChunkIndex = (uint32_t*)calloc(tiles, sizeof(uint32_t));
	set_max_chunks(v, numChunks);

	v->map = (OrbHandle*)calloc(tiles, sizeof(OrbHandle));
	v->gchunks = (OrbChunk*)calloc(tiles, sizeof(OrbChunk));
	v->voxelCap = 512 * numChunks / 2;
	if(v->voxelCap < 512)
		v->voxelCap = 512;
	v->voxels = (OrbVoxel*)calloc(v->voxelCap, sizeof(OrbVoxel));
	v->requestCap = 1024;
	v->requests = (uint32_t*)malloc(sizeof(uint32_t) * v->requestCap);

	v->camFOV = 90.0f;
	v->sunDir[0] = v->sunDir[1] = v->sunDir[2] = 1.0f;
	v->sunStrength[0] = v->sunStrength[1] = v->sunStrength[2] = 0.6f;
	v->ambientLightStrength[0] = v->ambientLightStrength[1] = v->ambientLightStrength[2] = 0.01f;
	v->diffuseBounceLimit = 5;
	v->specBounceLimit = 2;
	v->shadowSoftness = 10.0f;
	v->skyGradientBot[0] = 0.71f; v->skyGradientBot[1] = 0.85f; v->skyGradientBot[2] = 0.90f;
	v->skyGradientTop[0] = 0.00f; v->skyGradientTop[1] = 0.45f; v->skyGradientTop[2] = 0.74f;
	v->frameNum = 0;
	v->lastTime = 123.456f;
	return v;
}

void orh_destroy(OrhVolume* v)
{
	if(!v)
		return;
	free(v->cpuFlag); free(v->cpuChunkIndex); free(v->chunks);
	free(v->map); free(v->gchunks); free(v->voxels); free(v->requests);
	free(v);
}

/* V.c:659-706 */
static int add_chunk(OrhVolume* v, const int32_t pos[3])
{
	size_t mapIndex = (size_t)pos[0] + v->mapSize[0] * ((size_t)pos[1] + (size_t)pos[2] * v->mapSize[1]);
	size_t i = v->nextChunk;
	do
	{
		if(!orh_in_map_bounds(v, v->chunks[i].pos))
		{
			v->cpuChunkIndex[mapIndex] = (uint32_t)i;
			v->cpuFlag[mapIndex] = 1;
			memcpy(v->chunks[i].pos, pos, sizeof(int32_t) * 3);
			v->nextChunk = (i == v->chunkCap - 1) ? 0 : i + 1;
			return (int)i;
		}
		i++;
		if(i >= v->chunkCap)
			i = 0;
	} while(i != v->nextChunk);

	size_t newCap = v->chunkCap * 2;
	if(newCap > num_tiles(v))
		newCap = num_tiles(v);
	i = v->chunkCap;
	if(!set_max_chunks(v, newCap))
		return 0;
	v->cpuChunkIndex[mapIndex] = (uint32_t)i;
	v->cpuFlag[mapIndex] = 1;
	memcpy(v->chunks[i].pos, pos, sizeof(int32_t) * 3);
	v->nextChunk = (i == v->chunkCap - 1) ? 0 : i + 1;
	return (int)i;
}

/* V.c:708-714 */
static void remove_chunk(OrhVolume* v, size_t mapIndex)
{
	v->cpuFlag[mapIndex] = 0;
	v->nextChunk = v->cpuChunkIndex[mapIndex];
	clear_chunk(&v->chunks[v->cpuChunkIndex[mapIndex]]);
}

/* V.c:1126-1161: set one compressed voxel (material 255 in the top byte of `normal` = empty) */
void orh_set_voxel(OrhVolume* v, const int32_t mapPos[3], const int32_t chunkPos[3], uint32_t normal, uint32_t albedo)
{
	size_t mapIndex = (size_t)mapPos[0] + v->mapSize[0] * ((size_t)mapPos[1] + (size_t)mapPos[2] * v->mapSize[1]);
	if(v->cpuFlag[mapIndex] == 0)
	{
		if((normal >> 24) == 255)
			return;
		add_chunk(v, mapPos);
	}

	OrhChunk* c = &v->chunks[v->cpuChunkIndex[mapIndex]];
	uint32_t oldMat = c->voxels[chunkPos[0]][chunkPos[1]][chunkPos[2]][0] >> 24;
	uint32_t newMat = normal >> 24;

	if(oldMat == 255 && newMat != 255)
		c->numVoxels++;
	else if(oldMat != 255 && newMat == 255)
	{
		c->numVoxels--;
		if(c->numVoxels <= 0)
		{
			remove_chunk(v, mapIndex);
			return;
		}
	}

	c->voxels[chunkPos[0]][chunkPos[1]][chunkPos[2]][0] = normal;
	c->voxels[chunkPos[0]][chunkPos[1]][chunkPos[2]][1] = albedo;
	c->updated = 1;
}

/* bulk form of 512 orh_set_voxel calls: voxels[x][y][z][2] */
void orh_set_chunk(OrhVolume* v, const int32_t mapPos[3], const uint32_t* voxels)
{
	for(int x = 0; x < 8; x++)
		for(int y = 0; y < 8; y++)
			for(int z = 0; z < 8; z++)
			{
				int32_t cp[3] = {x, y, z};
				const uint32_t* w = voxels + ((x * 8 + y) * 8 + z) * 2;
				orh_set_voxel(v, mapPos, cp, w[0], w[1]);
			}
}

/* V.c:1289-1301 */
void orh_compress_voxel(uint8_t material, const float normal[3], const uint8_t albedo[3], uint32_t out[2])
{
	uint32_t n[3];
	for(int i = 0; i < 3; i++)
	{
		float c = normal[i];
		c = c < 1.0f ? c : 1.0f;
		c = c > -1.0f ? c : -1.0f;
		n[i] = (uint32_t)(((int)(c * 255.0f) + 255) / 2);
	}
	out[0] = ((uint32_t)material << 24) | (n[0] << 16) | (n[1] << 8) | n[2];
	out[1] = ((uint32_t)albedo[0] << 24) | ((uint32_t)albedo[1] << 16) | ((uint32_t)albedo[2] << 8);
}

/* ------------------------------------------------------------------ */
/* .voxvol reader: V.c:433-518 and V.c:520-593                          */

static void decompress_chunk(const uint8_t* mem, const OrhVolume* v, OrhChunk* chunk)
{
	memcpy(chunk->pos, mem, 12);
	mem += 12;
	chunk->updated = 0;
	chunk->numVoxels = 0;
	chunk->numVoxelsGpu = 0;
	if(!orh_in_map_bounds(v, chunk->pos))
		return;

	uint8_t numNormal = *mem++;
	const uint8_t* normalPalette = mem;
	mem += 3 * (size_t)numNormal;
	uint8_t numAlbedo = *mem++;
	const uint8_t* albedoPalette = mem;
	mem += 3 * (size_t)numAlbedo;

	int read = 0;
	while(read < 512)
	{
		uint8_t material = *mem++;
		uint8_t num = *mem++;
		for(int i = read; i < read + num; i++)
		{
			int x = i % 8, y = (i / 8) % 8, z = i / 64;
			if(material == 255)
			{
				chunk->voxels[x][y][z][0] = UINT32_MAX;
				continue;
			}
			const uint8_t* n;
			if(numNormal > 0) { n = normalPalette + 3 * (size_t)(*mem++); }
			else              { n = mem; mem += 3; }
			const uint8_t* a;
			if(numAlbedo > 0) { a = albedoPalette + 3 * (size_t)(*mem++); }
			else              { a = mem; mem += 3; }

			chunk->voxels[x][y][z][0] = ((uint32_t)material << 24) | ((uint32_t)n[0] << 16) | ((uint32_t)n[1] << 8) | n[2];
			chunk->voxels[x][y][z][1] = ((uint32_t)a[0] << 24) | ((uint32_t)a[1] << 16) | ((uint32_t)a[2] << 8);
			chunk->numVoxels++;
		}
		read += num;
	}
}

OrhVolume* orh_load_voxvol(const char* path, uint32_t minChunks)
{
	FILE* f = fopen(path, "rb");
	if(!f)
		return NULL;

	uint32_t mapSize[3];
	uint64_t chunkCap;
	if(fread(mapSize, 4, 3, f) != 3 || fread(&chunkCap, 8, 1, f) != 1)
	{
		fclose(f);
		return NULL;
	}
	OrhVolume* v = orh_create(mapSize[0], mapSize[1], mapSize[2], minChunks);
	set_max_chunks(v, (size_t)chunkCap);

	uint8_t* buf = (uint8_t*)malloc(sizeof(OrhChunk) * 2);
	for(size_t i = 0; i < chunkCap; i++)
	{
		uint16_t size;
		if(fread(&size, 2, 1, f) != 1 || fread(buf, 1, size, f) != size)
			break;
		decompress_chunk(buf, v, &v->chunks[i]);
		if(orh_in_map_bounds(v, v->chunks[i].pos))
		{
			const int32_t* p = v->chunks[i].pos;
			size_t mapIndex = (size_t)p[0] + v->mapSize[0] * ((size_t)p[1] + (size_t)p[2] * v->mapSize[1]);
			v->cpuFlag[mapIndex] = 1;
			v->cpuChunkIndex[mapIndex] = (uint32_t)i;
		}
	}
	free(buf);

	size_t ok = fread(v->materials, sizeof(OrbMaterial), 256, f);
	ok += fread(v->camPos, 4, 3, f);
	ok += fread(v->camOrient, 4, 3, f);
	ok += fread(&v->camFOV, 4, 1, f);
	ok += fread(&v->camViewMode, 4, 1, f);
	ok += fread(v->sunDir, 4, 3, f);
	ok += fread(v->sunStrength, 4, 3, f);
	ok += fread(v->ambientLightStrength, 4, 3, f);
	ok += fread(&v->diffuseBounceLimit, 4, 1, f);
	ok += fread(&v->specBounceLimit, 4, 1, f);
	ok += fread(&v->shadowSoftness, 4, 1, f);
	ok += fread(v->skyGradientBot, 4, 3, f);
	ok += fread(v->skyGradientTop, 4, 3, f);
	fclose(f);
	(void)ok;
	return v;
}

/* ------------------------------------------------------------------ */
/* packing: V.c:1391-1461                                               */

static int face_visible(const OrhVolume* v, const OrhChunk* c, int x, int y, int z)
{
	if(x < 0 || y < 0 || z < 0 || x >= 8 || y >= 8 || z >= 8)
		return 1;
	uint32_t mat = c->voxels[x][y][z][0] >> 24;
	return mat == 255 || v->materials[mat].opacity < 1.0f;
}

int orh_pack_chunk(const OrhVolume* v, const OrhChunk* c, OrbChunk* out, OrbVoxel* records)
{
	memset(out, 0, sizeof(*out));
	memcpy(out->pos, c->pos, 12);
	out->numIndirectSamples = 0;

	int n = 0;
	for(int z = 0; z < 8; z++)
		for(int y = 0; y < 8; y++)
			for(int x = 0; x < 8; x++)
			{
				unsigned index = (unsigned)(x + 8 * (y + 8 * z));
				if((index & 31) == 0 && index != 0 && ((index >> 5) & 3) == 0)
					out->partialCounts[(index >> 7) - 1] = (uint32_t)n;

				if((c->voxels[x][y][z][0] >> 24) == 255)
					continue;

				int visible = face_visible(v, c, x + 1, y, z) || face_visible(v, c, x - 1, y, z) ||
				              face_visible(v, c, x, y + 1, z) || face_visible(v, c, x, y - 1, z) ||
				              face_visible(v, c, x, y, z + 1) || face_visible(v, c, x, y, z - 1);
				if(!visible)
					continue;

				out->bitMask[index >> 5] |= 1u << (index & 31);

				/* linearise albedo: trunc(255 * (a * 0.00392156862)^2.2), V.c:1438-1448 */
				uint32_t a = c->voxels[x][y][z][1];
				uint8_t rgb[3] = {(uint8_t)(a >> 24), (uint8_t)(a >> 16), (uint8_t)(a >> 8)};
				uint32_t lin[3];
				for(int k = 0; k < 3; k++)
				{
					float f = (float)rgb[k] * 0.00392156862f;
					f = powf(f, 2.2f);
					f = f * 255.0f;
					lin[k] = (uint8_t)f;
				}

				records[n].normal = c->voxels[x][y][z][0];
				records[n].albedo = (lin[0] << 24) | (lin[1] << 16) | (lin[2] << 8);
				records[n].specLight = 0;
				records[n].diffuseLight = 0;
				n++;
			}
	return n;
}

/* ------------------------------------------------------------------ */
/* sync: V.c:719-786 + V.c:1463-1536, resident mode                     */

static size_t alloc_records(OrhVolume* v, int n)
{
	size_t node = 16;
	while(node < (size_t)n)
		node *= 2;
	if(v->voxelTop + node > v->voxelCap)
	{
		size_t cap = v->voxelCap;
		while(v->voxelTop + node > cap)
			cap *= 2;
		v->voxels = (OrbVoxel*)realloc(v->voxels, cap * sizeof(OrbVoxel));
		memset(v->voxels + v->voxelCap, 0, (cap - v->voxelCap) * sizeof(OrbVoxel));
		v->voxelCap = cap;
	}
	size_t at = v->voxelTop;
	v->voxelTop += node;
	return at;
}

/* op: 0 = DN_READ, 1 = DN_WRITE, 2 = DN_READ_WRITE (voxel.h:145-150) */
void orh_sync(OrhVolume* v, int op, int lightingSplit)
{
	v->frameNum++;
	if(v->frameNum >= (uint32_t)lightingSplit)
		v->frameNum = 0;

	v->numRequests = 0;

	for(uint32_t z = 0; z < v->mapSize[2]; z++)
	for(uint32_t y = 0; y < v->mapSize[1]; y++)
	for(uint32_t x = 0; x < v->mapSize[0]; x++)
	{
		size_t mapIndex = (size_t)x + v->mapSize[0] * ((size_t)y + (size_t)z * v->mapSize[1]);
		OrbHandle* h = &v->map[mapIndex];
		int gpuFlag = (int)(h->flags & 3);
		int gpuVisible = (h->flags & 4) > 0;
		h->lastUsed++;

		/* V.c:1463-1489 */
		if(op != 1 && gpuFlag == 2 && gpuVisible)
		{
			const OrhChunk* c = &v->chunks[v->cpuChunkIndex[mapIndex]];
			if(!(mapIndex % (size_t)lightingSplit != v->frameNum && !c->updated))
			{
				if(v->numRequests + 16 >= v->requestCap)
				{
					v->requestCap *= 2;
					v->requests = (uint32_t*)realloc(v->requests, sizeof(uint32_t) * v->requestCap);
				}
				for(uint32_t i = 0; i < c->numVoxelsGpu; i += 32)
					v->requests[v->numRequests++] = ((uint32_t)mapIndex << 4) | (i / 32);
			}
		}

		/* V.c:1491-1536 */
		if(op != 0)
		{
			if(v->cpuFlag[mapIndex] != 0 && gpuFlag == 0)
			{
				h->flags = 1;
				gpuFlag = 1;
			}
			else if(v->cpuFlag[mapIndex] == 0 && gpuFlag != 0)
			{
				h->flags = 0;
				gpuFlag = 0;
			}

			if(gpuFlag == 2 && v->chunks[v->cpuChunkIndex[mapIndex]].updated)
			{
				h->flags = 3;
				gpuFlag = 3;
			}

			/* resident mode: an unloaded tile is requested immediately */
			if(gpuFlag == 1)
			{
				h->flags = 3;
				gpuFlag = 3;
			}

			if(gpuFlag == 3 && v->cpuFlag[mapIndex] != 0)
			{
				OrhChunk* c = &v->chunks[v->cpuChunkIndex[mapIndex]];
				OrbVoxel records[512];
				OrbChunk header;
				int n = orh_pack_chunk(v, c, &header, records);
				c->numVoxelsGpu = (uint32_t)n;

				h->flags = 2;
				h->lastUsed = 0;
				v->gchunks[mapIndex] = header;
				h->voxelIndex = (uint32_t)alloc_records(v, n);
				memcpy(v->voxels + h->voxelIndex, records, sizeof(OrbVoxel) * (size_t)n);
				v->uploadBytes += sizeof(OrbChunk) + sizeof(OrbVoxel) * (size_t)n;
			}

			if(v->cpuFlag[mapIndex] != 0)
				v->chunks[v->cpuChunkIndex[mapIndex]].updated = 0;
		}
	}
}

/* ------------------------------------------------------------------ */
/* matrices: QM:1037-1124, 1214-1243, 1264-1283, 1301-1333              */

typedef struct { float m[4][4]; } M4;

static M4 m4_identity(void)
{
	M4 r;
	memset(&r, 0, sizeof(r));
	r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
	return r;
}

static M4 m4_mult(M4 a, M4 b)
{
	M4 r;
	for(int c = 0; c < 4; c++)
		for(int row = 0; row < 4; row++)
			r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2] + a.m[3][row] * b.m[c][3];
	return r;
}

static void v3_normalize(float v[3])
{
	float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
	if(len != 0.0f)
	{
		float inv = 1.0f / len;
		v[0] *= inv; v[1] *= inv; v[2] *= inv;
	}
	else
		v[0] = v[1] = v[2] = 0.0f;
}

static void v3_cross(const float a[3], const float b[3], float out[3])
{
	out[0] = (a[1] * b[2]) - (a[2] * b[1]);
	out[1] = (a[2] * b[0]) - (a[0] * b[2]);
	out[2] = (a[0] * b[1]) - (a[1] * b[0]);
}

static M4 m4_inv(M4 mat)
{
	M4 r;
	float t[6];
	float a = mat.m[0][0], b = mat.m[0][1], c = mat.m[0][2], d = mat.m[0][3],
	      e = mat.m[1][0], f = mat.m[1][1], g = mat.m[1][2], h = mat.m[1][3],
	      i = mat.m[2][0], j = mat.m[2][1], k = mat.m[2][2], l = mat.m[2][3],
	      m = mat.m[3][0], n = mat.m[3][1], o = mat.m[3][2], p = mat.m[3][3];

	t[0] = k * p - o * l; t[1] = j * p - n * l; t[2] = j * o - n * k;
	t[3] = i * p - m * l; t[4] = i * o - m * k; t[5] = i * n - m * j;
	r.m[0][0] =   f * t[0] - g * t[1] + h * t[2];
	r.m[1][0] = -(e * t[0] - g * t[3] + h * t[4]);
	r.m[2][0] =   e * t[1] - f * t[3] + h * t[5];
	r.m[3][0] = -(e * t[2] - f * t[4] + g * t[5]);
	r.m[0][1] = -(b * t[0] - c * t[1] + d * t[2]);
	r.m[1][1] =   a * t[0] - c * t[3] + d * t[4];
	r.m[2][1] = -(a * t[1] - b * t[3] + d * t[5]);
	r.m[3][1] =   a * t[2] - b * t[4] + c * t[5];

	t[0] = g * p - o * h; t[1] = f * p - n * h; t[2] = f * o - n * g;
	t[3] = e * p - m * h; t[4] = e * o - m * g; t[5] = e * n - m * f;
	r.m[0][2] =   b * t[0] - c * t[1] + d * t[2];
	r.m[1][2] = -(a * t[0] - c * t[3] + d * t[4]);
	r.m[2][2] =   a * t[1] - b * t[3] + d * t[5];
	r.m[3][2] = -(a * t[2] - b * t[4] + c * t[5]);

	t[0] = g * l - k * h; t[1] = f * l - j * h; t[2] = f * k - j * g;
	t[3] = e * l - i * h; t[4] = e * k - i * g; t[5] = e * j - i * f;
	r.m[0][3] = -(b * t[0] - c * t[1] + d * t[2]);
	r.m[1][3] =   a * t[0] - c * t[3] + d * t[4];
	r.m[2][3] = -(a * t[1] - b * t[3] + d * t[5]);
	r.m[3][3] =   a * t[2] - b * t[4] + c * t[5];

	float det = 1.0f / (a * r.m[0][0] + b * r.m[1][0] + c * r.m[2][0] + d * r.m[3][0]);
	for(int cc = 0; cc < 4; cc++)
		for(int rr = 0; rr < 4; rr++)
			r.m[cc][rr] = r.m[cc][rr] * det;
	return r;
}

/* V.c:788-810 */
void orh_view_projection(const OrhVolume* v, float aspectRatio, float nearPlane, float farPlane, float* view, float* projection)
{
	const float d2r = 0.01745329251f;
	float rx = v->camOrient[0] * d2r, ry = v->camOrient[1] * d2r, rz = v->camOrient[2] * d2r;
	float sinX = sinf(rx), cosX = cosf(rx), sinY = sinf(ry), cosY = cosf(ry), sinZ = sinf(rz), cosZ = cosf(rz);

	/* QM:1214-1243, top-left 3x3 only */
	float R[3][3];
	R[0][0] = cosY * cosZ;
	R[0][1] = cosY * sinZ;
	R[0][2] = -sinY;
	R[1][0] = sinX * sinY * cosZ - cosX * sinZ;
	R[1][1] = sinX * sinY * sinZ + cosX * cosZ;
	R[1][2] = sinX * cosY;
	R[2][0] = cosX * sinY * cosZ + sinX * sinZ;
	R[2][1] = cosX * sinY * sinZ - sinX * cosZ;
	R[2][2] = cosX * cosY;

	float fz;
	if(aspectRatio < 1.0f)
		fz = aspectRatio / tanf((v->camFOV * 0.5f) * d2r);
	else
		fz = 1.0f / tanf((v->camFOV * 0.5f) * d2r);

	float front[3];
	front[0] = R[0][0] * 0.0f + R[1][0] * 0.0f + R[2][0] * fz;
	front[1] = R[0][1] * 0.0f + R[1][1] * 0.0f + R[2][1] * fz;
	front[2] = R[0][2] * 0.0f + R[1][2] * 0.0f + R[2][2] * fz;

	/* lookat(pos, pos + front, up): QM:1301-1333 */
	float target[3] = {v->camPos[0] + front[0], v->camPos[1] + front[1], v->camPos[2] + front[2]};
	float dir[3] = {v->camPos[0] - target[0], v->camPos[1] - target[1], v->camPos[2] - target[2]};
	v3_normalize(dir);
	float up[3] = {0.0f, 1.0f, 0.0f};
	float r[3], u[3];
	v3_cross(up, dir, r);
	v3_normalize(r);
	v3_cross(dir, r, u);

	M4 RUD = m4_identity();
	RUD.m[0][0] = r[0]; RUD.m[1][0] = r[1]; RUD.m[2][0] = r[2];
	RUD.m[0][1] = u[0]; RUD.m[1][1] = u[1]; RUD.m[2][1] = u[2];
	RUD.m[0][2] = dir[0]; RUD.m[1][2] = dir[1]; RUD.m[2][2] = dir[2];
	M4 T = m4_identity();
	T.m[3][0] = -v->camPos[0]; T.m[3][1] = -v->camPos[1]; T.m[3][2] = -v->camPos[2];
	M4 V = m4_mult(RUD, T);

	/* perspective(fov, 1/aspectRatio, near, far): QM:1264-1283 */
	M4 P;
	memset(&P, 0, sizeof(P));
	float aspect = 1.0f / aspectRatio;
	float scale = tanf((v->camFOV * 0.5f) * d2r) * nearPlane;
	float right = aspect * scale;
	float top = scale;
	P.m[0][0] = nearPlane / right;
	P.m[1][1] = nearPlane / top;
	P.m[2][2] = -(farPlane + nearPlane) / (farPlane - nearPlane);
	P.m[3][2] = -2.0f * farPlane * nearPlane / (farPlane - nearPlane);
	P.m[2][3] = -1.0f;

	memcpy(view, &V, 64);
	memcpy(projection, &P, 64);
}

static void common_uniforms(const OrhVolume* v, OrbUniforms* u)
{
	memset(u, 0, sizeof(*u));
	memcpy(u->mapSize, v->mapSize, 12);
	u->useCubemap = 0;
	memcpy(u->skyGradientBot, v->skyGradientBot, 12);
	memcpy(u->skyGradientTop, v->skyGradientTop, 12);
	memcpy(u->sunStrength, v->sunStrength, 12);
	memcpy(u->ambientStrength, v->ambientLightStrength, 12);
}

/* V.c:845-876 */
void orh_draw_uniforms(const OrhVolume* v, const float* view, const float* projection, OrbUniforms* u)
{
	common_uniforms(v, u);
	M4 V, P, C;
	memcpy(&V, view, 64);
	memcpy(&P, projection, 64);
	C = V;
	C.m[3][0] = 0.0f; C.m[3][1] = 0.0f; C.m[3][2] = 0.0f;
	M4 iv = m4_inv(V), ic = m4_inv(C), ip = m4_inv(P);
	memcpy(u->invViewMat, &iv, 64);
	memcpy(u->invCenteredViewMat, &ic, 64);
	memcpy(u->invProjectionMat, &ip, 64);
	u->viewMode = v->camViewMode;
	u->composeRasterized = 0;
}

/* V.c:886-887, 936-947 */
void orh_light_uniforms(OrhVolume* v, int numDiffuseSamples, int maxDiffuseSamples, float time, OrbUniforms* u)
{
	if(v->frameNum == 0)
		v->lastTime = time;
	common_uniforms(v, u);
	memcpy(u->camPos, v->camPos, 12);
	u->time = v->lastTime;
	u->numDiffuseSamples = (uint32_t)numDiffuseSamples;
	u->maxDiffuseSamples = (uint32_t)maxDiffuseSamples;
	u->diffuseBounceLimit = v->diffuseBounceLimit;
	u->specularBounceLimit = v->specBounceLimit;
	float s[3] = {v->sunDir[0], v->sunDir[1], v->sunDir[2]};
	v3_normalize(s);
	memcpy(u->sunDir, s, 12);
	u->shadowSoftness = v->shadowSoftness;
}

static OrbBuffers buffers_of(OrhVolume* v)
{
	OrbBuffers b;
	b.map = v->map; b.chunks = v->gchunks; b.voxels = v->voxels; b.materials = v->materials;
	return b;
}

/* V.c:812-881 */
void orh_draw(OrhVolume* v, int w, int h, const float* view, const float* projection, float* image, OrbHit* hits)
{
	OrbUniforms u;
	orh_draw_uniforms(v, view, projection, &u);
	OrbBuffers b = buffers_of(v);
	orb_draw(&b, &u, w, h, image, hits, &v->drawCounters);
}

/* V.c:883-952 */
void orh_update_lighting(OrhVolume* v, int numDiffuseSamples, int maxDiffuseSamples, float time)
{
	OrbUniforms u;
	orh_light_uniforms(v, numDiffuseSamples, maxDiffuseSamples, time, &u);
	OrbBuffers b = buffers_of(v);
	orb_light(&b, &u, v->requests, v->numRequests, v->voxelCap, &v->lightCounters);
}

/* the same dispatch in two phases over a request range (see orb_light_compute / orb_light_commit) */
void orh_light_compute(OrhVolume* v, int numDiffuseSamples, int maxDiffuseSamples, float time, size_t first, size_t count, uint32_t* staging, uint8_t* propagate)
{
	OrbUniforms u;
	orh_light_uniforms(v, numDiffuseSamples, maxDiffuseSamples, time, &u);
	OrbBuffers b = buffers_of(v);
	orb_light_compute(&b, &u, v->requests, first, count, staging, propagate, &v->lightCounters);
}

void orh_light_commit(OrhVolume* v, const uint32_t* staging, const uint8_t* propagate)
{
	OrbUniforms u;
	common_uniforms(v, &u);
	OrbBuffers b = buffers_of(v);
	orb_light_commit(&b, &u, v->requests, v->numRequests, staging, propagate);
}

/* accessors for ctypes */
OrbHandle*   orh_map(OrhVolume* v)            { return v->map; }
OrbChunk*    orh_gpu_chunks(OrhVolume* v)     { return v->gchunks; }
OrbVoxel*    orh_voxels(OrhVolume* v)         { return v->voxels; }
size_t       orh_voxel_top(OrhVolume* v)      { return v->voxelTop; }
uint32_t*    orh_requests(OrhVolume* v)       { return v->requests; }
size_t       orh_num_requests(OrhVolume* v)   { return v->numRequests; }
OrbMaterial* orh_materials(OrhVolume* v)      { return v->materials; }
uint8_t*     orh_cpu_flags(OrhVolume* v)      { return v->cpuFlag; }
uint32_t*    orh_cpu_chunk_index(OrhVolume* v){ return v->cpuChunkIndex; }
OrhChunk*    orh_cpu_chunks(OrhVolume* v)     { return v->chunks; }
size_t       orh_chunk_cap(OrhVolume* v)      { return v->chunkCap; }
size_t       orh_sizeof_chunk(void)           { return sizeof(OrhChunk); }
OrbCounters* orh_draw_counters(OrhVolume* v)  { return &v->drawCounters; }
OrbCounters* orh_light_counters(OrhVolume* v) { return &v->lightCounters; }
size_t       orh_upload_bytes(OrhVolume* v)   { return v->uploadBytes; }
void orh_reset_counters(OrhVolume* v)
{
	memset(&v->drawCounters, 0, sizeof(OrbCounters));
	memset(&v->lightCounters, 0, sizeof(OrbCounters));
	v->uploadBytes = 0;
}

/* the public-parameter block, laid out like DNvolume's camera..sky fields, for bulk get/set from Python */
typedef struct OrhParams
{
	float camPos[3], camOrient[3], camFOV; uint32_t camViewMode;
	float sunDir[3], sunStrength[3], ambientLightStrength[3];
	uint32_t diffuseBounceLimit, specBounceLimit; float shadowSoftness;
	float skyGradientBot[3], skyGradientTop[3];
} OrhParams;

void orh_get_params(const OrhVolume* v, OrhParams* p)
{
	memcpy(p->camPos, v->camPos, 12); memcpy(p->camOrient, v->camOrient, 12); p->camFOV = v->camFOV; p->camViewMode = v->camViewMode;
	memcpy(p->sunDir, v->sunDir, 12); memcpy(p->sunStrength, v->sunStrength, 12); memcpy(p->ambientLightStrength, v->ambientLightStrength, 12);
	p->diffuseBounceLimit = v->diffuseBounceLimit; p->specBounceLimit = v->specBounceLimit; p->shadowSoftness = v->shadowSoftness;
	memcpy(p->skyGradientBot, v->skyGradientBot, 12); memcpy(p->skyGradientTop, v->skyGradientTop, 12);
}

void orh_set_params(OrhVolume* v, const OrhParams* p)
{
	memcpy(v->camPos, p->camPos, 12); memcpy(v->camOrient, p->camOrient, 12); v->camFOV = p->camFOV; v->camViewMode = p->camViewMode;
	memcpy(v->sunDir, p->sunDir, 12); memcpy(v->sunStrength, p->sunStrength, 12); memcpy(v->ambientLightStrength, p->ambientLightStrength, 12);
	v->diffuseBounceLimit = p->diffuseBounceLimit; v->specBounceLimit = p->specBounceLimit; v->shadowSoftness = p->shadowSoftness;
	memcpy(v->skyGradientBot, p->skyGradientBot, 12); memcpy(v->skyGradientTop, p->skyGradientTop, 12);
}
