/* engine.cpp -- the device half of the DN_* API: CUDA context, HBM allocations, chunk packing and upload,
 * DN_sync_gpu / DN_draw / DN_update_lighting, and the additive DN_b200_* entry points.
 *
 * Reference functions re-hosted here (all in /root/reference/src/DoonEngine/voxel.c):
 *   DN_init/DN_quit 119-163, DN_sync_gpu 719-786 (+ _DN_request_chunk_lighting 1463-1489 -> compact.cu,
 *   _DN_stream_to_gpu 1491-1536, _DN_chunk_to_gpu 1391-1461, _DN_stream_chunk/_DN_stream_voxels 1548-1640 -> upload.cu),
 *   DN_draw 812-881, DN_update_lighting 883-952, DN_set_max_voxels_gpu 1031-1082.
 * There is NO CPU fallback: without a CUDA device DN_init fails and nothing else works.
 */
#include "engine.h"
#include "hostmath.h"

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <math.h>
#include <time.h>
#include <stdlib.h>
#include <string.h>
#include <thread>

namespace dnb
{

static Context g_ctx;

Context& ctx() { return g_ctx; }
static int g_requestedDevice = -1;

bool cuda_ok(cudaError_t e, const char* what)
{
	if(e == cudaSuccess)
		return true;
	report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "CUDA error in %s: %s", what, cudaGetErrorString(e));
	return false;
}

static bool peer_barrier(VolumeImpl* v);

/* requests per replica in the host-driven (all-gather) sharding: ceil(total / world) rounded up to whole 4-request CTAs */
static inline size_t slice_len(size_t total, int world)
{
	const size_t per = (total + (size_t)world - 1) / (size_t)world;
	return (per + 3) & ~(size_t)3;
}

template <typename T> void device_free(DeviceArray<T>& a)
{
	if(a.ptr)
		cudaFree(a.ptr);
	a.ptr = nullptr;
	a.cap = 0;
}

/* grows `a` to at least `count` elements on the compute stream; `keep` copies the old contents, `zeroNew` clears the rest */
template <typename T> bool device_reserve(DeviceArray<T>& a, size_t count, bool keep, bool zeroNew, const char* what)
{
	if(count <= a.cap)
		return true;
	cudaStream_t s = ctx().stream();
	T* fresh = nullptr;
	if(!cuda_ok(cudaMalloc((void**)&fresh, count * sizeof(T)), what))
		return false;
	if(a.ptr)
	{
		/* the old block may still be written or read by queued work of either stream */
		cudaStreamSynchronize(ctx().uploadStream);
		cudaStreamSynchronize(s);
	}
	size_t kept = 0;
	if(a.ptr && keep)
	{
		kept = a.cap;
		cuda_ok(cudaMemcpyAsync(fresh, a.ptr, kept * sizeof(T), cudaMemcpyDeviceToDevice, s), what);
	}
	if(zeroNew)
		cuda_ok(cudaMemsetAsync(fresh + kept, 0, (count - kept) * sizeof(T), s), what);
	if(a.ptr)
	{
		cudaStreamSynchronize(s);
		cudaFree(a.ptr);
	}
	a.ptr = fresh;
	a.cap = count;
	return true;
}

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

bool device_create(VolumeImpl* v)
{
	const DNvolume* vol = &v->pub;
	const size_t tiles = num_tiles(vol);
	const size_t words = (tiles + 31) / 32;
	for(int a = 0; a < 3; a++)
		v->blocks[a] = div_up((&vol->mapSize.x)[a], 4);
	const size_t numBlocks = (size_t)v->blocks[0] * v->blocks[1] * v->blocks[2];

	bool ok = true;
	ok = ok && device_reserve(v->tileSlot, tiles ? tiles : 1, false, true, "tile table");
	ok = ok && device_reserve(v->occ64, numBlocks ? numBlocks : 1, false, true, "occupancy bitmap");
	ok = ok && device_reserve(v->visible, words ? words : 1, false, true, "visible bitmap");
	ok = ok && device_reserve(v->propagate, words ? words : 1, false, true, "propagate bitmap");
	ok = ok && device_reserve(v->forced, words ? words : 1, false, true, "forced bitmap");
	ok = ok && device_reserve(v->materials, DN_MAX_MATERIALS, false, true, "materials");
	v->materialsOnDevice.clear();
	ok = ok && device_reserve(v->slots, std::max<size_t>(vol->chunkCap, 64), false, true, "chunk slots");
	ok = ok && device_reserve(v->records, std::max<size_t>(vol->voxelCap, 4096), false, true, "voxel records");
	ok = ok && device_reserve(v->requests, 1024, false, false, "lighting requests");
	const uint32_t cb = dnb_compact_num_blocks((uint32_t)tiles);
	ok = ok && device_reserve(v->blockCounts, cb ? cb : 1, false, true, "compaction counts");
	ok = ok && device_reserve(v->blockOffsets, cb ? cb : 1, false, true, "compaction offsets");
	ok = ok && device_reserve(v->scalars, 32, false, true, "scalars");
	ok = ok && device_reserve(v->litCounter, 1, false, true, "lit counter");
	if(ok && !v->pinnedScalars)
		ok = cuda_ok(cudaMallocHost((void**)&v->pinnedScalars, 16 * sizeof(uint32_t)), "pinned scalars");
	if(!ok)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_FATAL, "failed to allocate device memory for a %ux%ux%u map", vol->mapSize.x, vol->mapSize.y, vol->mapSize.z);
		return false;
	}
	v->pub.voxelCap = v->records.cap;
	v->pub.glMapBufferID = 1;
	v->pub.glChunkBufferID = 2;
	v->pub.glVoxelBufferID = 3;
	return true;
}

void device_destroy(VolumeImpl* v)
{
	if(ctx().ready)
	{
		cudaStreamSynchronize(ctx().stream());
		cudaStreamSynchronize(ctx().uploadStream);
	}
	device_free(v->tileSlot); device_free(v->occ64); device_free(v->visible); device_free(v->propagate); device_free(v->forced);
	device_free(v->slots); device_free(v->records); device_free(v->materials); device_free(v->requests); device_free(v->staging);
	device_free(v->blockCounts); device_free(v->blockOffsets); device_free(v->scalars); device_free(v->forcedList); device_free(v->waveCtx); device_free(v->pickRays); device_free(v->pickHits); device_free(v->blobs[0]); device_free(v->blobs[1]); device_free(v->litCounter);
	device_free(v->mailbox); device_free(v->barrierStatus);
	v->peerAttached = false;
	for(int k = 0; k < 4; k++)
	{
		if(v->frameDone[k]) cudaEventDestroy(v->frameDone[k]);
		v->frameDone[k] = nullptr;
	}
	for(int i = 0; i < 4; i++)
	{
		if(v->tuner.ev[i]) cudaEventDestroy(v->tuner.ev[i]);
		v->tuner.ev[i] = nullptr;
	}
	v->tuner.probeCount = 0;
	for(int k = 0; k < 2; k++)
	{
		if(v->pinnedBlobs[k]) cudaFreeHost(v->pinnedBlobs[k]);
		if(v->blobDone[k]) cudaEventDestroy(v->blobDone[k]);
		v->pinnedBlobs[k] = nullptr;
		v->pinnedBlobCaps[k] = 0;
		v->blobDone[k] = nullptr;
	}
	if(v->pinnedScalars) cudaFreeHost(v->pinnedScalars);
	if(v->counters) cudaFree(v->counters);
	v->pinnedScalars = nullptr;
	v->counters = nullptr;
}

void fill_scene(VolumeImpl* v, DnbScene* s)
{
	const DNvolume* vol = &v->pub;
	memset(s, 0, sizeof(*s));
	s->mapSize[0] = vol->mapSize.x; s->mapSize[1] = vol->mapSize.y; s->mapSize[2] = vol->mapSize.z;
	s->blocks[0] = v->blocks[0]; s->blocks[1] = v->blocks[1]; s->blocks[2] = v->blocks[2];
	s->numTiles = (uint32_t)num_tiles(vol);
	s->maxMapSteps = 4u * (vol->mapSize.x + vol->mapSize.y + vol->mapSize.z) + 256u;
	for(int a = 0; a < 3; a++)
	{
		s->occMin[a] = v->occMin[a];
		s->occMax[a] = v->occMax[a];
	}
	s->occ64 = v->occ64.ptr;
	s->tileSlot = v->tileSlot.ptr;
	s->slots = v->slots.ptr;
	s->records = v->records.ptr;
	s->materials = v->materials.ptr;
	s->visible = v->visible.ptr;
	s->propagate = v->propagate.ptr;
	s->counters = v->counters;
	memcpy(s->skyBot, &vol->skyGradientBot, 12);
	memcpy(s->skyTop, &vol->skyGradientTop, 12);
	memcpy(s->sunStrength, &vol->sunStrength, 12);
	memcpy(s->ambient, &vol->ambientLightStrength, 12);
}

static void exact_opaque_bits(const DNvolume* vol, uint32_t bits[8]);

/* uploads the material table when it differs from the device's copy (upstream sends it with every draw and lighting call,
 * voxel.c:823-824; a 4 KB host compare is cheaper than a copy node between two kernels of the frame) and keeps the slots'
 * DNB_BBOX_OPAQUE flags true: they were derived from the opacities of v->opaqueBits; when the application has changed an opacity
 * across 1.0 since, every slot's flag is re-derived on the device before anything traces */
bool sync_materials(VolumeImpl* v, cudaStream_t s)
{
	DNvolume* vol = &v->pub;
	const size_t tableBytes = sizeof(DNmaterial) * DN_MAX_MATERIALS;
	if(v->materialsOnDevice.size() == tableBytes && memcmp(v->materialsOnDevice.data(), vol->materials, tableBytes) == 0)
		return true; /* (the opaque flags were settled when this table was sent) */
	bool ok = cuda_ok(cudaMemcpyAsync(v->materials.ptr, vol->materials, tableBytes, cudaMemcpyHostToDevice, s), "material upload");
	if(ok)
		v->materialsOnDevice.assign(reinterpret_cast<const unsigned char*>(vol->materials), reinterpret_cast<const unsigned char*>(vol->materials) + tableBytes);
	else
		v->materialsOnDevice.clear();
	uint32_t bits[8];
	exact_opaque_bits(vol, bits);
	if(!v->opaqueBitsValid || memcmp(bits, v->opaqueBits, sizeof(bits)) != 0)
	{
		memcpy(v->opaqueBits, bits, sizeof(bits));
		v->opaqueBitsValid = true;
		if(v->slotTop > 0)
		{
			/* the upload stream may still be scattering slots packed against the old table */
			cuda_ok(cudaEventRecord(ctx().evUploadDone, ctx().uploadStream), "event record");
			cuda_ok(cudaStreamWaitEvent(s, ctx().evUploadDone, 0), "stream wait");
			ok = ok && cuda_ok(dnb_launch_refresh_opaque(v->slots.ptr, v->slotTop, bits, s), "opaque-flag refresh");
		}
	}
	return ok;
}

/* ---- timing helpers: device time of a kernel group, only when DN_b200_enable_timing(true) ---- */
/* host wall-clock of the phases of a writing sync (always on: four clock reads per sync) */
static inline double host_now_ms()
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

struct ScopedTimer
{
	float* out;
	cudaStream_t s;
	ScopedTimer(float* dst, cudaStream_t stream) : out(dst), s(stream)
	{
		if(ctx().timing)
			cudaEventRecord(ctx().evT0, s);
	}
	~ScopedTimer()
	{
		if(ctx().timing)
		{
			cudaEventRecord(ctx().evT1, s);
			cudaEventSynchronize(ctx().evT1);
			cudaEventElapsedTime(out, ctx().evT0, ctx().evT1);
		}
	}
};

/* ------------------------------------------------------------------------------------------------ */
/* chunk packing: voxel.c:1391-1461                                                                   */

/* sRGB -> linear, truncated, with the reference's constants and libm powf (voxel.c:1441-1447); host-only, no CUDA needed */
static const uint8_t* albedo_lut()
{
	/* built once, by whichever thread gets here first: initialisation of a function-local static is synchronised (C++11), so the
	 * pack workers -- up to 32 call this at once -- can never see a half-filled table */
	struct Lut
	{
		uint8_t v[256];
		Lut()
		{
			for(int i = 0; i < 256; i++)
			{
				float f = (float)i * 0.00392156862f;
				f = powf(f, DN_GAMMA);
				f = f * 255.0f;
				v[i] = (uint8_t)f;
			}
		}
	};
	static const Lut lut;
	return lut.v;
}

/* per material: does a voxel of it hide the faces it touches?  (voxel.c:1391-1394: a face is open when the neighbour is empty or
 * its material has opacity < 1; NaN compares false, i.e. hides, exactly as upstream) */
static void opaque_table(const DNvolume* vol, uint8_t out[256])
{
	for(int m = 0; m < 256; m++)
		out[m] = (m != DN_MATERIAL_EMPTY && !(vol->materials[m].opacity < 1.0f)) ? 1 : 0;
}

/* per material: is a hit on it an OPAQUE hit?  voxelShared.comp:351 tests `material.opacity == 1.0` exactly (1.5 or NaN are "transparent"
 * there although they hide faces at packing): 256 bits, bit m of word m / 32 */
static void exact_opaque_bits(const DNvolume* vol, uint32_t bits[8])
{
	for(int w = 0; w < 8; w++)
		bits[w] = 0;
	for(int m = 0; m < 256; m++)
		if(vol->materials[m].opacity == 1.0f)
			bits[m >> 5] |= 1u << (m & 31);
}

/* transpose of an 8x8 bit matrix held as 8 bytes (byte i, bit j  <->  byte j, bit i) */
static inline uint64_t transpose8(uint64_t x)
{
	uint64_t t;
	t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x = x ^ t ^ (t << 7);
	t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
	t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
	return x;
}

/* fills the slot header (everything but voxelBase) and the records in local-index order; returns the record count.
 * Same result as _DN_chunk_to_gpu + _DN_check_face_visible (voxel.c:1391-1461) voxel by voxel, computed on bit rows: the chunk
 * is stored [x][y][z], so the 8 voxels of an (x, y) row are one cache line; each row becomes two bytes (solid, opaque) and the
 * six-neighbour test becomes a handful of ANDs per row instead of six dependent loads per voxel. */
static uint32_t pack_chunk(const DNvolume* vol, const uint8_t opaque[256], const uint32_t exactOpaque[8], const DNchunk* c, uint32_t mapIndex, DnbSlot* slot, uint4* records)
{
	const uint8_t* lut = albedo_lut();
	memset(slot, 0, sizeof(*slot));
	(void)mapIndex;
	uint32_t matIds = 0xFFFFFFFFu, numMats = 0; /* distinct materials of the surface voxels, first four */
	bool mixed = false;
	slot->pos[0] = c->pos.x; slot->pos[1] = c->pos.y; slot->pos[2] = c->pos.z;
	slot->numSamples = 0; /* an edit restarts the accumulation (voxel.c:1401) */

	/* bit z of S[x][y] / O[x][y]: voxel (x, y, z) is solid / hides its neighbours' faces */
	uint8_t S[8][8], O[10][10];
	memset(O, 0, sizeof(O)); /* O is padded by one row on each side; padding = "open" */
	for(int x = 0; x < 8; x++)
		for(int y = 0; y < 8; y++)
		{
			unsigned sBits = 0, oBits = 0;
			for(int z = 0; z < 8; z++)
			{
				const uint32_t mat = c->voxels[x][y][z].normal >> 24;
				sBits |= (unsigned)(mat != DN_MATERIAL_EMPTY) << z;
				oBits |= (unsigned)opaque[mat] << z;
			}
			S[x][y] = (uint8_t)sBits;
			O[x + 1][y + 1] = (uint8_t)oBits;
		}

	/* surface voxels, regrouped: byte z of T[y] holds bits x (a row of the local-index order x + 8*(y + 8*z)) */
	uint64_t T[8];
	for(int y = 0; y < 8; y++)
	{
		uint64_t rows = 0; /* byte x = surface bits over z of row (x, y) */
		for(int x = 0; x < 8; x++)
		{
			const unsigned o = O[x + 1][y + 1];
			/* hidden: both z neighbours inside the chunk and opaque, and the four x / y neighbours (padding rows are 0 = open) */
			const unsigned hidden = (o >> 1) & (o << 1) & 0x7Eu & O[x][y + 1] & O[x + 2][y + 1] & O[x + 1][y] & O[x + 1][y + 2];
			rows |= (uint64_t)(S[x][y] & ~hidden & 0xFFu) << (8 * x);
		}
		T[y] = transpose8(rows);
	}

	uint32_t n = 0;
	for(int z = 0; z < 8; z++)
		for(int q = 0; q < 2; q++)
		{
			/* mask word 2z + q: rows y = 4q .. 4q+3 of layer z */
			uint32_t word = 0;
			for(int k = 0; k < 4; k++)
				word |= (uint32_t)((T[4 * q + k] >> (8 * z)) & 0xFFu) << (8 * k);
			const int w = 2 * z + q;
			slot->mask[w] = word;
			slot->prefix[w] = (uint16_t)n;
			while(word)
			{
				const int bit = __builtin_ctz(word);
				word &= word - 1;
				const int x = bit & 7, y = 4 * q + (bit >> 3);
				const DNcompressedVoxel vx = c->voxels[x][y][z];
				const uint32_t mat = vx.normal >> 24;
				if(((matIds & 0xFFu) != mat) && (((matIds >> 8) & 0xFFu) != mat) && (((matIds >> 16) & 0xFFu) != mat) && ((matIds >> 24) != mat))
				{
					if(numMats < 4)
						matIds = (matIds & ~(0xFFu << (8 * numMats))) | (mat << (8 * numMats));
					else
						mixed = true;
					numMats++;
				}
				uint4 rec;
				rec.x = vx.normal;
				rec.y = ((uint32_t)lut[vx.albedo >> 24] << 24) | ((uint32_t)lut[(vx.albedo >> 16) & 0xFF] << 16) | ((uint32_t)lut[(vx.albedo >> 8) & 0xFF] << 8);
				rec.z = 0;
				rec.w = 0;
				records[n++] = rec;
			}
		}
	slot->numVoxels = n;

	/* bounding box of the surface voxels: T[y] byte z holds the x bits of row (y, z) */
	unsigned xb = 0, yb = 0, zb = 0;
	for(int y = 0; y < 8; y++)
	{
		if(T[y])
			yb |= 1u << y;
		for(int z = 0; z < 8; z++)
		{
			const unsigned row = (unsigned)((T[y] >> (8 * z)) & 0xFFu);
			xb |= row;
			if(row)
				zb |= 1u << z;
		}
	}
	slot->matIds = mixed ? 0xFFFFFFFFu : matIds;
	if(n == 0)
		slot->bbox = 0; /* nothing to hit; offsets of zero = no culling */
	else
	{
		const unsigned mn[3] = {(unsigned)__builtin_ctz(xb), (unsigned)__builtin_ctz(yb), (unsigned)__builtin_ctz(zb)};
		const unsigned mx[3] = {31u - (unsigned)__builtin_clz(xb), 31u - (unsigned)__builtin_clz(yb), 31u - (unsigned)__builtin_clz(zb)};
		uint32_t bb = 0;
		for(int a = 0; a < 3; a++)
			bb |= ((7u - mx[a]) << (3 * a)) | (mn[a] << (9 + 3 * a));
		slot->bbox = bb;
	}
	/* DNB_BBOX_OPAQUE (layout.h): every listed material is opaque in the table in force */
	if(mixed)
		slot->bbox |= DNB_BBOX_MIXED;
	else
	{
		bool all = true;
		for(uint32_t k = 0; k < numMats && k < 4; k++)
		{
			const uint32_t m = (matIds >> (8 * k)) & 0xFFu;
			all = all && ((exactOpaque[m >> 5] >> (m & 31)) & 1u);
		}
		if(all)
			slot->bbox |= DNB_BBOX_OPAQUE;
	}
	return n;
}

} // namespace dnb

/* host-only (no CUDA): packs the chunk at mapPos exactly as the next writing sync would upload it -- 128-byte slot header
 * (voxelBase = 0) and up to 512 records -- for tests and offline tools.  Returns the record count, -1 if the tile has no chunk. */
extern "C" int DN_b200_pack_chunk(DNvolume* vol, DNivec3 mapPos, void* slotOut128, void* recordsOut)
{
	if(!DN_in_map_bounds(vol, mapPos))
		return -1;
	const size_t mapIndex = DN_FLATTEN_INDEX(mapPos, vol->mapSize);
	if(vol->map[mapIndex].flag == 0)
		return -1;
	uint8_t opaque[256];
	uint32_t exact[8];
	dnb::opaque_table(vol, opaque);
	dnb::exact_opaque_bits(vol, exact);
	return (int)dnb::pack_chunk(vol, opaque, exact, &vol->chunks[vol->map[mapIndex].chunkIndex], (uint32_t)mapIndex, (DnbSlot*)slotOut128, (uint4*)recordsOut);
}

namespace dnb
{

/* persistent host workers for packing: spawning threads every sync costs more than packing a small batch */
class PackPool
{
public:
	static PackPool& get()
	{
		static PackPool pool;
		return pool;
	}
	unsigned size() const { return (unsigned)threads.size() + 1; }
	/* runs fn(t) for t in [0, n) on the pool (the caller takes part); n <= size() */
	void run(unsigned n, const std::function<void(unsigned)>& fn)
	{
		if(n <= 1)
		{
			fn(0);
			return;
		}
		{
			std::unique_lock<std::mutex> lock(m);
			job = &fn;
			jobCount = n;
			nextIndex = 1;
			pending = n - 1;
			generation++;
		}
		wake.notify_all();
		fn(0);
		std::unique_lock<std::mutex> lock(m);
		done.wait(lock, [&] { return pending == 0; });
		job = nullptr;
	}

private:
	PackPool()
	{
		unsigned n = std::thread::hardware_concurrency();
		if(n == 0) n = 4;
		n = std::min(n, 32u);
		for(unsigned i = 1; i < n; i++)
			threads.emplace_back([this] { loop(); });
	}
	~PackPool()
	{
		{
			std::unique_lock<std::mutex> lock(m);
			quit = true;
		}
		wake.notify_all();
		for(auto& t : threads)
			t.join();
	}
	void loop()
	{
		uint64_t seen = 0;
		for(;;)
		{
			unsigned index;
			const std::function<void(unsigned)>* fn;
			{
				std::unique_lock<std::mutex> lock(m);
				wake.wait(lock, [&] { return quit || (generation != seen && nextIndex < jobCount); });
				if(quit)
					return;
				index = nextIndex++;
				if(nextIndex >= jobCount)
					seen = generation;
				fn = job;
			}
			(*fn)(index);
			std::unique_lock<std::mutex> lock(m);
			if(--pending == 0)
				done.notify_one();
		}
	}
	std::vector<std::thread> threads;
	std::mutex m;
	std::condition_variable wake, done;
	const std::function<void(unsigned)>* job = nullptr;
	unsigned jobCount = 0, nextIndex = 0, pending = 0;
	uint64_t generation = 0;
	bool quit = false;
};

/* ---- record-pool allocator: record_pool.h (a buddy system over 512-record blocks: split on acquire, merge on release) ---- */
static void release_slot(VolumeImpl* v, uint32_t slot)
{
	const uint8_t cls = v->slotNodeClass[slot];
	if(cls != 0xFF)
	{
		v->pool.release(v->slotNodeStart[slot], cls);
		v->slotNodeClass[slot] = 0xFF;
		v->stats.residentRecords -= v->slotNumVoxels[slot];
		v->residentGroups -= (v->slotNumVoxels[slot] + 31) / 32;
		v->slotNumVoxels[slot] = 0;
	}
}

static uint32_t acquire_node(VolumeImpl* v, uint32_t slot, uint32_t n)
{
	const int cls = RecordPool::size_class(n);
	const uint32_t start = v->pool.acquire(cls);
	v->slotNodeStart[slot] = start;
	v->slotNodeClass[slot] = (uint8_t)cls;
	v->slotNumVoxels[slot] = n;
	v->residentGroups += (n + 31) / 32;
	v->stats.residentRecords += n;
	return start;
}

static uint32_t acquire_slot(VolumeImpl* v)
{
	uint32_t slot;
	if(!v->freeSlots.empty())
	{
		slot = v->freeSlots.back();
		v->freeSlots.pop_back();
	}
	else
	{
		slot = v->slotTop++;
		v->slotNodeStart.push_back(0);
		v->slotNodeClass.push_back(0xFF);
		v->slotNumVoxels.push_back(0);
		v->slotTile.push_back(0);
	}
	return slot;
}

/* ------------------------------------------------------------------------------------------------ */
/* the writing half of DN_sync_gpu: reconcile touched tiles with the device (voxel.c:1491-1536, resident mode) */

struct PendingItem
{
	uint32_t tile;
	uint32_t chunkIndex;
	bool     remove;
};

static const size_t UPLOAD_BATCH = 16384; /* chunks per batch: bounds the pinned blob at ~131 MiB */

static bool upload_batch(VolumeImpl* v, const PendingItem* items, size_t count)
{
	DNvolume* vol = &v->pub;
	Context& c = ctx();

	/* blob layout: [items][headers][records, 512 per item worst case], identical on host (pinned) and device */
	const size_t offItems = 0;
	const size_t offHeaders = (count * sizeof(DnbUploadItem) + 127) & ~(size_t)127;
	const size_t offRecords = offHeaders + count * sizeof(DnbSlot);
	const size_t blobBytes = offRecords + count * 512 * sizeof(uint4);

	/* this batch's buffer pair; wait only for the batch that used it last (two batches ago) */
	const unsigned turn = v->blobTurn++ & 1u;
	if(!v->blobDone[turn] && !cuda_ok(cudaEventCreateWithFlags(&v->blobDone[turn], cudaEventDisableTiming), "event create"))
		return false;
	cuda_ok(cudaEventSynchronize(v->blobDone[turn]), "upload batch"); /* (an event never recorded counts as complete) */
	if(blobBytes > v->pinnedBlobCaps[turn])
	{
		if(v->pinnedBlobs[turn])
			cudaFreeHost(v->pinnedBlobs[turn]);
		v->pinnedBlobs[turn] = nullptr;
		v->pinnedBlobCaps[turn] = 0;
		const size_t want = blobBytes + blobBytes / 4; /* batches of an edit stream vary in size: leave room instead of reallocating pinned memory every other frame */
		if(!cuda_ok(cudaMallocHost((void**)&v->pinnedBlobs[turn], want), "pinned upload blob"))
			return false;
		v->pinnedBlobCaps[turn] = want;
	}
	if(!device_reserve(v->blobs[turn], v->pinnedBlobCaps[turn], false, false, "device upload blob"))
		return false;
	unsigned char* const pinnedBlob = v->pinnedBlobs[turn];

	DnbUploadItem* hItems = reinterpret_cast<DnbUploadItem*>(pinnedBlob + offItems);
	DnbSlot* hHeaders = reinterpret_cast<DnbSlot*>(pinnedBlob + offHeaders);
	uint4* hRecords = reinterpret_cast<uint4*>(pinnedBlob + offRecords);

	/* parallel pack: worker t owns a contiguous range of items and packs their records back to back into its arena */
	uint8_t opaque[256];
	opaque_table(vol, opaque);
	/* the flags are packed against the table the DEVICE's flags are consistent with (v->opaqueBits); if the application has changed
	 * opacities since, the next draw / lighting call notices and re-derives every slot's flag, these included (sync_materials) */
	const uint32_t* exactOpaque = v->opaqueBits;
	unsigned workers = 1;
	if(count >= 128)
		workers = std::max(1u, std::min<unsigned>(PackPool::get().size(), (unsigned)(count / 64)));
	std::vector<size_t> arenaUsed(workers, 0);
	auto work = [&](unsigned t)
	{
		const size_t begin = count * t / workers, end = count * (t + 1) / workers;
		size_t cursor = begin * 512;
		for(size_t i = begin; i < end; i++)
		{
			hItems[i].mapIndex = items[i].tile;
			hItems[i].slotPlus1 = 0;
			hItems[i].recordOffset = (uint32_t)cursor;
			hItems[i].pad = 0;
			if(items[i].remove)
			{
				memset(&hHeaders[i], 0, sizeof(DnbSlot));
				continue;
			}
			cursor += pack_chunk(vol, opaque, exactOpaque, &vol->chunks[items[i].chunkIndex], items[i].tile, &hHeaders[i], hRecords + cursor);
		}
		arenaUsed[t] = cursor - begin * 512;
	};
	const double tPack0 = host_now_ms();
	PackPool::get().run(workers, work);
	const double tPack1 = host_now_ms();
	v->stats.lastPackHostMs += (float)(tPack1 - tPack0);

	/* serial: slots and record nodes */
	size_t recordBytes = 0;
	for(size_t i = 0; i < count; i++)
	{
		const uint32_t tile = items[i].tile;
		uint32_t slotPlus1 = v->tileSlotHost[tile];
		if(items[i].remove)
		{
			if(slotPlus1)
			{
				release_slot(v, slotPlus1 - 1);
				v->freeSlots.push_back(slotPlus1 - 1);
				v->tileSlotHost[tile] = 0;
				v->stats.chunksRemoved++;
				v->stats.residentChunks--;
			}
			continue;
		}
		const uint32_t n = hHeaders[i].numVoxels;
		if(slotPlus1)
			release_slot(v, slotPlus1 - 1);
		else
		{
			slotPlus1 = acquire_slot(v) + 1;
			v->tileSlotHost[tile] = slotPlus1;
			v->stats.residentChunks++;
		}
		v->slotTile[slotPlus1 - 1] = tile;
		hHeaders[i].voxelBase = acquire_node(v, slotPlus1 - 1, n);
		for(int a = 0; a < 3; a++)
		{
			v->occMin[a] = std::min(v->occMin[a], hHeaders[i].pos[a]);
			v->occMax[a] = std::max(v->occMax[a], hHeaders[i].pos[a]);
		}
		hItems[i].slotPlus1 = slotPlus1;
		vol->chunks[items[i].chunkIndex].numVoxelsGpu = n; /* voxel.c:1523 */
		v->stats.chunksUploaded++;
		recordBytes += (size_t)n * sizeof(uint4);
	}

	/* grow the pools before anything is scattered into them (automatic doubling, voxel.c:773-782) */
	if(v->slotTop > v->slots.cap)
	{
		size_t cap = v->slots.cap;
		while(cap < v->slotTop) cap *= 2;
		if(!device_reserve(v->slots, cap, true, true, "chunk slots"))
			return false;
	}
	if(v->pool.top > v->records.cap)
	{
		size_t cap = v->records.cap;
		while(cap < v->pool.top) cap *= 2;
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_NOTE, "automatically resizing voxel buffer to accomodate %zu GPU voxels (%zu bytes)", cap, cap * sizeof(uint4));
		if(!device_reserve(v->records, cap, true, true, "voxel records"))
			return false;
		vol->voxelCap = v->records.cap;
	}

	/* the scatter must not overtake kernels that still read the old chunk contents -- on this GPU, or (sharded over peer
	 * memory) the peers' merge kernels, which read this replica's visible bitmap that the scatter is about to edit */
	cudaStream_t cs = c.stream();
	if(v->peerAttached && v->peerMode == DN_B200_PEER_AUTO && v->peerMergeUnfenced)
		peer_barrier(v);
	cuda_ok(cudaEventRecord(c.evComputeDone, cs), "event record");
	cuda_ok(cudaStreamWaitEvent(c.uploadStream, c.evComputeDone, 0), "stream wait");

	unsigned char* dBlob = v->blobs[turn].ptr;
	bool ok = cuda_ok(cudaMemcpyAsync(dBlob + offItems, pinnedBlob + offItems, offRecords, cudaMemcpyHostToDevice, c.uploadStream), "upload headers");
	for(unsigned t = 0; t < workers && ok; t++)
	{
		if(arenaUsed[t] == 0)
			continue;
		const size_t at = offRecords + (count * t / workers) * 512 * sizeof(uint4);
		ok = cuda_ok(cudaMemcpyAsync(dBlob + at, pinnedBlob + at, arenaUsed[t] * sizeof(uint4), cudaMemcpyHostToDevice, c.uploadStream), "upload records");
	}
	const uint32_t mapSize[3] = {vol->mapSize.x, vol->mapSize.y, vol->mapSize.z};
	ok = ok && cuda_ok(dnb_launch_scatter(reinterpret_cast<const DnbUploadItem*>(dBlob + offItems), reinterpret_cast<const DnbSlot*>(dBlob + offHeaders),
	                                      reinterpret_cast<const uint4*>(dBlob + offRecords), (uint32_t)count, mapSize, v->blocks, v->tileSlot.ptr, v->occ64.ptr, v->visible.ptr,
	                                      v->slots.ptr, v->records.ptr, c.uploadStream), "scatter kernel");

	cuda_ok(cudaEventRecord(v->blobDone[turn], c.uploadStream), "event record");

	v->stats.bytesUploaded += count * (sizeof(DnbUploadItem) + sizeof(DnbSlot)) + recordBytes;
	v->stats.lastEnqueueHostMs += (float)(host_now_ms() - tPack1);
	return ok;
}

static void sync_write(VolumeImpl* v)
{
	DNvolume* vol = &v->pub;
	Context& c = ctx();
	if(v->touched.empty())
		return;

	if(vol->gpuVoxelLayout)
	{
		/* the pool is about to change: the snapshot made by DN_b200_mirror_voxel_layout is stale */
		DN_FREE(vol->gpuVoxelLayout);
		vol->gpuVoxelLayout = NULL;
		vol->numVoxelNodes = 0;
	}
	ScopedTimer timer(&v->stats.lastUploadMs, c.uploadStream);
	const double tScan0 = host_now_ms();
	v->stats.lastPackHostMs = v->stats.lastEnqueueHostMs = 0.0f;

	std::vector<PendingItem> pending;
	pending.reserve(v->touched.size());
	for(uint32_t tile : v->touched)
	{
		v->touchedFlag[tile] = 0;
		const bool onCpu = vol->map[tile].flag != 0;
		const bool onGpu = v->tileSlotHost[tile] != 0;
		PendingItem it;
		it.tile = tile;
		it.chunkIndex = vol->map[tile].chunkIndex;
		it.remove = false;
		if(onCpu)
		{
			DNchunk* chunk = &vol->chunks[it.chunkIndex];
			if(!onGpu || chunk->updated)
				pending.push_back(it);
			chunk->updated = false; /* voxel.c:1534-1535 */
		}
		else if(onGpu)
		{
			it.remove = true;
			pending.push_back(it);
		}
	}
	v->touched.clear();

	/* ascending tile order keeps slot / node assignment independent of edit order */
	std::sort(pending.begin(), pending.end(), [](const PendingItem& a, const PendingItem& b) { return a.tile < b.tile; });
	v->stats.lastScanHostMs = (float)(host_now_ms() - tScan0);

	for(size_t at = 0; at < pending.size(); at += UPLOAD_BATCH)
		if(!upload_batch(v, pending.data() + at, std::min(UPLOAD_BATCH, pending.size() - at)))
			break;

	/* later kernels on the compute stream see the new chunks */
	cuda_ok(cudaEventRecord(c.evUploadDone, c.uploadStream), "event record");
	cuda_ok(cudaStreamWaitEvent(c.stream(), c.evUploadDone, 0), "stream wait");
}

/* the reading half: build the request list on the device (compact.cu) */
static void sync_read(VolumeImpl* v, uint32_t split)
{
	DNvolume* vol = &v->pub;
	Context& c = ctx();
	cudaStream_t s = c.stream();
	ScopedTimer timer(&v->stats.lastCompactMs, s);

	DnbScene scene;
	fill_scene(v, &scene);

	/* chunks with pending edits are lit regardless of the split phase (voxel.c:1470) */
	const uint32_t* forced = nullptr;
	if(split > 1)
	{
		std::vector<uint32_t> list;
		for(uint32_t tile : v->touched)
			if(vol->map[tile].flag != 0 && v->tileSlotHost[tile] != 0 && vol->chunks[vol->map[tile].chunkIndex].updated)
				list.push_back(tile);
		if(!list.empty())
		{
			device_reserve(v->forcedList, list.size(), false, false, "forced tile list");
			cuda_ok(cudaMemcpyAsync(v->forcedList.ptr, list.data(), list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, s), "forced list upload");
			cudaStreamSynchronize(s); /* `list` is pageable and about to go out of scope */
			cuda_ok(dnb_launch_set_bits(v->forced.ptr, v->forcedList.ptr, (uint32_t)list.size(), s), "forced bits");
			v->forcedDirty = true;
		}
		if(v->forcedDirty)
			forced = v->forced.ptr;
	}

	/* the request buffer is sized for "every resident chunk visible" (the host knows that bound exactly), so the count,
	 * scan and write passes are queued back to back and the host waits once, for the total */
	if(v->residentGroups > v->requests.cap)
	{
		size_t cap = v->requests.cap ? v->requests.cap : 1024;
		while(cap < v->residentGroups) cap *= 2;
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_NOTE, "automatically resizing lighting request buffer to accomodate %zu requests (%zu bytes)", cap, cap * sizeof(uint32_t));
		device_reserve(v->requests, cap, false, false, "lighting requests");
	}
	/* the count kernel leaves the total in a device word (read there by the lighting and commit kernels) and also stores it straight
	 * into a pinned host word; nobody waits for it here: the host sizes everything from the bound it knows, and reads the exact
	 * number only when somebody asks (request_count) */
	Context& cx = ctx();
	v->syncSerial++; /* a ring of four host words: the lighting-kernel tuner reads the count of the dispatch it timed a frame or two later */
	cuda_ok(dnb_launch_compact_count(&scene, forced, split, vol->frameNum, v->blockCounts.ptr, v->blockOffsets.ptr, v->scalars.ptr, v->pinnedScalars + (v->syncSerial & 3u), s), "compaction count");
	cuda_ok(cudaEventRecord(cx.evCountDone, s), "event record");
	cuda_ok(dnb_launch_compact_write(&scene, forced, split, vol->frameNum, v->blockOffsets.ptr, v->requests.ptr, s), "compaction write");
	v->requestBound = v->residentGroups;
	v->countPending = true;

	if(forced)
	{
		const size_t words = (num_tiles(vol) + 31) / 32;
		cuda_ok(cudaMemsetAsync(v->forced.ptr, 0, words * sizeof(uint32_t), s), "forced bitmap clear");
		v->forcedDirty = false;
	}

	/* numLightingRequests: exact at return if the application asked for upstream's semantics (DN_b200_set_exact_sync) or the device
	 * happens to be done already; otherwise it keeps the last exact value until DN_b200_lighting_request_count / _fetch_ is called */
	request_count(v, v->exactSync);
}

size_t request_count(VolumeImpl* v, bool wait)
{
	if(v->countPending)
	{
		Context& cx = ctx();
		cudaError_t q = wait ? cudaEventSynchronize(cx.evCountDone) : cudaEventQuery(cx.evCountDone);
		if(q == cudaSuccess)
		{
			v->requestsValid = *reinterpret_cast<volatile uint32_t*>(v->pinnedScalars + (v->syncSerial & 3u));
			v->lastExactCount = v->requestsValid;
			v->pub.numLightingRequests = v->requestsValid;
			v->countPending = false;
		}
		else if(q != cudaErrorNotReady)
			cuda_ok(q, "request count");
		else
			cudaGetLastError(); /* not ready is not an error */
	}
	return v->countPending ? v->lastExactCount : v->requestsValid;
}

} // namespace dnb

using namespace dnb;

extern "C" { unsigned long long g_dnbKernelLaunches = 0; }

/* ------------------------------------------------------------------------------------------------ */

extern "C" int DN_b200_device_count(void)
{
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess)
	{
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" bool DN_b200_set_device(int ordinal)
{
	if(ctx().ready)
		return false;
	g_requestedDevice = ordinal;
	return true;
}

extern "C" bool DN_init(void)
{
	Context& c = ctx();
	if(c.ready)
		return true;

	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if(e != cudaSuccess || count == 0)
	{
		cudaGetLastError();
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_FATAL, "no CUDA device available (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		return false;
	}
	int device = g_requestedDevice;
	if(device < 0)
	{
		const char* env = getenv("DN_B200_DEVICE");
		device = env ? atoi(env) : 0;
	}
	if(device < 0 || device >= count || !cuda_ok(cudaSetDevice(device), "cudaSetDevice"))
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_FATAL, "cannot select CUDA device %d of %d", device, count);
		return false;
	}
	c.device = device;

	cudaDeviceProp prop;
	if(cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties") && prop.major != 10)
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_NOTE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

	bool ok = cuda_ok(cudaStreamCreateWithFlags(&c.ownStream, cudaStreamNonBlocking), "stream create");
	ok = ok && cuda_ok(cudaStreamCreateWithFlags(&c.uploadStream, cudaStreamNonBlocking), "stream create");
	ok = ok && cuda_ok(cudaStreamCreateWithFlags(&c.readStream, cudaStreamNonBlocking), "stream create");
	ok = ok && cuda_ok(cudaEventCreateWithFlags(&c.evDrawDone, cudaEventDisableTiming), "event create");
	ok = ok && cuda_ok(cudaEventCreateWithFlags(&c.evCountDone, cudaEventDisableTiming), "event create");
	ok = ok && cuda_ok(cudaEventCreateWithFlags(&c.evReadDone, cudaEventDisableTiming), "event create");
	ok = ok && cuda_ok(cudaEventCreateWithFlags(&c.evUploadDone, cudaEventDisableTiming), "event create");
	ok = ok && cuda_ok(cudaEventCreateWithFlags(&c.evComputeDone, cudaEventDisableTiming), "event create");
	ok = ok && cuda_ok(cudaEventCreate(&c.evT0), "event create") && cuda_ok(cudaEventCreate(&c.evT1), "event create");
	if(!ok)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_FATAL, "failed to create CUDA streams");
		return false;
	}

	c.framebuffers.clear();
	c.framebuffers.resize(1); /* handle 0 is never valid */
	c.ready = true;
	return true;
}

extern "C" void DN_quit(void)
{
	Context& c = ctx();
	if(!c.ready)
		return;
	cudaStreamSynchronize(c.ownStream);
	cudaStreamSynchronize(c.uploadStream);
	for(Framebuffer& fb : c.framebuffers)
	{
		if(fb.image) cudaFree(fb.image);
		if(fb.hits) cudaFree(fb.hits);
		if(fb.evRead) cudaEventDestroy(fb.evRead);
	}
	c.framebuffers.clear();
	cudaEventDestroy(c.evUploadDone); cudaEventDestroy(c.evComputeDone); cudaEventDestroy(c.evT0); cudaEventDestroy(c.evT1);
	cudaStreamDestroy(c.ownStream);
	cudaStreamDestroy(c.uploadStream);
	cudaStreamDestroy(c.readStream);
	cudaEventDestroy(c.evDrawDone); cudaEventDestroy(c.evReadDone); cudaEventDestroy(c.evCountDone);
	c.ownStream = c.uploadStream = nullptr;
	c.ready = false;
}

extern "C" void DN_b200_set_stream(void* cudaStream)
{
	Context& c = ctx();
	if(c.ready)
		cudaStreamSynchronize(c.stream());
	c.userStream = (cudaStream_t)cudaStream;
	c.useUserStream = cudaStream != nullptr;
}

extern "C" bool DN_b200_synchronize(void)
{
	Context& c = ctx();
	if(!c.ready)
		return false;
	bool ok = cuda_ok(cudaStreamSynchronize(c.uploadStream), "synchronize (upload stream)");
	return cuda_ok(cudaStreamSynchronize(c.stream()), "synchronize") && ok;
}

extern "C" uint64_t DN_b200_kernel_launches(void) { return g_dnbKernelLaunches; }
extern "C" void DN_b200_enable_timing(bool enable) { ctx().timing = enable; }

/* ------------------------------------------------------------------------------------------------ */
/* framebuffers                                                                                       */

static Framebuffer* find_fb(GLuint id)
{
	Context& c = ctx();
	if(id == 0 || id >= c.framebuffers.size() || !c.framebuffers[id].used)
		return nullptr;
	return &c.framebuffers[id];
}

extern "C" GLuint DN_b200_create_framebuffer(int width, int height)
{
	Context& c = ctx();
	if(!c.ready || width <= 0 || height <= 0)
		return 0;
	Framebuffer fb;
	fb.used = true;
	fb.width = width;
	fb.height = height;
	if(!cuda_ok(cudaMalloc((void**)&fb.image, (size_t)width * height * sizeof(float4)), "framebuffer"))
		return 0;
	cudaMemsetAsync(fb.image, 0, (size_t)width * height * sizeof(float4), c.stream());
	for(size_t i = 1; i < c.framebuffers.size(); i++)
		if(!c.framebuffers[i].used)
		{
			c.framebuffers[i] = fb;
			return (GLuint)i;
		}
	c.framebuffers.push_back(fb);
	return (GLuint)(c.framebuffers.size() - 1);
}

extern "C" void DN_b200_delete_framebuffer(GLuint id)
{
	Framebuffer* fb = find_fb(id);
	if(!fb)
		return;
	cudaStreamSynchronize(ctx().stream());
	cudaStreamSynchronize(ctx().readStream);
	cudaFree(fb->image);
	if(fb->hits) cudaFree(fb->hits);
	if(fb->evRead) cudaEventDestroy(fb->evRead);
	*fb = Framebuffer();
}

extern "C" bool DN_b200_framebuffer_size(GLuint id, int* width, int* height)
{
	Framebuffer* fb = find_fb(id);
	if(!fb)
		return false;
	*width = fb->width;
	*height = fb->height;
	return true;
}

extern "C" void* DN_b200_framebuffer_device_ptr(GLuint id)
{
	Framebuffer* fb = find_fb(id);
	return fb ? fb->image : nullptr;
}

extern "C" bool DN_b200_read_framebuffer(GLuint id, float* dst, size_t bytes)
{
	Framebuffer* fb = find_fb(id);
	const size_t need = fb ? (size_t)fb->width * fb->height * sizeof(float4) : 0;
	if(!fb || bytes < need)
		return false;
	cudaStream_t s = ctx().stream();
	return cuda_ok(cudaMemcpyAsync(dst, fb->image, need, cudaMemcpyDeviceToHost, s), "framebuffer read") && cuda_ok(cudaStreamSynchronize(s), "framebuffer read");
}

/* Starts copying the framebuffer to (preferably pinned) host memory on a side stream, ordered after everything queued
 * so far on the compute stream (i.e. after the DN_draw that produced it); later kernels overlap with the copy.
 * DN_b200_wait_framebuffer() blocks until the copy has landed and makes later draws wait for it too. */
extern "C" bool DN_b200_read_framebuffer_async(GLuint id, float* dst, size_t bytes)
{
	Framebuffer* fb = find_fb(id);
	const size_t need = fb ? (size_t)fb->width * fb->height * sizeof(float4) : 0;
	if(!fb || bytes < need)
		return false;
	Context& c = ctx();
	bool ok = cuda_ok(cudaEventRecord(c.evDrawDone, c.stream()), "event record");
	ok = ok && cuda_ok(cudaStreamWaitEvent(c.readStream, c.evDrawDone, 0), "stream wait");
	ok = ok && cuda_ok(cudaMemcpyAsync(dst, fb->image, need, cudaMemcpyDeviceToHost, c.readStream), "framebuffer read");
	ok = ok && cuda_ok(cudaEventRecord(c.evReadDone, c.readStream), "event record");
	if(ok && !fb->evRead)
		ok = cuda_ok(cudaEventCreateWithFlags(&fb->evRead, cudaEventDisableTiming), "event create");
	ok = ok && cuda_ok(cudaEventRecord(fb->evRead, c.readStream), "event record");
	fb->readPending = true; /* the next draw into this framebuffer waits for the copy (DN_draw) */
	return ok;
}

/* The same for the rows THIS replica of a sharded volume drew (peer mode: 16-pixel group rows rank, rank + world, ...;
 * host-driven mode: its band): `dst` is the whole-image buffer, preferably pinned memory shared by all the replicas' processes
 * (DN_b200_host_register on a shared mapping) -- every GPU then sends its share over its own PCIe link and the frame is
 * assembled in host memory without crossing NVLink. */
extern "C" bool DN_b200_read_framebuffer_rows_async(GLuint id, DNvolume* vol, float* dst, size_t bytes)
{
	Framebuffer* fb = find_fb(id);
	const size_t need = fb ? (size_t)fb->width * fb->height * sizeof(float4) : 0;
	if(!fb || !vol || bytes < need)
		return false;
	VolumeImpl* v = impl_of(vol);
	Context& c = ctx();
	const size_t groupBytes = (size_t)16 * fb->width * sizeof(float4);
	const int groupRows = fb->height / 16;
	int begin = 0, end = groupRows, stride = 1;
	if(v->peerAttached)
	{
		begin = v->shardRank;
		stride = v->shardWorld;
	}
	else if(v->shardWorld > 1)
	{
		const int per = (groupRows + v->shardWorld - 1) / v->shardWorld;
		begin = std::min(groupRows, per * v->shardRank);
		end = std::min(groupRows, per * (v->shardRank + 1));
	}
	bool ok = cuda_ok(cudaEventRecord(c.evDrawDone, c.stream()), "event record");
	ok = ok && cuda_ok(cudaStreamWaitEvent(c.readStream, c.evDrawDone, 0), "stream wait");
	const int count = end > begin ? (end - begin + stride - 1) / stride : 0;
	if(ok && count > 0)
	{
		/* plain 1-D copies: one for a contiguous band, one per 16-pixel group row for interleaved rows.  (cudaMemcpy2DAsync moved the
		 * same bytes at 24 GB/s on these boxes, a 1-D copy at 55 GB/s: tools/d2h_probe.py; the read-back bounds the end-to-end frame) */
		if(stride == 1)
		{
			const size_t at = (size_t)begin * groupBytes;
			ok = cuda_ok(cudaMemcpyAsync((char*)dst + at, (const char*)fb->image + at, groupBytes * (size_t)count, cudaMemcpyDeviceToHost, c.readStream), "framebuffer rows read");
		}
		else
			for(int g = begin; g < end && ok; g += stride)
			{
				const size_t at = (size_t)g * groupBytes;
				ok = cuda_ok(cudaMemcpyAsync((char*)dst + at, (const char*)fb->image + at, groupBytes, cudaMemcpyDeviceToHost, c.readStream), "framebuffer rows read");
			}
	}
	ok = ok && cuda_ok(cudaEventRecord(c.evReadDone, c.readStream), "event record");
	if(ok && !fb->evRead)
		ok = cuda_ok(cudaEventCreateWithFlags(&fb->evRead, cudaEventDisableTiming), "event create");
	ok = ok && cuda_ok(cudaEventRecord(fb->evRead, c.readStream), "event record");
	fb->readPending = true;
	return ok;
}

/* pins (page-locks) caller-owned host memory, e.g. a POSIX shared-memory mapping several replica processes write one frame into */
extern "C" bool DN_b200_host_register(void* ptr, size_t bytes)
{
	return ctx().ready && cuda_ok(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable), "cudaHostRegister");
}

extern "C" bool DN_b200_host_unregister(void* ptr)
{
	return ctx().ready && cuda_ok(cudaHostUnregister(ptr), "cudaHostUnregister");
}

extern "C" bool DN_b200_wait_framebuffer(void)
{
	return cuda_ok(cudaStreamSynchronize(ctx().readStream), "framebuffer read");
}

/* waits for the last asynchronous read-back of ONE framebuffer (the copies of other framebuffers keep flying): with two or
 * three framebuffers in rotation the copy of frame k overlaps the whole of frame k+1 */
extern "C" bool DN_b200_wait_framebuffer_read(GLuint id)
{
	Framebuffer* fb = find_fb(id);
	if(!fb)
		return false;
	if(!fb->evRead)
		return true;
	return cuda_ok(cudaEventSynchronize(fb->evRead), "framebuffer read");
}

extern "C" bool DN_b200_clear_framebuffer(GLuint id, float value)
{
	Framebuffer* fb = find_fb(id);
	if(!fb)
		return false;
	/* only 0 and bit patterns made of one repeated byte can be memset; everything else goes through the host */
	if(value == 0.0f)
		return cuda_ok(cudaMemsetAsync(fb->image, 0, (size_t)fb->width * fb->height * sizeof(float4), ctx().stream()), "framebuffer clear");
	std::vector<float> fill((size_t)fb->width * fb->height * 4, value);
	return cuda_ok(cudaMemcpy(fb->image, fill.data(), fill.size() * sizeof(float), cudaMemcpyHostToDevice), "framebuffer clear");
}

extern "C" bool DN_b200_capture_hits(GLuint id, bool enable)
{
	Framebuffer* fb = find_fb(id);
	if(!fb)
		return false;
	if(enable && !fb->hits)
	{
		if(!cuda_ok(cudaMalloc((void**)&fb->hits, (size_t)fb->width * fb->height * sizeof(DnbHit)), "hit buffer"))
			return false;
		cudaMemsetAsync(fb->hits, 0, (size_t)fb->width * fb->height * sizeof(DnbHit), ctx().stream());
	}
	else if(!enable && fb->hits)
	{
		cudaStreamSynchronize(ctx().stream());
		cudaFree(fb->hits);
		fb->hits = nullptr;
	}
	return true;
}

extern "C" bool DN_b200_read_hits(GLuint id, DNb200hit* dst, size_t count)
{
	Framebuffer* fb = find_fb(id);
	if(!fb || !fb->hits || count < (size_t)fb->width * fb->height)
		return false;
	cudaStream_t s = ctx().stream();
	return cuda_ok(cudaMemcpyAsync(dst, fb->hits, (size_t)fb->width * fb->height * sizeof(DnbHit), cudaMemcpyDeviceToHost, s), "hit read") && cuda_ok(cudaStreamSynchronize(s), "hit read");
}

/* ------------------------------------------------------------------------------------------------ */
/* frame                                                                                              */

/* every frame entry point refuses to run without device state: there is no CPU path to fall back to */
static bool device_ready(VolumeImpl* v, const char* who)
{
	if(ctx().ready && v->tileSlot.ptr)
		return true;
	report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_FATAL, "%s: no CUDA device state (DN_init failed or was not called); this library has no CPU path", who);
	return false;
}

extern "C" void DN_sync_gpu(DNvolume* vol, DNmemOp op, int lightingSplit)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_sync_gpu"))
		return;
	if(lightingSplit < 1)
		lightingSplit = 1;

	/* voxel.c:725-727 */
	vol->frameNum++;
	if(vol->frameNum >= (uint32_t)lightingSplit)
		vol->frameNum = 0;
	/* requests first: they are sized from what is resident BEFORE this call's uploads (voxel.c:757 precedes :761) */
	if(op != DN_WRITE)
		sync_read(v, (uint32_t)lightingSplit);
	else
	{
		/* a sync that does not read leaves an empty request list (voxel.c:729: numLightingRequests = 0) */
		vol->numLightingRequests = 0;
		v->requestsValid = 0;
		v->requestBound = 0;
		v->countPending = false;
		cuda_ok(cudaMemsetAsync(v->scalars.ptr, 0, sizeof(uint32_t), ctx().stream()), "request count clear");
	}
	if(op != DN_READ)
		sync_write(v);
}

extern "C" void DN_draw(DNvolume* vol, GLuint outputTexture, DNmat4 view, DNmat4 projection, int rasterColorTexture, int rasterDepthTexture)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_draw"))
		return;
	Framebuffer* fb = find_fb(outputTexture);
	if(!fb)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_draw: %u is not a framebuffer created by DN_b200_create_framebuffer", outputTexture);
		return;
	}
	if(rasterColorTexture >= 0 && rasterDepthTexture >= 0)
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_draw: composing with rasterized objects is not supported by the CUDA back end; drawing voxels only");
	if(vol->useCubemap)
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_draw: cubemap skies are not supported by the CUDA back end; using the sky gradient");

	cudaStream_t s = ctx().stream();
	/* frame pacing (engine.h): the draw of the frame maxFramesInFlight frames back -- and so everything queued before it -- has finished */
	if(v->maxFramesInFlight > 0 && v->framesCommitted >= (uint64_t)v->maxFramesInFlight)
	{
		cudaEvent_t ev = v->frameDone[(v->framesCommitted - (uint64_t)v->maxFramesInFlight) & 3u];
		if(ev)
			cuda_ok(cudaEventSynchronize(ev), "frame pacing");
	}
	if(fb->readPending)
	{
		/* pixels of the previous frame may still be on their way to the host */
		cuda_ok(cudaStreamWaitEvent(s, fb->evRead ? fb->evRead : ctx().evReadDone, 0), "stream wait");
		fb->readPending = false;
	}
	ScopedTimer timer(&v->stats.lastDrawMs, s);

	/* materials travel with every draw, as upstream (voxel.c:823-824) */
	sync_materials(v, s);

	/* voxel.c:845-853 */
	Mat4 V, P, C;
	memcpy(V.m, view.m, sizeof(V.m));
	memcpy(P.m, projection.m, sizeof(P.m));
	C = V;
	C.m[3][0] = 0.0f; C.m[3][1] = 0.0f; C.m[3][2] = 0.0f;
	const Mat4 invV = inverse(V), invC = inverse(C), invP = inverse(P);

	DnbScene scene;
	fill_scene(v, &scene);
	DnbDrawParams dp;
	memcpy(dp.invView, invV.m, 64);
	memcpy(dp.invCenteredView, invC.m, 64);
	memcpy(dp.invProjection, invP.m, 64);
	dp.viewMode = vol->camViewMode;
	dp.width = fb->width;
	dp.height = fb->height;
	const int groupRows = fb->height / 16;
	dp.rowStride = 1;
	if(v->peerAttached)
	{
		/* interleaved group rows: rank, rank + world, ... */
		dp.rowBegin = v->shardRank;
		dp.rowEnd = groupRows;
		dp.rowStride = v->shardWorld;
	}
	else if(v->shardWorld > 1)
	{
		const int per = (groupRows + v->shardWorld - 1) / v->shardWorld;
		dp.rowBegin = std::min(groupRows, per * v->shardRank);
		dp.rowEnd = std::min(groupRows, per * (v->shardRank + 1));
	}
	else
	{
		dp.rowBegin = 0;
		dp.rowEnd = groupRows;
	}
	cuda_ok(dnb_launch_draw(&scene, &dp, fb->image, v->peerAttached ? fb->mirror : nullptr, fb->hits, s), "draw kernel");

	/* sharded over peer memory: once every replica has drawn (and its mirrored pixels have landed), OR the peers' visible
	 * bits into this replica's, so that every replica compacts the identical request list */
	if(v->peerAttached && v->peerMode == DN_B200_PEER_AUTO)
	{
		peer_barrier(v);
		cuda_ok(dnb_launch_peer_or_visible(&v->peers, v->visible.ptr, (uint32_t)((num_tiles(vol) + 31) / 32), s), "visible merge");
		v->peerMergeUnfenced = true;
	}

	/* frame pacing: the end of this frame's draw (and with it of everything queued before it: the previous frame's lighting) */
	cudaEvent_t& done = v->frameDone[v->framesCommitted & 3u];
	if(!done)
		cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
	if(done)
		cudaEventRecord(done, s);
	v->framesCommitted++;
}

/* which lighting kernel runs: 0 = one warp per request (light.cu), 1 = persistent state machine (light_flat.cuh), 2 = whichever
 * is faster on this volume (default).  Initial value from $DN_B200_LIGHT_KERNEL ("warp" / "flat" / "auto"), overridden by
 * DN_b200_set_light_kernel. */
static int g_lightKernel = -1;
static int light_kernel_choice()
{
	if(g_lightKernel < 0)
	{
		const char* env = getenv("DN_B200_LIGHT_KERNEL");
		g_lightKernel = (env && strcmp(env, "warp") == 0) ? 0 : (env && strcmp(env, "flat") == 0) ? 1 : (env && strcmp(env, "wave") == 0) ? 3 : (env && strcmp(env, "spread") == 0) ? 4 : 2;
	}
	return g_lightKernel;
}

extern "C" void DN_b200_set_light_kernel(int which) { g_lightKernel = (which >= 0 && which <= 4) ? which : 2; }
extern "C" int DN_b200_get_light_kernel(void) { return light_kernel_choice(); }

extern "C" bool dnb_light_spread_usable(uint32_t numDiffuseSamples, uint32_t specularBounceLimit);
extern "C" cudaError_t dnb_launch_light_spread(const DnbScene* scene, const uint32_t* requests, const DnbWork* work, uint32_t gridCtas, const DnbStagingTargets* targets, cudaStream_t stream);

/* kernel indices used below: 0 warp per request (light.cu), 1 persistent state machine (light_flat.cuh), 2 wavefront pair
 * (light_wave.cuh), 3 spread: a warp per voxel with specular rays (light_spread.cuh).
 *
 * Auto mode.  The kernels are the same function of the map (tests/test_parity_gpu.py), so choosing between them is purely a question
 * of speed, and that depends on the scene: short rays that end together favour one warp per request, rays of very different length
 * the persistent state machine, and dispatches too small to fill the machine with one thread per voxel the spread kernel (a
 * candidate only up to SPREAD_MAX_REQUESTS requests).  Timing them on DIFFERENT dispatches does not compare them: every voxel of a
 * dispatch uses the same random directions (voxelLighting.comp:75,162-168), so one frame of the dense map costs five times another,
 * and while the visible set still grows the host does not even know a dispatch's size.  So the candidates are timed on the SAME
 * dispatch: a probe dispatch is split between them, CTA k of this process's share going to candidate k mod n -- statistically the same
 * work -- and the parts run back to back on the stream with an event in between (no synchronisation: read a dispatch or two later).
 * Every request is still lit exactly once, so a probe costs only what the slower candidates lose on their part.  Probes: the second
 * and fourth dispatch of a volume, then every 8th to 64th (the closer the best two, the sooner), and whenever the dispatch has grown or shrunk by a quarter since the last one (while
 * the visible set of a new camera position is still growing, that is every other dispatch); a
 * candidate that was more than twice as slow sits out fifteen of sixteen probes.  The estimates are smoothed over probes (one frame's
 * random directions can favour either kernel) and start over when the dispatch changes size; the fastest estimate runs, with 5 %
 * hysteresis.  The wavefront pair is not a candidate: after round 2's dropped-item fix it is slower than the persistent kernel on
 * every configuration (profiles/r2_light.md); it stays selectable explicitly. */
static const size_t SPREAD_MAX_REQUESTS = 8192;

/* reads a finished probe: per-CTA times of its kernels, from the count the device has published for that dispatch */
static void harvest_probe(VolumeImpl* v)
{
	VolumeImpl::LightTuner& t = v->tuner;
	if(t.probeCount == 0 || cudaEventQuery(t.ev[t.probeCount]) != cudaSuccess)
	{
		cudaGetLastError(); /* cudaErrorNotReady is not an error */
		return;
	}
	/* the CTAs the dispatch really had: the host sized it from a count that lags a frame behind; the device has published the real
	 * one in the meantime (a ring of four words), unless four more syncs have gone by */
	uint32_t ctas = t.probeCtas;
	if(!t.probeExact)
	{
		if(v->syncSerial - t.probeSerial >= 4u)
			ctas = 0;
		else
		{
			const uint32_t exact = *reinterpret_cast<volatile uint32_t*>(v->pinnedScalars + (t.probeSerial & 3u));
			const uint32_t total = (exact + 3u) / 4u;
			ctas = total > t.probeFirstCta ? (total - t.probeFirstCta + t.probeStride - 1u) / t.probeStride : 0u;
		}
	}
	const int n = t.probeCount;
	t.probeCount = 0;
	if(ctas < (uint32_t)n)
		return;
	double ns[3];
	for(int i = 0; i < n; i++)
	{
		float ms = 0.0f;
		if(cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]) != cudaSuccess)
			return;
		ns[i] = 1e6 * (double)ms / (double)((ctas - (uint32_t)i + (uint32_t)n - 1u) / (uint32_t)n);
	}
	/* one frame's random directions can favour either kernel (on the dense map the ratio swings between 0.7 and 1.9 from frame to
	 * frame), so the estimates are smoothed over probes -- and start over when the dispatch has changed size by a quarter, because
	 * per-CTA times of a half-empty machine say nothing about a full one */
	if(t.regimeCtas == 0 || ctas > t.regimeCtas + t.regimeCtas / 4u || ctas < t.regimeCtas - t.regimeCtas / 4u)
	{
		t.regimeCtas = ctas;
		for(int k = 0; k < 4; k++)
			t.samples[k] = 0;
	}
	for(int i = 0; i < n; i++)
	{
		const int k = t.probeKernels[i];
		t.nsPerCta[k] = t.samples[k] == 0 ? ns[i] : 0.5 * t.nsPerCta[k] + 0.5 * ns[i];
		t.samples[k]++;
	}
	/* the fastest estimate runs, with 5 % hysteresis in favour of the incumbent; the closer the runner-up, the sooner the next probe */
	int best = -1;
	double second = 0.0;
	for(int k = 0; k < 4; k++)
	{
		if(t.samples[k] == 0)
			continue;
		if(best < 0 || t.nsPerCta[k] < t.nsPerCta[best])
		{
			if(best >= 0)
				second = second == 0.0 ? t.nsPerCta[best] : std::min(second, t.nsPerCta[best]);
			best = k;
		}
		else
			second = second == 0.0 ? t.nsPerCta[k] : std::min(second, t.nsPerCta[k]);
	}
	if(best < 0)
		return;
	if(best != t.current && (t.samples[t.current] == 0 || t.nsPerCta[best] < 0.95 * t.nsPerCta[t.current]))
		t.current = best;
	const double ratio = second > 0.0 ? second / t.nsPerCta[best] : 1.0;
	t.probeInterval = ratio > 2.0 ? 64u : ratio > 1.3 ? 16u : 8u;
}

/* the kernels of this dispatch: one (returns 1), or the candidates of a probe (returns 2 or 3, the incumbent first) */
static int pick_light_kernels(VolumeImpl* v, uint32_t numCtas, bool spreadEligible, int kernels[3])
{
	VolumeImpl::LightTuner& t = v->tuner;
	const int mode = light_kernel_choice();
	if(mode != 2)
	{
		kernels[0] = mode == 3 ? 2 : mode == 4 ? (spreadEligible ? 3 : 0) : mode;
		return 1;
	}
	harvest_probe(v);
	const uint64_t n = t.dispatches++;
	if(t.current == 3 && !spreadEligible)
		t.current = 0;
	kernels[0] = t.current;

	const bool due = n == 1 || n == 3 || n - t.lastProbeAt >= t.probeInterval ||
	                 (t.lastProbeCtas > 0 && n - t.lastProbeAt >= 2u && (numCtas > t.lastProbeCtas + t.lastProbeCtas / 4u || numCtas < t.lastProbeCtas - t.lastProbeCtas / 4u));
	if(!due || t.probeCount != 0 || n == 0 || numCtas < 6u)
		return 1;
	const bool everyone = (t.probes & 15u) == 15u || t.probes < 2u;
	int count = 1;
	const int cand[3] = {0, 1, 3};
	for(int c = 0; c < 3; c++)
	{
		const int k = cand[c];
		if(k == t.current || (k == 3 && !spreadEligible))
			continue;
		/* a candidate that was more than twice as slow in its last probe sits most probes out: on the demo map one part of a probe run
		 * by the warp-per-request kernel costs as much as eight whole dispatches of the kernel that runs */
		const bool hopeless = t.samples[k] > 0 && t.nsPerCta[t.current] > 0.0 && t.nsPerCta[k] > 2.0 * t.nsPerCta[t.current];
		if(hopeless && !everyone)
			continue;
		kernels[count++] = k;
	}
	t.lastProbeAt = n;
	t.lastProbeCtas = numCtas;
	t.probes++;
	return count;
}

/* pool size of the wavefront kernels: every voxel of the dispatch in flight at once when that fits, else the cap set with
 * DN_b200_set_wave_slots / $DN_B200_WAVE_SLOTS (default 4 Mi slots = 960 MiB; the passes then stream the dispatch through the pool) */
static uint32_t g_waveSlots = 0;
extern "C" void DN_b200_set_wave_slots(uint32_t slots)
{
	g_waveSlots = slots == 0 ? 0 : (std::min<uint32_t>(std::max<uint32_t>(slots, 256u), 1u << 26) + 255u) & ~255u;
}
static uint32_t wave_pool_slots(uint32_t numCtas)
{
	if(g_waveSlots == 0)
	{
		const char* env = getenv("DN_B200_WAVE_SLOTS");
		DN_b200_set_wave_slots(env && atoll(env) >= 128 ? (uint32_t)std::min<long long>(atoll(env), 1ll << 26) : (1u << 22));
	}
	const unsigned long long items = ((unsigned long long)numCtas * 128ull + 255ull) & ~255ull; /* the serve kernel's CTAs hold 256 slots */
	return (uint32_t)std::min<unsigned long long>(items, g_waveSlots);
}

static bool light_compute(VolumeImpl* v, int numDiffuseSamples, int maxDiffuseSamples, float time)
{
	DNvolume* vol = &v->pub;
	if(!device_ready(v, "DN_update_lighting"))
		return false;
	cudaStream_t s = ctx().stream();

	/* voxel.c:886-887 */
	if(vol->frameNum == 0)
		vol->lastTime = time;

	/* the list's length stays on the device; the host works from the bound it knows (engine.h).  Only the host-driven sharding
	 * (contiguous slices exchanged with all-gathers) needs the exact number: its slice boundaries are derived from it. */
	const bool needExact = v->shardWorld > 1 && !v->peerAttached;
	const size_t bound = v->requestBound;
	const size_t total = needExact ? request_count(v, true) : bound;
	v->stagedRequests = needExact ? total : 0;
	v->stagedBound = bound;
	if(total == 0 && !v->peerAttached)
		return true; /* (attached replicas still fence and clear their propagate bitmap below) */
	/* grids are sized from the last exact count the host has seen (it lags a frame behind), with slack; the kernels stride or pull
	 * work from counters, so any size is correct */
	request_count(v, v->lastExactCount == 0); /* (the very first dispatches have nothing to go by: wait once) */
	const size_t seen = v->countPending ? v->lastExactCount + v->lastExactCount / 4 + 1024 : v->requestsValid;
	const size_t expect = std::max<size_t>(std::min<size_t>(seen, bound), 1);

	if(numDiffuseSamples < 0) numDiffuseSamples = 0;
	if(numDiffuseSamples > DNB_MAX_SAMPLES || vol->diffuseBounceLimit > DNB_MAX_BOUNCES)
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_update_lighting: at most %d diffuse samples and %d diffuse bounces per dispatch are supported; clamping", DNB_MAX_SAMPLES, DNB_MAX_BOUNCES);

	DnbLightParams lp;
	memset(&lp, 0, sizeof(lp));
	memcpy(lp.camPos, &vol->camPos, 12);
	float sun[3] = {vol->sunDir.x, vol->sunDir.y, vol->sunDir.z};
	normalize(sun); /* voxel.c:942 */
	memcpy(lp.sunDir, sun, 12);
	lp.shadowSoftness = vol->shadowSoftness;
	lp.time = vol->lastTime;
	lp.numDiffuseSamples = (uint32_t)std::min(numDiffuseSamples, (int)DNB_MAX_SAMPLES);
	lp.maxDiffuseSamples = (uint32_t)maxDiffuseSamples;
	lp.diffuseBounceLimit = std::min<uint32_t>(vol->diffuseBounceLimit, DNB_MAX_BOUNCES);
	lp.specularBounceLimit = vol->specBounceLimit;

	/* every random number of the dispatch (layout.h DnbLightParams).  NOTE the seeds use the UNclamped uniforms,
	 * exactly as the shader would form them. */
	const float t = lp.time;
	const float fLimit = (float)vol->diffuseBounceLimit;
	for(uint32_t i = 0; i < lp.numDiffuseSamples; i++)
	{
		const float seed = t * (float)(i + 1);                                  /* LI:259 */
		for(uint32_t b = 0; b < lp.diffuseBounceLimit; b++)
		{
			lp.glossyChoice[i][b] = (shader_rand(seed + fLimit + (float)b) + 1.0f) * 0.5f; /* LI:162 */
			shader_rand_unit_sphere(seed + (float)b, lp.diffuseBall[i][b]);           /* LI:168 */
		}
		shader_rand_unit_sphere(t * (float)(i + 1 + (uint32_t)numDiffuseSamples), lp.shadowBall[i]); /* LI:260,75 */
	}
	for(uint32_t b = 0; b < lp.diffuseBounceLimit; b++)
		shader_rand_unit_sphere(t + (float)b, lp.glossyBall[b]);                    /* LI:163 */

	DnbScene scene;
	fill_scene(v, &scene);

	/* the CTAs (4 requests each) this process lights, and where their staged words go */
	const uint32_t totalCtas = (uint32_t)(((needExact ? total : expect) + 3) / 4);
	uint32_t firstCta = 0, ctaStride = 1, numCtas = totalCtas;
	DnbWork work;
	memset(&work, 0, sizeof(work));
	work.count = needExact ? nullptr : v->scalars.ptr;
	work.ctaStride = 1;
	DnbStagingTargets targets;
	memset(&targets, 0, sizeof(targets));
	if(v->peerAttached)
	{
		if(total > v->peerRequestCap)
		{
			report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_FATAL, "DN_update_lighting: %zu lighting requests exceed the %zu the attached peer staging arrays hold; detach, DN_b200_peer_prepare with a larger capacity and attach again",
			       total, v->peerRequestCap);
			v->stagedRequests = 0;
			return false;
		}
		/* interleaved: rank, rank + world, ...; every replica's staging array is a target (stores over NVLink) */
		firstCta = (uint32_t)v->shardRank;
		ctaStride = (uint32_t)v->shardWorld;
		numCtas = totalCtas > firstCta ? (totalCtas - firstCta + ctaStride - 1) / ctaStride : 0;
		work.firstCta = firstCta;
		work.ctaStride = ctaStride;
		targets.count = v->peers.world;
		for(uint32_t p = 0; p < v->peers.world; p++)
			targets.dst[p] = v->peers.staging[p];
		/* the peers' previous commit read this replica's staging and propagate arrays: make sure they are done with them */
		if(v->peerMode == DN_B200_PEER_AUTO && !v->peerFenceSinceCommit)
			peer_barrier(v);
		const size_t words = (num_tiles(vol) + 31) / 32;
		cuda_ok(cudaMemsetAsync(v->propagate.ptr, 0, words * sizeof(uint32_t), s), "propagate bitmap clear");
	}
	else
	{
		size_t paddedTotal = total;
		if(v->shardWorld > 1)
		{
			/* contiguous slice [rank * per, (rank + 1) * per): the unit of the host-driven all-gather */
			const size_t per = slice_len(total, v->shardWorld);
			const size_t first = std::min(total, per * (size_t)v->shardRank);
			const size_t count = std::min(total, per * (size_t)(v->shardRank + 1)) - first;
			paddedTotal = per * (size_t)v->shardWorld;
			firstCta = (uint32_t)(first / 4);
			numCtas = (uint32_t)((count + 3) / 4);
			/* a slice that is not the last one ends on a CTA boundary (slice_len is a multiple of 4); the last CTA of the list is cut by the limit */
			work.limit = (uint32_t)std::min(total, per * (size_t)(v->shardRank + 1));
			work.firstCta = firstCta;
			work.numCtas = numCtas;
		}
		/* grown with half as much again in reserve: the request count creeps up frame by frame while a map is being edited, and
		 * every reallocation synchronises the device (config 4: a cudaMalloc + cudaFree per frame showed up as 6 ms of "lighting") */
		size_t want = paddedTotal * 96;
		if(want > v->staging.cap)
			want = std::max(want, v->staging.cap + v->staging.cap / 2);
		if(!device_reserve(v->staging, want, false, false, "lighting staging"))
			return false;
		targets.count = 1;
		targets.dst[0] = v->staging.ptr;
	}

	ScopedTimer timer(&v->stats.lastLightMs, s);
	bool ok = sync_materials(v, s);
	ok = ok && cuda_ok(dnb_upload_light_params(&lp, s), "lighting parameters");
	const bool spreadEligible = expect <= SPREAD_MAX_REQUESTS && dnb_light_spread_usable(lp.numDiffuseSamples, lp.specularBounceLimit);
	int kernels[3] = {0, -1, -1};
	const int parts = scene.counters ? 1 : pick_light_kernels(v, numCtas, spreadEligible, kernels); /* the instrumented build exists for the warp kernel only */
	VolumeImpl::LightTuner& tuner = v->tuner;
	bool probing = parts > 1;
	if(probing)
	{
		for(int i = 0; i <= parts && probing; i++)
			if(!tuner.ev[i] && cudaEventCreate(&tuner.ev[i]) != cudaSuccess)
				probing = false;
		if(!probing)
			cudaGetLastError();
	}
	const int launches = probing ? parts : 1;
	if(probing)
	{
		tuner.probeSerial = v->syncSerial;
		tuner.probeExact = needExact;
		tuner.probeCtas = numCtas;
		tuner.probeFirstCta = work.firstCta;
		tuner.probeStride = work.ctaStride ? work.ctaStride : 1u;
		for(int i = 0; i < parts; i++)
			tuner.probeKernels[i] = kernels[i];
		cudaEventRecord(tuner.ev[0], s);
	}
	for(int part = 0; part < launches; part++)
	{
		const int kernel = kernels[part];
		/* a probe's part: every `launches`-th CTA of this process's share, starting with the part's number */
		DnbWork w = work;
		uint32_t partCtas = numCtas;
		if(launches > 1)
		{
			const uint32_t stride = work.ctaStride ? work.ctaStride : 1u;
			w.firstCta = work.firstCta + (uint32_t)part * stride;
			w.ctaStride = stride * (uint32_t)launches;
			partCtas = numCtas > (uint32_t)part ? (numCtas - (uint32_t)part + (uint32_t)launches - 1u) / (uint32_t)launches : 0u;
			if(work.numCtas)
				w.numCtas = partCtas;
			if(partCtas == 0)
			{
				cudaEventRecord(tuner.ev[part + 1], s);
				continue;
			}
		}
		/* multi-GPU: the warp-per-request kernel stores its coalesced rows into every replica itself, and so do the persistent and
		 * spread kernels (4-byte stores as their lanes finish, but overlapped with the ray tracing: measured faster at 8 replicas than
		 * a push afterwards, 302 vs 349 ns per 4 requests on config 3); the wavefront kernels stage locally and their rows are pushed
		 * to the peers afterwards (light.cu dn_push_staging_kernel: 323 -> 261 ns at 4 replicas) */
		DnbStagingTargets to = targets;
		const bool pushAfter = v->peerAttached && kernel == 2 && targets.count > 1;
		if(pushAfter)
		{
			to.count = 1;
			to.dst[0] = targets.dst[v->peers.rank];
		}
		if(kernel == 2)
		{
			const uint32_t P = wave_pool_slots(partCtas);
			ok = ok && (P == 0 || device_reserve(v->waveCtx, (size_t)P * (dnb_wave_slot_bytes() / sizeof(uint4)), false, false, "wavefront lighting contexts"));
			ok = ok && cuda_ok(dnb_launch_light_wave(&scene, v->requests.ptr, &w, &to, v->waveCtx.ptr, P, v->scalars.ptr + 16, &tuner.lastWavePasses, s), "wavefront lighting kernels");
		}
		else if(kernel == 3)
			ok = ok && cuda_ok(dnb_launch_light_spread(&scene, v->requests.ptr, &w, std::max<uint32_t>(partCtas, 1u), &to, s), "lighting kernel (spread)");
		else
			ok = ok && cuda_ok(dnb_launch_light(&scene, v->requests.ptr, &w, std::max<uint32_t>(partCtas, 1u), &to, kernel == 1 ? v->scalars.ptr + 8 : nullptr, s), "lighting kernel");
		if(pushAfter)
			ok = ok && cuda_ok(dnb_launch_push_staging(&targets, v->peers.rank, &w, std::max<uint32_t>(partCtas, 1u), s), "staging push");
		if(probing)
			cudaEventRecord(tuner.ev[part + 1], s);
		tuner.launches[kernel]++;
	}
	if(probing)
		tuner.probeCount = parts;
	return ok;
}

static bool light_commit(VolumeImpl* v)
{
	if(!device_ready(v, "DN_b200_light_commit"))
		return false;
	cudaStream_t s = ctx().stream();
	DnbScene scene;
	fill_scene(v, &scene);
	ScopedTimer timer(&v->stats.lastCommitMs, s);
	if(v->peerAttached && v->peerMode == DN_B200_PEER_AUTO)
		peer_barrier(v); /* every replica has stored its staged words into this replica's staging array */
	/* every replica commits the WHOLE list from its own staging array; the length is read on the device */
	DnbWork work;
	memset(&work, 0, sizeof(work));
	work.count = v->scalars.ptr;
	work.ctaStride = 1;
	const size_t seen = v->countPending ? std::min<size_t>(v->lastExactCount + v->lastExactCount / 4 + 1024, v->stagedBound) : std::min(v->requestsValid, v->stagedBound);
	const bool ok = cuda_ok(dnb_launch_commit(&scene, v->slots.ptr, v->records.ptr, v->requests.ptr, &work, (uint32_t)(v->stagedBound ? std::max<size_t>(seen, 1) : 0), v->staging.ptr, v->litCounter.ptr,
	                                          v->peerAttached ? &v->peers : nullptr, s), "commit kernel");
	v->peerFenceSinceCommit = false;
	return ok;
}

extern "C" void DN_update_lighting(DNvolume* vol, int numDiffuseSamples, int maxDiffuseSamples, float time)
{
	VolumeImpl* v = impl_of(vol);
	if(v->shardWorld > 1 && !(v->peerAttached && v->peerMode == DN_B200_PEER_AUTO))
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_update_lighting on a sharded volume lights only this rank's slice; use DN_b200_light_compute + exchange + DN_b200_light_commit");
	if(light_compute(v, numDiffuseSamples, maxDiffuseSamples, time))
		light_commit(v);
}

extern "C" bool DN_b200_light_compute(DNvolume* vol, int numDiffuseSamples, int maxDiffuseSamples, float time)
{
	return light_compute(impl_of(vol), numDiffuseSamples, maxDiffuseSamples, time);
}

extern "C" bool DN_b200_light_commit(DNvolume* vol)
{
	return light_commit(impl_of(vol));
}

extern "C" size_t DN_b200_staging_slice_bytes(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	return slice_len(v->stagedRequests, v->shardWorld) * 96 * sizeof(uint32_t);
}

extern "C" bool DN_b200_set_shard(DNvolume* vol, int rank, int worldSize)
{
	if(worldSize < 1 || rank < 0 || rank >= worldSize)
		return false;
	VolumeImpl* v = impl_of(vol);
	v->shardRank = rank;
	v->shardWorld = worldSize;
	return true;
}

/* ------------------------------------------------------------------------------------------------ */
/* capacity of the record pool: voxel.c:1031-1082                                                     */

extern "C" bool DN_set_max_voxels_gpu(DNvolume* vol, size_t num)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_set_max_voxels_gpu"))
		return false;
	if(num < v->pool.top)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_set_max_voxels_gpu: %zu records are in use, cannot shrink to %zu", v->pool.top, num);
		return false;
	}
	if(num <= v->records.cap)
		return true; /* never shrinks: resident records keep their place */
	if(!device_reserve(v->records, num, true, true, "voxel records"))
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "failed to reallocate voxel buffer");
		return false;
	}
	vol->voxelCap = v->records.cap;
	return true;
}

/* ------------------------------------------------------------------------------------------------ */
/* state access                                                                                       */

static bool array_info(VolumeImpl* v, DNb200array which, void** ptr, size_t* bytes)
{
	const size_t tiles = num_tiles(&v->pub);
	switch(which)
	{
	case DN_B200_TILE_SLOTS: *ptr = v->tileSlot.ptr; *bytes = tiles * sizeof(uint32_t); return true;
	case DN_B200_VISIBLE:    *ptr = v->visible.ptr;  *bytes = ((tiles + 31) / 32) * sizeof(uint32_t); return true;
	case DN_B200_SLOTS:      *ptr = v->slots.ptr;    *bytes = (size_t)v->slotTop * sizeof(DnbSlot); return true;
	case DN_B200_RECORDS:    *ptr = v->records.ptr;  *bytes = v->pool.top * sizeof(uint4); return true;
	case DN_B200_REQUESTS:   *ptr = v->requests.ptr; *bytes = request_count(v, true) * sizeof(uint32_t); return true;
	case DN_B200_PROPAGATE:  *ptr = v->propagate.ptr; *bytes = ((tiles + 31) / 32) * sizeof(uint32_t); return true;
	case DN_B200_STAGING:
	{
		*ptr = v->staging.ptr;
		*bytes = (v->peerAttached || v->shardWorld <= 1 ? std::min(request_count(v, true), v->stagedBound) : slice_len(v->stagedRequests, v->shardWorld) * v->shardWorld) * 96 * sizeof(uint32_t);
		return true;
	}
	}
	return false;
}

extern "C" size_t DN_b200_array_bytes(DNvolume* vol, DNb200array which)
{
	void* p;
	size_t n;
	return array_info(impl_of(vol), which, &p, &n) ? n : 0;
}

extern "C" void* DN_b200_array_device_ptr(DNvolume* vol, DNb200array which)
{
	void* p;
	size_t n;
	return array_info(impl_of(vol), which, &p, &n) ? p : nullptr;
}

extern "C" size_t DN_b200_download(DNvolume* vol, DNb200array which, void* dst, size_t dstBytes)
{
	void* p;
	size_t n;
	if(!ctx().ready || !array_info(impl_of(vol), which, &p, &n) || n > dstBytes)
		return 0;
	if(n == 0 || !p)
		return 0;
	if(!DN_b200_synchronize())
		return 0;
	if(!cuda_ok(cudaMemcpy(dst, p, n, cudaMemcpyDeviceToHost), "download"))
		return 0;
	return n;
}

extern "C" size_t DN_b200_fetch_lighting_requests(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	const size_t total = request_count(v, true);
	if(total > vol->lightingRequestCap)
	{
		size_t cap = vol->lightingRequestCap ? vol->lightingRequestCap : 1;
		while(cap < total) cap *= 2;
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_NOTE, "automatically resizing lighting request memory to accomodate %zu requests (%zu bytes)", cap, cap * sizeof(GLuint));
		if(!DN_set_max_lighting_requests(vol, cap))
			return 0;
	}
	if(total && DN_b200_download(vol, DN_B200_REQUESTS, vol->lightingRequests, vol->lightingRequestCap * sizeof(GLuint)) == 0)
		return 0;
	return total;
}

/* the exact length of the request list of the last reading DN_sync_gpu (waits for the device if it is still in flight) */
extern "C" size_t DN_b200_lighting_request_count(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	if(!ctx().ready)
		return 0;
	return request_count(v, true);
}

extern "C" void DN_b200_set_exact_sync(DNvolume* vol, bool exact)
{
	impl_of(vol)->exactSync = exact;
}

extern "C" void DN_b200_set_max_frames_in_flight(DNvolume* vol, int frames)
{
	impl_of(vol)->maxFramesInFlight = frames < 0 ? 0 : (frames > 3 ? 3 : frames);
}

extern "C" bool DN_b200_enable_counters(DNvolume* vol, bool enable)
{
	VolumeImpl* v = impl_of(vol);
	if(!ctx().ready)
		return false;
	if(enable && !v->counters)
	{
		if(!cuda_ok(cudaMalloc((void**)&v->counters, sizeof(DnbCounters)), "counters"))
			return false;
		cudaMemsetAsync(v->counters, 0, sizeof(DnbCounters), ctx().stream());
	}
	else if(!enable && v->counters)
	{
		DN_b200_synchronize();
		cudaFree(v->counters);
		v->counters = nullptr;
	}
	return true;
}

extern "C" bool DN_b200_read_counters(DNvolume* vol, DNb200counters* out, bool reset)
{
	VolumeImpl* v = impl_of(vol);
	memset(out, 0, sizeof(*out));
	if(!v->counters || !DN_b200_synchronize())
		return false;
	DnbCounters h;
	if(!cuda_ok(cudaMemcpy(&h, v->counters, sizeof(h), cudaMemcpyDeviceToHost), "counter read"))
		return false;
	out->rays = h.rays; out->tiles = h.tiles; out->chunks = h.chunks; out->voxelSteps = h.voxelSteps;
	out->records = h.records; out->voxelsLit = h.voxelsLit; out->pixels = h.pixels;
	if(reset)
		cudaMemset(v->counters, 0, sizeof(DnbCounters));
	return true;
}

extern "C" void DN_b200_get_stats(DNvolume* vol, DNb200stats* out)
{
	VolumeImpl* v = impl_of(vol);
	v->stats.slotCap = v->slots.cap;
	v->stats.recordCap = v->records.cap;
	v->stats.lightLaunchesWarp = v->tuner.launches[0];
	v->stats.lightLaunchesFlat = v->tuner.launches[1];
	v->stats.nsPerCtaWarp = (float)v->tuner.nsPerCta[0];
	v->stats.nsPerCtaFlat = (float)v->tuner.nsPerCta[1];
	v->stats.usedNodes = v->pool.usedNodes;
	v->stats.freeNodes = v->pool.freeNodeCount;
	v->stats.recordTop = v->pool.top;
	v->stats.nodeSplits = v->pool.splits;
	v->stats.nodeMerges = v->pool.merges;
	v->stats.lightLaunchesWave = v->tuner.launches[2];
	v->stats.lightLaunchesSpread = v->tuner.launches[3];
	v->stats.nsPerCtaSpread = (float)v->tuner.nsPerCta[3];
	v->stats.nsPerCtaWave = (float)v->tuner.nsPerCta[2];
	v->stats.lastWavePasses = v->tuner.lastWavePasses;
	if(ctx().ready && v->litCounter.ptr && DN_b200_synchronize())
	{
		unsigned long long lit = 0;
		if(cuda_ok(cudaMemcpy(&lit, v->litCounter.ptr, sizeof(lit), cudaMemcpyDeviceToHost), "lit counter read"))
			v->stats.voxelsLit = lit;
	}
	*out = v->stats;
}

/* snapshot of the record-pool allocator in the reference's public form (voxel.h:74-79,108-117; DoonEngine/b200.h) */
extern "C" bool DN_b200_mirror_voxel_layout(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	DN_FREE(vol->gpuVoxelLayout);
	vol->gpuVoxelLayout = NULL;
	vol->numVoxelNodes = 0;
	const size_t count = v->pool.usedNodes + v->pool.freeNodeCount;
	DNvoxelNode* nodes = (DNvoxelNode*)DN_MALLOC(sizeof(DNvoxelNode) * (count ? count : 1));
	if(!nodes)
	{
		report(DN_MESSAGE_CPU_MEMORY, DN_MESSAGE_ERROR, "failed to allocate %zu voxel nodes", count);
		return false;
	}
	size_t n = 0;
	for(uint32_t slot = 0; slot < v->slotTop; slot++)
	{
		if(v->slotNodeClass[slot] == 0xFF)
			continue;
		const uint32_t tile = v->slotTile[slot];
		nodes[n].size = 16u << v->slotNodeClass[slot];
		nodes[n].startPos = v->slotNodeStart[slot];
		nodes[n].chunkPos.x = (int)(tile % vol->mapSize.x);
		nodes[n].chunkPos.y = (int)((tile / vol->mapSize.x) % vol->mapSize.y);
		nodes[n].chunkPos.z = (int)(tile / ((size_t)vol->mapSize.x * vol->mapSize.y));
		n++;
	}
	for(size_t at = 0; at < v->pool.nodeFree.size(); at++)
		if(v->pool.nodeFree[at] != 0xFF)
		{
			nodes[n].size = 16u << v->pool.nodeFree[at];
			nodes[n].startPos = at << 4;
			nodes[n].chunkPos.x = nodes[n].chunkPos.y = nodes[n].chunkPos.z = -1;
			n++;
		}
	std::sort(nodes, nodes + n, [](const DNvoxelNode& a, const DNvoxelNode& b) { return a.startPos < b.startPos; });
	vol->gpuVoxelLayout = nodes;
	vol->numVoxelNodes = n;
	return n == count;
}

extern "C" cudaError_t dnb_launch_or_bits(uint32_t* dst, const uint32_t* src, uint32_t words, cudaStream_t stream);

extern "C" bool DN_b200_or_bitmap(DNvolume* vol, DNb200array which, const void* deviceBitmap)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_b200_or_bitmap") || (which != DN_B200_VISIBLE && which != DN_B200_PROPAGATE))
		return false;
	const uint32_t words = (uint32_t)((num_tiles(vol) + 31) / 32);
	uint32_t* dst = which == DN_B200_VISIBLE ? v->visible.ptr : v->propagate.ptr;
	return cuda_ok(dnb_launch_or_bits(dst, (const uint32_t*)deviceBitmap, words, ctx().stream()), "bitmap merge");
}

/* ------------------------------------------------------------------------------------------------ */
/* multi-GPU over peer memory (DoonEngine/b200.h)                                                     */

namespace dnb
{
/* queues one device-side barrier over all replicas on the compute stream (peer.cu) */
static bool peer_barrier(VolumeImpl* v)
{
	v->peerEpoch++;
	v->peerFenceSinceCommit = true;
	v->peerMergeUnfenced = false;
	return cuda_ok(dnb_launch_peer_barrier(&v->peers, v->peerEpoch, v->barrierStatus.ptr, ctx().stream()), "peer barrier");
}
} // namespace dnb

extern "C" bool DN_b200_peer_prepare(DNvolume* vol, size_t requestCap, DNb200peerBuffers* mine)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_b200_peer_prepare") || !mine)
		return false;
	if(v->peerAttached)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_b200_peer_prepare: detach first");
		return false;
	}
	if(requestCap == 0)
		requestCap = v->residentGroups + v->residentGroups / 4 + 4096;
	if(!device_reserve(v->staging, requestCap * 96, false, false, "lighting staging"))
		return false;
	/* the mailbox is allocated once and never reset: epochs keep counting across re-attachments */
	if(!v->mailbox.ptr && !device_reserve(v->mailbox, 32, false, true, "peer mailbox"))
		return false;
	if(!v->barrierStatus.ptr && !device_reserve(v->barrierStatus, 2, false, true, "peer barrier status"))
		return false;
	if(!DN_b200_synchronize())
		return false;
	mine->staging = v->staging.ptr;
	mine->mailbox = v->mailbox.ptr;
	mine->visible = v->visible.ptr;
	mine->propagate = v->propagate.ptr;
	mine->stagingRequestCap = v->staging.cap / 96;
	return true;
}

extern "C" bool DN_b200_ipc_export(const void* devicePtr, void* handle64)
{
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 opaque bytes");
	cudaIpcMemHandle_t h;
	if(!cuda_ok(cudaIpcGetMemHandle(&h, const_cast<void*>(devicePtr)), "cudaIpcGetMemHandle"))
		return false;
	memcpy(handle64, &h, 64);
	return true;
}

extern "C" void* DN_b200_ipc_open(const void* handle64)
{
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	void* p = nullptr;
	if(!cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle"))
		return nullptr;
	return p;
}

extern "C" bool DN_b200_ipc_close(void* devicePtr)
{
	return cuda_ok(cudaIpcCloseMemHandle(devicePtr), "cudaIpcCloseMemHandle");
}

extern "C" bool DN_b200_peer_attach(DNvolume* vol, int rank, int worldSize, const DNb200peerBuffers* peers, DNb200peerMode mode)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_b200_peer_attach"))
		return false;
	if(worldSize < 1 || worldSize > DNB_MAX_PEERS || rank < 0 || rank >= worldSize || !peers)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_b200_peer_attach: bad rank %d / world %d (at most %d replicas)", rank, worldSize, DNB_MAX_PEERS);
		return false;
	}
	if(peers[rank].staging != v->staging.ptr || peers[rank].mailbox != v->mailbox.ptr || !v->mailbox.ptr)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_b200_peer_attach: entry %d must be this replica's own buffers (DN_b200_peer_prepare)", rank);
		return false;
	}
	DN_b200_synchronize();
	memset(&v->peers, 0, sizeof(v->peers));
	v->peers.world = (uint32_t)worldSize;
	v->peers.rank = (uint32_t)rank;
	size_t cap = (size_t)-1;
	for(int p = 0; p < worldSize; p++)
	{
		if(!peers[p].staging || !peers[p].mailbox || !peers[p].visible || !peers[p].propagate)
		{
			report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_b200_peer_attach: replica %d has unmapped buffers", p);
			return false;
		}
		v->peers.staging[p] = (uint32_t*)peers[p].staging;
		v->peers.mailbox[p] = (uint32_t*)peers[p].mailbox;
		v->peers.visible[p] = (const uint32_t*)peers[p].visible;
		v->peers.propagate[p] = (const uint32_t*)peers[p].propagate;
		cap = std::min<size_t>(cap, (size_t)peers[p].stagingRequestCap);
	}
	v->peerRequestCap = cap;
	v->peerMode = (int)mode;
	v->shardRank = rank;
	v->shardWorld = worldSize;
	v->peerFenceSinceCommit = false;
	v->peerMergeUnfenced = false;
	v->peerAttached = true;
	return true;
}

extern "C" void DN_b200_peer_detach(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	if(!v->peerAttached)
		return;
	if(ctx().ready)
		DN_b200_synchronize();
	v->peerAttached = false;
	v->shardRank = 0;
	v->shardWorld = 1;
}

extern "C" bool DN_b200_peer_capacity_ok(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	return !v->peerAttached || v->requestBound <= v->peerRequestCap;
}

extern "C" bool DN_b200_framebuffer_set_mirror(GLuint id, void* mirrorImage)
{
	Framebuffer* fb = find_fb(id);
	if(!fb)
		return false;
	cudaStreamSynchronize(ctx().stream());
	fb->mirror = (float4*)mirrorImage;
	return true;
}

extern "C" bool DN_b200_peer_exchange_visible(DNvolume* vol)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_b200_peer_exchange_visible") || !v->peerAttached)
		return false;
	return cuda_ok(dnb_launch_peer_or_visible(&v->peers, v->visible.ptr, (uint32_t)((num_tiles(vol) + 31) / 32), ctx().stream()), "visible merge");
}

extern "C" bool DN_b200_peer_barrier_status(DNvolume* vol, uint64_t* epochs, uint32_t* timeouts)
{
	VolumeImpl* v = impl_of(vol);
	if(!v->barrierStatus.ptr || !DN_b200_synchronize())
		return false;
	uint32_t st[2] = {0, 0};
	if(!cuda_ok(cudaMemcpy(st, v->barrierStatus.ptr, sizeof(st), cudaMemcpyDeviceToHost), "barrier status read"))
		return false;
	if(epochs) *epochs = st[0];
	if(timeouts) *timeouts = st[1];
	if(st[1] > v->peerTimeoutsReported)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_FATAL, "peer barrier timed out %u time(s): a replica did not arrive; results after that point are undefined", st[1] - v->peerTimeoutsReported);
		v->peerTimeoutsReported = st[1];
	}
	return true;
}

/* ------------------------------------------------------------------------------------------------ */
/* lit-state checkpoint: the reference saves only the map (voxel.c:595-654); after a load every chunk */
/* ------------------------------------------------------------------------------------------------ */
/* batched picking: DN_step_map (voxel.c:1195-1272) for `count` rays on the device map (pick.cu)      */

extern "C" size_t DN_b200_step_map_batch(DNvolume* vol, size_t count, const DNvec3* rayDirs, const DNvec3* rayPositions, int maxSteps, DNivec3* hitPositions, DNvoxel* hitVoxels,
                                         DNivec3* hitNormals, uint8_t* hitFlags)
{
	VolumeImpl* v = impl_of(vol);
	if(count == 0 || !device_ready(v, "DN_b200_step_map_batch"))
		return 0;
	if(count > 0x7FFFFFFFu)
	{
		report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_b200_step_map_batch: at most 2^31 - 1 rays per call");
		return 0;
	}
	/* DN_step_map sees every edit at once (it reads the CPU map); this call walks the DEVICE map, which is as of the last writing
	 * sync.  It does not upload pending edits behind the caller's back: that would consume the chunks' `updated` flags outside the
	 * DN_sync_gpu protocol, and a later DN_sync_gpu(READ*, lightingSplit > 1) would no longer force-light the edited chunks
	 * (voxel.c:1470) -- the lighting schedule would differ from a run that picked with DN_step_map.  So: refuse, and say what to do. */
	if(!v->touched.empty())
	{
		bool pending = false;
		for(uint32_t tile : v->touched)
		{
			const bool onCpu = vol->map[tile].flag != 0;
			const bool onGpu = v->tileSlotHost[tile] != 0;
			if((onCpu && (!onGpu || vol->chunks[vol->map[tile].chunkIndex].updated)) || (!onCpu && onGpu))
			{
				pending = true;
				break;
			}
		}
		if(pending)
		{
			report(DN_MESSAGE_GPU_MEMORY, DN_MESSAGE_ERROR, "DN_b200_step_map_batch: the map has edits that are not on the device yet; call DN_sync_gpu(vol, DN_WRITE or DN_READ_WRITE, split) first (or use DN_step_map, which reads the CPU map)");
			return 0;
		}
	}

	cudaStream_t s = ctx().stream();
	if(!device_reserve(v->pickRays, count * 6, false, false, "picking rays") || !device_reserve(v->pickHits, count, false, false, "picking results"))
		return 0;
	bool ok = cuda_ok(cudaMemcpyAsync(v->pickRays.ptr, rayDirs, count * sizeof(DNvec3), cudaMemcpyHostToDevice, s), "picking directions");
	ok = ok && cuda_ok(cudaMemcpyAsync(v->pickRays.ptr + count * 3, rayPositions, count * sizeof(DNvec3), cudaMemcpyHostToDevice, s), "picking origins");
	DnbScene scene;
	fill_scene(v, &scene);
	ok = ok && cuda_ok(dnb_launch_pick(&scene, v->pickRays.ptr, v->pickRays.ptr + count * 3, (uint32_t)count, maxSteps, v->pickHits.ptr, s), "picking kernel");
	std::vector<int4> res(count);
	ok = ok && cuda_ok(cudaMemcpyAsync(res.data(), v->pickHits.ptr, count * sizeof(int4), cudaMemcpyDeviceToHost, s), "picking read-back");
	ok = ok && cuda_ok(cudaStreamSynchronize(s), "picking");
	if(!ok)
		return 0;

	size_t hits = 0;
	for(size_t i = 0; i < count; i++)
	{
		bool hit = false;
		DNivec3 pos = {0, 0, 0}, normal = {-1000, -1000, -1000};
		if(maxSteps > 0)
		{
			/* the cell the ray starts in, from the CPU map: the only voxel a ray can find that the device (surface voxels only) may not hold */
			const DNivec3 start = {(int)floor(rayPositions[i].x * DN_CHUNK_SIZE), (int)floor(rayPositions[i].y * DN_CHUNK_SIZE), (int)floor(rayPositions[i].z * DN_CHUNK_SIZE)};
			if(start.x >= 0 && start.y >= 0 && start.z >= 0)
			{
				DNivec3 mapPos, chunkPos;
				DN_separate_position(start, &mapPos, &chunkPos);
				if(DN_in_map_bounds(vol, mapPos) && DN_does_chunk_exist(vol, mapPos) && DN_does_voxel_exist(vol, mapPos, chunkPos))
				{
					hit = true;
					pos = start;
				}
			}
			if(!hit)
			{
				const int4 r = res[i];
				const uint32_t code = (uint32_t)r.w;
				const uint32_t axis = (code >> 1) & 3u;
				if(axis < 3u)
				{
					normal.x = normal.y = normal.z = 0;
					(&normal.x)[axis] = (code & 16u) ? 0 : ((code & 8u) ? -1 : 1); /* -rayStep of the last step (voxel.c:1268-1269) */
				}
				if(code & 1u)
				{
					hit = true;
					pos.x = r.x; pos.y = r.y; pos.z = r.z;
				}
			}
		}
		if(hitNormals)
			hitNormals[i] = normal;
		if(hitFlags)
			hitFlags[i] = hit ? 1 : 0;
		if(hit)
		{
			hits++;
			if(hitPositions)
				hitPositions[i] = pos;
			if(hitVoxels)
			{
				DNivec3 mapPos, chunkPos;
				DN_separate_position(pos, &mapPos, &chunkPos);
				hitVoxels[i] = DN_get_voxel(vol, mapPos, chunkPos);
			}
		}
	}
	return hits;
}

/* re-accumulates its lighting from zero samples.  These two calls carry the accumulated lighting across a save / load.  */

namespace
{
struct LitFileHeader
{
	char     magic[8];   /* "DNLIT001" */
	uint32_t mapSize[3];
	uint32_t numChunks;
};
struct LitChunkHeader
{
	uint32_t mapIndex, numVoxels, numSamples;
	uint32_t visible;    /* the chunk's visible bit: specular hits of the last pass may have raised it (LI:101-105), so it is lighting state */
	uint32_t mask[16];   /* the surface mask the words belong to: a chunk edited since the checkpoint is skipped on load */
};
} // namespace

extern "C" bool DN_b200_save_lighting(DNvolume* vol, const char* filePath)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_b200_save_lighting") || !DN_b200_synchronize())
		return false;
	std::vector<DnbSlot> slots(v->slotTop);
	std::vector<uint4> records(v->pool.top);
	std::vector<uint32_t> visible((num_tiles(vol) + 31) / 32);
	if((v->slotTop && !cuda_ok(cudaMemcpy(slots.data(), v->slots.ptr, slots.size() * sizeof(DnbSlot), cudaMemcpyDeviceToHost), "slot download")) ||
	   (v->pool.top && !cuda_ok(cudaMemcpy(records.data(), v->records.ptr, records.size() * sizeof(uint4), cudaMemcpyDeviceToHost), "record download")) ||
	   (!visible.empty() && !cuda_ok(cudaMemcpy(visible.data(), v->visible.ptr, visible.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost), "visible bitmap download")))
		return false;
	FILE* f = fopen(filePath, "wb");
	if(!f)
	{
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "failed to open file \"%s\" for writing", filePath);
		return false;
	}
	LitFileHeader h;
	memcpy(h.magic, "DNLIT001", 8);
	h.mapSize[0] = vol->mapSize.x; h.mapSize[1] = vol->mapSize.y; h.mapSize[2] = vol->mapSize.z;
	h.numChunks = 0;
	const size_t tiles = num_tiles(vol);
	for(size_t t = 0; t < tiles; t++)
		h.numChunks += v->tileSlotHost[t] != 0;
	bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
	for(size_t t = 0; t < tiles && ok; t++)
	{
		if(!v->tileSlotHost[t])
			continue;
		const DnbSlot& s = slots[v->tileSlotHost[t] - 1];
		LitChunkHeader c;
		c.mapIndex = (uint32_t)t; c.numVoxels = s.numVoxels; c.numSamples = s.numSamples;
		c.visible = (visible[t >> 5] >> (t & 31)) & 1u;
		memcpy(c.mask, s.mask, sizeof(c.mask));
		ok = fwrite(&c, sizeof(c), 1, f) == 1 && (s.numVoxels == 0 || fwrite(&records[s.voxelBase], sizeof(uint4), s.numVoxels, f) == s.numVoxels);
	}
	fclose(f);
	if(!ok)
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "failed to write lighting checkpoint \"%s\"", filePath);
	return ok;
}

/* returns the number of chunks whose lighting was restored, -1 on error.  Call after the map is resident (a writing sync). */
extern "C" int DN_b200_load_lighting(DNvolume* vol, const char* filePath)
{
	VolumeImpl* v = impl_of(vol);
	if(!device_ready(v, "DN_b200_load_lighting") || !DN_b200_synchronize())
		return -1;
	FILE* f = fopen(filePath, "rb");
	if(!f)
	{
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "failed to open file \"%s\" for reading", filePath);
		return -1;
	}
	LitFileHeader h;
	if(fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "DNLIT001", 8) != 0 || h.mapSize[0] != vol->mapSize.x || h.mapSize[1] != vol->mapSize.y || h.mapSize[2] != vol->mapSize.z)
	{
		fclose(f);
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "\"%s\" is not a lighting checkpoint of a %ux%ux%u map", filePath, vol->mapSize.x, vol->mapSize.y, vol->mapSize.z);
		return -1;
	}
	std::vector<DnbSlot> slots(v->slotTop);
	std::vector<uint4> records(v->pool.top);
	std::vector<uint32_t> visible((num_tiles(vol) + 31) / 32);
	if((v->slotTop && !cuda_ok(cudaMemcpy(slots.data(), v->slots.ptr, slots.size() * sizeof(DnbSlot), cudaMemcpyDeviceToHost), "slot download")) ||
	   (v->pool.top && !cuda_ok(cudaMemcpy(records.data(), v->records.ptr, records.size() * sizeof(uint4), cudaMemcpyDeviceToHost), "record download")) ||
	   (!visible.empty() && !cuda_ok(cudaMemcpy(visible.data(), v->visible.ptr, visible.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost), "visible bitmap download")))
	{
		fclose(f);
		return -1;
	}
	int restored = 0;
	bool ok = true;
	std::vector<uint4> saved(512);
	const size_t tiles = num_tiles(vol);
	for(uint32_t k = 0; k < h.numChunks && ok; k++)
	{
		LitChunkHeader c;
		ok = fread(&c, sizeof(c), 1, f) == 1 && c.numVoxels <= 512;
		if(!ok)
			break;
		ok = c.numVoxels == 0 || fread(saved.data(), sizeof(uint4), c.numVoxels, f) == c.numVoxels;
		if(!ok || c.mapIndex >= tiles || !v->tileSlotHost[c.mapIndex])
			continue;
		DnbSlot& s = slots[v->tileSlotHost[c.mapIndex] - 1];
		/* edited since the checkpoint (surface, material, normal or albedo of any voxel): its lighting restarts, as after any edit (voxel.c:1401) */
		bool same = s.numVoxels == c.numVoxels && memcmp(s.mask, c.mask, sizeof(c.mask)) == 0;
		for(uint32_t i = 0; i < c.numVoxels && same; i++)
			same = records[s.voxelBase + i].x == saved[i].x && (records[s.voxelBase + i].y >> 8) == (saved[i].y >> 8);
		if(!same)
			continue;
		memcpy(&records[s.voxelBase], saved.data(), (size_t)c.numVoxels * sizeof(uint4));
		s.numSamples = c.numSamples;
		if(c.visible)
			visible[c.mapIndex >> 5] |= 1u << (c.mapIndex & 31u);
		restored++;
	}
	fclose(f);
	if(!ok)
	{
		report(DN_MESSAGE_FILE_IO, DN_MESSAGE_ERROR, "lighting checkpoint \"%s\" is truncated", filePath);
		return -1;
	}
	if(!visible.empty() && !cuda_ok(cudaMemcpy(v->visible.ptr, visible.data(), visible.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "visible bitmap upload"))
		return -1;
	if((v->slotTop && !cuda_ok(cudaMemcpy(v->slots.ptr, slots.data(), slots.size() * sizeof(DnbSlot), cudaMemcpyHostToDevice), "slot upload")) ||
	   (v->pool.top && !cuda_ok(cudaMemcpy(v->records.ptr, records.data(), records.size() * sizeof(uint4), cudaMemcpyHostToDevice), "record upload")))
		return -1;
	return restored;
}
