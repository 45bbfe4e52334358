/* engine.h -- private state behind a DNvolume and the process-wide CUDA context. */
#ifndef DN_B200_ENGINE_H
#define DN_B200_ENGINE_H

#include "DoonEngine/b200.h"
#include "DoonEngine/voxel.h"
#include "kernels.h"
#include "record_pool.h"

#include <cuda_runtime.h>
#include <vector>

namespace dnb
{


template <typename T> struct DeviceArray
{
	T*     ptr = nullptr;
	size_t cap = 0; /* elements */
};

struct Framebuffer
{
	bool    used = false;
	int     width = 0, height = 0;
	float4* image = nullptr;
	DnbHit* hits = nullptr;
	bool    readPending = false; /* an asynchronous read-back of this image has been queued */
	cudaEvent_t evRead = nullptr; /* completion of the last asynchronous read-back of THIS image */
	float4* mirror = nullptr;    /* second destination of every pixel drawn (peer memory of the root replica), or NULL */
};

struct Context
{
	bool         ready = false;
	int          device = 0;
	cudaStream_t ownStream = nullptr;    /* kernels */
	cudaStream_t uploadStream = nullptr; /* host->device copies of edited chunks + their scatter */
	cudaStream_t readStream = nullptr;   /* asynchronous framebuffer read-back */
	cudaEvent_t  evDrawDone = nullptr, evReadDone = nullptr, evCountDone = nullptr;
	cudaStream_t userStream = nullptr;   /* DN_b200_set_stream */
	bool         useUserStream = false;
	cudaEvent_t  evUploadDone = nullptr, evComputeDone = nullptr;
	cudaEvent_t  evT0 = nullptr, evT1 = nullptr;
	bool         timing = false;
	std::vector<Framebuffer> framebuffers;

	cudaStream_t stream() const { return useUserStream ? userStream : ownStream; }
};

Context& ctx();

/* the public struct comes first so that DNvolume* == VolumeImpl* */
struct VolumeImpl
{
	DNvolume pub;
	uint32_t magic;

	/* ---- host mirror of what is resident ---- */
	std::vector<uint32_t> tileSlotHost;   /* per tile: 0 or slot+1 */
	std::vector<uint32_t> freeSlots;
	uint32_t              slotTop = 0;    /* slots [0, slotTop) have been handed out at least once */
	std::vector<uint32_t> slotNodeStart;  /* per slot: first record of its node */
	std::vector<uint8_t>  slotNodeClass;  /* per slot: size class of its node, 0xFF = none */
	std::vector<uint32_t> slotNumVoxels;  /* per slot: records in use */
	RecordPool            pool;           /* the record-pool allocator (record_pool.h) */
	std::vector<uint32_t> slotTile;       /* per slot: owner tile (for the gpuVoxelLayout mirror) */
	size_t                residentGroups = 0; /* sum over resident chunks of ceil(records/32): upper bound of the request count */

	/* ---- edit tracking ---- */
	std::vector<uint32_t> touched;        /* tiles to reconcile at the next writing sync */
	std::vector<uint8_t>  touchedFlag;

	/* ---- device ---- */
	uint32_t blocks[3];
	int32_t  occMin[3] = {0x3FFFFFFF, 0x3FFFFFFF, 0x3FFFFFFF}, occMax[3] = {-0x3FFFFFFF, -0x3FFFFFFF, -0x3FFFFFFF};
	DeviceArray<uint32_t>           tileSlot;
	DeviceArray<unsigned long long> occ64;
	DeviceArray<uint32_t>           visible, propagate, forced;
	DeviceArray<DnbSlot>            slots;
	DeviceArray<uint4>              records;
	DeviceArray<DnbMaterial>        materials;
	DeviceArray<uint32_t>           requests;
	DeviceArray<uint32_t>           staging;   /* 96 words per request */
	DeviceArray<uint32_t>           blockCounts, blockOffsets, scalars;
	DeviceArray<uint32_t>           forcedList;
	DeviceArray<float>              pickRays;  /* DN_b200_step_map_batch: 3 floats of direction, 3 of origin per ray */
	DeviceArray<int4>               pickHits;
	DeviceArray<uint4>              waveCtx;   /* context pool of the wavefront lighting kernels (light_wave.cuh), allocated on first use */
	DeviceArray<unsigned long long> litCounter; /* voxel lighting updates committed since creation */
	/* upload batches travel through two (pinned host, device) buffer pairs in rotation: the host packs batch k+1 while the copy and
	 * scatter of batch k are still queued behind the GPU's frame (with one pair the host had to wait for the GPU at every batch) */
	DeviceArray<unsigned char>      blobs[2];
	unsigned char* pinnedBlobs[2] = {nullptr, nullptr};
	size_t         pinnedBlobCaps[2] = {0, 0};
	cudaEvent_t    blobDone[2] = {nullptr, nullptr}; /* the batch that last used the pair has been scattered */
	unsigned       blobTurn = 0;
	uint32_t*      pinnedScalars = nullptr;    /* read-back of the request count: a ring of four words, [syncSerial & 3] */
	uint64_t       syncSerial = 0;             /* reading syncs so far */
	DnbCounters*   counters = nullptr;         /* device, NULL when instrumentation is off */
	bool           forcedDirty = false;
	uint32_t       opaqueBits[8] = {0, 0, 0, 0, 0, 0, 0, 0}; /* materials with opacity == 1.0 in the table the slots' DNB_BBOX_OPAQUE flags were derived from */
	bool           opaqueBitsValid = false;
	std::vector<unsigned char> materialsOnDevice;            /* the table as last uploaded: an unchanged table is not sent again */

	/* ---- the lighting-request list of the last reading sync.  Its length stays on the device (scalars[0], read there by the lighting
	 * and commit kernels: layout.h DnbWork); the host knows an upper bound at once and the exact number when it asks (request_count) ---- */
	size_t requestBound = 0;                   /* the list cannot be longer than this: the 32-voxel groups of everything resident at that sync */
	bool   countPending = false;               /* the exact length has not been read back yet (evCountDone / pinnedScalars[0]) */
	size_t requestsValid = 0;                  /* exact length, once known (0 after a sync that did not read) */
	size_t lastExactCount = 0;                 /* the most recent exact length seen: sizes grids while the current one is still in flight */
	bool   exactSync = false;                  /* DN_sync_gpu waits for the exact length (numLightingRequests valid at return, as upstream) */
	/* frame pacing: nothing in the frame calls waits for the device any more, so the host could queue arbitrarily many frames ahead;
	 * DN_draw therefore waits until the draw of the frame `maxFramesInFlight` frames back (and everything queued before it) has finished */
	cudaEvent_t frameDone[4] = {nullptr, nullptr, nullptr, nullptr};
	uint64_t    framesCommitted = 0;       /* draws queued so far */
	int         maxFramesInFlight = 2;
	size_t stagedBound = 0;                    /* requestBound of the last compute phase (sizes the staging array and the commit grid) */
	size_t stagedRequests = 0;                 /* exact length used by the last compute phase, where it was needed (collective sharding), else 0 */
	int    shardRank = 0, shardWorld = 1;

	/* ---- lighting-kernel selection (engine.cpp pick_light_kernel): both kernels give the same bits, so the faster one for
	 * THIS map and camera is found by timing them on live dispatches ---- */
	struct LightTuner
	{
		/* a PROBE dispatch is split between the candidate kernels (interleaved CTAs: same frame, same random directions,
		 * statistically the same work), run back to back on the stream with an event between them; read a frame or two later */
		cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
		int      probeKernels[3] = {-1, -1, -1};
		int      probeCount = 0;               /* kernels of the probe in flight, 0 = none */
		uint64_t probeSerial = 0;              /* the reading sync whose list the probe worked on (pinnedScalars ring) */
		bool     probeExact = false;           /* probeCtas is exact (host-driven sharding) */
		uint32_t probeCtas = 0, probeFirstCta = 0, probeStride = 1;
		double   nsPerCta[4] = {0.0, 0.0, 0.0, 0.0}; /* per kernel, smoothed over the probes it took part in: 0 warp per request, 1 persistent, 2 wavefront, 3 spread */
		uint32_t probeInterval = 8;            /* dispatches between probes: short while the best two are close, long when one is far ahead */
		uint32_t regimeCtas = 0;               /* dispatch size the estimates belong to; they start over when it changes by a quarter */
		uint32_t samples[4] = {0, 0, 0, 0};
		uint64_t dispatches = 0, probes = 0;
		uint64_t lastProbeAt = 0;
		uint32_t lastProbeCtas = 0;
		uint64_t launches[4] = {0, 0, 0, 0};
		int      current = 0;                  /* the kernel that runs between probes */
		uint32_t lastWavePasses = 0;
	} tuner;

	/* ---- multi-GPU over peer memory (DoonEngine/b200.h) ---- */
	bool         peerAttached = false;
	int          peerMode = 0;              /* DNb200peerMode */
	DnbPeerTable peers;
	size_t       peerRequestCap = 0;        /* requests every replica's staging array can hold */
	uint32_t     peerEpoch = 0;             /* barriers issued so far; identical on every replica (SPMD) */
	bool         peerFenceSinceCommit = false; /* a barrier has been passed since this replica's last commit */
	bool         peerMergeUnfenced = false; /* peers may still be reading this replica's visible bitmap (merge not yet fenced) */
	uint32_t     peerTimeoutsReported = 0;
	DeviceArray<uint32_t> mailbox, barrierStatus;

	DNb200stats stats;
};

const uint32_t VOLUME_MAGIC = 0xD00EB200u;

void report(DNmessageType type, DNmessageSeverity severity, const char* fmt, ...);
bool cuda_ok(cudaError_t e, const char* what);

inline VolumeImpl* impl_of(DNvolume* vol) { return reinterpret_cast<VolumeImpl*>(vol); }
inline size_t      num_tiles(const DNvolume* vol) { return (size_t)vol->mapSize.x * vol->mapSize.y * vol->mapSize.z; }

void touch_tile(VolumeImpl* v, size_t mapIndex);
bool device_create(VolumeImpl* v);   /* allocates the per-tile arrays for pub.mapSize */
void device_destroy(VolumeImpl* v);
void fill_scene(VolumeImpl* v, DnbScene* s);
bool sync_materials(VolumeImpl* v, cudaStream_t s);
size_t request_count(VolumeImpl* v, bool wait); /* exact length of the current request list; !wait: the last known one if it is still in flight */

} // namespace dnb

#endif
