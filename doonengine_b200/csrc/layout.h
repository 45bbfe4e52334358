/* layout.h -- how a DNvolume lives in B200 HBM, shared by the host code and the kernels.
 *
 * The reference keeps three SSBOs: a 12-byte handle per tile, a dense 96-byte
 * chunk header per TILE, and a pool of 16-byte voxel records.  This layout keeps the record (it is the
 * unit the lighting kernel updates) and re-designs everything a ray touches on the way to it:
 *
 *   occ64[]     1 bit per tile, one uint64 per 4x4x4 block of tiles.  The outer DDA tests this bit and
 *               touches nothing else for an empty tile; a ray keeps the current block's word in a
 *               register.  (replaces reading map[i].flags, voxelShared.comp:440-441)
 *   tileSlot[]  uint32 per tile: 0 = nothing resident, else chunk-pool slot + 1.
 *   slots[]     POOLED 128-byte chunk slots (one cache line / one warp-wide coalesced load):
 *               the 512-bit surface mask, the record base, per-mask-word record prefix counts
 *               (generalises partialCounts[3], voxelShared.comp:156), sample count, owner tile.
 *   records[]   uint4 per surface voxel, words as the reference packs them
 *               (voxelShared.comp:30-36 / decode :236-256): w0 material|normal, w1 albedo|spec.x,
 *               w2 spec.yz|diffuse.x, w3 diffuse.yz.
 *   visible[]   1 bit per tile in flat map-index order (the reference's flags bit 2).
 *   propagate[] visible bits raised by specular hits during a lighting pass, merged afterwards (N3).
 *
 * Flat tile index = x + sx*(y + sy*z) (voxel.h:31).
 */
#ifndef DN_B200_LAYOUT_H
#define DN_B200_LAYOUT_H

#include <stdint.h>

#define DNB_EPSILON 0.0001f          /* voxelShared.comp:7 */
#define DNB_MAX_CHUNK_STEPS 4096u    /* loop guard; a ray that trips it is a miss (oracle.h N11) */
#define DNB_MAX_BOUNCES 16           /* capacity of the per-dispatch random table */
#define DNB_MAX_SAMPLES 8

#if defined(__CUDACC__)
#define DNB_ALIGN(n) __align__(n)
#else
#define DNB_ALIGN(n) __attribute__((aligned(n)))
#endif

/* one resident chunk; 128 bytes */
typedef struct DNB_ALIGN(128) DnbSlot
{
	uint32_t mask[16];    /* bit (x + 8*(y + 8*z)) set = surface voxel with a record */
	uint32_t voxelBase;   /* index of the chunk's first record */
	uint32_t numVoxels;   /* number of records */
	uint32_t numSamples;  /* lighting dispatches accumulated (Chunk.numIndirectSamples) */
	uint32_t matIds;      /* the distinct material ids of the chunk's surface voxels, one per byte, unused bytes = 0xFF (never a voxel's material:
	                         255 = empty); more than four: 0xFFFFFFFF and DNB_BBOX_MIXED in bbox */
	uint16_t prefix[16];  /* records before mask word i */
	int32_t  pos[3];      /* owner tile position */
	uint32_t bbox;        /* bounding box of the surface voxels, ready to use as cell offsets (trace.cuh "exact chunk cull"):
	                         bits [3a, 3a+3) = 7 - max on axis a (rays stepping up), bits [9+3a, 12+3a) = min on axis a (rays stepping down);
	                         DNB_BBOX_MIXED: more than four materials (matIds does not list them);
	                         DNB_BBOX_OPAQUE: every material listed in matIds has opacity == 1.0 in the material table the device holds, so
	                         a set voxel bit is an opaque hit (voxelShared.comp:351) without looking at the record -- kept true by the host:
	                         set at packing, re-derived for every slot by dn_refresh_opaque_kernel when the table's opacities change */
} DnbSlot;

#define DNB_BBOX_MIXED  0x40000000u
#define DNB_BBOX_OPAQUE 0x80000000u

/* material as the kernels read it; 32 bytes like DNmaterial, same field order (voxel.h:82-94) */
typedef struct DnbMaterial
{
	float    pad[2];
	uint32_t emissive;
	float    opacity;
	float    refractIndex;
	float    specular;
	uint32_t reflectType;
	uint32_t shininess;
} DnbMaterial;

/* per-pixel first-hit record for parity tests (same fields as the oracle's OrbHit) */
typedef struct DnbHit
{
	int32_t  status;      /* 0 = ray misses the map box, 1 = enters but hits nothing, 2 = hit */
	uint32_t mapIndex;
	uint32_t localIndex;
	uint32_t recordIndex; /* relative to the chunk's first record */
} DnbHit;

/* traversal counters (optional instrumentation; same meaning as the oracle's OrbCounters) */
typedef struct DnbCounters
{
	unsigned long long rays, tiles, chunks, voxelSteps, records, voxelsLit, pixels;
} DnbCounters;

/* everything a kernel needs to walk the map */
typedef struct DnbScene
{
	uint32_t mapSize[3];
	uint32_t blocks[3];          /* ceil(mapSize / 4) */
	uint32_t numTiles;
	uint32_t maxMapSteps;        /* 4*(sx+sy+sz)+256 */
	int32_t  occMin[3], occMax[3]; /* inclusive tile bounding box of everything ever resident (conservative) */
	const unsigned long long* occ64;
	const uint32_t*    tileSlot;
	const DnbSlot*     slots;
	const uint4*       records;
	const DnbMaterial* materials;
	uint32_t*          visible;
	uint32_t*          propagate;
	DnbCounters*       counters; /* NULL unless instrumentation is on */
	float skyBot[3], skyTop[3];
	float sunStrength[3];
	float ambient[3];
} DnbScene;

typedef struct DnbDrawParams
{
	float    invView[16];        /* column-major */
	float    invCenteredView[16];
	float    invProjection[16];
	uint32_t viewMode;
	int32_t  width, height;      /* image size; (width/16)x(height/16) 16x16 groups are drawn (voxel.c:879) */
	int32_t  rowBegin, rowEnd;   /* 16-pixel group rows rowBegin, rowBegin+rowStride, ... < rowEnd are drawn by this launch */
	int32_t  rowStride;          /* 1 = a contiguous band; world size = rows interleaved over the replicas (screen-tile split) */
} DnbDrawParams;

/* multi-GPU over peer memory (DoonEngine/b200.h): where replica r's exchange buffers are mapped in THIS process */
#define DNB_MAX_PEERS 8
typedef struct DnbPeerTable
{
	uint32_t        world, rank;
	uint32_t*       staging[DNB_MAX_PEERS];
	uint32_t*       mailbox[DNB_MAX_PEERS];
	const uint32_t* visible[DNB_MAX_PEERS];
	const uint32_t* propagate[DNB_MAX_PEERS];
} DnbPeerTable;

/* scheduling knobs of the persistent lighting kernel (light_flat.cuh) */
typedef struct DnbFlatTuning
{
	int budget, endLanes, patience;
	int endMax; /* 1: finished rays are also served as soon as they are the most populated state of the warp */
	int keep;   /* a stepping phase goes on while at least keep/8 of the lanes it started with are still in it */
	int run;    /* > 0: a stepping phase is exactly this many steps instead (no vote inside the loop) */
} DnbFlatTuning;

/* the staging arrays one lighting launch stores into: its own, or every replica's */
typedef struct DnbStagingTargets
{
	uint32_t  count;
	uint32_t* dst[DNB_MAX_PEERS];
} DnbStagingTargets;

/* which lighting requests a launch covers.  The request COUNT normally stays on the device: the compaction kernels leave it in a
 * device word and the lighting / commit kernels read it there, so DN_sync_gpu need not wait for it (the host only knows an upper
 * bound -- the groups of everything resident -- and sizes buffers and grids from that and from the last count it has seen). */
typedef struct DnbWork
{
	const uint32_t* count;    /* device word holding the length of the request list, or NULL: `limit` is the length */
	uint32_t limit;           /* count == NULL: the length; else, if non-zero, an upper clamp (requests [0, limit) only) */
	uint32_t firstCta, ctaStride; /* the launch handles the 4-request CTAs firstCta, firstCta + ctaStride, ... */
	uint32_t numCtas;         /* how many of them; 0 = as many as the request count gives */
} DnbWork;

/* random numbers of one lighting dispatch.  Every seed in voxelLighting.comp is a function of
 * (time, sample, bounce) only, never of the voxel (LI:75,162-168,259-260), so the host evaluates
 * rand()/rand_unit_sphere() once per dispatch with libm sinf and every thread reads the same table. */
typedef struct DnbLightParams
{
	float    camPos[3];
	float    sunDir[3];          /* normalised on the host (voxel.c:942) */
	float    shadowSoftness;
	float    time;
	uint32_t numDiffuseSamples;
	uint32_t maxDiffuseSamples;
	uint32_t diffuseBounceLimit;
	uint32_t specularBounceLimit;
	float glossyChoice[DNB_MAX_SAMPLES][DNB_MAX_BOUNCES];     /* (rand(seed_i + limit + b) + 1) * 0.5 */
	float diffuseBall[DNB_MAX_SAMPLES][DNB_MAX_BOUNCES][3];   /* rand_unit_sphere(seed_i + b) */
	float shadowBall[DNB_MAX_SAMPLES][3];                     /* rand_unit_sphere(time * (i + 1 + n)) */
	float glossyBall[DNB_MAX_BOUNCES][3];                     /* rand_unit_sphere(time + b) */
} DnbLightParams;

#endif
