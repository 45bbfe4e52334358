/* DoonEngine/b200.h -- additive entry points of libdoon_b200.so.  Nothing here changes the meaning of a
 * reference symbol (DoonEngine/voxel.h); these exist because the CUDA back end has no OpenGL objects to hand out,
 * and because tests, benchmarks and the multi-GPU host need to see device state.
 *
 * Reference interface each group stands in for:
 *   framebuffers      the GL_RGBA32F texture the application creates and passes to DN_draw as `outputTexture`
 *                     (reference main.c:304-357; glBindImageTexture at voxel.c:820)
 *   device state      glMapBuffer / glGetBufferSubData on vol->gl{Map,Chunk,Voxel}BufferID (voxel.c:731)
 *   lighting phases   the single glDispatchCompute at voxel.c:950, split into compute + commit so the lit records
 *                     can be exchanged between GPUs in between
 */
#ifndef DN_B200_H
#define DN_B200_H

#include "voxel.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device selection (call before DN_init; default = device 0 or $DN_B200_DEVICE) ---- */
int  DN_b200_device_count(void);
bool DN_b200_set_device(int ordinal);
/* run all kernels of this library on a caller-owned CUDA stream (cudaStream_t), NULL = the library's own */
void DN_b200_set_stream(void* cudaStream);
/* blocks until everything queued by this library has finished; false + message on a CUDA error */
bool DN_b200_synchronize(void);

/* ---- framebuffers: linear RGBA32F, row 0 = screen y -1, usable as DN_draw's outputTexture ---- */
GLuint DN_b200_create_framebuffer(int width, int height);
void   DN_b200_delete_framebuffer(GLuint fb);
bool   DN_b200_framebuffer_size(GLuint fb, int* width, int* height);
void*  DN_b200_framebuffer_device_ptr(GLuint fb);                     /* float4[width*height] in device memory */
bool   DN_b200_read_framebuffer(GLuint fb, float* dst, size_t bytes); /* synchronous device -> host copy */
bool   DN_b200_clear_framebuffer(GLuint fb, float value);
/* asynchronous read-back on a side stream: ordered after the work queued so far, overlaps with what is queued next */
bool   DN_b200_read_framebuffer_async(GLuint fb, float* dst, size_t bytes);
/* the rows THIS replica of a sharded volume drew, into a whole-image buffer (e.g. pinned memory shared by the replicas' processes) */
bool   DN_b200_read_framebuffer_rows_async(GLuint fb, DNvolume* vol, float* dst, size_t bytes);
bool   DN_b200_host_register(void* ptr, size_t bytes);   /* cudaHostRegister(portable) on caller-owned memory */
bool   DN_b200_host_unregister(void* ptr);
bool   DN_b200_wait_framebuffer(void);            /* every read-back queued so far */
bool   DN_b200_wait_framebuffer_read(GLuint fb);  /* the last read-back of this framebuffer only */

/* per-pixel first-hit capture for parity tests: status 0 box miss / 1 no hit / 2 hit, tile, voxel, record-in-chunk */
typedef struct DNb200hit { int32_t status; uint32_t mapIndex, localIndex, recordIndex; } DNb200hit;
bool DN_b200_capture_hits(GLuint fb, bool enable);
bool DN_b200_read_hits(GLuint fb, DNb200hit* dst, size_t count);

/* ---- lighting kernel choice (all of them produce the same bits):
 * 0 = one warp per lighting request (the reference's work-group shape, voxel.c:950 / LI:3),
 * 1 = persistent warps running every voxel as a state machine with dynamic work fetch,
 * 3 = wavefront: voxel contexts in device memory, a full-warp shading / ray set-up kernel alternating with a persistent
 *     ray-stepping kernel (selectable; no longer a candidate of the automatic choice),
 * 4 = spread, for small dispatches: a voxel with specular rays gets a whole warp, its rays side by side (falls back to 0 when a
 *     dispatch asks for more than 8 diffuse samples or 4 specular bounces),
 * 2 = auto (default): now and then a dispatch is SPLIT between the candidates (0, 1 and, up to 8 192 requests, 4), CTA by CTA, each
 *     part timed; the fastest runs until the next such probe (engine.cpp pick_light_kernels).
 * Initial value from $DN_B200_LIGHT_KERNEL = "warp" | "flat" | "wave" | "spread" | "auto". ---- */
void DN_b200_set_light_kernel(int which);
int  DN_b200_get_light_kernel(void);
/* context-pool size of the wavefront kernels (rounded to a multiple of 128; 0 = default / $DN_B200_WAVE_SLOTS = 4 Mi slots of 240
 * bytes).  A dispatch with more voxels than slots is streamed through the pool; results do not depend on the size. */
void DN_b200_set_wave_slots(uint32_t slots);
/* scheduling knobs of the persistent kernel (see csrc/light_flat.cuh); results do not depend on them.  0 = defaults / environment
 * ($DN_B200_FLAT_BUDGET 24, $DN_B200_FLAT_END 20, $DN_B200_FLAT_PATIENCE 48; also $DN_B200_FLAT_KEEP 2, in eighths) */
void DN_b200_set_flat_tuning(int budget, int endLanes, int patience);

/* ---- request list ---- */
/* copies the device-built request list into vol->lightingRequests (growing it like voxel.c:1474-1484); returns the count */
size_t DN_b200_fetch_lighting_requests(DNvolume* vol);

/* ---- device state download (synchronous).  Returns bytes written, 0 on error / too small a buffer ---- */
typedef enum DNb200array
{
	DN_B200_TILE_SLOTS = 0, /* uint32 per tile: 0 = not resident, else chunk slot + 1 */
	DN_B200_VISIBLE    = 1, /* uint32 words, 1 bit per tile, flat index order */
	DN_B200_SLOTS      = 2, /* 128-byte chunk slots (csrc/layout.h DnbSlot), slot-cap entries */
	DN_B200_RECORDS    = 3, /* 16-byte voxel records, record-cap entries */
	DN_B200_REQUESTS   = 4, /* uint32 request words of the last reading sync */
	DN_B200_STAGING    = 5, /* 96 uint32 per request: the lit words of the last lighting compute phase */
	DN_B200_PROPAGATE  = 6  /* uint32 words, 1 bit per tile: visible bits raised by specular hits, merged at commit */
} DNb200array;
size_t DN_b200_array_bytes(DNvolume* vol, DNb200array which);
size_t DN_b200_download(DNvolume* vol, DNb200array which, void* dst, size_t dstBytes);
void*  DN_b200_array_device_ptr(DNvolume* vol, DNb200array which);

/* ---- the lighting-request count.  Upstream DN_sync_gpu blocks on a glMapBuffer of the whole map and leaves the request list and
 * vol->numLightingRequests on the host (voxel.c:719-786).  Here the list is built on the device and STAYS there: the lighting kernels
 * read its length from device memory, so a reading DN_sync_gpu returns without waiting for the GPU and the next DN_update_lighting is
 * queued right behind it.  vol->numLightingRequests then holds the exact length only if the device happened to be done already;
 * otherwise it keeps the last exact value it had.  Whoever needs the number asks:
 *   DN_b200_lighting_request_count()   the exact length (waits for the compaction kernel if necessary) -- also refreshes the field
 *   DN_b200_fetch_lighting_requests()  the list itself, into vol->lightingRequests (above)
 *   DN_b200_set_exact_sync(vol, true)  restores upstream's contract: every reading DN_sync_gpu waits, the field is exact at return ---- */
size_t DN_b200_lighting_request_count(DNvolume* vol);
void   DN_b200_set_exact_sync(DNvolume* vol, bool exact);
/* frame pacing: with no call of the frame loop waiting for the device, DN_draw holds the host back until the draw of the frame
 * `frames` frames earlier -- and so everything queued before it -- has finished (default 2: the device never runs dry, latency stays
 * bounded); 0 = no limit, at most 3 */
void   DN_b200_set_max_frames_in_flight(DNvolume* vol, int frames);

/* ---- instrumentation ---- */
typedef struct DNb200counters { uint64_t rays, tiles, chunks, voxelSteps, records, voxelsLit, pixels; } DNb200counters;
/* when enabled the kernels count traversal work (slower); counters accumulate until read with reset */
bool DN_b200_enable_counters(DNvolume* vol, bool enable);
bool DN_b200_read_counters(DNvolume* vol, DNb200counters* out, bool reset);

typedef struct DNb200stats
{
	uint64_t chunksUploaded, chunksRemoved, bytesUploaded; /* since creation */
	uint64_t residentChunks, residentRecords;              /* now */
	uint64_t slotCap, recordCap;
	uint64_t voxelsLit;                                     /* voxel lighting updates committed since creation */
	float    lastDrawMs, lastCompactMs, lastUploadMs, lastLightMs, lastCommitMs; /* device time of the last call of each kind, when timing is on */
	uint64_t lightLaunchesWarp, lightLaunchesFlat;          /* lighting dispatches run by each kernel */
	float    nsPerCtaWarp, nsPerCtaFlat;                    /* auto mode: running estimate of each kernel's time per 4 requests */
	float    lastScanHostMs, lastPackHostMs, lastEnqueueHostMs; /* host wall-clock of the last writing sync: dirty-tile scan + sort, packing, allocation + enqueue */
	uint64_t lightLaunchesWave;                             /* lighting dispatches run by the wavefront kernels */
	float    nsPerCtaWave;                                  /* auto mode: their time per 4 requests */
	uint32_t lastWavePasses;                                /* serve + step passes the last wavefront dispatch queued */
	uint32_t pad0;
	uint64_t nodeSplits, nodeMerges;                        /* record-pool allocator: buddy splits / merges since creation */
	uint64_t usedNodes, freeNodes;                          /* record-pool nodes now */
	uint64_t recordTop;                                     /* records of the pool handed out to the allocator (multiple of 512) */
	uint64_t lightLaunchesSpread;                           /* lighting dispatches run by the one-warp-per-voxel kernel (small dispatches) */
	float    nsPerCtaSpread;                                /* auto mode: its time per 4 requests */
	uint32_t pad1;
} DNb200stats;
void DN_b200_get_stats(DNvolume* vol, DNb200stats* out); /* synchronises (reads the device-side lit counter) */
void DN_b200_enable_timing(bool enable);
uint64_t DN_b200_kernel_launches(void); /* CUDA kernels this library has launched since it was loaded (all volumes) */ /* record CUDA events around each kernel group (adds a sync when read) */

/* vol->gpuVoxelLayout / vol->numVoxelNodes (voxel.h:108-117): upstream this array IS the allocator (O(nodes) scans per upload,
 * voxel.c:1560-1592).  Here the allocator is a buddy system keyed by address (csrc/engine.cpp), so the public array is only a
 * MIRROR, built when asked for: after this call gpuVoxelLayout holds every node of the record pool -- used ones with their owner's
 * chunkPos, free ones with chunkPos.x = -1 -- in ascending startPos order, and numVoxelNodes their count.  The next call that changes
 * the pool (a writing DN_sync_gpu with pending edits, DN_set_map_size) frees the mirror again and resets numVoxelNodes to 0; between
 * mirrors the two fields are NULL and 0, never a count without an array.  Returns false on allocation failure. */
bool DN_b200_mirror_voxel_layout(DNvolume* vol);

/* tiles whose host-side state was changed WITHOUT going through a DN_* call (e.g. writing vol->chunks[i].voxels
 * directly and setting .updated) must be announced, because DN_sync_gpu does not scan the whole map */
void DN_b200_touch_tile(DNvolume* vol, DNivec3 mapPos);
void DN_b200_rescan(DNvolume* vol);
/* host-only (works without a CUDA device): packs the chunk at mapPos exactly as the next writing sync would upload it (surface
 * culling, bit mask, prefix counts, albedo linearisation: voxel.c:1391-1461) into a 128-byte slot header (csrc/layout.h DnbSlot,
 * voxelBase = 0) and up to 512 16-byte records; returns the record count, -1 if the tile has no chunk */
int DN_b200_pack_chunk(DNvolume* vol, DNivec3 mapPos, void* slotOut128, void* recordsOut);
/* `count` DN_set_compressed_voxel / DN_remove_voxel calls in one (voxel.c:1126-1183): positions in voxel units, a voxel with
 * material DN_MATERIAL_EMPTY removes; positions outside the map are skipped; returns the number of edits applied */
size_t DN_b200_set_voxels(DNvolume* vol, size_t count, const DNivec3* positions, const DNcompressedVoxel* voxels); /* marks every tile touched: the next writing sync reconciles the whole map */

/* `count` whole chunks in one call (the loader's bulk path, voxel.c:560-640, for procedurally built maps): voxels = count x [8][8][8]
 * DNcompressedVoxel in the chunk's own [x][y][z] order; a chunk without a solid voxel removes the tile's chunk; tiles outside the map
 * are skipped; returns the number of chunks present afterwards among those given */
size_t DN_b200_set_chunks(DNvolume* vol, size_t count, const DNivec3* mapPositions, const DNcompressedVoxel* voxels);

/* ---- batched picking: DN_step_map (voxel.c:1195-1272) for `count` rays in one call, walked on the DEVICE map (csrc/pick.cu: the
 * reference's single-axis voxel DDA over the occupancy bit-grid and the chunks' surface masks).  Per ray the outputs are exactly
 * what DN_step_map would return: hitFlags[i] = its return value; hitNormals[i] is always written (-1000 if no step was taken);
 * hitPositions[i] / hitVoxels[i] only on a hit.  Any output array may be NULL.  The device map is as of the last writing DN_sync_gpu:
 * with edits pending the call reports an error and returns 0 without touching the outputs (it never syncs behind the caller's back --
 * that would consume the chunks' `updated` flags outside the DN_sync_gpu protocol and change which chunks a lightingSplit > 1 schedule
 * force-lights, voxel.c:1470).  Returns the number of hits. ---- */
size_t DN_b200_step_map_batch(DNvolume* vol, size_t count, const DNvec3* rayDirs, const DNvec3* rayPositions, int maxSteps, DNivec3* hitPositions, DNvoxel* hitVoxels,
                              DNivec3* hitNormals, uint8_t* hitFlags);

/* ---- lit-state checkpoint: DN_save_volume / DN_load_volume persist only the map (voxel.c:520-654), so upstream every chunk
 * re-accumulates its lighting from zero samples after a load.  These carry the accumulated lighting (the three lit words of every
 * record + each chunk's sample count) across: save any time; load after DN_load_volume + a writing DN_sync_gpu.  Chunks whose
 * surface mask changed since the checkpoint are skipped (their lighting restarts, as after any edit).  load returns the number of
 * chunks restored, -1 on error. ---- */
bool DN_b200_save_lighting(DNvolume* vol, const char* filePath);
int  DN_b200_load_lighting(DNvolume* vol, const char* filePath);

/* ---- multi-GPU: one process per GPU, the same volume replicated in each ---- */
/* this process lights requests [rank*ceil(R/world), ...) and draws the rank-th band of 16-pixel rows */
bool DN_b200_set_shard(DNvolume* vol, int rank, int worldSize);
/* DN_update_lighting == light_compute + light_commit.  Between the two, a sharded host all-gathers the staging
 * array (DN_B200_STAGING; each rank's slice is DN_b200_staging_slice_bytes() long at rank*that offset) and ORs
 * the ranks' DN_B200_PROPAGATE bitmaps together. */
bool   DN_b200_light_compute(DNvolume* vol, int numDiffuseSamples, int maxDiffuseSamples, float time);
bool   DN_b200_light_commit(DNvolume* vol);
size_t DN_b200_staging_slice_bytes(DNvolume* vol);
/* bitmap[] |= other[] for a bitmap gathered from another rank (device pointer, same length); `which` is
 * DN_B200_VISIBLE (after a sharded draw) or DN_B200_PROPAGATE (after a sharded lighting compute phase) */
bool   DN_b200_or_bitmap(DNvolume* vol, DNb200array which, const void* deviceBitmap);

/* ---- multi-GPU over peer memory (NVLink / NVSwitch): the kernels exchange their results themselves ----
 * Replaces the host-driven all-gathers above.  Every replica maps the others' exchange buffers (cudaIpc handles
 * between processes, plain device pointers inside one process) and then the ordinary frame calls run SPMD:
 *   DN_draw            draws the 16-pixel rows  rank, rank+world, ...  (interleaved: balances sky against terrain),
 *                      stores each pixel into its own framebuffer AND the framebuffer's mirror (peer memory of the
 *                      root replica) from inside the kernel; then a device-side barrier over NVLink and a merge
 *                      kernel that ORs the peers' visible bitmaps into this replica's
 *   DN_sync_gpu        unchanged: every replica compacts the identical bitmap into the identical request list
 *   DN_update_lighting lights request CTAs  rank, rank+world, ...  and stores the three staged words of every voxel
 *                      straight into EVERY replica's staging array (peer stores from the lighting kernel, so the
 *                      exchange overlaps the ray tracing); device-side barrier; every replica commits everything
 * Results are bit-identical to the unsharded run for any number of replicas (snapshot semantics, oracle.h N1-N3).
 * No host synchronisation and no NCCL call is on the frame path; the barrier is a one-CTA kernel that posts an epoch
 * word into every peer's mailbox (st.release.sys) and spins on its own (ld.acquire.sys), bounded by a time-out. */
#define DN_B200_MAX_PEERS 8
typedef struct DNb200peerBuffers
{
	void* staging;   /* uint32[96 * stagingRequestCap]: lit words, written by every replica */
	void* mailbox;   /* uint32[DN_B200_MAX_PEERS]: barrier epochs, slot r written by replica r */
	void* visible;   /* this replica's visible bitmap (read by the peers' merge kernel) */
	void* propagate; /* this replica's propagate bitmap (read by the peers' commit) */
	uint64_t stagingRequestCap;
} DNb200peerBuffers;
typedef enum DNb200peerMode
{
	DN_B200_PEER_AUTO   = 0, /* one process per GPU: DN_draw / DN_update_lighting run the barriers and merges themselves */
	DN_B200_PEER_MANUAL = 1  /* several replicas driven by ONE host thread on one stream (tests): no device barriers; the host
	                            sequences the phases with DN_b200_peer_exchange_visible / DN_b200_light_compute / DN_b200_light_commit */
} DNb200peerMode;
/* sizes this replica's exchange buffers for `requestCap` lighting requests (0 = every resident chunk fully lit, plus slack) and
 * returns their device pointers; call after the map is resident and before exporting handles */
bool  DN_b200_peer_prepare(DNvolume* vol, size_t requestCap, DNb200peerBuffers* mine);
/* cudaIpcGetMemHandle / cudaIpcOpenMemHandle(lazy peer access) / cudaIpcCloseMemHandle on 64-byte opaque handles */
bool  DN_b200_ipc_export(const void* devicePtr, void* handle64);
void* DN_b200_ipc_open(const void* handle64);
bool  DN_b200_ipc_close(void* devicePtr);
/* peers[world]: entry r = replica r's buffers as seen from THIS process (entry `rank` = the result of peer_prepare) */
bool  DN_b200_peer_attach(DNvolume* vol, int rank, int worldSize, const DNb200peerBuffers* peers, DNb200peerMode mode);
void  DN_b200_peer_detach(DNvolume* vol);
/* false when the request list of the last reading sync no longer fits the attached staging arrays (all replicas see the
 * same answer): detach, prepare with a larger capacity, exchange handles and attach again */
bool  DN_b200_peer_capacity_ok(DNvolume* vol);
/* every pixel DN_draw writes into `fb` is also stored to `mirrorImage` (same size, usually the root replica's framebuffer
 * in peer memory); NULL = off.  The root must not read a mirrored framebuffer before the post-draw barrier. */
bool  DN_b200_framebuffer_set_mirror(GLuint fb, void* mirrorImage);
/* DN_B200_PEER_MANUAL only: visible |= every peer's visible (call once every replica has drawn) */
bool  DN_b200_peer_exchange_visible(DNvolume* vol);
/* epochs completed and time-outs seen by this replica's device barrier (a time-out is also reported through the callback) */
bool  DN_b200_peer_barrier_status(DNvolume* vol, uint64_t* epochs, uint32_t* timeouts);

#ifdef __cplusplus
}
#endif

#endif
