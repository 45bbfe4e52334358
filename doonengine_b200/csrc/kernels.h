/* kernels.h -- launch wrappers of the CUDA kernels (draw.cu, light.cu, compact.cu, upload.cu). */
#ifndef DN_B200_KERNELS_H
#define DN_B200_KERNELS_H

#include "layout.h"
#include <cuda_runtime.h>

/* one edited chunk in an upload batch (upload.cu) */
typedef struct DnbUploadItem
{
	uint32_t mapIndex;
	uint32_t slotPlus1;    /* 0 = the tile lost its chunk */
	uint32_t recordOffset; /* first record of this chunk inside the batch's record blob */
	uint32_t pad;
} DnbUploadItem;

#ifdef __cplusplus
extern "C" {
#endif

/* kernels this library has launched since it was loaded (every launch wrapper counts; read with DN_b200_kernel_launches) */
extern unsigned long long g_dnbKernelLaunches;
#define DNB_LAUNCHED(n) (g_dnbKernelLaunches += (n))

/* mirror: second destination of every pixel (peer memory of the root replica), or NULL */
cudaError_t dnb_launch_draw(const DnbScene* scene, const DnbDrawParams* params, float4* image, float4* mirror, DnbHit* hits, cudaStream_t stream);

cudaError_t dnb_upload_light_params(const DnbLightParams* params, cudaStream_t stream);
/* lights this launch's share of the request list (DnbWork, layout.h: the count is read on the device) and stores the staged words of
 * request r at word 96 r of every array in `targets`.  gridCtas: CTAs to launch (the warp-per-request kernel strides over its share,
 * so any number >= 1 is correct) / the ceiling of the persistent kernel's grid.
 * flatCounter: NULL = dn_light_kernel (one warp per request); else a device word used as the work counter of the persistent
 * state-machine kernel dn_light_flat_kernel (light_flat.cuh).  Both give identical results. */
cudaError_t dnb_launch_light(const DnbScene* scene, const uint32_t* requests, const DnbWork* work, uint32_t gridCtas,
                             const DnbStagingTargets* targets, uint32_t* flatCounter, cudaStream_t stream);
/* the same dispatch as a wavefront over a pool of P context slots in global memory (light_wave.cuh): ctx = dnb_wave_slot_bytes() * P
 * bytes, P a multiple of 256; counters = 8 device words.  Queues its passes on `stream` and returns once the last voxel is staged
 * (the host follows the live-slot count a few passes behind); passesOut = passes queued. */
size_t      dnb_wave_slot_bytes(void);
cudaError_t dnb_launch_light_wave(const DnbScene* scene, const uint32_t* requests, const DnbWork* work,
                                  const DnbStagingTargets* targets, uint4* ctx, uint32_t P, uint32_t* counters, uint32_t* passesOut, cudaStream_t stream);
/* copies the staged rows of this launch's CTAs from replica `self`'s staging array (peers->dst[self]) into every other replica's
 * (coalesced 16-byte stores over NVLink); used after the wavefront kernels, which stage locally */
cudaError_t dnb_launch_push_staging(const DnbStagingTargets* peers, uint32_t self, const DnbWork* work, uint32_t gridCtas, cudaStream_t stream);
/* peers: NULL, or the table whose propagate bitmaps (all replicas') are ORed into visible instead of only the local one.
 * boundRequests: what the host knows the request count cannot exceed (sizes the persistent grid; 0 = no requests) */
cudaError_t dnb_launch_commit(const DnbScene* scene, DnbSlot* slots, uint4* records, const uint32_t* requests, const DnbWork* work, uint32_t boundRequests, const uint32_t* staging,
                              unsigned long long* litCounter, const DnbPeerTable* peers, cudaStream_t stream);

/* pick.cu: `count` DN_step_map rays (voxel.c:1195-1272) on the device map; out[i] = {cell, code} (see dn_pick_kernel) */
cudaError_t dnb_launch_pick(const DnbScene* scene, const float* dirs, const float* origins, uint32_t count, int maxSteps, int4* out, cudaStream_t stream);

/* peer.cu: device-side barrier over the replicas' mailboxes, and visible |= every peer's visible */
cudaError_t dnb_launch_peer_barrier(const DnbPeerTable* peers, uint32_t epoch, uint32_t* status, cudaStream_t stream);
cudaError_t dnb_launch_peer_or_visible(const DnbPeerTable* peers, uint32_t* visible, uint32_t words, cudaStream_t stream);

/* upload.cu: re-derives DNB_BBOX_OPAQUE of slots[0, numSlots) from their matIds and the 256-bit table opaqueBits (layout.h) */
cudaError_t dnb_launch_refresh_opaque(DnbSlot* slots, uint32_t numSlots, const uint32_t opaqueBits[8], cudaStream_t stream);

uint32_t    dnb_compact_num_blocks(uint32_t numTiles);
/* hostTotal: NULL or a device-accessible pinned host word that also receives the total (zero-copy store) */
cudaError_t dnb_launch_compact_count(const DnbScene* scene, const uint32_t* forced, uint32_t split, uint32_t frameNum, uint32_t* blockCounts, uint32_t* blockOffsets, uint32_t* grandTotal,
                                     uint32_t* hostTotal, cudaStream_t stream);
cudaError_t dnb_launch_compact_write(const DnbScene* scene, const uint32_t* forced, uint32_t split, uint32_t frameNum, const uint32_t* blockOffsets, uint32_t* requests, cudaStream_t stream);
cudaError_t dnb_launch_set_bits(uint32_t* bits, const uint32_t* tiles, uint32_t n, cudaStream_t stream);

cudaError_t dnb_launch_scatter(const DnbUploadItem* items, const DnbSlot* headers, const uint4* blobRecords, uint32_t numItems, const uint32_t mapSize[3], const uint32_t blocks[3],
                               uint32_t* tileSlot, unsigned long long* occ64, uint32_t* visible, DnbSlot* slots, uint4* records, cudaStream_t stream);

#ifdef __cplusplus
}
#endif

#endif
