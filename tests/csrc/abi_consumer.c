/* A C consumer of the public headers (tests/test_host_hygiene.py).  Compiled twice by gcc as plain C:
 *   - against the headers under include/DoonEngine and linked with libdoon_b200.so: prints the layout table and drives a host-only volume
 *     (no CUDA device needed) through the DN_* calls, so a typo in a header shows up as a compile error or a wrong number;
 *   - with -DUSE_REFERENCE_HEADER against the reference's own voxel.h where /root/reference exists (not linked, layout table only):
 *     the two tables must be identical line for line. */
#ifdef USE_REFERENCE_HEADER
#include "DoonEngine/voxel.h"
#else
#include "DoonEngine/voxel.h"
#include "DoonEngine/b200.h"
#endif
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#define SZ(T) printf("sizeof(" #T ") = %zu\n", sizeof(T))
#define OFF(T, f) printf("offsetof(" #T ", " #f ") = %zu\n", offsetof(T, f))

static void layout_table(void)
{
	SZ(DNcolor); SZ(DNvoxel); SZ(DNcompressedVoxel); SZ(DNchunk); SZ(DNchunkHandle); SZ(DNvoxelNode); SZ(DNmaterial); SZ(DNvolume);
	SZ(DNvec3); SZ(DNivec3); SZ(DNuvec3); SZ(DNmat4);
	OFF(DNvoxel, material); OFF(DNvoxel, normal); OFF(DNvoxel, albedo);
	OFF(DNchunk, pos); OFF(DNchunk, updated); OFF(DNchunk, numVoxels); OFF(DNchunk, numVoxelsGpu); OFF(DNchunk, voxels);
	OFF(DNchunkHandle, flag); OFF(DNchunkHandle, chunkIndex);
	OFF(DNvoxelNode, size); OFF(DNvoxelNode, startPos); OFF(DNvoxelNode, chunkPos);
	OFF(DNmaterial, emissive); OFF(DNmaterial, opacity); OFF(DNmaterial, refractIndex); OFF(DNmaterial, specular); OFF(DNmaterial, reflectType); OFF(DNmaterial, shininess);
	OFF(DNvolume, glMapBufferID); OFF(DNvolume, glChunkBufferID); OFF(DNvolume, glVoxelBufferID); OFF(DNvolume, mapSize); OFF(DNvolume, chunkCap); OFF(DNvolume, nextChunk);
	OFF(DNvolume, voxelCap); OFF(DNvolume, numVoxelNodes); OFF(DNvolume, numLightingRequests); OFF(DNvolume, lightingRequestCap); OFF(DNvolume, map); OFF(DNvolume, chunks);
	OFF(DNvolume, materials); OFF(DNvolume, lightingRequests); OFF(DNvolume, gpuVoxelLayout); OFF(DNvolume, camPos); OFF(DNvolume, camOrient); OFF(DNvolume, camFOV);
	OFF(DNvolume, camViewMode); OFF(DNvolume, sunDir); OFF(DNvolume, sunStrength); OFF(DNvolume, ambientLightStrength); OFF(DNvolume, diffuseBounceLimit);
	OFF(DNvolume, specBounceLimit); OFF(DNvolume, shadowSoftness); OFF(DNvolume, useCubemap); OFF(DNvolume, glCubemapTex); OFF(DNvolume, skyGradientBot);
	OFF(DNvolume, skyGradientTop); OFF(DNvolume, frameNum); OFF(DNvolume, lastTime);
	printf("DN_READ = %d, DN_WRITE = %d, DN_READ_WRITE = %d\n", (int)DN_READ, (int)DN_WRITE, (int)DN_READ_WRITE);
	printf("DN_MESSAGE_CPU_MEMORY = %d, DN_MESSAGE_GPU_MEMORY = %d, DN_MESSAGE_SHADER = %d, DN_MESSAGE_FILE_IO = %d\n", (int)DN_MESSAGE_CPU_MEMORY, (int)DN_MESSAGE_GPU_MEMORY,
	       (int)DN_MESSAGE_SHADER, (int)DN_MESSAGE_FILE_IO);
	printf("DN_MESSAGE_NOTE = %d, DN_MESSAGE_ERROR = %d, DN_MESSAGE_FATAL = %d\n", (int)DN_MESSAGE_NOTE, (int)DN_MESSAGE_ERROR, (int)DN_MESSAGE_FATAL);
	printf("DN_CHUNK_SIZE = %d, DN_CHUNK_LENGTH = %d, DN_MAX_MATERIALS = %d, DN_MATERIAL_EMPTY = %d\n", DN_CHUNK_SIZE, DN_CHUNK_LENGTH, DN_MAX_MATERIALS, DN_MATERIAL_EMPTY);
}

#ifndef USE_REFERENCE_HEADER
static int g_messages = 0;
static void on_message(DNmessageType type, DNmessageSeverity severity, const char* text)
{
	(void)type; (void)severity; (void)text;
	g_messages++;
}

static int drive(const char* tmpPath)
{
	g_DN_message_callback = on_message;
	DNuvec3 size = {4, 3, 2};
	DNvolume* vol = DN_create_volume(size, 8); /* no DN_init: a host-only volume (edit / save / load for offline tools) */
	if(!vol || vol->mapSize.x != 4 || vol->mapSize.z != 2 || vol->gpuVoxelLayout != NULL || vol->numVoxelNodes != 0)
		return 1;
	DNivec3 tile = {1, 2, 1}, cell = {3, 4, 5};
	DNvoxel vx;
	memset(&vx, 0, sizeof(vx));
	vx.material = 7; vx.normal.x = 0.0f; vx.normal.y = 1.0f; vx.normal.z = 0.0f; vx.albedo.r = 10; vx.albedo.g = 200; vx.albedo.b = 30;
	DN_set_voxel(vol, tile, cell, vx);
	if(!DN_does_chunk_exist(vol, tile) || !DN_does_voxel_exist(vol, tile, cell))
		return 2;
	DNvoxel back = DN_get_voxel(vol, tile, cell);
	if(back.material != 7 || back.albedo.g != 200 || back.normal.y != 1.0f)
		return 3;
	DNvec3 dir = {{0.0f, -1.0f, 0.0f}}, from = {{1.0f + 3.5f / 8.0f, 2.99f, 1.0f + 5.5f / 8.0f}};
	DNivec3 hitPos, hitNormal;
	DNvoxel hitVoxel;
	if(!DN_step_map(vol, dir, from, 64, &hitPos, &hitVoxel, &hitNormal) || hitPos.x != 11 || hitPos.y != 20 || hitPos.z != 13 || hitNormal.y != 1 || hitVoxel.material != 7)
		return 4;
	unsigned char slot[128];
	unsigned char records[512 * 16];
	if(DN_b200_pack_chunk(vol, tile, slot, records) != 1)
		return 5;
	if(!DN_save_volume(tmpPath, vol))
		return 6;
	DN_delete_volume(vol);
	vol = DN_load_volume(tmpPath, 8);
	if(!vol || !DN_does_voxel_exist(vol, tile, cell) || DN_get_voxel(vol, tile, cell).albedo.b != 30)
		return 7;
	if(!DN_b200_mirror_voxel_layout(vol) || vol->numVoxelNodes != 0 || vol->gpuVoxelLayout == NULL)
		return 8; /* nothing is resident on a host-only volume: an empty, but non-NULL, mirror */
	DN_remove_voxel(vol, tile, cell);
	if(DN_does_voxel_exist(vol, tile, cell))
		return 9;
	DN_delete_volume(vol);
	return 0;
}
#endif

int main(int argc, char** argv)
{
	layout_table();
#ifndef USE_REFERENCE_HEADER
	if(argc > 1)
	{
		const int rc = drive(argv[1]);
		printf("drive = %d\n", rc);
		return rc;
	}
#else
	(void)argc; (void)argv;
#endif
	return 0;
}
