/* DoonEngine/voxel.h -- the DN_* volume API, exported by libdoon_b200.so (B200 / sm_100a CUDA back end).
 *
 * Drop-in for /root/reference/src/DoonEngine/voxel.h: the same 30 entry points (voxel.h:156-371), the same
 * public structs with identical field order, types and sizes (DNvoxel 20 B, DNcompressedVoxel 8 B,
 * DNchunk 4120 B, DNchunkHandle 8 B, DNvoxelNode 32 B, DNmaterial 32 B, DNvolume 232 B on x86-64), the same
 * enum values.  What differs is behind the struct:
 *
 *   - no OpenGL.  GLuint/GLfloat are plain uint32_t/float here.  The three gl*BufferID fields of DNvolume
 *     are opaque non-zero handles; `outputTexture` of DN_draw is a framebuffer handle made by
 *     DN_b200_create_framebuffer() (DoonEngine/b200.h), a linear RGBA32F image in device memory.
 *   - the map is RESIDENT: every chunk the CPU map holds is uploaded by DN_sync_gpu(DN_WRITE / DN_READ_WRITE);
 *     there is no demand streaming / LRU eviction (reference voxel.c:1554-1640), so `minChunks` only sizes
 *     the initial pools.  `gpuVoxelLayout` / `numVoxelNodes` are NULL / 0 except right after
 *     DN_b200_mirror_voxel_layout() (DoonEngine/b200.h), which builds a snapshot of this library's own record allocator.
 *   - the request list is built on the device and stays there: `lightingRequests` is mirrored to the host on demand
 *     (DN_b200_fetch_lighting_requests) and `numLightingRequests` is exact after DN_b200_lighting_request_count() -- or after
 *     every reading DN_sync_gpu once DN_b200_set_exact_sync(vol, true) has asked for upstream's blocking contract (b200.h).
 *   - raster composition (rasterColorTexture/rasterDepthTexture >= 0) and cubemap skies are not implemented:
 *     DN_draw reports DN_MESSAGE_ERROR and draws without them.
 *
 * Threading contract is the reference's: all DN_* calls of a process from one thread; every call that
 * exposes data is synchronous with respect to that data.
 */
#ifndef DN_VOXEL_H
#define DN_VOXEL_H

#include "globals.h"
#include "mathtypes.h"
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DN_NO_GL_TYPEDEFS
#ifndef __glad_h_
typedef uint32_t GLuint;
typedef float    GLfloat;
#endif
#endif

#define DN_CHUNK_SIZE 8          /* voxels along one edge of a chunk            (reference voxel.h:18) */
#define DN_CHUNK_LENGTH 512      /* voxels in a chunk                           (voxel.h:20) */
#define DN_MAX_MATERIALS 256     /*                                             (voxel.h:23) */
#define DN_MATERIAL_EMPTY 255    /* material id meaning "no voxel"              (voxel.h:25) */
#define DN_GAMMA 2.2f            /* sRGB -> linear exponent applied at upload   (voxel.h:28) */

/* x + sx * (y + z * sy)  (voxel.h:31; unparenthesised upstream, parenthesised here) */
#define DN_FLATTEN_INDEX(p, s) ((p.x) + (s.x) * ((p.y) + (p.z) * (s.y)))

typedef struct DNcolor { uint8_t r, g, b; } DNcolor;

typedef struct DNvoxel
{
	uint8_t material; /* 0..254; 255 = empty */
	DNvec3  normal;
	DNcolor albedo;   /* sRGB */
} DNvoxel;

typedef struct DNcompressedVoxel
{
	uint32_t normal; /* material<<24 | nx<<16 | ny<<8 | nz, n* = (int(n*255)+255)/2 */
	uint32_t albedo; /* r<<24 | g<<16 | b<<8 */
} DNcompressedVoxel;

typedef struct DNchunk
{
	DNivec3  pos;          /* tile position, (-1,-1,-1) when the slot is unused */
	bool     updated;      /* edits not yet pushed to the device */
	uint32_t numVoxels;    /* solid voxels */
	uint32_t numVoxelsGpu; /* surface voxels resident on the device after the last upload */
	DNcompressedVoxel voxels[DN_CHUNK_SIZE][DN_CHUNK_SIZE][DN_CHUNK_SIZE]; /* [x][y][z] */
} DNchunk;

typedef struct DNchunkHandle
{
	uint8_t  flag;       /* 0 = no chunk, 1 = chunk present in `chunks` */
	uint32_t chunkIndex;
} DNchunkHandle;

typedef struct DNvoxelNode
{
	uint32_t size;     /* records, power of two in [16, 512] */
	size_t   startPos; /* first record */
	DNivec3  chunkPos; /* owner tile, x = -1 when free */
} DNvoxelNode;

typedef struct DNmaterial
{
	DNvec2  padding;
	GLuint  emissive;
	GLfloat opacity;
	GLfloat refractIndex;
	GLfloat specular;
	GLuint  reflectType; /* 0 = no sky reflection, 1 = reflects the sky, 2 = highlight only */
	GLuint  shininess;
} DNmaterial;

typedef struct DNvolume
{
	GLuint glMapBufferID;   /* opaque device handles (read only) */
	GLuint glChunkBufferID;
	GLuint glVoxelBufferID;

	DNuvec3 mapSize;             /* tiles; read only */
	size_t  chunkCap;            /* length of `chunks` */
	size_t  nextChunk;
	size_t  voxelCap;            /* device record pool capacity */
	size_t  numVoxelNodes;
	size_t  numLightingRequests; /* 32-voxel groups queued by the last reading DN_sync_gpu */
	size_t  lightingRequestCap;

	DNchunkHandle* map;       /* mapSize.x*y*z handles, index DN_FLATTEN_INDEX */
	DNchunk*       chunks;
	DNmaterial*    materials; /* DN_MAX_MATERIALS entries, read-write */
	GLuint*        lightingRequests;
	DNvoxelNode*   gpuVoxelLayout;

	DNvec3   camPos;      /* in tiles; read-write */
	DNvec3   camOrient;   /* degrees: pitch, yaw, roll */
	float    camFOV;      /* degrees */
	uint32_t camViewMode; /* 0 lit, 1 albedo, 2 diffuse, 3 specular, 4 voxel normal, 5 face normal */

	DNvec3   sunDir;
	DNvec3   sunStrength;
	DNvec3   ambientLightStrength;
	uint32_t diffuseBounceLimit;
	uint32_t specBounceLimit;
	float    shadowSoftness;

	bool   useCubemap;
	GLuint glCubemapTex;
	DNvec3 skyGradientBot;
	DNvec3 skyGradientTop;

	uint32_t frameNum; /* lighting-split phase, read only */
	float    lastTime; /* time latched at phase 0, read only */
} DNvolume;

typedef enum DNmemOp { DN_READ = 0, DN_WRITE = 1, DN_READ_WRITE = 2 } DNmemOp;

/* ---- lifecycle (reference voxel.h:156-184) ---- */
bool      DN_init(void);   /* selects the CUDA device, creates streams; false + FATAL message if no usable GPU */
void      DN_quit(void);
DNvolume* DN_create_volume(DNuvec3 mapSize, unsigned int minChunks);
void      DN_delete_volume(DNvolume* vol);
DNvolume* DN_load_volume(const char* filePath, unsigned int minChunks);
bool      DN_save_volume(const char* filePath, DNvolume* vol);

/* ---- frame: call order per frame is draw -> sync -> update_lighting (reference main.c:503-505) ---- */
void DN_set_view_projection_matrices(DNvolume* vol, float aspectRatio /* height / width */, float nearPlane, float farPlane, DNmat4* view, DNmat4* projection);
void DN_draw(DNvolume* vol, GLuint outputTexture, DNmat4 view, DNmat4 projection, int rasterColorTexture, int rasterDepthTexture);
void DN_update_lighting(DNvolume* vol, int numDiffuseSamples, int maxDiffuseSamples, float time);
void DN_sync_gpu(DNvolume* vol, DNmemOp op, int lightingSplit);

/* ---- capacity (voxel.h:225-267) ---- */
int  DN_add_chunk(DNvolume* vol, DNivec3 pos);
void DN_remove_chunk(DNvolume* vol, DNivec3 pos);
bool DN_set_map_size(DNvolume* vol, DNuvec3 size);
bool DN_set_max_chunks(DNvolume* vol, size_t num);
bool DN_set_max_voxels_gpu(DNvolume* vol, size_t num);
bool DN_set_max_lighting_requests(DNvolume* vol, size_t num);

/* ---- voxels (voxel.h:277-344); no bounds checking, as upstream ---- */
bool              DN_in_map_bounds(DNvolume* vol, DNivec3 pos);
bool              DN_in_chunk_bounds(DNivec3 pos);
DNvoxel           DN_get_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos);
DNcompressedVoxel DN_get_compressed_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos);
void              DN_set_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos, DNvoxel voxel);
void              DN_set_compressed_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos, DNcompressedVoxel voxel);
void              DN_remove_voxel(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos);
bool              DN_does_chunk_exist(DNvolume* vol, DNivec3 pos);
bool              DN_does_voxel_exist(DNvolume* vol, DNivec3 mapPos, DNivec3 chunkPos);
bool              DN_step_map(DNvolume* vol, DNvec3 rayDir, DNvec3 rayPos, int maxSteps, DNivec3* hitPos, DNvoxel* hitVoxel, DNivec3* hitNormal);

/* ---- utility (voxel.h:354-371) ---- */
void              DN_separate_position(DNivec3 pos, DNivec3* mapPos, DNivec3* chunkPos);
DNvec3            DN_cam_dir(DNvec3 orient);
DNcompressedVoxel DN_compress_voxel(DNvoxel voxel);
DNvoxel           DN_decompress_voxel(DNcompressedVoxel voxel);

#ifdef __cplusplus
}
#endif

#endif
