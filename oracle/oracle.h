/* oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the DoonEngine per-voxel lighting + ray-cast draw path.
 * Nothing in the product (doonengine_b200/, include/) may include, link or
 * execute anything under oracle/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * The oracle works on buffers in the REFERENCE's GPU layout (the SSBOs the
 * GLSL shaders bind), so that the bytes produced by the reference's own host
 * code (oracle/_ref, "Oracle A") can be fed to it unchanged:
 *
 *   binding 0  map[]     OrbHandle   12 B   assets/shaders/voxelShared.comp:49-54,74-77
 *   binding 1  chunks[]  OrbChunk    96 B   voxelShared.comp:39-46,80-83 ; voxel.c:25-34
 *   binding 2  materials OrbMaterial 32 B   voxelShared.comp:57-69,86-89 ; voxel.h:82-94
 *   binding 3  requests  uint32             voxelLighting.comp:7-10
 *   binding 4  voxels[]  OrbVoxel    16 B   voxelShared.comp:30-36,92-95 ; voxel.c:16-22
 *
 * Parity status: PINNED, both halves, by the reference's own code compiled where it lies (oracle/Makefile; outputs only under
 * oracle/_ref/, git-ignored; exists only where /root/reference does):
 *   host half    oracle/_ref/libdoon_ref.so = the reference's voxel.c + glad.c behind the fake-GL shim (oracle/fake_gl.c);
 *                tests/test_oracle_vs_reference.py: host_cpu.c produces the same buffers, requests, uniforms and matrices, bit for bit.
 *   shader half  oracle/_ref/libglsl_ref.so = the reference's assets/shaders/voxel{Shared,Lighting,Draw}.comp compiled AS C++
 *                (oracle/glsl/translate.py does the syntax, glsl_compat.h the types and built-ins, glsl_harness.cpp the dispatch);
 *                tests/test_glsl_pin.py: shader_cpu.c produces the same pixels, first hits, lit words, visible flags and sample
 *                counts, bit for bit, over the demo map, a glass / edit / lightingSplit scene and the three synthetic maps.
 *   fixtures     tests/golden/ (the .npz files) are written by the two together (reference host dispatching reference shaders, nothing
 *                restated: tests/golden/make_golden.py) and checked wherever the suite runs, incl. the GPU box.
 * The reference ships no tests, golden vectors or KATs of its own.  What remains DEFINED rather than pinned
 * are the points where GLSL itself leaves the result to the implementation or the shaders race -- rules N1-N11 below; both
 * implementations (and the CUDA kernels) follow the same rules, so they can be compared bit for bit.
 *
 * Determinism rules added where the reference is racy / implementation
 * defined:
 *   N1  rays read voxel records as they were BEFORE the lighting dispatch.
 *   N2  every work-group sees the pre-dispatch numIndirectSamples; it is
 *       incremented once per chunk afterwards.
 *   N3  visible bit: lighting first clears it for every chunk that had at
 *       least one live invocation, then ORs in specular propagations whose
 *       source chunk was visible pre-dispatch.  Draw ORs.
 *   N4  lastUsed = 0 writes are performed (benign).
 *   N5  sin() = glibc sinf on the host.
 *   N6  normalize(v) = v * (1/sqrt(dot(v,v))), IEEE div and sqrt; dot is
 *       x*x + y*y + z*z left to right; no FMA contraction anywhere;
 *       min/max = IEEE minNum/maxNum (fminf/fmaxf); round = nearest-even;
 *       reflect = I - 2*dot(N,I)*N; refract per the GLSL spec formula;
 *       mix(a,b,t) = a*(1-t) + b*t; pow = powf; sign(NaN) = 0.
 *       vec3(mask)*x in iterate_DDA is evaluated as a select (mask ? x : 0),
 *       which differs from the literal product only for infinite deltaDist
 *       (a ray direction component that is exactly 0), where GLSL is undefined.
 *   N7  alpha of pixels whose ray misses the map box is written as -1.
 *   N8  the `voxel` out-parameter keeps its previous value when not written.
 *   N9  `time` is supplied by the harness.
 *   N11 (added) every DDA loop carries an iteration guard (ORB_MAX_*_STEPS);
 *       the reference would spin forever on NaN directions.  A ray that
 *       trips the guard is a miss.
 */
#ifndef DN_ORACLE_H
#define DN_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrbHandle   { uint32_t flags, lastUsed, voxelIndex; } OrbHandle;
typedef struct OrbChunk    { int32_t pos[3]; uint32_t numIndirectSamples; uint32_t partialCounts[3]; uint32_t bitMask[16]; uint32_t pad; } OrbChunk;
typedef struct OrbVoxel    { uint32_t normal, albedo, specLight, diffuseLight; } OrbVoxel;
typedef struct OrbMaterial { float pad[2]; uint32_t emissive; float opacity, refractIndex, specular; uint32_t reflectType, shininess; } OrbMaterial;

/* every uniform the two programs read (voxel.c:856-876 and voxel.c:922-947) */
typedef struct OrbUniforms
{
	uint32_t mapSize[3];
	uint32_t useCubemap;              /* must be 0: cubemap sampling is out of scope */
	float    skyGradientBot[3];
	float    skyGradientTop[3];
	float    sunStrength[3];
	float    ambientStrength[3];
	/* draw only */
	uint32_t viewMode;
	uint32_t composeRasterized;       /* must be 0 */
	float    invViewMat[16];          /* column-major, m[col][row] flattened */
	float    invCenteredViewMat[16];
	float    invProjectionMat[16];
	/* lighting only */
	float    time;
	uint32_t numDiffuseSamples;
	uint32_t maxDiffuseSamples;
	uint32_t diffuseBounceLimit;
	uint32_t specularBounceLimit;
	float    sunDir[3];               /* already normalised on the host, voxel.c:942 */
	float    shadowSoftness;
	float    camPos[3];
} OrbUniforms;

typedef struct OrbBuffers
{
	OrbHandle*         map;
	OrbChunk*          chunks;
	OrbVoxel*          voxels;
	const OrbMaterial* materials;     /* 256 entries */
} OrbBuffers;

/* traversal counters used for the ALGORITHMIC byte count */
typedef struct OrbCounters
{
	uint64_t rays;        /* step_map calls */
	uint64_t tiles;       /* T: map tiles visited (get_map_tile calls) */
	uint64_t chunks;      /* C: loaded chunks entered (step_chunk calls) */
	uint64_t voxelSteps;  /* voxel-level DDA iterations */
	uint64_t records;     /* H: voxel records fetched inside step_chunk */
	uint64_t voxelsLit;   /* V: live lighting invocations */
	uint64_t pixels;      /* P: pixels written */
} OrbCounters;

/* per-pixel first-hit record written by orb_draw */
typedef struct OrbHit
{
	int32_t  status;      /* 0 = ray misses the map box, 1 = enters the box but hits nothing, 2 = hit */
	uint32_t mapIndex;    /* tile of the hit voxel (flattened) */
	uint32_t localIndex;  /* x + 8*(y + 8*z) inside the chunk */
	uint32_t recordIndex; /* index into voxels[] */
} OrbHit;

#define ORB_EPSILON 0.0001f
#define ORB_MAX_CHUNK_STEPS 4096   /* N11 */

/* voxelDraw.comp main() over (w/16)x(h/16) work-groups of 16x16.
 * image: h*w*4 floats, row-major, row 0 = screen y -1. Pixels outside the
 * dispatched area are left untouched.  hits may be NULL. */
void orb_draw(const OrbBuffers* buf, const OrbUniforms* u, int w, int h, float* image, OrbHit* hits, OrbCounters* counters);

/* voxelLighting.comp main() over numRequests work-groups of 32, with the
 * snapshot semantics N1-N3.  Updates voxels[], chunks[].numIndirectSamples
 * and map[].flags in place. */
void orb_light(const OrbBuffers* buf, const OrbUniforms* u, const uint32_t* requests, size_t numRequests, size_t numVoxelRecords, OrbCounters* counters);

/* the two phases orb_light is made of, exposed so that a dispatch can be split over request ranges (multi-GPU
 * sharding emulation in tests/): compute writes 96 staged words per request and ORs propagate[]; commit applies. */
void orb_light_compute(const OrbBuffers* buf, const OrbUniforms* u, const uint32_t* requests, size_t first, size_t count, uint32_t* staging, uint8_t* propagate, OrbCounters* counters);
void orb_light_commit(const OrbBuffers* buf, const OrbUniforms* u, const uint32_t* requests, size_t numRequests, const uint32_t* staging, const uint8_t* propagate);

/* helpers exposed for unit tests */
uint32_t orb_get_voxel_index(const OrbBuffers* buf, uint32_t mapIndex, int x, int y, int z);
void     orb_get_voxel_position(const OrbBuffers* buf, uint32_t chunk, uint32_t voxNum, int out[3]);
float    orb_rand(float seed);
void     orb_rand_unit_sphere(float seed, float out[3]);
int      orb_num_threads(void);
void     orb_set_num_threads(int n); /* n <= 0: every online processor (ignores OMP_NUM_THREADS) */

#ifdef __cplusplus
}
#endif

#endif
