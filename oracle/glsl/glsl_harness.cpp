/* glsl_harness.cpp -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Runs the reference's OWN compute shaders on the CPU: the text of assets/shaders/voxel{Shared,Lighting,Draw}.comp is compiled as C++
 * (oracle/glsl/translate.py makes it syntactically C++ at build time, oracle/glsl/glsl_compat.h supplies the GLSL types and
 * built-ins), one object per shader invocation, dispatched over the same buffers and uniforms as oracle/shader_cpu.c's orb_draw /
 * orb_light.  This is what PINS the hand restatement shader_cpu.c: tests/test_glsl_pin.py drives both over the same maps and
 * compares every pixel, first hit, lit word, visible bit and sample count bit for bit.
 *
 * The shaders are racy where oracle.h says so; the buffer proxies below give one dispatch the determinism rules N1-N3 WITHOUT touching
 * the shader text:
 *   voxels[i]        field stores are held back in a side array and applied after the dispatch, so every read sees the pre-dispatch
 *                    value (N1) -- without copying the whole pool per dispatch, which would cost more than the dispatch itself
 *   chunks[i]        numIndirectSamples reads see the pre-dispatch value, `++` is applied once after the dispatch (N2)
 *   map[i].flags     reads see the pre-dispatch value; `&= ~4` / `|= 4` are recorded and applied after the dispatch, clears first (N3);
 *                    `= 3` (a stream request, voxelShared.comp:464) is applied directly -- it never happens in resident mode
 *   map[i].lastUsed  live (N4: every writer stores 0)
 * The built library is oracle/_ref/libglsl_ref.so; it exists only where /root/reference does (oracle/Makefile).
 */
#include "glsl_compat.h"
#include "../oracle.h"

#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace glsl
{

struct Device
{
	const OrbBuffers*  buf;
	const OrbUniforms* u;
	OrbVoxel*          pending;      /* held-back voxel stores (lighting only), indexed like voxels[]; a record is valid once written */
	uint8_t*           setVisible;   /* per tile */
	uint8_t*           clearVisible; /* per tile */
	uint32_t*          sampleIncrements; /* per tile */
	const uint32_t*    requests;
	float*             image;
	int                w, h;
};

struct ShaderBase;

/* ---- voxels[] ---- */
struct VoxelField
{
	ShaderBase* inv;
	uint index;
	int field; /* word of the record */
	void operator=(uint v);
	operator uint() const;
};
struct VoxelRef
{
	ShaderBase* inv;
	uint index;
	VoxelField normal, albedo, specLight, diffuseLight;
	template <class T> operator T() const;
};
struct VoxelArr
{
	ShaderBase* inv;
	VoxelRef operator[](uint i) const;
};

/* ---- map[] ---- */
struct FlagsRef
{
	ShaderBase* inv;
	Device* d;
	uint index;
	operator uint() const { return __atomic_load_n(&d->buf->map[index].flags, __ATOMIC_RELAXED); } /* not written during the dispatch */
	void operator|=(uint m);
	void operator&=(uint m) { if(!(m & 4u)) d->clearVisible[index] = 1; if((~m) & ~4u) abort(); }
	void operator=(uint v) { (void)v; abort(); /* a stream request: impossible with every chunk resident */ }
};
struct LastUsedRef
{
	Device* d;
	uint index;
	operator uint() const { return __atomic_load_n(&d->buf->map[index].lastUsed, __ATOMIC_RELAXED); }
	void operator=(uint v) { __atomic_store_n(&d->buf->map[index].lastUsed, v, __ATOMIC_RELAXED); }
};
struct MapRef
{
	FlagsRef flags;
	LastUsedRef lastUsed;
	uint voxelIndex;
	template <class T> operator T() const
	{
		T h;
		h.flags = (uint)flags;
		h.lastUsed = (uint)lastUsed;
		h.voxelIndex = voxelIndex;
		return h;
	}
};
struct MapArr
{
	ShaderBase* inv;
	Device* d;
	MapRef operator[](uint i) const
	{
		MapRef r = {{inv, d, i}, {d, i}, d->buf->map[i].voxelIndex};
		return r;
	}
};

/* ---- chunks[] ---- */
struct SampleRef
{
	Device* d;
	uint index;
	operator uint() const { return d->buf->chunks[index].numIndirectSamples; }
	void operator++(int) { __atomic_fetch_add(&d->sampleIncrements[index], 1u, __ATOMIC_RELAXED); }
};
struct ChunkRef
{
	ivec3 pos;
	SampleRef numIndirectSamples;
	const uint* partialCounts;
	const uint* bitMask;
};
struct ChunkArr
{
	Device* d;
	ChunkRef operator[](uint i) const
	{
		const OrbChunk& c = d->buf->chunks[i];
		ChunkRef r = {ivec3(c.pos[0], c.pos[1], c.pos[2]), {d, i}, c.partialCounts, c.bitMask};
		return r;
	}
};

/* ---- materials[] (std140 block of 32-byte entries; `bool emissive` is a 32-bit word there) ---- */
struct MaterialRef
{
	const OrbMaterial* m;
	template <class T> operator T() const
	{
		T r;
		r.padding = vec2(m->pad[0], m->pad[1]);
		r.emissive = m->emissive != 0u;
		r.opacity = m->opacity;
		r.refractIndex = m->refractIndex;
		r.specular = m->specular;
		r.reflectType = m->reflectType;
		r.shininess = m->shininess;
		return r;
	}
};
struct MaterialArr
{
	Device* d;
	MaterialRef operator[](uint i) const { MaterialRef r = {d->buf->materials + i}; return r; }
};

struct image2D {};

/* everything the shader text refers to but does not define: uniforms, buffers, built-in variables, image functions */
struct ShaderBase
{
	Device* dev;
	/* voxelShared.comp */
	uvec3 mapSize;
	bool useCubemap;
	samplerCube skyCubemap;
	vec3 skyGradientBot, skyGradientTop;
	vec3 sunStrength, ambientStrength;
	MapArr map;
	ChunkArr chunks;
	MaterialArr materials;
	VoxelArr voxels;
	/* voxelLighting.comp */
	const uint* chunkIndices;
	float time;
	uint numDiffuseSamples, maxDiffuseSamples, diffuseBounceLimit, specularBounceLimit;
	vec3 sunDir;
	float shadowSoftness;
	vec3 camPos;
	/* voxelDraw.comp */
	image2D colorOutput;
	sampler2D colorSample, depthSample;
	uint viewMode;
	bool composeRasterized;
	mat4 invViewMat, invCenteredViewMat, invProjectionMat;
	/* built-in variables */
	uvec3 gl_WorkGroupID, gl_LocalInvocationID, gl_WorkGroupSize, gl_GlobalInvocationID;
	/* instrumentation (not visible to the shader text) */
	uint lastVoxelRead;
	int  visibleSetAt;
	uint writtenVoxel; /* the one voxel this invocation has stored to, 0xFFFFFFFF = none yet */

	explicit ShaderBase(Device* d) : dev(d)
	{
		const OrbUniforms* u = d->u;
		mapSize = uvec3(u->mapSize[0], u->mapSize[1], u->mapSize[2]);
		useCubemap = u->useCubemap != 0;
		skyGradientBot = vec3(u->skyGradientBot[0], u->skyGradientBot[1], u->skyGradientBot[2]);
		skyGradientTop = vec3(u->skyGradientTop[0], u->skyGradientTop[1], u->skyGradientTop[2]);
		sunStrength = vec3(u->sunStrength[0], u->sunStrength[1], u->sunStrength[2]);
		ambientStrength = vec3(u->ambientStrength[0], u->ambientStrength[1], u->ambientStrength[2]);
		map.inv = this; map.d = d; chunks.d = d; materials.d = d; voxels.inv = this;
		chunkIndices = d->requests;
		time = u->time;
		numDiffuseSamples = u->numDiffuseSamples; maxDiffuseSamples = u->maxDiffuseSamples;
		diffuseBounceLimit = u->diffuseBounceLimit; specularBounceLimit = u->specularBounceLimit;
		sunDir = vec3(u->sunDir[0], u->sunDir[1], u->sunDir[2]);
		shadowSoftness = u->shadowSoftness;
		camPos = vec3(u->camPos[0], u->camPos[1], u->camPos[2]);
		viewMode = u->viewMode;
		composeRasterized = u->composeRasterized != 0;
		memcpy(invViewMat.m, u->invViewMat, 64);
		memcpy(invCenteredViewMat.m, u->invCenteredViewMat, 64);
		memcpy(invProjectionMat.m, u->invProjectionMat, 64);
		gl_WorkGroupID = gl_LocalInvocationID = gl_GlobalInvocationID = uvec3(0, 0, 0);
		gl_WorkGroupSize = uvec3(1, 1, 1);
		lastVoxelRead = 0xFFFFFFFFu;
		visibleSetAt = -1;
		writtenVoxel = 0xFFFFFFFFu;
	}

	ivec2 imageSize(const image2D&) const { return ivec2(dev->w, dev->h); }
	void imageStore(const image2D&, const ivec2& p, const vec4& v) const
	{
		float* px = dev->image + ((size_t)p.y * (size_t)dev->w + (size_t)p.x) * 4;
		px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
	}
};

inline void FlagsRef::operator|=(uint m)
{
	if(m & 4u)
	{
		d->setVisible[index] = 1;
		inv->visibleSetAt = (int)index;
	}
	if(m & ~4u)
		abort();
}

inline VoxelRef VoxelArr::operator[](uint i) const
{
	VoxelRef r = {inv, i, {inv, i, 0}, {inv, i, 1}, {inv, i, 2}, {inv, i, 3}};
	return r;
}

inline VoxelField::operator uint() const
{
	return reinterpret_cast<const uint32_t*>(inv->dev->buf->voxels + index)[field];
}

inline void VoxelField::operator=(uint v)
{
	Device* d = inv->dev;
	if(!d->pending || (inv->writtenVoxel != 0xFFFFFFFFu && inv->writtenVoxel != index))
		abort(); /* the draw shader stores no voxel; a lighting invocation stores only its own */
	if(inv->writtenVoxel == 0xFFFFFFFFu)
	{
		d->pending[index] = d->buf->voxels[index];
		inv->writtenVoxel = index;
	}
	reinterpret_cast<uint32_t*>(d->pending + index)[field] = v;
}

template <class T> VoxelRef::operator T() const
{
	inv->lastVoxelRead = index;
	T r;
	r.normal = (uint)normal;
	r.albedo = (uint)albedo;
	r.specLight = (uint)specLight;
	r.diffuseLight = (uint)diffuseLight;
	return r;
}

#include "glsl_shaders.inc"

} // namespace glsl

using namespace glsl;

static void apply_flags(const OrbBuffers* buf, size_t tiles, const uint8_t* clearV, const uint8_t* setV, const uint32_t* inc)
{
	for(size_t i = 0; i < tiles; i++)
	{
		if(clearV && clearV[i])
			buf->map[i].flags &= ~4u;
		if(setV[i])
			buf->map[i].flags |= 4u;
		if(inc && inc[i])
			buf->chunks[i].numIndirectSamples += inc[i];
	}
}

/* voxelDraw.comp main() over (w/16) x (h/16) work-groups of 16 x 16 (voxel.c:879); hits may be NULL: status 2 = the pixel's ray hit an
 * opaque voxel (it executed `map[index].flags |= 4`, voxelDraw.comp:121), mapIndex = that tile, recordIndex = the last voxels[] element
 * the invocation read (the hit record); status 1 otherwise */
extern "C" void glsl_draw(const OrbBuffers* buf, const OrbUniforms* u, int w, int h, float* image, OrbHit* hits)
{
	const size_t tiles = (size_t)u->mapSize[0] * u->mapSize[1] * u->mapSize[2];
	std::vector<uint8_t> setV(tiles ? tiles : 1, 0);
	Device dev;
	memset(&dev, 0, sizeof(dev));
	dev.buf = buf; dev.u = u; dev.pending = nullptr; dev.setVisible = setV.data(); dev.clearVisible = nullptr; dev.sampleIncrements = nullptr;
	dev.image = image; dev.w = w; dev.h = h;
	const int gx = w / 16, gy = h / 16;
	#pragma omp parallel for schedule(dynamic, 1) collapse(2)
	for(int by = 0; by < gy; by++)
		for(int bx = 0; bx < gx; bx++)
			for(int ly = 0; ly < 16; ly++)
				for(int lx = 0; lx < 16; lx++)
				{
					DrawShader s(&dev);
					s.gl_WorkGroupID = uvec3((uint)bx, (uint)by, 0);
					s.gl_LocalInvocationID = uvec3((uint)lx, (uint)ly, 0);
					s.gl_WorkGroupSize = uvec3(16, 16, 1);
					s.gl_GlobalInvocationID = uvec3((uint)(bx * 16 + lx), (uint)(by * 16 + ly), 0);
					s.shader_main();
					if(hits)
					{
						OrbHit& hit = hits[(size_t)(by * 16 + ly) * (size_t)w + (size_t)(bx * 16 + lx)];
						hit.status = 1;
						hit.mapIndex = hit.localIndex = hit.recordIndex = 0;
						if(s.visibleSetAt >= 0)
						{
							hit.status = 2;
							hit.mapIndex = (uint32_t)s.visibleSetAt;
							hit.localIndex = 0xFFFFFFFFu;
							hit.recordIndex = s.lastVoxelRead;
						}
					}
				}
	apply_flags(buf, tiles, nullptr, setV.data(), nullptr);
}

/* voxelLighting.comp main() over numRequests work-groups of 32, snapshot semantics as described at the top */
extern "C" void glsl_light(const OrbBuffers* buf, const OrbUniforms* u, const uint32_t* requests, size_t numRequests, size_t numVoxelRecords)
{
	const size_t tiles = (size_t)u->mapSize[0] * u->mapSize[1] * u->mapSize[2];
	std::vector<uint8_t> setV(tiles ? tiles : 1, 0), clearV(tiles ? tiles : 1, 0);
	std::vector<uint32_t> inc(tiles ? tiles : 1, 0);
	/* untouched pages of this array are never mapped: only the records the dispatch stores to cost anything */
	OrbVoxel* pending = static_cast<OrbVoxel*>(malloc((numVoxelRecords ? numVoxelRecords : 1) * sizeof(OrbVoxel)));
	if(!pending)
		abort();
	Device dev;
	memset(&dev, 0, sizeof(dev));
	dev.buf = buf; dev.u = u; dev.pending = pending; dev.setVisible = setV.data(); dev.clearVisible = clearV.data(); dev.sampleIncrements = inc.data();
	dev.requests = requests;
	std::vector<std::vector<uint32_t>> written((size_t)omp_get_max_threads());
	#pragma omp parallel for schedule(dynamic, 4)
	for(long long g = 0; g < (long long)numRequests; g++)
		for(uint lane = 0; lane < 32; lane++)
		{
			LightingShader s(&dev);
			s.gl_WorkGroupID = uvec3((uint)g, 0, 0);
			s.gl_LocalInvocationID = uvec3(lane, 0, 0);
			s.gl_WorkGroupSize = uvec3(32, 1, 1);
			s.gl_GlobalInvocationID = uvec3((uint)g * 32u + lane, 0, 0);
			s.shader_main();
			if(s.writtenVoxel != 0xFFFFFFFFu)
				written[(size_t)omp_get_thread_num()].push_back(s.writtenVoxel);
		}
	for(const std::vector<uint32_t>& list : written)
		for(uint32_t i : list)
			buf->voxels[i] = pending[i];
	free(pending);
	apply_flags(buf, tiles, clearV.data(), setV.data(), inc.data());
}

extern "C" int glsl_ref_version(void) { return 1; }
