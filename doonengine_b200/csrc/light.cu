/* light.cu -- the per-voxel lighting update, in two phases so that it is deterministic and shardable.
 *
 * Phase 1  dn_light_kernel   one warp per lighting request (= 32 consecutive surface voxels of one chunk,
 *                            voxelLighting.comp:3,212-214), one thread per voxel.  Traces the 15 fixed
 *                            specular rays, the diffuse paths and the shadow rays (LI:65-203,224-264),
 *                            forms the running mean and packs the three lighting words (LI:266-278) into a
 *                            STAGING buffer in request order.  It reads the map strictly read-only, so every
 *                            ray sees the pre-dispatch lighting (oracle.h N1/N2) and a slice of the request
 *                            list can run on any GPU.
 * Phase 2  dn_commit_kernel  scatters the staged words into the voxel records (coalesced 16-byte stores),
 *                            clears the visible bit of every chunk that had a live invocation (LI:281) and
 *                            bumps the chunk's sample count once (LI:284-285).
 *          dn_merge_visible_kernel  ORs in the visible bits raised by specular hits (LI:101-105) after the
 *                            clears (N3).
 *
 * Staging layout: request r owns 96 words: [32 x w1][32 x w2][32 x w3] -> three fully coalesced 128-byte
 * stores per warp, and a contiguous byte range per rank for the multi-GPU all-gather.
 *
 * Multi-GPU over peer memory (DoonEngine/b200.h): the lighting kernel is given the staging array of EVERY replica
 * (peer pointers mapped over NVLink) and stores each warp's three 128-byte rows into all of them as soon as the
 * warp has finished tracing -- the exchange is fused into the kernel and overlaps the ray tracing of the other
 * warps; no all-gather follows.  Replica `rank` of `world` runs the CTAs rank, rank + world, ... (interleaved, so
 * expensive and cheap regions of the map spread evenly).
 */
#include "kernels.h"
#include "trace.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <cuda_pipeline.h>

/* LI:23 */
__constant__ float c_spherePoints[15][3] = {
	{0.000000f, 1.000000f, 0.000000f}, {-0.379803f, 0.857143f, 0.347931f}, {0.061185f, 0.714286f, -0.697174f},
	{0.499316f, 0.571429f, 0.651270f}, {-0.889696f, 0.428571f, -0.157375f}, {0.808584f, 0.285714f, -0.514354f},
	{-0.256942f, 0.142857f, 0.955810f}, {-0.460906f, 0.000000f, -0.887449f}, {0.929687f, -0.142857f, 0.339521f},
	{-0.885815f, -0.285714f, 0.365650f}, {0.382949f, -0.428571f, -0.818338f}, {0.245607f, -0.571429f, 0.783037f},
	{-0.605521f, -0.714286f, -0.350913f}, {0.503065f, -0.857143f, -0.110596f}, {-0.000000f, -1.000000f, 0.000000f}};

/* per-dispatch parameters incl. the host-evaluated random table (layout.h) */
__constant__ DnbLightParams c_light;

struct LightCtx
{
	RayState    st;
	DnbCounters lc;
	bool        firstSample;    /* LI:62 */
	bool        sourceVisible;  /* visible bit of the chunk being lit, pre-dispatch (N3) */
};

/* where a ray's light goes.  The shader adds every contribution straight into the voxel's running sum (`color += ...`, LI:78,117,
 * 123,139,143,186,197); AccSum does exactly that.  AccList keeps the addends apart instead, so that rays traced by DIFFERENT lanes
 * (light_spread.cuh) can be replayed into the sum afterwards in the shader's order -- float addition is not associative, the order
 * is part of the result. */
struct AccSum
{
	f3 c;
	DNB_FN void add(f3 a) { c = c + a; }
	/* LI:101-105: a visible chunk makes the chunks it reflects visible */
	DNB_FN void hit_tile(const DnbScene& S, uint32_t hitIndex)
	{
		const uint32_t bit = 1u << (hitIndex & 31u);
		if(!(__ldcg(S.propagate + (hitIndex >> 5)) & bit))
			atomicOr(S.propagate + (hitIndex >> 5), bit);
	}
};
#define DNB_MAX_ADDENDS 4
struct AccList
{
	f3 a[DNB_MAX_ADDENDS];
	uint32_t n;
	uint32_t tiles[DNB_MAX_ADDENDS], numTiles; /* visible-bit propagations, applied only once the voxel's rays are known to stand */
	DNB_FN void hit_tile(const DnbScene&, uint32_t hitIndex)
	{
#pragma unroll
		for(uint32_t k = 0; k < DNB_MAX_ADDENDS; k++)
			if(k == numTiles)
				tiles[k] = hitIndex;
		numTiles++;
	}
	DNB_FN void add(f3 x)
	{
#pragma unroll
		for(uint32_t k = 0; k < DNB_MAX_ADDENDS; k++) /* static indexing keeps the list in registers */
			if(k == n)
				a[k] = x;
		n++;
	}
};

DNB_FN uint32_t encode_rgba(uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
	return ((x & 0xFFu) << 24) | ((y & 0xFFu) << 16) | ((z & 0xFFu) << 8) | (w & 0xFFu);
}

/* LI:65-80 */
template <bool COUNT, class ACC>
DNB_FN void shadow_ray(const DnbScene& S, LightCtx& cx, f3 rayPos, uint32_t sample, ACC& color)
{
	const f3 sunDir = ld3(c_light.sunDir);
	f3 dir;
	if(cx.firstSample)
		dir = sunDir + DNB_EPSILON;
	else
		dir = normalize3(sunDir * c_light.shadowSoftness + ld3(c_light.shadowBall[sample])) + DNB_EPSILON;

	f3 tmpNormal = splat3(0.0f), colorAdd;
	float colorMult;
	if(!trace_ray<false, COUNT, true>(S, cx.st, cx.lc, dir, rcp3(dir), rayPos, true, tmpNormal, colorAdd, colorMult))
		color.add(ld3(S.sunStrength) * colorMult + colorAdd);
}

/* LI:83-146 */
template <bool COUNT, class ACC>
DNB_FN void specular_ray(const DnbScene& S, LightCtx& cx, f3 rayPos, f3 rayDir, f3 albedo, uint32_t reflectType, ACC& color)
{
	f3 lastPos = rayPos;
	f3 multiplier = albedo;
	const f3 sunDir = ld3(c_light.sunDir);

	for(uint32_t i = 0; i < c_light.specularBounceLimit; i++)
	{
		f3 colorAdd, tmpNormal = splat3(0.0f);
		float colorMult;
		if(trace_ray<false, COUNT>(S, cx.st, cx.lc, rayDir, rcp3(rayDir), rayPos, true, tmpNormal, colorAdd, colorMult))
		{
			/* LI:101-105: a visible chunk makes the chunks it reflects visible */
			i3 hp = toi3(rayPos);
			if(cx.sourceVisible && in_map_bounds(S, hp))
			{
				color.hit_tile(S, (uint32_t)hp.x + S.mapSize[0] * ((uint32_t)hp.y + S.mapSize[1] * (uint32_t)hp.z));
			}

			/* LI:108-110: adjacent hit = occluded */
			f3 dist = abs3(floor3(rayPos * 8.0f) - floor3(lastPos * 8.0f));
			if(dot3(dist, dist) <= 1.0f)
				return;

			const uint4 rec = cx.st.vox;
			DnbMaterial hitMaterial = load_material(S, vox_material(rec));
			f3 hitAlbedo = vox_albedo(rec);
			f3 hitDiffuse = vox_diffuse(rec) * (1.0f - hitMaterial.specular);

			if(hitMaterial.emissive)
			{
				color.add(((hitAlbedo * colorMult + colorAdd) * multiplier) * albedo);
				return;
			}

			f3 hitColor = hitDiffuse * hitAlbedo;
			color.add((hitColor * colorMult + colorAdd) * multiplier);
			if(hitMaterial.specular == 0.0f)
				return;

			multiplier = multiplier * ((hitAlbedo * colorMult) * hitMaterial.specular);
			reflectType = hitMaterial.reflectType;
			lastPos = rayPos;
			rayDir = reflect3(rayDir, vox_normal(rec));
		}
		else if(dot3(rayDir, sunDir) > 0.99f)
		{
			color.add(ld3(S.sunStrength) * colorMult + colorAdd);
			return;
		}
		else
		{
			f3 base = (reflectType == 1u) ? sky_color(S, rayDir) : ld3(S.sunStrength);
			color.add((base * colorMult + colorAdd) * multiplier);
			return;
		}
	}
}

/* LI:149-203 */
template <bool COUNT, class ACC>
DNB_FN void diffuse_ray(const DnbScene& S, LightCtx& cx, f3 normal, f3 rayPos, uint4 initialVoxel, uint32_t sample, ACC& color)
{
	cx.st.vox = initialVoxel;
	f3 hitNormal = normal;
	DnbMaterial hitMaterial;
	hitMaterial.specular = 0.0f;
	hitMaterial.shininess = 0;
	hitMaterial.emissive = 0;

	f3 newColor = splat3(1.0f);
	const f3 lastPos = rayPos; /* never advanced: adjacency is tested against the first origin (LI:157,182) */
	f3 lastDir = splat3(0.0f);
	const f3 sunDir = ld3(c_light.sunDir);

	for(uint32_t i = 0; i < c_light.diffuseBounceLimit; i++)
	{
		f3 dir;
		if(i > 0 && c_light.glossyChoice[sample][i] < hitMaterial.specular)
			dir = normalize3(reflect3(lastDir, hitNormal) * (float)hitMaterial.shininess + ld3(c_light.glossyBall[i]));
		else if(cx.firstSample)
			dir = normalize3(hitNormal) + DNB_EPSILON;
		else
			dir = normalize3(hitNormal + ld3(c_light.diffuseBall[sample][i])) + DNB_EPSILON;

		f3 colorAdd, tmpNormal = splat3(0.0f);
		float colorMult;
		bool hit = trace_ray<false, COUNT>(S, cx.st, cx.lc, dir, rcp3(dir), rayPos, true, tmpNormal, colorAdd, colorMult);

		const uint4 rec = cx.st.vox;
		hitNormal = vox_normal(rec);
		hitMaterial = load_material(S, vox_material(rec));

		if(hit)
		{
			f3 dist = abs3(floor3(lastPos * 8.0f) - floor3(rayPos * 8.0f));
			if(dot3(dist, dist) < 1.0f)
				return;

			f3 through = vox_albedo(rec) * colorMult + colorAdd;
			if(hitMaterial.emissive)
			{
				color.add(newColor * through);
				return;
			}
			newColor = newColor * through;
		}
		else
		{
			float ndl = fmaxf(dot3(dir, sunDir), 0.0f);
			color.add(((newColor * ndl) * ld3(S.sunStrength)) * colorMult + colorAdd);
			return;
		}

		lastDir = dir;
	}
}

/* n-th set bit of the 512-bit mask -> local voxel index (same answer as get_voxel_position, SH:172-224,
 * whose `<` partial-count quirk only changes where its scan starts) */
DNB_FN int nth_voxel(const uint32_t* mask, const uint16_t* prefix, uint32_t numVoxels, uint32_t voxNum)
{
	if(voxNum >= numVoxels)
		return -1;
	int w = 0;
#pragma unroll
	for(int i = 1; i < 16; i++)
		if((uint32_t)prefix[i] <= voxNum)
			w = i;
	/* prefix is non-decreasing, so w is the last word starting at or before voxNum; skip empty words backwards is not
	 * needed: the word containing the voxel is the last one whose prefix <= voxNum AND has a bit there */
	uint32_t inWord = voxNum - (uint32_t)prefix[w];
	return w * 32 + (int)__fns(mask[w], 0, (int)inWord + 1);
}

/* the three staged words of one voxel -> row `at` of every target array (1 target, or every replica's over NVLink) */
DNB_FN void stage_words(const DnbStagingTargets& T, size_t at, uint32_t w1, uint32_t w2, uint32_t w3)
{
	/* static indices (the table lives in the kernel's parameter bank; a dynamic index would copy it to local memory) */
#pragma unroll
	for(uint32_t p = 0; p < DNB_MAX_PEERS; p++)
		if(p < T.count)
		{
			uint32_t* out = T.dst[p] + at;
			out[0] = w1;
			out[32] = w2;
			out[64] = w3;
		}
}

/* the request count and this launch's share of the 4-request CTAs (layout.h DnbWork) */
DNB_FN uint32_t work_requests(const DnbWork& W)
{
	if(!W.count)
		return W.limit;
	const uint32_t n = *reinterpret_cast<const volatile uint32_t*>(W.count);
	return (W.limit && n > W.limit) ? W.limit : n;
}
DNB_FN uint32_t work_ctas(const DnbWork& W, uint32_t numRequests)
{
	if(W.numCtas)
		return W.numCtas;
	const uint32_t total = (numRequests + 3u) / 4u;
	return total > W.firstCta ? (total - W.firstCta + W.ctaStride - 1u) / W.ctaStride : 0u;
}

/* registers: ptxas settles at 96 (5 CTAs = 20 warps per SM) with a few spills; both fewer registers (more warps, more spills) and
 * more registers (no spills, 16 warps) measured slower on B200 (0.49 / 0.54 ms vs 0.45 ms on config 2) */
/* LI:266-278: clamp, quantise (round() = nearest even, N6) and pack the three lit words of a voxel */
DNB_FN void pack_lit_words(uint4 rec, f3 specLight, f3 diffuseLight, uint32_t& w1, uint32_t& w2, uint32_t& w3)
{
	specLight = clamp01(specLight);
	diffuseLight = clamp01(diffuseLight);
	const f3 albedo = vox_albedo(rec);
	const uint32_t wx = (uint32_t)rintf(diffuseLight.x * 65535.0f);
	const uint32_t wy = (uint32_t)rintf(diffuseLight.y * 65535.0f);
	const uint32_t wz = (uint32_t)rintf(diffuseLight.z * 65535.0f);
	w1 = encode_rgba((uint32_t)rintf(albedo.x * 255.0f), (uint32_t)rintf(albedo.y * 255.0f), (uint32_t)rintf(albedo.z * 255.0f), (uint32_t)rintf(specLight.x * 255.0f));
	w2 = encode_rgba((uint32_t)rintf(specLight.y * 255.0f), (uint32_t)rintf(specLight.z * 255.0f), (wx >> 8) & 0xFFu, wx & 0xFFu);
	w3 = encode_rgba((wy >> 8) & 0xFFu, wy & 0xFFu, (wz >> 8) & 0xFFu, wz & 0xFFu);
}

/* the whole of one shader invocation after its set-up (LI:234-278), every ray traced by THIS thread in the shader's order */
template <bool COUNT>
DNB_FN void light_voxel(const DnbScene& S, LightCtx& cx, uint4 rec, const DnbMaterial& material, f3 rayPos, float indirectSamples, uint32_t& w1, uint32_t& w2, uint32_t& w3)
{
	const f3 normal = vox_normal(rec);
	const f3 albedo = vox_albedo(rec);
	AccSum spec, diff;
	spec.c = splat3(0.0f);
	diff.c = splat3(0.0f);

	/* LI:239-251 */
	const f3 viewDir = rayPos - ld3(c_light.camPos);
	if(material.specular > 0.0f && dot3(viewDir, normal) < 0.0f && material.reflectType <= 1u)
	{
		const f3 reflected = reflect3(normalize3(viewDir), normal);
		for(int i = 0; i < 15; i++)
		{
			f3 specDir = normalize3(reflected * (float)material.shininess + ld3(c_spherePoints[i])) + DNB_EPSILON;
			specular_ray<COUNT>(S, cx, rayPos, specDir, albedo, material.reflectType, spec);
		}
		spec.c = div3(spec.c, 15.0f);
	}

	/* LI:254-264 */
	if(material.specular < 1.0f)
	{
		const f3 ambient = ld3(S.ambient);
		for(uint32_t i = 0; i < c_light.numDiffuseSamples; i++)
		{
			diff.c = diff.c + ambient;
			diffuse_ray<COUNT>(S, cx, normal + DNB_EPSILON, rayPos, rec, i, diff);
			shadow_ray<COUNT>(S, cx, rayPos, i, diff);
		}
		const float denom = indirectSamples + (float)c_light.numDiffuseSamples;
		diff.c = div3(vox_diffuse(rec) * indirectSamples + diff.c, denom);
	}
	pack_lit_words(rec, spec.c, diff.c, w1, w2, w3);
}

/* SKIP_SPECULAR: leave the voxels that trace specular rays to somebody else (light_spread.cuh) */
template <bool COUNT, bool SKIP_SPECULAR = false>
DNB_FN void light_request(const DnbScene& S, const uint32_t* __restrict__ requests, uint32_t r, uint32_t warp, uint32_t lane, const DnbStagingTargets& T, DnbSlot* s_slot);

/* LI:239-242: does this voxel trace its 15 specular rays? */
DNB_FN bool traces_specular(const DnbMaterial& material, f3 rayPos, f3 normal)
{
	const f3 viewDir = rayPos - ld3(c_light.camPos);
	return material.specular > 0.0f && dot3(viewDir, normal) < 0.0f && material.reflectType <= 1u;
}

/* grid-stride over this launch's CTAs of 4 requests: the grid is sized from what the host knows (an upper bound and the last count it
 * saw), the real count is read on the device */
template <bool COUNT>
__global__ void __launch_bounds__(128) dn_light_kernel(DnbScene S, const uint32_t* __restrict__ requests, DnbWork W, DnbStagingTargets T)
{
	__shared__ DnbSlot s_slot[4];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const uint32_t numRequests = work_requests(W);
	const uint32_t numCtas = work_ctas(W, numRequests);
	for(uint32_t k = blockIdx.x; k < numCtas; k += gridDim.x)
	{
		const uint32_t r = (W.firstCta + k * W.ctaStride) * 4 + warp;
		if(r < numRequests)
			light_request<COUNT>(S, requests, r, warp, lane, T, s_slot);
		__syncwarp();
	}
}

template <bool COUNT, bool SKIP_SPECULAR>
DNB_FN void light_request(const DnbScene& S, const uint32_t* __restrict__ requests, uint32_t r, uint32_t warp, uint32_t lane, const DnbStagingTargets& T, DnbSlot* s_slot)
{
	const uint32_t request = __ldg(requests + r);
	const uint32_t mapIndex = request >> 4;
	const size_t at = (size_t)r * 96u + lane;

	const uint32_t slotId = __ldg(S.tileSlot + mapIndex) - 1u;
	if(slotId == 0xFFFFFFFFu)
	{
		/* the chunk was removed between the request and the dispatch: nothing to light */
		stage_words(T, at, 0, 0, 0);
		return;
	}

	/* stage the 128-byte chunk slot: one coalesced load per warp */
	reinterpret_cast<uint32_t*>(&s_slot[warp])[lane] = __ldg(reinterpret_cast<const uint32_t*>(S.slots + slotId) + lane);
	__syncwarp();
	const DnbSlot& slot = s_slot[warp];

	const uint32_t voxNum = lane + (request & 15u) * 32u;
	const int local = nth_voxel(slot.mask, slot.prefix, slot.numVoxels, voxNum);
	if(local < 0)
	{
		stage_words(T, at, 0, 0, 0);
		return;
	}

	LightCtx cx;
	ray_state_reset(cx.st);
	cx.lc = DnbCounters{0, 0, 0, 0, 0, 0, 0};
	cx.firstSample = false;
	cx.sourceVisible = (__ldg(S.visible + (mapIndex >> 5)) >> (mapIndex & 31u)) & 1u;

	const i3 chunkPos = {local & 7, (local >> 3) & 7, local >> 6};
	const i3 mapPos = {slot.pos[0], slot.pos[1], slot.pos[2]};

	const uint4 rec = __ldg(S.records + (slot.voxelBase + voxNum));
	const f3 normal = vox_normal(rec);
	const DnbMaterial material = load_material(S, vox_material(rec));

	const uint32_t ns = slot.numSamples; /* pre-dispatch value for every group of the chunk (N2) */
	const float indirectSamples = (float)(ns < c_light.maxDiffuseSamples ? ns : c_light.maxDiffuseSamples);
	if(indirectSamples == 0.0f)
		cx.firstSample = true;

	/* LI:231-232 */
	f3 rayPos = (tof3(chunkPos) * 0.125f + tof3(mapPos)) + 0.0625f;
	rayPos = rayPos + normal * (0.0625f - DNB_EPSILON);
	if(SKIP_SPECULAR && traces_specular(material, rayPos, normal))
		return;

	uint32_t w1, w2, w3;
	light_voxel<COUNT>(S, cx, rec, material, rayPos, indirectSamples, w1, w2, w3);
	stage_words(T, at, w1, w2, w3);

	if(COUNT)
	{
		atomicAdd(&S.counters->rays, cx.lc.rays);
		atomicAdd(&S.counters->tiles, cx.lc.tiles);
		atomicAdd(&S.counters->chunks, cx.lc.chunks);
		atomicAdd(&S.counters->voxelSteps, cx.lc.voxelSteps);
		atomicAdd(&S.counters->records, cx.lc.records);
		atomicAdd(&S.counters->voxelsLit, 1ull);
	}
}

/* phase 2: one warp per request, persistent grid-stride warps.  The lit-voxel count (the metric's numerator) is summed per warp and
 * per CTA and added with ONE atomic per CTA: with one atomic per request the 10^7 same-address atomics of a full-size dispatch
 * WERE the kernel (11.4 ms for 10.9 M requests on B200, 1.2 TB/s; the streaming itself needs a quarter of that). */
__global__ void __launch_bounds__(256) dn_commit_kernel(const uint32_t* __restrict__ tileSlot, DnbSlot* __restrict__ slots, uint4* __restrict__ records, uint32_t* __restrict__ visible,
                                                        const uint32_t* __restrict__ requests, DnbWork W, const uint32_t* __restrict__ staging,
                                                        unsigned long long* __restrict__ litCounter)
{
	__shared__ uint32_t s_lit[8];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t numRequests = work_requests(W);
	uint32_t lit = 0;
	for(uint32_t r = blockIdx.x * 8 + warp; r < numRequests; r += gridDim.x * 8)
	{
		const uint32_t request = __ldg(requests + r);
		const uint32_t mapIndex = request >> 4, group = request & 15u;
		const uint32_t slotId = __ldg(tileSlot + mapIndex) - 1u;
		if(slotId == 0xFFFFFFFFu)
			continue;

		DnbSlot* slot = slots + slotId;
		const uint32_t numVoxels = slot->numVoxels, base = slot->voxelBase;
		const uint32_t voxNum = lane + group * 32u;
		if(voxNum < numVoxels)
		{
			const uint32_t* in = staging + (size_t)r * 96u;
			uint4 rec;
			rec.x = records[base + voxNum].x;
			rec.y = __ldg(in + lane);
			rec.z = __ldg(in + 32 + lane);
			rec.w = __ldg(in + 64 + lane);
			records[base + voxNum] = rec;
		}
		if(group * 32u < numVoxels)
		{
			if(lane == 0)
			{
				/* LI:281; the groups of a chunk all clear the same bit: whoever still sees it set clears it */
				const uint32_t bit = 1u << (mapIndex & 31u);
				if(*reinterpret_cast<volatile uint32_t*>(visible + (mapIndex >> 5)) & bit)
					atomicAnd(visible + (mapIndex >> 5), ~bit);
				if(group == 0)
					slot->numSamples = slot->numSamples + 1u;
			}
			const uint32_t left = numVoxels - group * 32u;
			lit += left < 32u ? left : 32u;
		}
	}
	if(lane == 0)
		s_lit[warp] = lit;
	__syncthreads();
	if(threadIdx.x == 0)
	{
		unsigned long long total = 0;
		for(int w = 0; w < 8; w++)
			total += s_lit[w];
		if(total)
			atomicAdd(litCounter, total); /* the metric's numerator: voxel lighting updates */
	}
}

__global__ void dn_merge_visible_kernel(uint32_t* __restrict__ visible, uint32_t* __restrict__ propagate, uint32_t numWords)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= numWords)
		return;
	const uint32_t p = propagate[i];
	if(p)
	{
		visible[i] |= p;
		propagate[i] = 0;
	}
}

/* the same over peer memory: visible |= every replica's propagate bitmap.  Nothing is cleared here -- the peers may still be
 * reading this replica's bitmap; the host clears it at the start of the next lighting pass, behind a barrier (engine.cpp). */
__global__ void dn_merge_visible_peers_kernel(DnbPeerTable T, uint32_t* __restrict__ visible, uint32_t numWords)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if(i >= numWords)
		return;
	uint32_t add = 0;
	for(uint32_t p = 0; p < T.world; p++)
		add |= *reinterpret_cast<const volatile uint32_t*>(T.propagate[p] + i);
	if(add)
		visible[i] |= add;
}

#include "light_flat.cuh"
#include "light_wave.cuh"
#include "light_spread.cuh"

static DnbFlatTuning g_flatTuning = {0, 0, 0, 0, 0, 0};

/* scheduling knobs of the persistent kernel (experiments; results do not depend on them) */
extern "C" void DN_b200_set_flat_tuning(int budget, int endLanes, int patience)
{
	g_flatTuning.budget = budget;
	g_flatTuning.endLanes = endLanes;
	g_flatTuning.patience = patience;
}

extern "C" cudaError_t dnb_upload_light_params(const DnbLightParams* params, cudaStream_t stream)
{
	return cudaMemcpyToSymbolAsync(c_light, params, sizeof(DnbLightParams), 0, cudaMemcpyHostToDevice, stream);
}

/* gridCtas: CTAs to launch for the warp-per-request kernel (any number >= 1 is correct: the kernel strides over its share) and the
 * ceiling of the persistent kernel's grid */
extern "C" cudaError_t dnb_launch_light(const DnbScene* scene, const uint32_t* requests, const DnbWork* work, uint32_t gridCtas,
                                        const DnbStagingTargets* targets, uint32_t* flatCounter, cudaStream_t stream)
{
	if(gridCtas == 0)
		return cudaSuccess;
	if(scene->counters)
		{ DNB_LAUNCHED(1); dn_light_kernel<true><<<gridCtas, 128, 0, stream>>>(*scene, requests, *work, *targets); }
	else if(flatCounter)
	{
		/* persistent warps: enough CTAs to fill the machine, each lane pulls voxels from the work counter */
		static int ctasPerDevice = 0;
		if(ctasPerDevice == 0)
		{
			int dev = 0, sms = 148, perSm = 4;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			if(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, dn_light_flat_kernel, FLAT_WARPS * 32, 0) != cudaSuccess || perSm < 1)
				perSm = 4;
			ctasPerDevice = sms * perSm;
		}
		cudaError_t e = cudaMemsetAsync(flatCounter, 0, sizeof(uint32_t), stream);
		if(e != cudaSuccess)
			return e;
		const uint32_t grid = gridCtas < (uint32_t)ctasPerDevice ? gridCtas : (uint32_t)ctasPerDevice;
		DnbFlatTuning& tuning = g_flatTuning;
		if(tuning.budget == 0)
		{
			auto knob = [](const char* name, int dflt) { const char* e = getenv(name); return e && atoi(e) > 0 ? atoi(e) : dflt; };
			tuning.budget = knob("DN_B200_FLAT_BUDGET", 24);
			tuning.endLanes = knob("DN_B200_FLAT_END", 20);
			tuning.patience = knob("DN_B200_FLAT_PATIENCE", 48);
		}
		static int endMax = -1;
		if(endMax < 0)
		{
			const char* e = getenv("DN_B200_FLAT_ENDMAX");
			endMax = e ? atoi(e) != 0 : 0;
		}
		tuning.endMax = endMax;
		static int keep = -1;
		if(keep < 0)
		{
			const char* e = getenv("DN_B200_FLAT_KEEP");
			keep = e && atoi(e) > 0 && atoi(e) <= 8 ? atoi(e) : 2;
		}
		tuning.keep = keep;
		static int run = -1;
		if(run < 0)
		{
			const char* e = getenv("DN_B200_FLAT_RUN");
			run = e && atoi(e) > 0 ? atoi(e) : 0;
		}
		tuning.run = run;
		{ DNB_LAUNCHED(1); dn_light_flat_kernel<<<grid, FLAT_WARPS * 32, 0, stream>>>(*scene, requests, *work, flatCounter, *targets, tuning); }
	}
	else
		{ DNB_LAUNCHED(1); dn_light_kernel<false><<<gridCtas, 128, 0, stream>>>(*scene, requests, *work, *targets); }
	return cudaGetLastError();
}

/* Multi-GPU, wavefront kernels: their serve kernel finishes voxels one at a time, so storing each voxel's three words straight into
 * every replica means 4-byte stores scattered over NVLink (one packet each; measured 45 % slower at 4 replicas than at 1).  They
 * therefore stage into their own replica only, and this kernel then pushes the rows of the CTAs the replica owns (4 requests = 1536
 * contiguous bytes each) to the other replicas with coalesced 16-byte stores.  (The warp-per-request kernel keeps its fused 128-byte
 * row stores, and the persistent kernel its per-voxel stores, which overlap with its ray tracing.) */
__global__ void __launch_bounds__(128) dn_push_staging_kernel(const uint4* __restrict__ own, DnbStagingTargets T, uint32_t self, DnbWork W)
{
	const uint32_t numRequests = work_requests(W);
	const uint32_t numCtas = work_ctas(W, numRequests);
	for(uint32_t k = blockIdx.x; k < numCtas; k += gridDim.x)
	{
		const uint32_t cta = W.firstCta + k * W.ctaStride;
		const uint32_t firstRequest = cta * 4u;
		if(firstRequest >= numRequests)
			break;
		const uint32_t rows = numRequests - firstRequest < 4u ? numRequests - firstRequest : 4u;
		if(threadIdx.x < rows * 24u) /* 96 words = 24 uint4 per request */
		{
			const size_t at = (size_t)cta * 96u + threadIdx.x;
			const uint4 v = own[at];
#pragma unroll
			for(uint32_t p = 0; p < DNB_MAX_PEERS; p++)
				if(p < T.count && p != self)
					reinterpret_cast<uint4*>(T.dst[p])[at] = v;
		}
	}
}

extern "C" cudaError_t dnb_launch_push_staging(const DnbStagingTargets* peers, uint32_t self, const DnbWork* work, uint32_t gridCtas, cudaStream_t stream)
{
	if(gridCtas == 0 || peers->count < 2)
		return cudaSuccess;
	const uint32_t grid = gridCtas < 148u * 16u ? gridCtas : 148u * 16u;
	{ DNB_LAUNCHED(1); dn_push_staging_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const uint4*>(peers->dst[self]), *peers, self, *work); }
	return cudaGetLastError();
}

/* boundRequests: what the host knows the request count cannot exceed (sizes the persistent grid; 0 = nothing to commit) */
extern "C" cudaError_t dnb_launch_commit(const DnbScene* scene, DnbSlot* slots, uint4* records, const uint32_t* requests, const DnbWork* work, uint32_t boundRequests, const uint32_t* staging,
                                         unsigned long long* litCounter, const DnbPeerTable* peers, cudaStream_t stream)
{
	const uint32_t numRequests = boundRequests;
	if(numRequests > 0)
	{
		/* persistent: 8 CTAs of 8 warps per SM at most (the kernel streams; its warps only wait on memory) */
		static int maxCtas = 0;
		if(maxCtas == 0)
		{
			int dev = 0, sms = 148;
			cudaGetDevice(&dev);
			cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
			maxCtas = sms * 8;
		}
		const uint32_t ctas = std::min<uint32_t>((numRequests + 7) / 8, (uint32_t)maxCtas);
		{ DNB_LAUNCHED(1); dn_commit_kernel<<<ctas, 256, 0, stream>>>(scene->tileSlot, slots, records, scene->visible, requests, *work, staging, litCounter); }
		cudaError_t e = cudaGetLastError();
		if(e != cudaSuccess)
			return e;
	}
	const uint32_t words = (scene->numTiles + 31) / 32;
	if(peers)
		{ DNB_LAUNCHED(1); dn_merge_visible_peers_kernel<<<(words + 255) / 256, 256, 0, stream>>>(*peers, scene->visible, words); }
	else
		{ DNB_LAUNCHED(1); dn_merge_visible_kernel<<<(words + 255) / 256, 256, 0, stream>>>(scene->visible, scene->propagate, words); }
	return cudaGetLastError();
}
