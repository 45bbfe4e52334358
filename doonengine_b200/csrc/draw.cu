/* draw.cu -- per-pixel primary ray cast, shaded from the lit voxel it hits.
 *
 * Follows /root/reference/assets/shaders/voxelDraw.comp main() (DR:63-147), background_color (DR:24-30, sky
 * branch) and voxel_color (DR:32-61); dispatched like voxel.c:879 over (w/16) x (h/16) groups of 16x16 pixels
 * (pixels outside are left untouched).  Writes a LINEAR RGBA32F framebuffer in device memory:
 * RGB = pow(colour, 0.4545), A = distance to the hit in tiles (-1 when nothing is hit: the reference leaves
 * it undefined for rays that miss the map box, oracle.h N7).
 *
 * Launch shape: one thread per pixel, 256-thread CTAs covering 32x8 pixels as eight 8x4 warp footprints, so
 * the 32 primary rays of a warp stay spatially compact (they walk the same tiles and chunk slots and share
 * L1 lines); the pixel a thread owns does not change any result.  Bound: latency of dependent gathers
 * (occupancy word -> slot id -> mask word -> record), not HBM; see DESIGN.md.
 */
#include "kernels.h"
#include "trace.cuh"

/* SH:261-273 */
DNB_FN void intersect_aabb(f3 invRayDir, f3 rayPos, f3 boxMin, f3 boxMax, float& tNear, float& tFar)
{
	f3 tMin = (boxMin - rayPos) * invRayDir;
	f3 tMax = (boxMax - rayPos) * invRayDir;
	f3 t1 = min3v(tMin, tMax);
	f3 t2 = max3v(tMin, tMax);
	tNear = fmaxf(fmaxf(t1.x, t1.y), t1.z);
	tFar = fminf(fminf(t2.x, t2.y), t2.z);
}

/* SH:276-284 */
DNB_FN f3 normal_aabb(f3 hitPos, f3 boxMin, f3 boxMax)
{
	f3 c = (boxMin + boxMax) * 0.5f;
	f3 p = hitPos - c;
	f3 d = (boxMax - boxMin) * 0.5f;
	const float bias = 1.0f + DNB_EPSILON;
	f3 q = mk3(p.x / d.x * bias, p.y / d.y * bias, p.z / d.z * bias);
	return normalize3(trunc3(q));
}

/* DR:24-30 with maxDepth < 0 */
DNB_FN f3 background_color(const DnbScene& S, f3 rayDir)
{
	f3 s = sky_color(S, rayDir);
	return mk3(powf(s.x, 2.2f), powf(s.y, 2.2f), powf(s.z, 2.2f));
}

/* DR:32-61 */
DNB_FN f3 voxel_color(const DnbScene& S, uint32_t viewMode, uint4 rec, f3 colorAdd, float colorMult, f3 hitNormal)
{
	DnbMaterial material = load_material(S, vox_material(rec));
	f3 albedo = vox_albedo(rec);
	f3 spec = vox_spec(rec) * material.specular;
	f3 diffuse = vox_diffuse(rec) * (1.0f - material.specular);

	switch(viewMode)
	{
	case 0:
	{
		f3 solid = material.emissive ? albedo : (diffuse * albedo + spec);
		return solid * colorMult + colorAdd;
	}
	case 1: return albedo;
	case 2: return material.emissive ? albedo : diffuse;
	case 3: return material.emissive ? albedo : spec;
	case 4: return abs3(vox_normal(rec));
	case 5: return abs3(hitNormal);
	}
	return splat3(0.0f);
}

/* 4 CTAs per SM = 64 registers (32 warps per SM) with ~100 bytes of spills: 3-7 % faster than 80 registers / 3 CTAs on B200 */
#ifndef DRAW_MIN_BLOCKS
#define DRAW_MIN_BLOCKS 4
#endif
template <bool COUNT>
__global__ void __launch_bounds__(256, DRAW_MIN_BLOCKS) dn_draw_kernel(DnbScene S, DnbDrawParams P, float4* __restrict__ image, float4* __restrict__ mirror, DnbHit* __restrict__ hits)
{
	/* 32x8 pixel CTA made of 4x2 warp footprints of 8x4 pixels; blockIdx.y counts 8-row strips of the 16-pixel group rows
	 * rowBegin, rowBegin + rowStride, ... this launch owns */
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int px = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
	const int groupRow = P.rowBegin + (int)(blockIdx.y >> 1) * P.rowStride;
	const int py = groupRow * 16 + (int)(blockIdx.y & 1u) * 8 + (warp >> 2) * 4 + (lane >> 3);
	if(px >= (P.width / 16) * 16 || groupRow >= P.rowEnd)
		return;

	RayState st;
	ray_state_reset(st);
	DnbCounters lc = {0, 0, 0, 0, 0, 0, 0};

	f3 finalColorAdd = splat3(0.0f);
	float finalColorMult = 1.0f;
	f3 finalColor;
	float finalDepth = -1.0f;

	/* DR:75-83 */
	float sx = (float)px / (float)P.width * 2.0f - 1.0f;
	float sy = (float)py / (float)P.height * 2.0f - 1.0f;

	const float origin4[4] = {0.0f, 0.0f, 0.0f, 1.0f};
	float rp[4];
	mat4_mul_vec4(P.invView, origin4, rp);
	f3 rayPos = mk3(rp[0], rp[1], rp[2]);
	const f3 orgRayPos = rayPos;

	/* (invCenteredViewMat * invProjectionMat) * vec4(screenPos, 0, 1): the product is formed first, as GLSL does */
	float m[16];
#pragma unroll
	for(int c = 0; c < 4; c++)
#pragma unroll
		for(int r = 0; r < 4; r++)
			m[c * 4 + r] = P.invCenteredView[0 * 4 + r] * P.invProjection[c * 4 + 0] + P.invCenteredView[1 * 4 + r] * P.invProjection[c * 4 + 1] +
			               P.invCenteredView[2 * 4 + r] * P.invProjection[c * 4 + 2] + P.invCenteredView[3 * 4 + r] * P.invProjection[c * 4 + 3];
	const float sp4[4] = {sx, sy, 0.0f, 1.0f};
	float rd[4];
	mat4_mul_vec4(m, sp4, rd);
	f3 rayDir = normalize3(mk3(rd[0], rd[1], rd[2])) + DNB_EPSILON;
	f3 invRayDir = rcp3(rayDir);

	const f3 boxMax = mk3((float)S.mapSize[0], (float)S.mapSize[1], (float)S.mapSize[2]);
	float tNear, tFar;
	intersect_aabb(invRayDir, rayPos, splat3(0.0f), boxMax, tNear, tFar);

	DnbHit hit = {0, 0, 0, 0};

	if(tNear > tFar || tFar < 0.0f)
	{
		finalColor = background_color(S, rayDir);
	}
	else
	{
		if(tNear > 0.0f)
			rayPos = rayPos + rayDir * (tNear + DNB_EPSILON);
		f3 finalNormal = normal_aabb(rayPos, splat3(0.0f), boxMax);

		if(trace_ray<true, COUNT>(S, st, lc, rayDir, invRayDir, rayPos, false, finalNormal, finalColorAdd, finalColorMult))
		{
			/* DR:119-121: mark the tile under the hit position visible */
			f3 fl = floor3(rayPos);
			if(fl.x >= 0.0f && fl.y >= 0.0f && fl.z >= 0.0f && fl.x < boxMax.x && fl.y < boxMax.y && fl.z < boxMax.z)
			{
				uint32_t index = (uint32_t)fl.x + S.mapSize[0] * ((uint32_t)fl.y + S.mapSize[1] * (uint32_t)fl.z);
				uint32_t bit = 1u << (index & 31u);
				/* most pixels of a warp hit an already-marked tile: test before the atomic */
				if(!(__ldcg(S.visible + (index >> 5)) & bit))
					atomicOr(S.visible + (index >> 5), bit);
			}

			f3 dv = rayPos - orgRayPos;
			finalDepth = sqrtf(dot3(dv, dv));
			finalColor = voxel_color(S, P.viewMode, st.vox, finalColorAdd, finalColorMult, finalNormal);

			hit.status = 2;
			hit.mapIndex = st.hitMapIndex;
			hit.localIndex = st.hitLocalIndex;
			hit.recordIndex = st.hitRecord;
		}
		else
		{
			finalColor = background_color(S, rayDir) * finalColorMult + finalColorAdd;
			finalDepth = -1.0f;
			hit.status = 1;
		}
	}

	finalColor = mk3(powf(finalColor.x, 0.4545f), powf(finalColor.y, 0.4545f), powf(finalColor.z, 0.4545f));

	const size_t at = (size_t)py * (size_t)P.width + (size_t)px;
	const float4 pixel = make_float4(finalColor.x, finalColor.y, finalColor.z, finalDepth);
	image[at] = pixel;
	if(mirror)
		mirror[at] = pixel; /* the root replica's framebuffer in peer memory: the gather is fused into the draw */
	if(hits)
		hits[at] = hit;

	if(COUNT)
	{
		lc.pixels = 1;
		atomicAdd(&S.counters->rays, lc.rays);
		atomicAdd(&S.counters->tiles, lc.tiles);
		atomicAdd(&S.counters->chunks, lc.chunks);
		atomicAdd(&S.counters->voxelSteps, lc.voxelSteps);
		atomicAdd(&S.counters->records, lc.records);
		atomicAdd(&S.counters->pixels, lc.pixels);
	}
}

extern "C" cudaError_t dnb_launch_draw(const DnbScene* scene, const DnbDrawParams* params, float4* image, float4* mirror, DnbHit* hits, cudaStream_t stream)
{
	const int cols = (params->width / 16) * 16;
	const int stride = params->rowStride > 0 ? params->rowStride : 1;
	const int groups = params->rowEnd > params->rowBegin ? (params->rowEnd - params->rowBegin + stride - 1) / stride : 0;
	if(cols <= 0 || groups <= 0)
		return cudaSuccess;
	DnbDrawParams p = *params;
	p.rowStride = stride;
	dim3 grid((unsigned)((cols + 31) / 32), (unsigned)(groups * 2));
	if(scene->counters)
		{ DNB_LAUNCHED(1); dn_draw_kernel<true><<<grid, 256, 0, stream>>>(*scene, p, image, mirror, hits); }
	else
		{ DNB_LAUNCHED(1); dn_draw_kernel<false><<<grid, 256, 0, stream>>>(*scene, p, image, mirror, hits); }
	return cudaGetLastError();
}
